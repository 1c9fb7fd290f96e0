#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sdf.py -x -q -s -k tc > gpurun_out/pytest_e.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_e.log
tail -40 gpurun_out/pytest_e.log
