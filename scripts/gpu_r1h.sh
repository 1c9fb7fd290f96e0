#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_plugin.py -q > gpurun_out/pytest_h.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_h.log
tail -40 gpurun_out/pytest_h.log
