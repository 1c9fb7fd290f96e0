#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sdf.py tests/test_gpu_render.py -x -q > gpurun_out/pytest_c.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_c.log
tail -40 gpurun_out/pytest_c.log
