#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_plugin.py -x -q > gpurun_out/pytest_i.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_i.log
tail -30 gpurun_out/pytest_i.log | cut -c1-300
