#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -x > gpurun_out/r1y_pytest.log 2>&1; echo "pytest exit=$?"; tail -40 gpurun_out/r1y_pytest.log | cut -c1-500
