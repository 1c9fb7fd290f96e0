#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_render.py -q -k config1 > gpurun_out/r3c_pytest.log 2>&1; echo "pytest exit=$?"; tail -8 gpurun_out/r3c_pytest.log | cut -c1-400
