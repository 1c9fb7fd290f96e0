#!/bin/bash
# light-sum shaders (f4) parity vs the reference's values and autograd gradients
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shaders.py -q > gpurun_out/r1v_pytest.log 2>&1; echo "pytest exit=$?"; tail -40 gpurun_out/r1v_pytest.log | cut -c1-700
