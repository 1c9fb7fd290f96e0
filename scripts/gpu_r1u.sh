#!/bin/bash
# DDF fitting pass parity (f2) + regression of the training path after the DDF autograd refactor
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ddf_fit.py tests/test_gpu_train.py -q > gpurun_out/r1u_pytest.log 2>&1; echo "pytest exit=$?"; tail -30 gpurun_out/r1u_pytest.log | cut -c1-600
