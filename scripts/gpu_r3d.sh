#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3d_pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r3d_pytest_gpu.log; tail -4 gpurun_out/r3d_pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3d_smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/r3d_smoke.log; tail -2 gpurun_out/r3d_smoke.log
