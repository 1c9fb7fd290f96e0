#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q > gpurun_out/pytest_tc1.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_tc1.log
tail -25 gpurun_out/pytest_tc1.log
timeout 120 python scripts/quick_tc_bench.py 20000 2048 > gpurun_out/quick_tc1.log 2>&1; echo "exit=$?" >> gpurun_out/quick_tc1.log
timeout 120 python scripts/quick_tc_bench.py 100000 2048 >> gpurun_out/quick_tc1.log 2>&1; echo "exit=$?" >> gpurun_out/quick_tc1.log
cat gpurun_out/quick_tc1.log
