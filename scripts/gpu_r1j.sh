#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q > gpurun_out/pytest_j.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_j.log
tail -5 gpurun_out/pytest_j.log | cut -c1-300
for i in tc tc2; do timeout 120 python scripts/quick_tc_bench.py 400000 2048 $i; done
