#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q > gpurun_out/pytest_tc4.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_tc4.log
tail -30 gpurun_out/pytest_tc4.log
timeout 120 python scripts/quick_tc_bench.py 200000 2048 > gpurun_out/quick_tc4.log 2>&1; echo "exit=$?" >> gpurun_out/quick_tc4.log
cat gpurun_out/quick_tc4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sky_shade_tc -s 1 -c 1 -o gpurun_out/prof_tc4 python scripts/quick_tc_bench.py 20000 2048 > gpurun_out/ncu_tc4.log 2>&1
