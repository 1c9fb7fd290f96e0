#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q > gpurun_out/pytest_tc2.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_tc2.log
tail -40 gpurun_out/pytest_tc2.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks_tc2.csv &
SMI=$!
timeout 120 python scripts/quick_tc_bench.py 200000 2048 > gpurun_out/quick_tc2.log 2>&1; echo "exit=$?" >> gpurun_out/quick_tc2.log
kill $SMI
cat gpurun_out/quick_tc2.log
sort gpurun_out/clocks_tc2.csv | uniq -c | sort -rn | head -5
# launch list + full capture of the tensor-core kernel
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_tc2.csv python scripts/quick_tc_bench.py 20000 2048 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sky_shade_tc -s 1 -c 1 -o gpurun_out/prof_tc2 python scripts/quick_tc_bench.py 20000 2048 > gpurun_out/ncu_tc2.log 2>&1
ls -la gpurun_out/
