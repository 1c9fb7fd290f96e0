#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/tf32_trunc_probe.py > gpurun_out/r1x_tf32_probe.log 2>&1; cat gpurun_out/r1x_tf32_probe.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1x_eval_launches.csv python bench.py --workload eval --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1x_eval_under_ncu.log 2>&1; echo "ncu exit=$?"
python scripts/summarise_launches.py gpurun_out/r1x_eval_launches.csv gpurun_out/r1x_eval_launch_summary.txt
