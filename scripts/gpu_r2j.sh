#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j_train_launches.csv python bench.py --workload train --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2j_train_under_ncu.log 2>&1; echo "ncu exit=$?"
python scripts/summarise_launches.py gpurun_out/r2j_train_launches.csv gpurun_out/r2j_train_launch_summary.txt | head -48
