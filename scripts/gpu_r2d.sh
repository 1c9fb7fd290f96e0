#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_train.json 2> gpurun_out/r2d_train.err; echo "train exit=$?"; cut -c1-300 gpurun_out/r2d_bench_train.json; tail -3 gpurun_out/r2d_train.err
timeout 600 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline --no-ddf-fit > gpurun_out/r2d_bench_train_nofit.json 2> gpurun_out/r2d_train2.err; echo "train exit=$?"; cut -c1-300 gpurun_out/r2d_bench_train_nofit.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_train_launches.csv python bench.py --workload train --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2d_train_under_ncu.log 2>&1; echo "ncu exit=$?"
python scripts/summarise_launches.py gpurun_out/r2d_train_launches.csv gpurun_out/r2d_train_launch_summary.txt | head -30
