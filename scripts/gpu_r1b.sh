#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/k4_phase_profile.py 200000 > gpurun_out/k4_phase.log 2>&1; echo "exit=$?" >> gpurun_out/k4_phase.log
cat gpurun_out/k4_phase.log
# launch list of the bench command restricted to the step's kernels (the torch weight-packing kernels at start-up are skipped)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'reni_|lambert|sky_shade|shade_finalize|FillFunctor|index|gather' -c 300 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# DRAM traffic of K4 at the full bench launch size (one pass)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sky_shade_tc -s 1 -c 1 --csv --log-file gpurun_out/k4_dram_full.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/k4_dram_full.log 2>&1
tail -3 gpurun_out/k4_dram_full.csv
