#!/bin/bash
# round 2, session 3: DRAM traffic of a full-size K4 launch on the FINAL kernel + launch list of the headline bench command
NCU=/usr/local/cuda/bin/ncu
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sky_shade_tc2 -s 1 -c 1 --csv \
  --log-file gpurun_out/r5v_k4_tc2_dram_full.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-kernels > gpurun_out/r5v_b1.log 2>&1
tail -4 gpurun_out/r5v_k4_tc2_dram_full.csv
$NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r5v_bench_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-kernels > gpurun_out/r5v_b2.log 2>&1
python scripts/summarise_launches.py gpurun_out/r5v_bench_launches.csv > gpurun_out/r5v_bench_launch_summary.txt 2>&1
head -12 gpurun_out/r5v_bench_launch_summary.txt
