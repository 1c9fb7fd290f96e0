#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sky_shade_tc2 -s 1 -c 1 -o gpurun_out/r3a_prof_k4_tc2 -f python bench.py --steps 1 --warmup 1 --points 40000 --no-cpu-baseline > gpurun_out/r3a_ncu_k4.log 2>&1; echo "ncu k4 exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sdf_field_tc -s 1 -c 1 -o gpurun_out/r3a_prof_k2_tc -f python bench.py --workload eval --steps 1 --warmup 1 --height 360 --width 640 > gpurun_out/r3a_ncu_k2.log 2>&1; echo "ncu k2 exit=$?"
