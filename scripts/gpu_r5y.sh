#!/bin/bash
# round 2, session 3: ncu --set full of the FINAL K4 kernel (CTA-pair tcgen05 kernel after the issuer rework), 100 k points x 1024 DDF directions
timeout 600 /usr/local/cuda/bin/ncu --set full --clock-control none --import-source on -k regex:sky_shade_tc2 -s 2 -c 1 -o gpurun_out/r5y_prof_k4_tc2 -f \
  python scripts/quick_tc_bench.py 100000 2048 tc2 > gpurun_out/r5y_ncu_k4.log 2>&1; echo "ncu k4 exit=$?"
ls -la gpurun_out/r5y_*.ncu-rep; tail -2 gpurun_out/r5y_ncu_k4.log
