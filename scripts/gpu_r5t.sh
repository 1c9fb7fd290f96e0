#!/bin/bash
# round 2, session 3: issue rate of the CTA-pair MMA with the A operand in TMEM (".ts" form) next to the shared-memory form, N = 128 / 256
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I neusky_b200/csrc -o /tmp/tc_probe tools/tc_probe.cu 2> gpurun_out/r5t_build.err || { tail gpurun_out/r5t_build.err; exit 1; }
{ for n in 128 256 64; do for m in 0 3; do timeout 60 /tmp/tc_probe rate2 $n 0 $m | grep probe; done; done; timeout 60 /tmp/tc_probe rate2 128 8 3 | grep probe; } > gpurun_out/r5t_rate2_ts.log 2>&1
cat gpurun_out/r5t_rate2_ts.log
