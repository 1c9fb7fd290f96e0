#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I neusky_b200/csrc -o /tmp/tc_probe tools/tc_probe.cu 2> gpurun_out/r4p_build.err || { tail gpurun_out/r4p_build.err; exit 1; }
{ for m in 0 1 2; do for w in 4 8 16; do timeout 60 /tmp/tc_probe sin $m $w | grep probe; done; done; } > gpurun_out/r4p_sin.log 2>&1
cat gpurun_out/r4p_sin.log
