#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I neusky_b200/csrc -o /tmp/tc_probe tools/tc_probe.cu 2> gpurun_out/r4l_build.err || { tail gpurun_out/r4l_build.err; exit 1; }
{ for w in 4 8 16; do for x in 8 16 32; do for d in 1 2 4; do /tmp/tc_probe ldtm2 $w $x $d | grep probe; done; done; done; } > gpurun_out/r4l_ldtm2.log 2>&1
cat gpurun_out/r4l_ldtm2.log
