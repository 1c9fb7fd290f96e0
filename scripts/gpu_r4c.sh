#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/k4_phase_profile.py 200000 tc2 > gpurun_out/r4c_k4_phase.log 2>&1; echo rc=$?; cat gpurun_out/r4c_k4_phase.log
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I neusky_b200/csrc -o /tmp/tc_probe tools/tc_probe.cu 2> gpurun_out/r4c_probe_build.err && {
for a in "rate2 256 0 0" "rate2 128 0 0" "rate2 256 4 0" "rate2 128 8 0" "rate2 128 8 1" "rate 0 128 0" "rate 0 256 0"; do /tmp/tc_probe $a; done; } > gpurun_out/r4c_probe.log 2>&1; cat gpurun_out/r4c_probe.log; tail -3 gpurun_out/r4c_probe_build.err
