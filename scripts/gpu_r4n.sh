#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/k4_phase_profile.py 200000 tc2 > gpurun_out/r4n_k4_phase.log 2>&1; echo rc=$?; cat gpurun_out/r4n_k4_phase.log
