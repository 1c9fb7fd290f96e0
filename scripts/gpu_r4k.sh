#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/train_gap_probe.py > gpurun_out/r4k_train_gap.log 2>&1; echo rc=$?; tail -25 gpurun_out/r4k_train_gap.log | cut -c1-200
