#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for p in 0 1; do NSK_K4_PACE_NS=$p timeout 600 python bench.py --steps 3 --warmup 3 --no-extras --no-kernels --no-cpu-baseline > gpurun_out/r4z_bench.json 2> gpurun_out/r4z_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r4z_bench.json')); print('pace $p', d['value'], d['roofline']['frac'], d['clocks']['sm_mhz'])"; done
