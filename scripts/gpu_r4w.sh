#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for i in 1 2; do timeout 600 python bench.py --steps 3 --warmup 3 --no-extras --no-kernels --no-cpu-baseline > gpurun_out/r4w_bench$i.json 2> gpurun_out/r4w_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r4w_bench$i.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])"; done; tail -3 gpurun_out/r4w_bench.err
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q 2>&1 | tail -2
