#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 300 2>&1 | tail -2
for i in 1 2; do timeout 600 python bench.py --steps 3 --warmup 3 --no-extras --no-kernels --no-cpu-baseline > gpurun_out/r5a_bench.json 2> gpurun_out/r5a_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r5a_bench.json')); print(d['value'], d['roofline']['frac'], d['clocks']['sm_mhz'])"; done; tail -2 gpurun_out/r5a_bench.err
