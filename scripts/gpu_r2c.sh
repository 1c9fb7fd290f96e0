#!/bin/bash
# round 2, call C: the default bench line (headline + extras + kernels) on 1 GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/r2c_bench.json; tail -8 gpurun_out/r2c_bench.err
