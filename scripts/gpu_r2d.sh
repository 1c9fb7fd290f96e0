#!/bin/bash
# round 2, call D: the default bench line on 2 GPUs (extras exercise all-gather and the gradient all-reduce)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-kernels ) > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2d_bench_n2.json; tail -8 gpurun_out/r2d_bench_n2.err
