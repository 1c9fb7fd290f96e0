#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
timeout 300 $TR bench.py --gpus 2 --workload relight --steps 2 --warmup 3 > gpurun_out/r3g_relight_n2.json 2> gpurun_out/r3g_relight.err; echo "relight exit=$?"; grep '^{' gpurun_out/r3g_relight_n2.json | cut -c1-200
