#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519"
timeout 600 $TR bench.py --gpus 2 --workload train --steps 5 --warmup 3 > gpurun_out/r2y_train_n2.json 2> gpurun_out/r2y_train.err; echo "train exit=$?"; grep '^{' gpurun_out/r2y_train_n2.json | cut -c1-330; tail -3 gpurun_out/r2y_train.err | cut -c1-200
