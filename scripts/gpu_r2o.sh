#!/bin/bash
# 2-GPU runs of the supplementary workloads (configs 3, 4, 5) + the headline at N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 2 --workload train --steps 5 --warmup 3 > gpurun_out/r2o_train_n2.json 2> gpurun_out/r2o_train.err; echo "train exit=$?"; grep '^{' gpurun_out/r2o_train_n2.json | cut -c1-260; tail -2 gpurun_out/r2o_train.err
timeout 600 $TR bench.py --gpus 2 --workload eval --steps 3 --warmup 3 > gpurun_out/r2o_eval_n2.json 2> gpurun_out/r2o_eval.err; echo "eval exit=$?"; grep '^{' gpurun_out/r2o_eval_n2.json | cut -c1-260; tail -2 gpurun_out/r2o_eval.err
timeout 600 $TR bench.py --gpus 2 --workload relight --steps 2 --warmup 3 > gpurun_out/r2o_relight_n2.json 2> gpurun_out/r2o_relight.err; echo "relight exit=$?"; grep '^{' gpurun_out/r2o_relight_n2.json | cut -c1-260; tail -2 gpurun_out/r2o_relight.err
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_shade_n2.json 2> gpurun_out/r2o_shade.err; echo "shade exit=$?"; grep '^{' gpurun_out/r2o_shade_n2.json | cut -c1-260; tail -2 gpurun_out/r2o_shade.err
