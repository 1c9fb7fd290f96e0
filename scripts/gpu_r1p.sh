#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -k "train_step" 2>&1 | tail -30 > gpurun_out/r1p_pytest_train.log
tail -30 gpurun_out/r1p_pytest_train.log
timeout 600 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/r1p_bench_train.json 2> gpurun_out/r1p_bench_train.err; echo "bench exit=$?"
cat gpurun_out/r1p_bench_train.json; tail -20 gpurun_out/r1p_bench_train.err
