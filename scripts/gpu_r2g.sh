#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r2g_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_train.json 2> gpurun_out/r2g_train.err; echo "train exit=$?"; cut -c1-300 gpurun_out/r2g_bench_train.json; tail -3 gpurun_out/r2g_train.err
timeout 600 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline --no-ddf-fit > gpurun_out/r2g_bench_train_nofit.json 2> gpurun_out/r2g_train2.err; echo "train exit=$?"; cut -c1-300 gpurun_out/r2g_bench_train_nofit.json
timeout 300 python scripts/gemm_bench.py > gpurun_out/r2g_gemm_bench.jsonl 2> /dev/null
