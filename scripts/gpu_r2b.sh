#!/bin/bash
# full GPU parity suite + train bench after the tf32 GEMM rework
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/r2b_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_train.json 2> gpurun_out/r2b_train.err; echo "train exit=$?"; cut -c1-400 gpurun_out/r2b_bench_train.json
cp gpurun_out/r1z_gemm_bench.jsonl /dev/null 2>&1
timeout 300 python scripts/gemm_bench.py > gpurun_out/r2b_gemm_bench.jsonl 2> /dev/null; echo "gemm exit=$?"
