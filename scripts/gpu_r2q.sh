#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r2q_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_train.json 2> gpurun_out/r2q_train.err; echo "train exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2q_bench_train.json')); print({k:d[k] for k in ('value','ms_per_step','wall_ms_per_step','gpu_launches')})"
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline --no-ddf-fit > gpurun_out/r2q_bench_train_nofit.json 2> gpurun_out/r2q_train2.err; python -c "
import json; d=json.load(open('gpurun_out/r2q_bench_train_nofit.json')); print({k:d[k] for k in ('value','ms_per_step','wall_ms_per_step','gpu_launches')})"
timeout 300 python scripts/gemm_bench.py > gpurun_out/r2q_gemm_bench.jsonl 2> /dev/null
timeout 120 python scripts/tn_debug.py > gpurun_out/r2q_tn_probe.log 2>&1; tail -4 gpurun_out/r2q_tn_probe.log
