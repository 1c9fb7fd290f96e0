#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_loaders.py -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/r2k_pytest.log | cut -c1-300
timeout 900 python bench.py --workload relight --steps 2 --warmup 3 > gpurun_out/r2k_bench_relight.json 2> gpurun_out/r2k_relight.err; echo "relight exit=$?"; cut -c1-900 gpurun_out/r2k_bench_relight.json; tail -5 gpurun_out/r2k_relight.err
