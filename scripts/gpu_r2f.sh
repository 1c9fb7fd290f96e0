#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -x -k "gemm" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit=$?"; tail -15 gpurun_out/r2f_pytest.log | cut -c1-400
timeout 300 python scripts/gemm_bench.py > gpurun_out/r2f_gemm_bench.jsonl 2> gpurun_out/r2f_gemm.err; echo "gemm exit=$?"; grep '"split": 3' gpurun_out/r2f_gemm_bench.jsonl | grep '"nt"' | cut -c1-200; tail -3 gpurun_out/r2f_gemm.err
