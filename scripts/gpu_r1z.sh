#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -x > gpurun_out/r1z_pytest.log 2>&1; echo "pytest exit=$?"; tail -15 gpurun_out/r1z_pytest.log | cut -c1-400
timeout 300 python scripts/gemm_bench.py > gpurun_out/r1z_gemm_bench.jsonl 2> gpurun_out/r1z_gemm.err; echo "gemm exit=$?"; grep '"split": 3' gpurun_out/r1z_gemm_bench.jsonl | cut -c1-200
