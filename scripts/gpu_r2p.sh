#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -k "gemm_tn" > gpurun_out/r2p_pytest.log 2>&1; echo "pytest exit=$?"; tail -12 gpurun_out/r2p_pytest.log | cut -c1-300
timeout 300 python scripts/gemm_bench.py > gpurun_out/r2p_gemm_bench.jsonl 2> gpurun_out/r2p_gemm.err; echo "gemm exit=$?"; grep '"split": 3' gpurun_out/r2p_gemm_bench.jsonl | grep '"tn"' | cut -c1-200; tail -3 gpurun_out/r2p_gemm.err
