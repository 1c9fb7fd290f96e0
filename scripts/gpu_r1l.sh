#!/bin/bash
# first GPU run of the training-path ops: parity tests + contraction throughput
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q 2>&1 | tail -60 > gpurun_out/r1l_pytest_train.log
cat gpurun_out/r1l_pytest_train.log | tail -25
timeout 300 python scripts/gemm_bench.py > gpurun_out/r1l_gemm_bench.jsonl 2> gpurun_out/r1l_gemm_bench.err
cat gpurun_out/r1l_gemm_bench.jsonl; tail -5 gpurun_out/r1l_gemm_bench.err
