#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -k "train_step or ddf_visibility or sdf_field" 2>&1 | tail -40 > gpurun_out/r1m_pytest_train.log
tail -30 gpurun_out/r1m_pytest_train.log
timeout 300 python scripts/gemm_bench.py > gpurun_out/r1m_gemm_bench.jsonl 2> gpurun_out/r1m_gemm_bench.err
cat gpurun_out/r1m_gemm_bench.jsonl; tail -5 gpurun_out/r1m_gemm_bench.err
