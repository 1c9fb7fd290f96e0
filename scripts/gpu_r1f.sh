#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sdf.py tests/test_gpu_render.py -q -s > gpurun_out/pytest_f.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_f.log
grep -E "^n=|^\{|passed|failed|exit" gpurun_out/pytest_f.log | tail -20
timeout 600 python scripts/kernel_bench.py > gpurun_out/kernel_bench.jsonl 2> gpurun_out/kernel_bench.err; echo "kb exit=$?"
grep sdf_field gpurun_out/kernel_bench.jsonl; tail -5 gpurun_out/kernel_bench.err
