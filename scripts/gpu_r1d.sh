#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sdf.py tests/test_gpu_render.py -q > gpurun_out/pytest_d.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_d.log
tail -30 gpurun_out/pytest_d.log
timeout 600 python scripts/kernel_bench.py > gpurun_out/kernel_bench.jsonl 2> gpurun_out/kernel_bench.err; echo "kb exit=$?"
cat gpurun_out/kernel_bench.jsonl; tail -5 gpurun_out/kernel_bench.err
