#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/kernel_bench.py > gpurun_out/r3b_kernel_bench.jsonl 2> gpurun_out/r3b_kb.err; echo "kb exit=$?"; grep reni gpurun_out/r3b_kernel_bench.jsonl | cut -c1-260; tail -3 gpurun_out/r3b_kb.err
