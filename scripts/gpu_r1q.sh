#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_train.py tests/test_gpu_parity.py tests/test_gpu_plugin.py -q -k "reni or train_step or Reni or RENI or plugin" > gpurun_out/r1q_pytest.log 2>&1
grep -v "^  \|^$" gpurun_out/r1q_pytest.log | cut -c1-1800 | tail -40
for sp in 1 3; do
timeout 600 python bench.py --workload train --steps 5 --warmup 3 --split $sp > gpurun_out/r1q_bench_train_split$sp.json 2> gpurun_out/r1q_bench_train_split$sp.err; echo "bench exit=$?"
cut -c1-330 gpurun_out/r1q_bench_train_split$sp.json; tail -5 gpurun_out/r1q_bench_train_split$sp.err
done
