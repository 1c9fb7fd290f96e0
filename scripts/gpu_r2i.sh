#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_parity.py -q -x > gpurun_out/r2i_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r2i_pytest.log | cut -c1-200
for i in 1 2; do
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_train_$i.json 2> gpurun_out/r2i_train.err; echo "train exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_train_$i.json')); print({k:d[k] for k in ('ms_per_step','wall_ms_per_step','host_launch_ms_per_step','gpu_launches')})"
done
