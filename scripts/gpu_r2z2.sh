#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ddf_fit.py -q > gpurun_out/r2z2_pytest.log 2>&1; echo "pytest exit=$?"; tail -12 gpurun_out/r2z2_pytest.log | cut -c1-400
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2z2_bench_train.json 2> gpurun_out/r2z2_train.err; echo "train exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2z2_bench_train.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','loss')})"; tail -3 gpurun_out/r2z2_train.err
