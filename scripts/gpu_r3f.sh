#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_render.py -q -k "reni or relight" > gpurun_out/r3f_pytest.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/r3f_pytest.log | cut -c1-300
timeout 900 python bench.py --workload relight --steps 2 --warmup 3 > gpurun_out/r3f_bench_relight.json 2> gpurun_out/r3f_relight.err; echo "relight exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/r3f_bench_relight.json')); print(d['value'], d['config']['ms_per_latent_frame'])"; tail -3 gpurun_out/r3f_relight.err
