#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_render.py tests/test_gpu_parity.py tests/test_gpu_plugin.py -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest exit=$?"; tail -8 gpurun_out/r2s_pytest.log | cut -c1-300
timeout 600 python bench.py --workload eval --steps 3 --warmup 3 > gpurun_out/r2s_bench_eval.json 2> gpurun_out/r2s_e1.err; echo "eval exit=$?"; cut -c1-230 gpurun_out/r2s_bench_eval.json; tail -3 gpurun_out/r2s_e1.err
timeout 600 python bench.py --workload eval --samples 48 --sampler proposal --steps 3 --warmup 3 > gpurun_out/r2s_bench_eval_p48.json 2> gpurun_out/r2s_e2.err; echo "eval-prop exit=$?"; cut -c1-230 gpurun_out/r2s_bench_eval_p48.json
