#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_render.py tests/test_gpu_plugin.py -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest exit=$?"; tail -12 gpurun_out/r2m_pytest.log | cut -c1-400
timeout 900 python bench.py --workload relight --steps 2 --warmup 3 > gpurun_out/r2m_bench_relight.json 2> gpurun_out/r2m_relight.err; echo "relight exit=$?"; cut -c1-330 gpurun_out/r2m_bench_relight.json; tail -4 gpurun_out/r2m_relight.err
timeout 600 python bench.py --workload eval --steps 3 --warmup 3 > gpurun_out/r2m_bench_eval.json 2> gpurun_out/r2m_e1.err; echo "eval exit=$?"; cut -c1-250 gpurun_out/r2m_bench_eval.json; tail -4 gpurun_out/r2m_e1.err
