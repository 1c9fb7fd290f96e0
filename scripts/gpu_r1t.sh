#!/bin/bash
# proposal-sampler integration: render parity, microbench of P1/P2, eval bench with the shipped S=48 proposal placement
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_sampler.py tests/test_gpu_train.py -q > gpurun_out/r1t_pytest.log 2>&1; echo "pytest exit=$?"; tail -8 gpurun_out/r1t_pytest.log
timeout 600 python scripts/kernel_bench.py > gpurun_out/r1t_kernel_bench.jsonl 2> gpurun_out/r1t_kernel_bench.err; echo "kb exit=$?"; tail -4 gpurun_out/r1t_kernel_bench.jsonl | cut -c1-400
timeout 600 python bench.py --workload eval --samples 48 --sampler proposal --steps 3 --warmup 3 > gpurun_out/r1t_bench_eval_proposal48.json 2> gpurun_out/r1t_e1.err; echo "eval-prop exit=$?"; cut -c1-300 gpurun_out/r1t_bench_eval_proposal48.json
timeout 600 python bench.py --workload eval --samples 48 --steps 3 --warmup 3 > gpurun_out/r1t_bench_eval_uniform48.json 2> gpurun_out/r1t_e2.err; echo "eval-uni48 exit=$?"; cut -c1-300 gpurun_out/r1t_bench_eval_uniform48.json
timeout 600 python bench.py --workload eval --steps 3 --warmup 3 > gpurun_out/r1t_bench_eval_uniform128.json 2> gpurun_out/r1t_e3.err; echo "eval-uni128 exit=$?"; cut -c1-300 gpurun_out/r1t_bench_eval_uniform128.json
