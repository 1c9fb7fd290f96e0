#!/bin/bash
# round-1 FINAL measurement pass on the final tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2v_pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r2v_pytest_gpu.log; tail -3 gpurun_out/r2v_pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/r2v_smoke.log; tail -2 gpurun_out/r2v_smoke.log
timeout 900 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench exit=$?"; cut -c1-200 gpurun_out/r2v_bench.json; tail -2 gpurun_out/r2v_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2v_bench_reference.json 2> gpurun_out/r2v_ref.err; echo "ref exit=$?"; cut -c1-160 gpurun_out/r2v_bench_reference.json
timeout 600 python bench.py --workload eval --steps 3 --warmup 3 > gpurun_out/r2v_bench_eval.json 2> gpurun_out/r2v_e1.err; echo "eval exit=$?"; cut -c1-200 gpurun_out/r2v_bench_eval.json
timeout 600 python bench.py --workload eval --samples 48 --sampler proposal --steps 3 --warmup 3 > gpurun_out/r2v_bench_eval_p48.json 2> gpurun_out/r2v_e2.err; echo "eval-prop exit=$?"; cut -c1-200 gpurun_out/r2v_bench_eval_p48.json
timeout 600 python bench.py --workload eval --samples 48 --steps 3 --warmup 3 > gpurun_out/r2v_bench_eval_u48.json 2> gpurun_out/r2v_e3.err; echo "eval-u48 exit=$?"; cut -c1-200 gpurun_out/r2v_bench_eval_u48.json
timeout 900 python bench.py --workload relight --steps 2 --warmup 3 > gpurun_out/r2v_bench_relight.json 2> gpurun_out/r2v_relight.err; echo "relight exit=$?"; cut -c1-200 gpurun_out/r2v_bench_relight.json
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_bench_train.json 2> gpurun_out/r2v_train.err; echo "train exit=$?"; cut -c1-260 gpurun_out/r2v_bench_train.json
timeout 600 python scripts/kernel_bench.py > gpurun_out/r2v_kernel_bench.jsonl 2> gpurun_out/r2v_kb.err; echo "kb exit=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2v_eval_launches.csv python bench.py --workload eval --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu exit=$?"
python scripts/summarise_launches.py gpurun_out/r2v_eval_launches.csv gpurun_out/r2v_eval_launch_summary.txt | head -14
