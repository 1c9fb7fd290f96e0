#!/bin/bash
# full GPU suite + smoke + headline bench + reference arm on the restored HEAD
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1s_pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/r1s_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1s_smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/r1s_smoke.log
timeout 600 python bench.py > gpurun_out/r1s_bench_n1.json 2> gpurun_out/r1s_bench_n1.err; echo "bench exit=$?"; cat gpurun_out/r1s_bench_n1.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1s_bench_ref.json 2> gpurun_out/r1s_bench_ref.err; echo "ref exit=$?"; cat gpurun_out/r1s_bench_ref.json | cut -c1-600
timeout 600 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/r1s_bench_train.json 2> gpurun_out/r1s_bench_train.err; echo "train exit=$?"; cat gpurun_out/r1s_bench_train.json | cut -c1-900
