#!/bin/bash
# round 2 FINAL: full GPU parity suite + smoke + default bench (both arms) + per-workload lines on the final tree
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/test_errors.jsonl
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r4u_pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r4u_pytest_gpu.log; tail -4 gpurun_out/r4u_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4u_smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/r4u_smoke.log; tail -2 gpurun_out/r4u_smoke.log
S=$(date +%s); timeout 1200 python bench.py > gpurun_out/r4u_bench.json 2> gpurun_out/r4u_bench.err; echo "bench exit=$? in $(( $(date +%s) - S )) s"; cut -c1-200 gpurun_out/r4u_bench.json; tail -2 gpurun_out/r4u_bench.err
S=$(date +%s); timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r4u_bench_reference.json 2> gpurun_out/r4u_ref.err; echo "ref exit=$? in $(( $(date +%s) - S )) s"; cut -c1-200 gpurun_out/r4u_bench_reference.json
timeout 600 python bench.py --workload relight --steps 2 --warmup 3 > gpurun_out/r4u_relight.json 2> gpurun_out/r4u_relight.err; echo "relight exit=$?"; cut -c1-160 gpurun_out/r4u_relight.json
timeout 600 python bench.py --workload eval --samples 48 --sampler proposal --steps 3 --warmup 2 > gpurun_out/r4u_eval_p48.json 2> gpurun_out/r4u_e2.err; echo "eval-prop exit=$?"; cut -c1-160 gpurun_out/r4u_eval_p48.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4u_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-kernels --no-cpu-baseline > /dev/null 2>&1; echo "ncu exit=$?"
python scripts/summarise_launches.py gpurun_out/r4u_bench_launches.csv gpurun_out/r4u_bench_launch_summary.txt | head -8
