#!/bin/bash
# round 2, session 2: full GPU parity suite + smoke + default bench (both arms) on the current tree
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/test_errors.jsonl
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r4f_pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r4f_pytest_gpu.log; tail -4 gpurun_out/r4f_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4f_smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/r4f_smoke.log; tail -2 gpurun_out/r4f_smoke.log
S=$(date +%s); timeout 1200 python bench.py > gpurun_out/r4f_bench.json 2> gpurun_out/r4f_bench.err; echo "bench exit=$? in $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/r4f_bench.json; tail -2 gpurun_out/r4f_bench.err
S=$(date +%s); timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r4f_bench_reference.json 2> gpurun_out/r4f_ref.err; echo "ref exit=$? in $(( $(date +%s) - S )) s"; cut -c1-200 gpurun_out/r4f_bench_reference.json
