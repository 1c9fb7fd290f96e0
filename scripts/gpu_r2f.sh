#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/test_errors.jsonl
timeout 300 python scripts/reni_fused_phase_profile.py > gpurun_out/r02_reni_fused_phase_cycles.log 2>&1; cat gpurun_out/r02_reni_fused_phase_cycles.log | tail -13
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E  " gpurun_out/r2f_pytest.log | head -30
timeout 600 python bench.py --workload relight --steps 2 --warmup 3 > gpurun_out/r2f_relight.json 2> gpurun_out/r2f_relight.err; tail -c 1500 gpurun_out/r2f_relight.json; tail -3 gpurun_out/r2f_relight.err
