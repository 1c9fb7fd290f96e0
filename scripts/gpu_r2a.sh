#!/bin/bash
# round 2, call A: full GPU test suite (no -x: list every failure), smoke, short default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/test_errors.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -40 gpurun_out/r2a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2a_smoke.log
echo skip-bench
