#!/bin/bash
# round 2, call B: full GPU test suite again (tightened tolerances, drop-in model tests, full-size frame tests)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/test_errors.jsonl
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --durations=8 > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
grep -E "passed|failed|FAILED|ERROR|^E  " gpurun_out/r2b_pytest.log | head -40
