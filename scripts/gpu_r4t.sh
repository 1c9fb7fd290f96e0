#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/test_errors.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "compact_relight" > gpurun_out/r4t_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E  " gpurun_out/r4t_pytest.log | head -30
cat gpurun_out/test_errors.jsonl
