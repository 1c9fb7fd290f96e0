#!/bin/bash
# relight pass on warp-level mma (16-row tiles, 32 codes per read of the compact cache)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/test_errors.jsonl
timeout 900 python -m pytest tests/test_gpu_fullsize_frame.py tests/test_gpu_render.py -m gpu -q --timeout 600 > gpurun_out/r4j_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E  " gpurun_out/r4j_pytest.log | head -30
grep config5 gpurun_out/test_errors.jsonl
timeout 600 python bench.py --workload relight --steps 2 --warmup 3 > gpurun_out/r4j_relight.json 2> gpurun_out/r4j_relight.err; tail -c 1900 gpurun_out/r4j_relight.json; tail -3 gpurun_out/r4j_relight.err
