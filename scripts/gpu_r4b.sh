#!/bin/bash
# round 2, session 2: plain-tf32 GEMM on the TMA operand path (NT pair shape / TN MN-major boxes) + tf32 DDF backward option
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/test_errors.jsonl
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_gemm_loaders.py tests/test_gpu_ddf_fit.py -m gpu -q --timeout 600 > gpurun_out/r4b_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E  " gpurun_out/r4b_pytest.log | head -30
grep ddf_train gpurun_out/test_errors.jsonl
timeout 300 python scripts/gemm_bench.py > gpurun_out/r4b_gemm_bench.jsonl 2> gpurun_out/r4b_gemm.err; echo "gemm rc=$?"; cat gpurun_out/r4b_gemm_bench.jsonl | cut -c1-200; tail -3 gpurun_out/r4b_gemm.err
for m in 0 1; do
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline --ddf-split-bwd $m > gpurun_out/r4b_bench_train_bwd$m.json 2> gpurun_out/r4b_train$m.err; echo "train rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r4b_bench_train_bwd$m.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','loss')})"; tail -3 gpurun_out/r4b_train$m.err
done
