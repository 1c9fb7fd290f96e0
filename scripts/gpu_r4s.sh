#!/bin/bash
# per-warp mbarrier arrivals in K2 and the tf32 GEMM
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/test_errors.jsonl
timeout 900 python -m pytest tests/test_gpu_sdf.py tests/test_gpu_render.py tests/test_gpu_train.py tests/test_gpu_gemm_loaders.py tests/test_gpu_ddf_fit.py -m gpu -q -x --timeout 600 > gpurun_out/r4s_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E  " gpurun_out/r4s_pytest.log | head -30
timeout 600 python scripts/kernel_bench.py 2> gpurun_out/r4s_kb.err | grep -E "sdf_field_tc" | cut -c1-230
timeout 300 python scripts/gemm_bench.py 2>/dev/null | grep -E '"split": 3' | cut -c1-150
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r4s_train.json 2> gpurun_out/r4s_train.err; python -c "
import json; d=json.load(open('gpurun_out/r4s_train.json')); print({k:d[k] for k in ('value','ms_per_step','loss')})"; tail -2 gpurun_out/r4s_train.err
