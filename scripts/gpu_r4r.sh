#!/bin/bash
# K2 (fused SDF / colour field) with 16 epilogue warps (4 per scheduler), 768 threads
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/test_errors.jsonl
timeout 900 python -m pytest tests/test_gpu_sdf.py tests/test_gpu_render.py tests/test_gpu_fullsize_frame.py tests/test_gpu_plugin.py -m gpu -q -x --timeout 600 > gpurun_out/r4r_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E  " gpurun_out/r4r_pytest.log | head -30
timeout 600 python scripts/kernel_bench.py > gpurun_out/r4r_kernel_bench.jsonl 2> gpurun_out/r4r_kb.err; echo "kb rc=$?"
grep -E "sdf_field|relight|reni_rows_fused" gpurun_out/r4r_kernel_bench.jsonl | cut -c1-260
timeout 600 python bench.py --workload eval --steps 3 --warmup 2 > gpurun_out/r4r_eval.json 2> gpurun_out/r4r_eval.err; python -c "
import json; d=json.load(open('gpurun_out/r4r_eval.json')); print(d['value'], d['ms_per_step'])"; tail -3 gpurun_out/r4r_eval.err
