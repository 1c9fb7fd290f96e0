#!/bin/bash
# K4 (CTA-pair kernel) with 16 epilogue warps (4 per scheduler), 768 threads
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/test_errors.jsonl
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_plugin.py tests/test_gpu_render.py -m gpu -q -x --timeout 600 > gpurun_out/r4q_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E  " gpurun_out/r4q_pytest.log | head -30
timeout 300 python scripts/k4_phase_profile.py 200000 tc2 > gpurun_out/r4q_k4_phase.log 2>&1; echo rc=$?; cat gpurun_out/r4q_k4_phase.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-extras --no-kernels --no-cpu-baseline > gpurun_out/r4q_bench.json 2> gpurun_out/r4q_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r4q_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])"; tail -3 gpurun_out/r4q_bench.err
