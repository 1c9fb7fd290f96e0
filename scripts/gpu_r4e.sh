#!/bin/bash
# K4 (CTA-pair kernel): 5-stage weight ring paid for by aliasing the next tile's inputs onto ACT_H
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/test_errors.jsonl
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 600 > gpurun_out/r4e_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E  " gpurun_out/r4e_pytest.log | head -30
timeout 300 python scripts/k4_phase_profile.py 200000 tc2 > gpurun_out/r4e_k4_phase.log 2>&1; echo rc=$?; cat gpurun_out/r4e_k4_phase.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-extras --no-kernels --no-cpu-baseline > gpurun_out/r4e_bench.json 2> gpurun_out/r4e_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r4e_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])"; tail -3 gpurun_out/r4e_bench.err
