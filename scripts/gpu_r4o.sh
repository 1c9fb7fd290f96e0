#!/bin/bash
# K4: register reallocation (service warps 64, epilogue warps 160) + all TMEM loads of a FiLM chunk / two 32-column loads of a mapping layer in flight;
# RENI++ fused rows: encoding warps 56, epilogue warps 104
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/test_errors.jsonl
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_plugin.py -m gpu -q -x --timeout 600 > gpurun_out/r4o_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E  " gpurun_out/r4o_pytest.log | head -30
timeout 300 python scripts/k4_phase_profile.py 200000 tc2 > gpurun_out/r4o_k4_phase.log 2>&1; echo rc=$?; cat gpurun_out/r4o_k4_phase.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-extras --no-kernels --no-cpu-baseline > gpurun_out/r4o_bench.json 2> gpurun_out/r4o_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r4o_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])"; tail -3 gpurun_out/r4o_bench.err
timeout 600 python bench.py --workload relight --steps 2 --warmup 3 > gpurun_out/r4o_relight.json 2> gpurun_out/r4o_relight.err; python -c "
import json; d=json.load(open('gpurun_out/r4o_relight.json')); print(d['value'], d['config']['ms_per_latent_by_stage'], d['roofline']['frac'], d['roofline_cache_pass'])"; tail -3 gpurun_out/r4o_relight.err
