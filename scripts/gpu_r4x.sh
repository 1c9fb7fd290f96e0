#!/bin/bash
# imported tiny-cuda-nn grids through K1 / K2 / K4 (GridMode) + regression of the plain paths
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/test_errors.jsonl
timeout 900 python -m pytest tests/test_tcnn_import.py -m gpu -q --timeout 600 > gpurun_out/r4x_pytest_tcnn.log 2>&1; echo "tcnn pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E  " gpurun_out/r4x_pytest_tcnn.log | head -30
timeout 1200 python -m pytest tests/test_gpu_sdf.py tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_plugin.py tests/test_gpu_render.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 600 > gpurun_out/r4x_pytest.log 2>&1; echo "regression pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E  " gpurun_out/r4x_pytest.log | head -30
timeout 600 python bench.py --steps 3 --warmup 3 --no-kernels --no-cpu-baseline > gpurun_out/r4x_bench.json 2> gpurun_out/r4x_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r4x_bench.json')); print(d['value'], d['roofline']['frac'], d['extra']['eval']['value'], d['extra']['eval']['ms_per_step'])"; tail -3 gpurun_out/r4x_bench.err
