#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tcnn_import.py -m gpu -q --timeout 600 > gpurun_out/r4y_pytest_tcnn.log 2>&1; echo "tcnn pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E  assert" gpurun_out/r4y_pytest_tcnn.log | head -10
for i in 1 2; do timeout 600 python bench.py --steps 3 --warmup 3 --no-extras --no-kernels --no-cpu-baseline > gpurun_out/r4y_bench$i.json 2> gpurun_out/r4y_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r4y_bench$i.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])"; done
nvidia-smi --query-gpu=name,serial,uuid,power.limit --format=csv
