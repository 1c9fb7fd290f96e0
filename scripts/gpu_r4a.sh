#!/bin/bash
# round 2, session 2: K1 backward rewrite (thread = point) -- parity, kernel bench, and fresh launch lists of the train / eval steps
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_parity.py tests/test_gpu_train.py tests/test_gpu_sdf.py -m gpu -q --timeout 600 > gpurun_out/r4a_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E  " gpurun_out/r4a_pytest.log | head -30
timeout 600 python scripts/kernel_bench.py > gpurun_out/r4a_kernel_bench.jsonl 2> gpurun_out/r4a_kb.err; echo "kb rc=$?"
grep -E "hash_encode|pdf_resample|neus" gpurun_out/r4a_kernel_bench.jsonl | cut -c1-330
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r4a_train_launches.csv python bench.py --workload train --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r4a_train_ncu.log 2>&1; echo "ncu train rc=$?"
python scripts/summarise_launches.py gpurun_out/r4a_train_launches.csv > gpurun_out/r4a_train_launch_summary.txt 2>&1; head -45 gpurun_out/r4a_train_launch_summary.txt | cut -c1-200
