#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_train.json 2> gpurun_out/r2h_train.err; echo "train exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2h_bench_train.json')); print({k:d[k] for k in ('ms_per_step','wall_ms_per_step','host_launch_ms_per_step','gpu_launches')})"
timeout 600 python - <<'PY' > gpurun_out/r2h_profile.txt 2>&1
import sys, torch, cProfile, pstats, io, os
sys.argv = ["bench.py", "--workload", "train", "--steps", "3", "--warmup", "2", "--no-cpu-baseline"]
import bench
pr = cProfile.Profile(); pr.enable()
bench.main()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35); print(s.getvalue())
PY
head -70 gpurun_out/r2h_profile.txt | cut -c1-180
