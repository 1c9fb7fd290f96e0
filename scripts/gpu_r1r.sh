#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r1r_launches_train.csv python bench.py --workload train --steps 1 --warmup 1 > gpurun_out/r1r_bench_under_ncu.log 2>&1
echo "ncu exit=$?"; wc -l gpurun_out/r1r_launches_train.csv
timeout 300 python - <<'PY' > gpurun_out/r1r_torch_profile.txt 2>&1
import os, sys, subprocess
sys.argv = ["bench.py", "--workload", "train", "--steps", "3", "--warmup", "3"]
import torch
from torch.profiler import profile, ProfilerActivity
import runpy
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    runpy.run_path("bench.py", run_name="__main__")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
PY
tail -70 gpurun_out/r1r_torch_profile.txt | cut -c1-220
