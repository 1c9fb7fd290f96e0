#!/bin/bash
mkdir -p gpurun_out
for t in 32768 65536 131072; do
timeout 600 python bench.py --workload eval --steps 2 --warmup 2 --tile $t > gpurun_out/r2n_bench_eval_$t.json 2> gpurun_out/r2n_e.err; echo "tile $t exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_eval_$t.json')); print($t, d['value'], d['ms_per_step'], d['gpu_launches'])"; tail -2 gpurun_out/r2n_e.err
done
