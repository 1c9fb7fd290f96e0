#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_render.py -q > gpurun_out/pytest_g.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_g.log
tail -5 gpurun_out/pytest_g.log
timeout 600 python bench.py --workload eval --steps 2 --warmup 1 > gpurun_out/bench_eval_n1.json 2> gpurun_out/bench_eval_n1.err; echo "eval exit=$?"
cat gpurun_out/bench_eval_n1.json; tail -5 gpurun_out/bench_eval_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'nsk|sdf_|sky_|reni_|lambert|neus_|shade_' -c 400 --csv --log-file gpurun_out/launches_eval.csv python bench.py --workload eval --steps 1 --warmup 0 --height 180 --width 320 > /dev/null 2>&1
