#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_render.py -q -k relight > gpurun_out/r2l_pytest.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/r2l_pytest.log | cut -c1-300
timeout 900 python bench.py --workload relight --steps 2 --warmup 3 > gpurun_out/r2l_bench_relight.json 2> gpurun_out/r2l_relight.err; echo "relight exit=$?"; cut -c1-1400 gpurun_out/r2l_bench_relight.json; tail -5 gpurun_out/r2l_relight.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:'relight|reni_|shade_fin' -c 600 --log-file gpurun_out/r2l_relight_launches.csv python bench.py --workload relight --steps 1 --warmup 3 --latents 4 > /dev/null 2>&1
python scripts/summarise_launches.py gpurun_out/r2l_relight_launches.csv gpurun_out/r2l_relight_launch_summary.txt | head
