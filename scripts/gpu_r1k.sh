#!/bin/bash
# full GPU suite + smoke + headline bench + ncu launch list + full captures of the two tensor-core kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke.log; cat gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json | cut -c1-600; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'reni_|lambert|sky_shade|shade_finalize|FillFunctor|index' -c 300 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sky_shade_tc2 -s 1 -c 1 --csv --log-file gpurun_out/k4_dram_full.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/k4_dram_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sky_shade_tc2 -s 1 -c 1 -o gpurun_out/prof_k4_tc2 python bench.py --steps 1 --warmup 1 --points 40000 --no-cpu-baseline > gpurun_out/ncu_k4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sdf_field_tc -s 1 -c 1 -o gpurun_out/prof_k2_tc python bench.py --workload eval --steps 1 --warmup 1 --height 360 --width 640 > gpurun_out/ncu_k2.log 2>&1
ls -la gpurun_out | tail -20
