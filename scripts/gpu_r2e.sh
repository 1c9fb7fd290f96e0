#!/bin/bash
# round-1 closing measurement pass on the current tree: smoke, headline bench (+ reference arm), eval benches, per-kernel rooflines,
# ncu launch list of the headline command
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/r2e_smoke.log; tail -2 gpurun_out/r2e_smoke.log
timeout 900 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench exit=$?"; cut -c1-300 gpurun_out/r2e_bench.json; tail -2 gpurun_out/r2e_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2e_bench_reference.json 2> gpurun_out/r2e_ref.err; echo "ref exit=$?"; cut -c1-200 gpurun_out/r2e_bench_reference.json
timeout 600 python bench.py --workload eval --steps 3 --warmup 3 > gpurun_out/r2e_bench_eval.json 2> gpurun_out/r2e_e1.err; echo "eval exit=$?"; cut -c1-250 gpurun_out/r2e_bench_eval.json
timeout 600 python bench.py --workload eval --samples 48 --sampler proposal --steps 3 --warmup 3 > gpurun_out/r2e_bench_eval_proposal48.json 2> gpurun_out/r2e_e2.err; echo "eval-prop exit=$?"; cut -c1-250 gpurun_out/r2e_bench_eval_proposal48.json
timeout 600 python scripts/kernel_bench.py > gpurun_out/r2e_kernel_bench.jsonl 2> gpurun_out/r2e_kb.err; echo "kb exit=$?"; wc -l gpurun_out/r2e_kernel_bench.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2e_bench_under_ncu.log 2>&1; echo "ncu exit=$?"
python scripts/summarise_launches.py gpurun_out/r2e_launches_bench.csv gpurun_out/r2e_bench_launch_summary.txt | head -12
