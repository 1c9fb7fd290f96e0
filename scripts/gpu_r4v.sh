#!/bin/bash
# ncu --set full captures of the kernels changed this round: K2 (16 epilogue warps), fused RENI++ rows, compact relight pass
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sdf_field_tc -s 1 -c 1 -o gpurun_out/r4v_prof_k2_tc -f python bench.py --workload eval --steps 1 --warmup 1 --height 360 --width 640 > gpurun_out/r4v_ncu_k2.log 2>&1; echo "ncu k2 exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'reni_rows_fused|relight_h16' -s 2 -c 2 -o gpurun_out/r4v_prof_relight -f python bench.py --workload relight --steps 1 --warmup 1 --height 360 --width 640 --latents 32 > gpurun_out/r4v_ncu_relight.log 2>&1; echo "ncu relight exit=$?"
ls -la gpurun_out/r4v_*.ncu-rep
