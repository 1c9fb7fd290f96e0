#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 2 -c 1 -o gpurun_out/r2t_gemm_nt_after -f python scripts/gemm_one.py 256 256 3 nt > gpurun_out/r2t_ncu1.log 2>&1; echo "ncu1 exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 2 -c 1 -o gpurun_out/r2t_gemm_nt2560_after -f python scripts/gemm_one.py 2560 256 3 nt > gpurun_out/r2t_ncu2.log 2>&1; echo "ncu2 exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 2 -c 1 -o gpurun_out/r2t_gemm_tn_after -f python scripts/gemm_one.py 256 256 3 tn > gpurun_out/r2t_ncu3.log 2>&1; echo "ncu3 exit=$?"
