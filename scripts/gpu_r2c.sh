#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 2 -c 1 -o gpurun_out/r2c_gemm_nt_256_256_s3 -f python scripts/gemm_one.py 256 256 3 > gpurun_out/r2c_ncu.log 2>&1; echo "ncu exit=$?"
