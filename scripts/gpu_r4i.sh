#!/bin/bash
# round 2, session 2: ncu DRAM-traffic metrics refreshed (K1 backward rewrite, compact relight pass, fused RENI++ rows) + compute-sanitizer on the kernels touched this session
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r4i_ncu_hbm_kernels.csv \
  -k regex:'hash_encode|neus_composite|proposal_density|pdf_resample|reni_rows_fused|relight_' \
  python scripts/ncu_kernels.py > gpurun_out/r4i_ncu_hbm_kernels_algorithmic.jsonl 2> gpurun_out/r4i_ncu_hbm.err
echo "ncu rc=$?"; tail -3 gpurun_out/r4i_ncu_hbm.err
python scripts/summarise_ncu_metrics.py gpurun_out/r4i_ncu_hbm_kernels.csv > gpurun_out/r4i_ncu_hbm_kernels_summary.txt 2>&1; tail -n 30 gpurun_out/r4i_ncu_hbm_kernels_summary.txt
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 900 $CS --tool $tool --error-exitcode 7 --print-limit 20 python -m pytest -q -x -p no:cacheprovider \
     tests/test_gpu_backward.py tests/test_gpu_gemm_loaders.py "tests/test_gpu_train.py::test_gemm_nt" "tests/test_gpu_train.py::test_gemm_tn" "tests/test_gpu_render.py" -m gpu > gpurun_out/r4i_sanitizer_${tool}.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/r4i_sanitizer_${tool}.log | tail -3
done
