#!/bin/bash
# round 2, call E: ncu DRAM-traffic metrics for the bandwidth-bound kernels + compute-sanitizer runs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_hbm_kernels.csv \
  -k regex:'hash_encode|neus_composite|proposal_density|pdf_resample|reni_|relight_collapsed|gemm_tf32' \
  python scripts/ncu_kernels.py > gpurun_out/r02_ncu_hbm_kernels_algorithmic.jsonl 2> gpurun_out/r02_ncu_hbm.err
echo "ncu rc=$?"; tail -3 gpurun_out/r02_ncu_hbm.err
python scripts/summarise_ncu_metrics.py gpurun_out/r02_ncu_hbm_kernels.csv > gpurun_out/r02_ncu_hbm_kernels_summary.txt 2>&1; tail -n 40 gpurun_out/r02_ncu_hbm_kernels_summary.txt
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 600 $CS --tool $tool --error-exitcode 7 --print-limit 20 python -m pytest -q -x -p no:cacheprovider \
     tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_sampler.py tests/test_gpu_shaders.py -m gpu > gpurun_out/r02_sanitizer_${tool}.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/r02_sanitizer_${tool}.log | tail -3
done
timeout 600 $CS --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest -q -x -p no:cacheprovider tests/test_gpu_train.py -m gpu -k "not proposal" > gpurun_out/r02_sanitizer_memcheck_train.log 2>&1
echo "memcheck train rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_sanitizer_memcheck_train.log | tail -3
timeout 600 $CS --tool initcheck --error-exitcode 7 --print-limit 20 python -m pytest -q -x -p no:cacheprovider tests/test_gpu_tc.py tests/test_gpu_sdf.py -m gpu > gpurun_out/r02_sanitizer_initcheck_tc.log 2>&1
echo "initcheck tc rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_sanitizer_initcheck_tc.log | tail -3
