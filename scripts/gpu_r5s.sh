#!/bin/bash
# round 2, session 3: compute-sanitizer on the kernels touched in this session (film_sin_bwd with fused column sums: shared-memory
# fold + atomics; neus_composite with the device-resident inv_s) through the training parity tests
CS=/usr/local/cuda/bin/compute-sanitizer
OUT=gpurun_out/r5s_sanitizer.txt
: > $OUT
for tool in memcheck racecheck; do
  echo "=== $tool: tests/test_gpu_train.py (colsum / film_sin / ddf_visibility split=3 / train step) + tests/test_gpu_backward.py" >> $OUT
  timeout 500 $CS --tool $tool --error-exitcode 7 --print-limit 20 python -m pytest -q -x -p no:cacheprovider \
    tests/test_gpu_train.py -k "colsum_and_film or (ddf_visibility and 3-None) or train_step_losses_and_gradients_vs_oracle_autograd and 3-" 2>&1 | grep -v "^$" | tail -8 >> $OUT
  echo "exit=$?" >> $OUT
done
cat $OUT
