#!/bin/bash
# GPU call 1: parity tests of the straightforward kernels + tcgen05 primitive probes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/probe1_smi.txt 2>&1
P=tools/bin/tc_probe
{
for args in "mma_ss 256 16" "mma_ss 256 64" "mma_ss 256 256" "mma_ss 128 256" "mma_ss 64 64" "mma_ss 256 48" \
            "mma_ts 256 16" "mma_ts 256 64" "mma_ts 256 256" "mma_ts 128 128" \
            "rate 0 256 0" "rate 0 128 0" "rate 0 64 0" "rate 1 256 0" "rate 1 128 0" "rate 1 64 0" \
            "rate 0 256 1" "rate 0 128 1" "rate 1 256 1" "rate 1 128 1" \
            "stream 32 148" "stream 16 148" "stream 32 74" "stream 64 148" "ldtm"; do
  echo "== $args"; timeout 60 $P $args 2>&1 | tail -8; echo "exit=$?"
done
} > gpurun_out/probe1.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu1.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu1.log
tail -5 gpurun_out/pytest_gpu1.log
grep -E "probe|exit=[1-9]" gpurun_out/probe1.log | head -60
