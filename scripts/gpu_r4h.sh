#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/nccl_probe.py > gpurun_out/r4h_nccl_n$N.log 2>&1; echo rc=$?
grep -E "^world|via|NVLS|Using network|Channel 00|P2P|SHM" gpurun_out/r4h_nccl_n$N.log | sort | uniq -c | sort -rn | head -20
nvidia-smi topo -m 2>/dev/null | head -12
