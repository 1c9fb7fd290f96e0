"""One tf32 GEMM shape for ncu: python scripts/gemm_one.py N K split [nt|tn]"""
import sys

import torch

sys.path.insert(0, ".")
from neusky_b200 import ops  # noqa: E402

N, K, split = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else "nt"
dev = torch.device("cuda:0")
M = 1024 * 321
if mode == "nt":
    A = torch.randn(M, K, device=dev)
    B = torch.randn(N, K, device=dev)
    C = torch.empty(M, N, device=dev)
    for _ in range(3):
        ops.gemm_nt(A, B, out=C, split=split)
else:
    A = torch.randn(M, N, device=dev)
    B = torch.randn(M, K, device=dev)
    C = torch.zeros(N, K, device=dev)
    for _ in range(3):
        ops.gemm_tn(A, B, C, split=split)
torch.cuda.synchronize()
