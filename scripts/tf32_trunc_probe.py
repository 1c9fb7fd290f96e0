"""Does tcgen05.mma.kind::tf32 truncate or round the low 13 mantissa bits of its fp32 operands?  C = A . I through the
split=1 GEMM returns tf32(A) exactly (one non-zero product per output, fp32 accumulate)."""
import sys

import torch

sys.path.insert(0, ".")
from neusky_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
A = torch.randn(256, 64, generator=g).to(dev)
I = torch.eye(64, device=dev)
C = ops.gemm_nt(A, I, split=1)
bits = A.view(torch.int32)
trunc = (bits & -8192).view(torch.float32)                      # clear the low 13 bits
rn = ((bits + 4096) & -8192).view(torch.float32)                # round half up in magnitude
print("A as the A operand:  == trunc:", bool(torch.equal(C, trunc)), " == round:", bool(torch.equal(C, rn)), " max|C-A|/|A|:", float(((C - A).abs() / A.abs()).max()))
C2 = ops.gemm_nt(I.contiguous(), A[:64].t().contiguous(), split=1)   # C2[i, j] = sum_k I[i,k] * B[j,k] with B = A[:64]^T -> B[j, i] = A[i, j]
A64 = A[:64]
print("A as the B operand:  == trunc:", bool(torch.equal(C2, (A64.view(torch.int32) & -8192).view(torch.float32))),
      " == round:", bool(torch.equal(C2, ((A64.view(torch.int32) + 4096) & -8192).view(torch.float32))))
