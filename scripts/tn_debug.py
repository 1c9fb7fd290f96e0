"""Bring-up probe of the MN-major tensor-map operands of the 3xTF32 TN GEMM: one-hot inputs must land on exactly one output row /
column and pair only equal reduction indices.  (History: with CU_TENSOR_MAP_SWIZZLE_128B + UMMA layout type 2 the MMA returned
zeros; 32-bit MN-major operands need SWIZZLE_128B_ATOM_32B + layout type 1 (SWIZZLE_128B_BASE32B), whose k atom is 4 rows:
SBO = 512 B.  With SBO = 1024 B rows 4-7 of every 8 were skipped and rows 8-11 counted twice.)"""
import sys

import torch
sys.path.insert(0, ".")
from neusky_b200 import ops
dev = torch.device("cuda:0")
M, P, Q = 16, 128, 256
def run(A, B):
    C = torch.zeros(P, Q, device=dev)
    ops.gemm_tn(A.to(dev), B.to(dev), C, split=3)
    return C.cpu()

A = torch.ones(M, P); B = torch.ones(M, Q)
C = run(A, B)
print(" all-ones: C min/max", float(C.min()), float(C.max()), "(expect 16)")
for m0, p0 in ((0, 0), (0, 5), (0, 33), (1, 5), (9, 7)):
    A = torch.zeros(M, P); A[m0, p0] = 1.0
    C = run(A, torch.ones(M, Q))
    rows = torch.nonzero(C.abs().sum(1) > 0).flatten().tolist()
    print(f"  A one-hot m={m0} p={p0}: rows {rows[:6]} vals {[round(float(C[r,0]),2) for r in rows[:3]]}")
for m0, q0 in ((0, 5), (1, 33), (9, 200)):
    B = torch.zeros(M, Q); B[m0, q0] = 1.0
    C = run(torch.ones(M, P), B)
    cols = torch.nonzero(C.abs().sum(0) > 0).flatten().tolist()
    print(f"  B one-hot m={m0} q={q0}: cols {cols[:6]}")
for m0, m1 in ((0, 0), (0, 1), (5, 5), (9, 9)):
    A = torch.zeros(M, P); A[m0, 0] = 1.0
    B = torch.zeros(M, Q); B[m1, 0] = 1.0
    C = run(A, B)
    print(f"  k pairing {m0},{m1}: C[0,0]={float(C[0,0])} sum={float(C.abs().sum())}")
