"""Throughput of the training-path tf32 contractions at the config-4 shapes (R=1024 rays x ~321 directions)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from neusky_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M = 1024 * 321


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


rows = []
for split in (1, 3):
    for (N, K) in ((256, 256), (2560, 256), (256, 2560), (256, 40), (256, 16), (32, 256)):
        A = torch.randn(M, K, device=dev)
        B = torch.randn(N, K, device=dev)
        C = torch.empty(M, N, device=dev)
        ms = timeit(lambda: ops.gemm_nt(A, B, out=C, split=split))
        rows.append({"op": "nt", "split": split, "M": M, "N": N, "K": K, "ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9, "GBps": 4.0 * (M * K + M * N) / ms / 1e6})
    for (P, Q) in ((256, 256), (2560, 256), (256, 40)):
        A = torch.randn(M, P, device=dev)
        B = torch.randn(M, Q, device=dev)
        C = torch.zeros(P, Q, device=dev)
        ms = timeit(lambda: ops.gemm_tn(A, B, C, split=split))
        rows.append({"op": "tn", "split": split, "M": M, "P": P, "Q": Q, "ms": ms, "tflops": 2.0 * M * P * Q / ms / 1e9, "GBps": 4.0 * (M * P + M * Q) / ms / 1e6})
# cuBLAS tf32 for scale (library baseline, not on the product path)
torch.backends.cuda.matmul.allow_tf32 = True
A = torch.randn(M, 256, device=dev); B = torch.randn(256, 256, device=dev)
ms = timeit(lambda: A @ B.T)
rows.append({"op": "cublas_tf32_nt", "M": M, "N": 256, "K": 256, "ms": ms, "tflops": 2.0 * M * 256 * 256 / ms / 1e9})
for r in rows:
    print(json.dumps(r))
