"""The 3xTF32 NT GEMM has two operand loaders: tensor-map TMA boxes (default) and the LDGSTS fallback used when a tensor map
cannot be encoded.  Both must give the same fp32-accurate result, bit for bit (same MMA order, same operands)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from neusky_b200 import _lib

    _lib.load()
    return torch.device("cuda:0")


@pytest.mark.parametrize("M,N,K", [(1000, 256, 256), (4099, 256, 72), (777, 2560, 256), (513, 40, 2560), (255, 32, 40)])
def test_tma_and_ldgsts_loaders_agree(dev, M, N, K):
    from neusky_b200 import ops

    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev)
    B = (torch.randn(N, K, generator=g) * 0.2).to(dev)
    bias = (torch.randn(N, generator=g) * 0.1).to(dev)
    os.environ.pop("NSK_GEMM_NO_TMA", None)
    tma = ops.gemm_nt(A, B, bias=bias, act="leaky", split=3)
    os.environ["NSK_GEMM_NO_TMA"] = "1"
    try:
        ldg = ops.gemm_nt(A, B, bias=bias, act="leaky", split=3)
    finally:
        os.environ.pop("NSK_GEMM_NO_TMA", None)
    ref = torch.nn.functional.leaky_relu(A.double() @ B.double().T + bias.double(), 0.2)
    scale = float((A.double().norm(dim=1)[:, None] * B.double().norm(dim=1)[None, :]).mean())
    assert float((tma.double() - ref).abs().max()) <= 2e-6 * scale
    assert float((ldg.double() - ref).abs().max()) <= 2e-6 * scale
    assert torch.equal(tma, ldg), "the two loaders feed the same MMAs in the same order: results must be identical"


@pytest.mark.parametrize("M,P,Q", [(5000, 256, 256), (333, 2560, 256), (40000, 128, 64), (17, 256, 32)])
def test_tn_tma_and_register_loaders_agree(dev, M, P, Q):
    """dW = dY^T X, 3xTF32: MN-major tensor-map boxes (default when P and Q are multiples of 32) vs the register-transposing fallback."""
    from neusky_b200 import ops

    g = torch.Generator().manual_seed(M + P + Q)
    A = (torch.randn(M, P, generator=g) * 0.3).to(dev)
    B = (torch.randn(M, Q, generator=g) * 0.3).to(dev)
    C0 = torch.randn(P, Q, generator=g).to(dev)
    os.environ.pop("NSK_GEMM_NO_TMA", None)
    tma = ops.gemm_tn(A, B, C0.clone(), split=3)
    os.environ["NSK_GEMM_NO_TMA"] = "1"
    try:
        reg = ops.gemm_tn(A, B, C0.clone(), split=3)
    finally:
        os.environ.pop("NSK_GEMM_NO_TMA", None)
    ref = C0.double() + A.double().T @ B.double()
    scale = float((A.double().norm(dim=0)[:, None] * B.double().norm(dim=0)[None, :]).mean())
    tol = 3e-6 * scale + 1e-5
    assert float((tma.double() - ref).abs().max()) <= tol
    assert float((reg.double() - ref).abs().max()) <= tol
