"""K2 (SDF / albedo field with analytic normals) vs the CPU oracle (torch autograd gradient).
fp32 path tolerance: 1e-3 relative (north_star); observed ~1e-5."""
import pytest
import torch

from neusky_b200 import init as nb_init

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from neusky_b200 import _lib

    _lib.load()
    return torch.device("cuda:0")


def _trained_like(p, seed):
    """Geometric init leaves the PE / hash columns of glin0 at zero and the hash table at 1e-3; perturb them so
    every term of the analytic gradient (PE, hash interpolation, contraction) is exercised."""
    g = torch.Generator().manual_seed(seed)
    p = {k: v.clone() for k, v in p.items()}
    p["glin0.weight_v"][:, 3:] = 0.05 * torch.randn(256, 68, generator=g)
    p["glin0.weight_g"] = p["glin0.weight_v"].norm(dim=1, keepdim=True) * (0.8 + 0.4 * torch.rand(256, 1, generator=g))
    p["encoding.hash_table"] = p["encoding.hash_table"] * 200.0
    p["glin1.bias"] = 0.02 * torch.randn(256, generator=g)
    return p


def _points(n, seed, spread=0.9):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(n, 3, generator=g) * 2 - 1) * spread


@pytest.mark.parametrize("n,spread", [(1, 0.9), (31, 0.9), (1000, 0.9), (257, 1.8)])
def test_sdf_field_simt_vs_oracle(dev, n, spread):
    from neusky_b200 import ops, packing
    from oracle import neusky_oracle as O

    log2_T = 15
    p = _trained_like(nb_init.init_sdf_params(3, log2_T=log2_T), 4)
    x = _points(n, n, spread)   # spread 1.8: some points outside the unit cube -> L-inf contraction branch
    ref = O.sdf_field(x, p, O.hash_scalings(), log2_T)
    blob = packing.pack_sdf_simt({k: v.to(dev) for k, v in p.items()})
    out = ops.sdf_field(x.to(dev), blob, p["encoding.hash_table"].to(dev), O.hash_scalings().to(dev), log2_T, want_geo=True)
    torch.cuda.synchronize()
    assert torch.allclose(out["sdf"].cpu(), ref["sdf"], rtol=1e-4, atol=1e-5), (out["sdf"].cpu() - ref["sdf"]).abs().max()
    assert torch.allclose(out["geo"].cpu(), ref["geo"], rtol=1e-4, atol=1e-5)
    gerr = (out["gradient"].cpu() - ref["gradient"]).abs().max() / ref["gradient"].abs().max()
    assert float(gerr) <= 1e-3, float(gerr)
    assert float(ref["gradient"].norm(dim=-1).min()) > 1e-3
    assert torch.allclose(out["albedo"].cpu(), ref["albedo"], rtol=1e-4, atol=1e-5)


def test_sdf_field_geometric_init_is_a_sphere(dev):
    """nerfstudio geometric init: sdf(x) ~ |x| - bias, gradient ~ x/|x| (sanity of the analytic gradient's sign)."""
    from neusky_b200 import ops, packing
    from oracle import neusky_oracle as O

    p = nb_init.init_sdf_params(5, log2_T=12)
    x = _points(512, 9)
    blob = packing.pack_sdf_simt({k: v.to(dev) for k, v in p.items()})
    out = ops.sdf_field(x.to(dev), blob, p["encoding.hash_table"].to(dev), O.hash_scalings().to(dev), 12, want_albedo=False)
    n = torch.nn.functional.normalize(out["gradient"].cpu(), dim=-1)
    cos = (n * torch.nn.functional.normalize(x, dim=-1)).sum(-1)
    assert float(cos.mean()) > 0.8
    assert "albedo" not in out


# ------------------------------------------------------------------------------------------------ tensor-core path
# Tolerances of the fp16-operand tcgen05 path, stated separately from the fp32 path (north_star): sdf |err| <= 2e-3
# (scene units; x enters as fp16 hi+lo and the sdf row is an fp32 dot), gradient direction cos >= 0.999 and
# magnitude within 2 %, albedo |err| <= 5e-3.
@pytest.mark.parametrize("n,spread", [(1, 0.9), (127, 0.9), (128, 0.9), (5000, 0.9), (40000, 0.9), (300, 1.8)])
def test_sdf_field_tc_vs_oracle(dev, n, spread):
    from neusky_b200 import ops, packing
    from oracle import neusky_oracle as O

    log2_T = 15
    p = _trained_like(nb_init.init_sdf_params(3, log2_T=log2_T), 4)
    x = _points(n, n + 1, spread)
    ref = O.sdf_field(x[:2048], p, O.hash_scalings(), log2_T)
    pd = {k: v.to(dev) for k, v in p.items()}
    out = ops.sdf_field(x.to(dev), packing.pack_sdf_tc(pd), pd["encoding.hash_table"], O.hash_scalings().to(dev), log2_T, impl="tc")
    exact = ops.sdf_field(x.to(dev), packing.pack_sdf_simt(pd), pd["encoding.hash_table"], O.hash_scalings().to(dev), log2_T, impl="simt")
    torch.cuda.synchronize()
    m = min(n, 2048)
    e_sdf = (out["sdf"].cpu()[:m] - ref["sdf"]).abs().max()
    e_alb = (out["albedo"].cpu()[:m] - ref["albedo"]).abs().max()
    g, gr = out["gradient"].cpu()[:m], ref["gradient"]
    cos = torch.nn.functional.cosine_similarity(g, gr, dim=-1).min()
    mag = (g.norm(dim=-1) / gr.norm(dim=-1) - 1).abs().max()
    print(f"n={n}: sdf err {float(e_sdf):.2e}, albedo err {float(e_alb):.2e}, grad cos min {float(cos):.6f}, grad |mag-1| {float(mag):.2e}")
    assert float(e_sdf) <= 2e-3 and float(e_alb) <= 5e-3 and float(cos) >= 0.999 and float(mag) <= 2e-2
    # every row (also beyond the oracle slice) against the exact fp32 kernel: exercises the persistent multi-tile loop
    assert float((out["sdf"] - exact["sdf"]).abs().max()) <= 2e-3
    assert float((out["albedo"] - exact["albedo"]).abs().max()) <= 5e-3
    assert float(torch.nn.functional.cosine_similarity(out["gradient"], exact["gradient"], dim=-1).min()) >= 0.999
