"""Backward kernels vs torch autograd through the CPU oracle (fp64), driven through neusky_b200.autograd."""
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


def _case(R, S, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.sort(torch.rand(R, S + 1, generator=g) * 2.0 + 0.05, dim=1).values
    starts, ends = t[:, :-1, None].contiguous(), t[:, 1:, None].contiguous()
    ray_dirs = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    sdf = (1.0 - (starts + ends) / 2) * 0.3 + 0.02 * torch.randn(R, S, 1, generator=g)
    grad = -ray_dirs[:, None, :] * (0.7 + 0.6 * torch.rand(R, S, 1, generator=g)) + 0.3 * torch.randn(R, S, 3, generator=g)
    albedo = torch.rand(R, S, 3, generator=g)
    dnorm = 1.0 + torch.rand(R, 1, generator=g)
    return sdf, grad, albedo, ray_dirs, starts, ends, dnorm


@pytest.mark.parametrize("R,S,rho", [(3, 5, 1.0), (17, 48, 1.0), (9, 128, 0.3), (4, 200, 1.0)])
def test_neus_composite_backward_vs_oracle_autograd(dev, R, S, rho):
    from neusky_b200 import autograd as nba
    from oracle import neusky_oracle as O

    sdf, grad, albedo, ray_dirs, starts, ends, dnorm = _case(R, S, R * 31 + S)
    inv_s0 = 12.0
    g = torch.Generator().manual_seed(5)
    # random cotangents for every output
    cot = {"weights": torch.randn(R, S, generator=g), "wa": torch.randn(R, S, 3, generator=g), "normals": torch.randn(R, S, 3, generator=g),
           "accumulation": torch.randn(R, generator=g), "p2p_raw": torch.randn(R, generator=g), "normal": torch.randn(R, 3, generator=g),
           "albedo": torch.randn(R, 3, generator=g), "bg_transmittance": torch.randn(R, generator=g)}

    # ---- reference: fp64 autograd through the oracle ----
    a = [t.double().requires_grad_(True) for t in (sdf, grad, albedo)]
    inv_ref = torch.tensor(inv_s0, dtype=torch.float64, requires_grad=True)
    alpha = O.neus_alpha(a[0], a[1], ray_dirs.double()[:, None, :], (ends - starts).double(), inv_ref, rho)
    w, T = O.weights_from_alphas(alpha)
    normals = torch.nn.functional.normalize(a[1], dim=-1)
    steps = ((starts + ends) / 2).double()
    acc = w.sum(-2)
    outs = {"weights": w[..., 0], "wa": w * a[2], "normals": normals, "accumulation": acc[:, 0], "p2p_raw": ((w * steps).sum(-2) / (acc + 1e-10))[:, 0],
            "normal": (w * normals).sum(-2), "albedo": (w * a[2]).sum(-2) + (1.0 - acc), "bg_transmittance": T[:, -1, 0]}
    loss = sum((outs[k] * cot[k].double()).sum() for k in cot)
    loss.backward()

    # ---- CUDA ----
    dv = lambda t: t.to(dev).requires_grad_(True)
    sd, gr, al = dv(sdf), dv(grad), dv(albedo)
    inv_t = torch.tensor(inv_s0, device=dev, requires_grad=True)
    o = nba.neus_composite(sd, gr, al, inv_t, ray_dirs.to(dev), starts.to(dev), ends.to(dev), (ends - starts).to(dev), dnorm.to(dev), rho)
    names = ("weights", "wa", "normals", "accumulation", "p2p_raw", "normal", "albedo", "bg_transmittance")
    for k, t in zip(names, o):
        assert torch.allclose(t.detach().cpu().double(), outs[k].detach().reshape(t.shape), rtol=2e-4, atol=2e-6), k
    sum((t * cot[k].to(dev).reshape(t.shape)).sum() for k, t in zip(names, o)).backward()
    for name, got, ref in (("sdf", sd.grad, a[0].grad), ("grad", gr.grad, a[1].grad), ("albedo", al.grad, a[2].grad)):
        ref = ref.float()
        scale = ref.abs().max().clamp_min(1e-6)
        err = (got.cpu().reshape(ref.shape) - ref).abs().max() / scale
        assert float(err) <= 2e-3, (name, float(err))
    assert abs(float(inv_t.grad) - float(inv_ref.grad)) <= 2e-3 * max(1.0, abs(float(inv_ref.grad)))


def test_hash_encode_autograd_table_gradient(dev):
    from neusky_b200 import autograd as nba
    from oracle import neusky_oracle as O

    log2_T, L = 12, 16
    table = nb_init.init_hash_table(4, L, log2_T)
    sc = O.hash_scalings(L)
    x = torch.rand(2000, 3, generator=torch.Generator().manual_seed(2)) * 2 - 1
    t_ref = table.clone().requires_grad_(True)
    (O.hash_encode(x, t_ref, sc, log2_T) ** 2).sum().backward()
    t_gpu = table.to(dev).requires_grad_(True)
    (nba.hash_encode(x.to(dev), t_gpu, sc.to(dev), log2_T) ** 2).sum().backward()
    assert torch.allclose(t_gpu.grad.cpu(), t_ref.grad, rtol=1e-4, atol=1e-8)


def test_lambert_shade_and_finalize_backward_vs_oracle_autograd(dev):
    from neusky_b200 import autograd as nba, ops
    from oracle import neusky_oracle as O

    R, S, D, K = 13, 6, 70, 3
    g = torch.Generator().manual_seed(8)
    normals = torch.nn.functional.normalize(torch.randn(R, S, 3, generator=g), dim=-1)
    wa = torch.rand(R, S, 3, generator=g) / S
    dirs = torch.nn.functional.normalize(torch.randn(D, 3, generator=g), dim=-1)
    rad = torch.exp(0.5 * torch.randn(K, D, 3, generator=g))
    mask = dirs[:, 2] > 0
    Dp = int(mask.sum())
    vis_sel = torch.rand(R, Dp, generator=g)
    cam = torch.randint(0, K, (R,), generator=g)
    bg = torch.rand(R, 3, generator=g)
    acc = torch.rand(R, generator=g)
    cot = torch.randn(R, 3, generator=g)

    # reference: fp64 autograd through the oracle's renderer pieces
    n64, w64, r64, v64, b64, a64 = (t.double().requires_grad_(True) for t in (normals, wa, rad, vis_sel, bg, acc))
    vis_full = torch.ones(R, D, dtype=torch.float64)
    vis_full = vis_full.index_put((torch.arange(R)[:, None], torch.nonzero(mask)[:, 0][None, :].expand(R, Dp)), v64)
    lin = torch.zeros(R, 3, dtype=torch.float64)
    for s in range(S):
        lin = lin + w64[:, s] * O.lambertian_radiance(torch.ones(R, 3, dtype=torch.float64), n64[:, s], dirs.double(), r64[cam], vis_full)
    rgb_ref = O.linear_to_srgb(lin + b64 * (1.0 - a64[:, None]))
    (rgb_ref * cot.double()).sum().backward()

    dv = lambda t: t.to(dev).requires_grad_(True)
    n, w, r, v, b, a = dv(normals), dv(wa), dv(rad), dv(vis_sel), dv(bg), dv(acc)
    dirs_d = dirs.to(dev)
    m = mask.to(dev)
    sel_index = torch.where(m, torch.cumsum(m.to(torch.int32), 0, dtype=torch.int32) - 1, torch.full_like(m, -1, dtype=torch.int32)).to(torch.int32)
    inv_count, _ = ops.lambert_prep(n.detach(), w.detach(), dirs_d, m.to(torch.uint8), r.detach(), cam.to(dev, torch.int32), 1.0)
    lin_g = nba.lambert_shade(n, w, r, v, inv_count, dirs_d, sel_index, cam.to(dev, torch.int32))
    rgb = nba.shade_finalize(lin_g, b, a)
    assert torch.allclose(rgb.detach().cpu().double(), rgb_ref.detach(), rtol=1e-4, atol=1e-5)
    (rgb * cot.to(dev)).sum().backward()
    for name, got, ref in (("normals", n.grad, n64.grad), ("wa", w.grad, w64.grad), ("radiance", r.grad, r64.grad), ("vis", v.grad, v64.grad), ("bg", b.grad, b64.grad), ("acc", a.grad, a64.grad)):
        ref = ref.float()
        err = (got.cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-6)
        assert float(err) <= 2e-3, (name, float(err))


def test_hash_encode_double_backward_normals_path(dev):
    """The reference's normals: n = d sdf / d x by autograd with create_graph, and a loss on n back-propagated into the
    hash table (sdf_albedo_field.py:235-238 + the eikonal loss).  Here sdf = sum(W . feat) with a fixed random W."""
    from neusky_b200 import autograd as nba
    from oracle import neusky_oracle as O

    log2_T, L = 10, 16
    table = nb_init.init_hash_table(6, L, log2_T) * 300.0
    sc = O.hash_scalings(L)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1500, 3, generator=g) * 1.6 - 0.8
    W = torch.randn(2 * L, generator=g)
    cot = torch.randn(1500, 3, generator=g)

    def run(enc, xx, tt, Wd, cotd):
        xx = xx.clone().requires_grad_(True)
        sdf = (enc(xx, tt) * Wd).sum(-1)
        (n,) = torch.autograd.grad(sdf.sum(), xx, create_graph=True)
        loss = (n * cotd).sum() + 0.1 * (sdf ** 2).sum()
        loss.backward()
        return n.detach(), tt.grad

    t_ref = table.double().requires_grad_(True)
    n_ref, gt_ref = run(lambda a, b: O.hash_encode(a, b, sc, log2_T), x.double(), t_ref, W.double(), cot.double())
    t_gpu = table.to(dev).requires_grad_(True)
    n_gpu, gt_gpu = run(lambda a, b: nba.hash_encode(a, b, sc.to(dev), log2_T), x.to(dev), t_gpu, W.to(dev), cot.to(dev))
    assert float((n_gpu.cpu() - n_ref.float()).abs().max() / n_ref.abs().max()) <= 1e-4
    assert float((gt_gpu.cpu() - gt_ref.float()).abs().max() / gt_ref.abs().max()) <= 1e-3


@pytest.mark.parametrize("mode,rot,log_domain,with_scale", [("table", False, True, True), ("table", True, True, True), ("rows", False, True, True),
                                                           ("rows", True, False, True), ("table", False, True, False)])
def test_reni_radiance_backward_vs_oracle_autograd(dev, mode, rot, log_domain, with_scale):
    """d/d latent codes and d/d scale of the RENI++ radiance (decoder frozen) against fp64 torch autograd through the oracle's
    restatement of RENIField.get_outputs + unnormalise (reni_illumination_field.py:493-573, base_spherical_field.py:143-154);
    table mode = every (code, direction) pair (neusky_model.py:488-504), rows mode = one code per camera ray (:535-549)."""
    from neusky_b200 import autograd as nba, packing
    from oracle import neusky_oracle as O

    K, D = 3, 37            # D not a multiple of the 8-row block
    p = nb_init.init_reni_params(8)
    g = torch.Generator().manual_seed(17)
    Z = torch.randn(K, 100, 3, generator=g)
    sc = 0.2 * torch.randn(K, generator=g)
    dirs = torch.nn.functional.normalize(torch.randn(D, 3, generator=g), dim=-1)
    R = None
    if rot:
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
        R = q.contiguous()
    cam = torch.randint(0, K, (D,), generator=g)
    cot = torch.randn((K, D, 3) if mode == "table" else (D, 3), generator=g)

    # ---- oracle, fp64 autograd
    Zr, scr = Z.double().requires_grad_(True), sc.double().requires_grad_(True)
    pd = {k: v.double() for k, v in p.items()}
    Rd = None if R is None else R.double()
    if mode == "table":
        ref = torch.stack([O.reni_unnormalise(O.reni_field(dirs.double(), Zr[k : k + 1].expand(D, -1, -1), scr[k : k + 1].expand(D) if with_scale else None, pd, Rd, log_domain), log_domain)
                           for k in range(K)], 0)
    else:
        ref = O.reni_unnormalise(O.reni_field(dirs.double(), Zr[cam], scr[cam] if with_scale else None, pd, Rd, log_domain), log_domain)
    (ref * cot.double()).sum().backward()

    # ---- CUDA
    pc = {k: v.to(dev) for k, v in p.items()}
    blob, blob_b = packing.pack_reni(pc), packing.pack_reni_bwd(pc)
    Zc = Z.to(dev).requires_grad_(True)
    scc = sc.to(dev).requires_grad_(True) if with_scale else None
    out = nba.reni_radiance(dirs.to(dev), Zc, scc, blob, blob_b, row_cam=cam.to(dev, torch.int32) if mode == "rows" else None,
                            rotation=None if R is None else R.to(dev), log_domain=log_domain)
    assert float((out.detach().cpu().double() - ref.detach()).abs().max() / ref.detach().abs().max()) <= 1e-3
    (out * cot.to(dev)).sum().backward()
    rel = lambda a, b: float((a.double().cpu() - b).norm() / (b.norm() + 1e-30))
    assert rel(Zc.grad, Zr.grad) <= 2e-3, rel(Zc.grad, Zr.grad)
    if with_scale:
        assert rel(scc.grad, scr.grad) <= 2e-3, rel(scc.grad, scr.grad)
