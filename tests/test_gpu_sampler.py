"""GPU parity of the proposal-network sampler (csrc/proposal_sampler.cu through the C ABI) against
oracle/sampler_oracle.py.  Bars: spacing / euclidean bins BIT-EXACT given the same weights (north_star: "sample placement must
match bit-exactly"); densities and weights within 1e-5 relative (fp32 MLP sum order, expf ulp); end-to-end placement within 2e-5
absolute; gradients within 1e-3 relative of fp64 autograd through the oracle."""
import numpy as np
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


@pytest.fixture(scope="module")
def SO():
    from oracle import sampler_oracle

    return sampler_oracle


def _rays(n, seed=0):
    from oracle import neusky_oracle as O

    c2w = O.look_at_camera((0.0, -0.9, 0.25))
    o, d, _ = O.pinhole_rays(n, n, float(n), float(n), n / 2, n / 2, c2w)
    near, far = O.sphere_collider(o, d)
    return o.contiguous(), d.contiguous(), near.reshape(-1).contiguous(), far.reshape(-1).contiguous()


@pytest.mark.parametrize("S", [256, 96, 7])
def test_uniform_bins_bit_exact(dev, SO, S):
    from neusky_b200 import proposal as P

    R = 33
    assert np.array_equal(P.uniform_bins(R, S, dev).cpu().numpy(), SO.uniform_bins(R, S))
    jit = torch.rand(R, generator=torch.Generator().manual_seed(S))
    assert np.array_equal(P.uniform_bins(R, S, dev, jit.to(dev)).cpu().numpy(), SO.uniform_bins(R, S, jit.numpy()))
    assert P.uniform_bins(0, S, dev).shape == (0, S + 1)


@pytest.mark.parametrize("S,N", [(256, 96), (96, 48), (37, 5), (1, 3), (512, 200)])
@pytest.mark.parametrize("train", [False, True])
def test_pdf_resample_bit_exact_given_weights(dev, SO, S, N, train):
    from neusky_b200 import proposal as P

    rng = np.random.default_rng(S * 1000 + N)
    R = 131
    bins = SO.uniform_bins(R, S, rng.random(R).astype(np.float32))
    w = (rng.random((R, S)) ** 8).astype(np.float32)
    w /= np.maximum(w.sum(-1, keepdims=True), 1e-6)
    w[0] = 0.0                                # empty ray
    w[1] = 0.0; w[1, S // 2] = 1.0            # delta histogram
    w[2] = 1.0 / S                            # uniform
    near = rng.random(R).astype(np.float32) * 0.5
    far = near + 0.5 + rng.random(R).astype(np.float32)
    jit = rng.random(R).astype(np.float32) if train else None
    ref = SO.pdf_resample(bins, w, N, jit)
    ref_e = SO.spacing_to_euclidean(ref, near[:, None], far[:, None])
    t = lambda a: None if a is None else torch.from_numpy(a).to(dev)
    nb, ne, _ = P.pdf_resample(t(bins), t(near), t(far), N, weights=t(w), jitter=t(jit))
    assert np.array_equal(nb.cpu().numpy(), ref), f"max |d| = {np.abs(nb.cpu().numpy() - ref).max()}"
    assert np.array_equal(ne.cpu().numpy(), ref_e)


def test_pdf_resample_anneal_and_zero_padding(dev, SO):
    from neusky_b200 import proposal as P

    rng = np.random.default_rng(3)
    R, S, N = 64, 96, 48
    bins = SO.uniform_bins(R, S)
    w = (rng.random((R, S)) ** 4).astype(np.float32) * 0.05
    near, far = np.zeros(R, np.float32), np.ones(R, np.float32) * 2
    t = lambda a: torch.from_numpy(a).to(dev)
    ref = SO.pdf_resample(bins, np.power(w, np.float32(0.5), dtype=np.float32), N)
    nb, _, _ = P.pdf_resample(t(bins), t(near), t(far), N, weights=t(w), anneal=0.5)
    assert np.abs(nb.cpu().numpy() - ref).max() <= 1e-6            # powf differs by ulps between libm and CUDA
    # histogram_padding = 0 with all-zero weights: the eps padding path (uniform result, nothing NaN)
    nb0, _, _ = P.pdf_resample(t(bins), t(near), t(far), N, weights=t(np.zeros((R, S), np.float32)), histogram_padding=0.0)
    ref0 = SO.pdf_resample(bins, np.zeros((R, S), np.float32), N, histogram_padding=0.0)
    assert np.array_equal(nb0.cpu().numpy(), ref0)


@pytest.mark.parametrize("max_res", [64, 256])
def test_proposal_density_vs_oracle(dev, SO, max_res):
    from neusky_b200 import proposal as P

    o, d, near, far = _rays(12)
    R, S = o.shape[0], 64
    p = SO.init_proposal_net(max_res, table_scale=1.0)
    f = P.HashMLPDensityField(p, max_res, device=dev)
    bins = SO.uniform_bins(R, S, np.random.default_rng(0).random(R).astype(np.float32))
    e = SO.spacing_to_euclidean(bins, near.numpy()[:, None], far.numpy()[:, None])
    mids = torch.from_numpy((e[:, :-1] + e[:, 1:]) / 2)
    pos = o[:, None, :] + d[:, None, :] * mids[..., None]
    ref = SO.proposal_density(pos, p, SO.proposal_scalings(max_res))
    got = f.density_on_rays(o.to(dev), d.to(dev), near.to(dev), far.to(dev), torch.from_numpy(bins).to(dev)).cpu()
    assert float(((got - ref).abs() / (ref.abs() + 1e-6)).max()) <= 2e-5
    # positions mode (density_fn), including points outside the cube (contraction) and far away (selector -> 0)
    g = torch.Generator().manual_seed(1)
    x = torch.cat([torch.randn(500, 3, generator=g) * 1.5, torch.tensor([[1e9, 0.0, 0.0], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [-3.0, 2.0, 0.5]])])
    refx = SO.proposal_density(x, p, SO.proposal_scalings(max_res))
    gotx = f.density_fn(x.to(dev)).cpu()[:, 0]
    assert float(((gotx - refx).abs() / (refx.abs() + 1e-6)).max()) <= 2e-5
    assert float(gotx[500]) == 0.0


def test_proposal_sampler_end_to_end_eval_and_train(dev, SO):
    from neusky_b200 import proposal as P

    o, d, near, far = _rays(16)
    R = o.shape[0]
    nets = [SO.init_proposal_net(1, table_scale=1.0, density_bias=1.0), SO.init_proposal_net(2, table_scale=1.0, density_bias=2.0)]
    fields = [P.HashMLPDensityField(nets[0], 64, device=dev), P.HashMLPDensityField(nets[1], 256, device=dev)]
    sampler = P.ProposalNetworkSampler()
    for train in (False, True):
        jit = [np.random.default_rng(7 + i).random(R).astype(np.float32) for i in range(3)] if train else None
        anneal = 0.7 if train else 1.0
        e_ref, wl_ref, sl_ref, el_ref = SO.proposal_sample(o, d, near, far, nets, anneal=anneal, jitters=jit)
        sampler.training = train
        sampler.set_anneal(anneal)
        rs, wl, sl = sampler.generate_ray_samples(o.to(dev), d.to(dev), near.to(dev), far.to(dev), fields,
                                                  jitters=None if jit is None else [torch.from_numpy(j).to(dev) for j in jit])
        assert rs.euclidean_bins.shape == (R, 49) and [tuple(w.shape) for w in wl] == [(R, 256, 1), (R, 96, 1)]
        assert np.array_equal(sl[0].spacing_bins.cpu().numpy(), sl_ref[0])                      # level 0 is bit-exact by construction
        for w, wr in zip(wl, wl_ref):
            assert float(np.abs(w[..., 0].cpu().numpy() - wr).max()) <= 2e-5
        assert float(np.abs(sl[1].spacing_bins.cpu().numpy() - sl_ref[1]).max()) <= 2e-5
        assert float(np.abs(rs.euclidean_bins.cpu().numpy() - e_ref).max()) <= 5e-5
        eb = rs.euclidean_bins.cpu().numpy()
        assert np.all(np.diff(eb, axis=-1) >= 0)
        assert np.array_equal(rs.frustums.starts[..., 0].cpu().numpy(), eb[:, :-1]) and np.array_equal(rs.frustums.ends[..., 0].cpu().numpy(), eb[:, 1:])


def test_density_weights_and_interlevel_backward_vs_fp64_autograd(dev, SO):
    from neusky_b200 import proposal as P

    rng = np.random.default_rng(11)
    R, Sp, Sf = 40, 96, 48
    near = (rng.random(R) * 0.3).astype(np.float32)
    far = (near + 1.0 + rng.random(R)).astype(np.float32)
    cp = np.sort(rng.random((R, Sp + 1)).astype(np.float32), -1); cp[:, 0] = 0; cp[:, -1] = 1
    c = np.sort(rng.random((R, Sf + 1)).astype(np.float32), -1); c[:, 0] = 0; c[:, -1] = 1
    dens = np.exp(rng.normal(size=(R, Sp)) * 1.5 + 1.0).astype(np.float32)
    w = rng.random((R, Sf)).astype(np.float32); w /= w.sum(-1, keepdims=True) * 1.2
    t = lambda a: torch.from_numpy(a).to(dev)
    # oracle, fp64
    e = torch.from_numpy(SO.spacing_to_euclidean(cp, near[:, None], far[:, None])).double()
    dref = torch.from_numpy(dens).double().requires_grad_(True)
    wp_ref = SO.density_weights_torch(dref, e[:, 1:] - e[:, :-1])
    loss_ref = SO.interlevel_loss([wp_ref, torch.from_numpy(w).double()], [torch.from_numpy(cp).double(), torch.from_numpy(c).double()])
    loss_ref.backward()
    # CUDA
    _, _, wp = P.pdf_resample(t(cp), t(near), t(far), Sf, density=t(dens), want_euclid=False)
    assert float((wp.cpu().double() - wp_ref.detach()).abs().max()) <= 1e-6
    loss, g_wp = P.interlevel_loss_level(t(c), t(w), t(cp), wp, want_grad=True)
    assert abs(float(loss) - float(loss_ref)) <= 1e-5 * abs(float(loss_ref)) + 1e-9
    g_d = P.density_weights_bwd(t(cp), t(dens), t(near), t(far), g_wp)
    rel = float((g_d.cpu().double() - dref.grad).norm() / dref.grad.norm())
    assert rel <= 1e-3, rel


def test_proposal_density_backward_vs_fp64_autograd(dev, SO):
    from neusky_b200 import proposal as P
    from oracle import neusky_oracle as O

    o, d, near, far = _rays(10)
    R, S = o.shape[0], 32
    p = SO.init_proposal_net(5, log2_T=12, table_scale=1.0)
    f = P.HashMLPDensityField(p, 64, log2_hashmap_size=12, device=dev)
    bins = SO.uniform_bins(R, S)
    g = torch.randn(R, S, generator=torch.Generator().manual_seed(3))
    e = SO.spacing_to_euclidean(bins, near.numpy()[:, None], far.numpy()[:, None])
    mids = torch.from_numpy((e[:, :-1] + e[:, 1:]) / 2)
    pos = (o[:, None, :] + d[:, None, :] * mids[..., None]).double()
    pd = {k: v.double().requires_grad_(True) for k, v in p.items()}
    x = O.scene_contraction_linf(pos.reshape(-1, 3))
    x = (x + 2.0) / 4.0
    sel = ((x > 0.0) & (x < 1.0)).all(dim=-1)
    feat = O.hash_encode((x * sel[:, None]).float(), pd["encoding.hash_table"], SO.proposal_scalings(64), 12)
    h = torch.relu(feat @ pd["mlp.0.weight"].T + pd["mlp.0.bias"])
    dens = (torch.exp((h @ pd["mlp.1.weight"].T + pd["mlp.1.bias"])[:, 0]) * sel).reshape(R, S)
    (dens * g.double()).sum().backward()
    f.backward_on_rays(o.to(dev), d.to(dev), near.to(dev), far.to(dev), torch.from_numpy(bins).to(dev), g.to(dev))
    for k in p:
        got, ref = f.params[k].grad.cpu().double(), pd[k].grad
        rel = float((got - ref).norm() / (ref.norm() + 1e-30))
        assert rel <= 1e-3, (k, rel)
