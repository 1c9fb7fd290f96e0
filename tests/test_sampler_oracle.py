"""CPU tests of the proposal-sampler oracle (oracle/sampler_oracle.py): internal consistency, agreement of the explicit-order
numpy statement with a restatement in the torch calls nerfstudio itself makes, and edge cases.  The sampler lives in
nerfstudio (absent, un-pinned): parity unpinned, see the oracle's header."""
import numpy as np
import torch

from oracle import neusky_oracle as O
from oracle import sampler_oracle as SO


def _rays(n=8):
    c2w = O.look_at_camera((0.0, -0.9, 0.25))
    o, d, _ = O.pinhole_rays(n, n, float(n), float(n), n / 2, n / 2, c2w)
    near, far = O.sphere_collider(o, d)
    return o, d, near, far


def test_uniform_bins_eval_is_linspace_and_jitter_stays_inside():
    b = SO.uniform_bins(3, 256)
    assert b.shape == (3, 257) and b[0, 0] == 0.0 and b[0, -1] == 1.0
    assert np.array_equal(b[0], torch.linspace(0, 1, 257).numpy())
    j = SO.uniform_bins(4, 16, np.array([0.0, 0.25, 0.5, 0.999], np.float32))
    assert np.all(np.diff(j, axis=-1) > 0) and j.min() >= 0.0 and j.max() <= 1.0
    base = torch.linspace(0, 1, 17).numpy()
    assert np.array_equal(j[0, 1:], (base[1:] + base[:-1]) / np.float32(2)) and j[0, 0] == 0.0   # t_rand = 0: the lower brackets


def test_pdf_resample_matches_torch_restatement_to_a_few_ulp():
    rng = np.random.default_rng(0)
    for S, N in ((256, 96), (96, 48), (37, 5)):
        bins = SO.uniform_bins(40, S, rng.random(40).astype(np.float32))
        w = (rng.random((40, S)) ** 8).astype(np.float32)
        w /= w.sum(-1, keepdims=True)
        a = SO.pdf_resample(bins, w, N)
        b = SO.pdf_resample_torch(torch.from_numpy(bins), torch.from_numpy(w), N).numpy()
        assert a.shape == (40, N + 1)
        assert np.abs(a - b).max() <= 4e-7          # torch.sum's vectorised order moves the pdf by <= 1 ulp
        assert np.all(np.diff(a, axis=-1) >= 0)


def test_pdf_resample_edge_cases():
    S, N = 16, 7
    bins = SO.uniform_bins(3, S)
    w = np.zeros((3, S), np.float32)
    w[1, 5] = 1.0                     # delta histogram
    w[2] = 1e-12                      # (almost) empty ray: uniform after padding
    out = SO.pdf_resample(bins, w, N)
    u = np.linspace(0, 1 - 1 / (N + 1), N + 1) + 1 / (2 * (N + 1))
    assert np.allclose(out[0], u, atol=1e-6) and np.allclose(out[2], u, atol=1e-6)   # uniform pdf -> the stratified positions themselves
    inside = (out[1] >= bins[1, 5]) & (out[1] <= bins[1, 6])
    assert inside.sum() >= N - 1       # nearly all mass in bin 5 (padding 0.01 leaves 14 % outside)
    # zero histogram_padding and all-zero weights: the eps padding path
    out0 = SO.pdf_resample(bins, np.zeros((3, S), np.float32), N, histogram_padding=0.0)
    assert np.all(np.isfinite(out0)) and np.allclose(out0[0], u, atol=1e-6)


def test_density_weights_match_torch_and_sum_below_one():
    rng = np.random.default_rng(1)
    dens = np.exp(rng.normal(size=(5, 64)) * 3).astype(np.float32)
    deltas = np.full((5, 64), 0.02, np.float32)
    w = SO.density_weights(dens, deltas)
    wt = SO.density_weights_torch(torch.from_numpy(dens), torch.from_numpy(deltas)).numpy()
    assert np.abs(w - wt).max() <= 1e-6
    assert np.all(w >= 0) and np.all(w.sum(-1) <= 1 + 1e-5)


def test_proposal_density_selector_and_contraction():
    p = SO.init_proposal_net(3, table_scale=1.0)
    sc = SO.proposal_scalings(64)
    assert sc.tolist() == [16.0, 22.0, 31.0, 45.0, 63.0]       # float32 pow: the top level is 63, not 64 (cf. hash_scalings)
    assert SO.proposal_scalings(256).tolist() == [16.0, 32.0, 64.0, 128.0, 256.0]
    pos = torch.tensor([[0.1, 0.2, 0.3], [1e9, 0.0, 0.0], [0.0, 0.0, 0.0], [-0.999, 0.999, 0.5], [3.0, -2.0, 1.0]])
    d = SO.proposal_density(pos, p, sc)
    assert d.shape == (5,) and torch.all(d >= 0) and torch.isfinite(d).all()
    # far away: contraction maps to the cube boundary 2 - 1/mag -> (x+2)/4 = 1.0 exactly -> selector kills the density
    assert float(d[1]) == 0.0


def test_proposal_sample_eval_and_training():
    o, d, near, far = _rays(8)
    nets = [SO.init_proposal_net(1, table_scale=1.0), SO.init_proposal_net(2, table_scale=1.0)]
    e, wl, sl, el = SO.proposal_sample(o, d, near, far, nets)
    assert e.shape == (64, 49) and [w.shape for w in wl] == [(64, 256), (64, 96)] and [s.shape for s in sl] == [(64, 257), (64, 97), (64, 49)]
    assert np.all(np.diff(e, axis=-1) >= 0)
    nr, fr = near.numpy(), far.numpy()
    assert np.all(e >= nr - 1e-6) and np.all(e <= fr + 1e-6)
    rng = np.random.default_rng(5)
    jit = [rng.random(64).astype(np.float32) for _ in range(3)]
    e2, *_ = SO.proposal_sample(o, d, near, far, nets, anneal=0.6, jitters=jit)
    assert np.all(np.diff(e2, axis=-1) >= 0) and not np.array_equal(e, e2)


def test_interlevel_loss_zero_when_proposal_bounds_fine():
    # a proposal histogram that is an upper envelope of the fine one gives zero loss; a deficient one gives a positive loss
    c = torch.linspace(0, 1, 9)[None]
    w = torch.full((1, 8), 0.1)
    cp = torch.linspace(0, 1, 5)[None]
    wp_hi = torch.full((1, 4), 0.25)
    wp_lo = torch.full((1, 4), 0.01)
    assert float(SO.interlevel_loss([wp_hi, w], [cp, c])) == 0.0
    assert float(SO.interlevel_loss([wp_lo, w], [cp, c])) > 0.0
