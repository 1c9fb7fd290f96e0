"""Product-side direction samplers (neusky_b200/samplers.py) against fixtures produced by the reference's own
IcosahedronSampler / EquirectangularSampler (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import torch

from neusky_b200 import samplers


def test_icosphere_bit_exact_vs_reference(golden):
    g = golden("icosphere")
    for n, D, up in ((100, 162, 73), (256, 362, 169), (512, 642, 308)):
        d = samplers.IcosahedronSampler(n)().frustums.directions.numpy()
        assert d.shape == (D, 3)
        assert np.array_equal(d.view(np.uint32), g[f"dirs_{n}"].view(np.uint32)), f"icosphere {n}: vertex order / rounding differs from the reference"
        assert int((d[:, 2] > 0).sum()) == up


def test_icosphere_random_rotation_is_a_rotation():
    s = samplers.IcosahedronSampler(100, apply_random_rotation=True, seed=1)
    a, b = s().frustums.directions, s().frustums.directions
    assert not torch.allclose(a, b)
    assert torch.allclose(a.norm(dim=-1), torch.ones(162), atol=1e-5)
    G = s.directions @ s.directions.T
    assert torch.allclose(a @ a.T, G, atol=1e-5)      # pairwise angles preserved


def test_equirect_matches_oracle_and_bench():
    from oracle import neusky_oracle as O
    import bench

    d = samplers.EquirectangularSampler(64)().frustums.directions
    assert d.shape == (2048, 3) and int((d[:, 2] > 0).sum()) == 1024 and int((d[:, 2] == 0).sum()) == 0
    assert torch.equal(d, O.equirect_directions(64))
    assert torch.equal(d, bench._equirect_directions(64))


def test_icosphere_upper_hemisphere_count_is_rotation_invariant():
    """The subdivided icosahedron is centrally symmetric (v in the set <=> -v in the set), so under ANY rotation exactly half of its
    directions have z > 0 (up to vertices exactly on the equator, a measure-zero event).  The training iteration's shapes are therefore
    static -- D' = D / 2 -- which is what lets neusky_b200/graphed.py capture it once as a CUDA graph."""
    from scipy.spatial.transform import Rotation

    base = samplers.IcosahedronSampler(512)().frustums.directions.to(torch.float64).numpy()      # [642, 3]
    assert base.shape == (642, 3)
    # central symmetry: every direction has its antipode in the set
    d = np.abs(base[:, None, :] + base[None, :, :]).max(-1)                                       # |v_i + v_j|
    assert float(d.min(axis=1).max()) < 1e-6
    rots = Rotation.random(50, random_state=np.random.RandomState(3)).as_matrix()
    for Rm in rots:
        z = (base @ Rm).astype(np.float32)[:, 2]
        assert int((z > 0).sum()) == 321
    # the sampler's own rotated draws (what the training loop consumes)
    smp = samplers.IcosahedronSampler(512, apply_random_rotation=True, seed=5)
    for _ in range(10):
        dirs = smp().frustums.directions
        assert int((dirs[:, 2] > 0).sum()) == 321
    # the UN-rotated set has 26 vertices exactly on the equator: 308 strictly above it (the eval configuration's D')
    assert int((samplers.IcosahedronSampler(512)().frustums.directions[:, 2] > 0).sum()) == 308
