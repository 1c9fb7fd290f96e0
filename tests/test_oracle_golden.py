"""The oracle restatement vs. fixtures produced by the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import torch

from neusky_b200 import init as nb_init
from oracle import neusky_oracle as O


def _sha(params):
    h = hashlib.sha256()
    for k in sorted(params):
        h.update(k.encode())
        h.update(params[k].detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def test_icosphere_bit_exact(golden):
    g = golden("icosphere")
    for n, D, up in ((100, 162, 73), (256, 362, 169), (512, 642, 308)):
        ours = O.icosphere_directions(n).numpy()
        ref = g[f"dirs_{n}"]
        assert ours.shape == (D, 3)
        assert np.array_equal(ours.view(np.uint32), ref.view(np.uint32)), f"icosphere {n} differs bitwise"
        assert int((ours[:, 2] > 0).sum()) == up


def test_lambert_renderer(golden):
    g = golden("lambert")
    t = {k: torch.from_numpy(g[k]) for k in g.files}
    rgb = O.lambertian_render(t["albedo"], t["normals"], t["dirs"], t["light"], t["visibility"], t["bg"], t["weights"], training=False)
    assert torch.allclose(rgb, t["rgb"], rtol=1e-5, atol=1e-6)


def test_reni_field(golden):
    g = golden("reni")
    p = nb_init.init_reni_params(int(g["seed"]))
    assert _sha(p) == str(g["weights_sha256"]), "seeded RNG stream drifted: regenerate goldens"
    dirs, Z, sc = (torch.from_numpy(g[k]) for k in ("dirs", "latents", "scale"))
    rad = O.reni_radiance_table(dirs, Z, sc, p)
    ref = torch.from_numpy(g["radiance"])
    assert torch.allclose(rad, ref, rtol=2e-4, atol=1e-6), (rad - ref).abs().max()
    rad_r = O.reni_radiance_table(dirs, Z, sc, p, rotation=torch.from_numpy(g["rotation"]))
    assert torch.allclose(rad_r, torch.from_numpy(g["radiance_rot"]), rtol=2e-4, atol=1e-6)


def test_ddf_model(golden):
    g = golden("ddf_model")
    p = nb_init.init_ddf_params(int(g["seed"]), final_gain=float(g["final_gain"]))
    assert _sha(p) == str(g["weights_sha256"]), "seeded RNG stream drifted: regenerate goldens"
    out = O.ddf_model(torch.from_numpy(g["positions"]), torch.from_numpy(g["directions"]), p, O.hash_scalings(), 19, 1.0)
    ref = torch.from_numpy(g["expected_termination_dist"])
    assert ref.std() > 0.05  # the fixture exercises a non-trivial output range
    assert torch.allclose(out, ref, rtol=1e-4, atol=2e-5), (out - ref).abs().max()


def test_visibility(golden):
    g = golden("visibility")
    p = nb_init.init_ddf_params(int(g["seed"]), final_gain=float(g["final_gain"]))
    o, d, p2p, dirs = (torch.from_numpy(g[k]) for k in ("origins", "ray_dirs", "p2p", "dirs"))
    pts = O.surface_points(o, d, p2p, 1.0)
    assert int((pts.norm(dim=-1) < 1.0).sum()) == pts.shape[0]  # hack branch pulled every point inside
    v = O.compute_visibility(pts, dirs, p, O.hash_scalings(), 19, 1.0, float(g["threshold"]), float(g["sigmoid_scale"]))
    assert torch.allclose(v["termination_dist"], torch.from_numpy(g["termination_dist"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(v["expected_termination_dist"], torch.from_numpy(g["expected_termination_dist"]), rtol=1e-4, atol=2e-5)
    ref = torch.from_numpy(g["visibility"])
    assert 0.05 < float(ref.mean()) < 0.999
    assert torch.allclose(v["visibility"], ref, rtol=0, atol=2e-4), (v["visibility"] - ref).abs().max()
