"""The reference-named plugin surface (neusky_b200/plugin.py) driven the way the reference drives its own classes,
against fixtures produced by the reference's code (tests/golden) and the CPU oracle."""
from types import SimpleNamespace

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


def _ray_samples(origins, directions, starts, ends, cam=None):
    fr = SimpleNamespace(origins=origins, directions=directions, starts=starts, ends=ends)
    return SimpleNamespace(frustums=fr, deltas=None if ends is None else ends - starts, camera_indices=cam)


def test_compute_visibility_signature_vs_reference_golden(dev, golden):
    """Same call as tests/golden/make_golden.py makes on the reference's NeuSkyFactoModel.compute_visibility."""
    from neusky_b200.plugin import NeuSkyVisibility

    g = golden("visibility")
    p = nb_init.init_ddf_params(int(g["seed"]), final_gain=float(g["final_gain"]))
    o, d, p2p, dirs = (torch.from_numpy(g[k]).to(dev) for k in ("origins", "ray_dirs", "p2p", "dirs"))
    R, S, D = o.shape[0], 3, dirs.shape[0]
    rs = _ray_samples(o[:, None].expand(R, S, 3).contiguous(), d[:, None].expand(R, S, 3).contiguous(), torch.zeros(R, S, 1, device=dev), torch.ones(R, S, 1, device=dev))
    illum = dirs[None].expand(R * S, D, 3)
    for impl, tol in (("simt", 5e-4), ("tc", 2e-2), ("tc2", 2e-2)):
        m = NeuSkyVisibility(p, device=dev, impl=impl)
        vd = m.compute_visibility(rs, p2p, illum, float(g["threshold"]), float(g["sigmoid_scale"]))
        assert vd["visibility"].shape == (R * S, D, 1)
        vis = vd["visibility"].reshape(R, S, D)
        assert torch.equal(vis[:, 0], vis[:, 2])
        assert float((vis[:, 0].cpu() - torch.from_numpy(g["visibility"])).abs().max()) <= tol
        assert torch.allclose(vd["visibility_batch"]["termination_dist"].cpu(), torch.from_numpy(g["termination_dist"]), rtol=1e-5, atol=2e-6)


def test_reni_field_forward_vs_reference_golden(dev, golden):
    from neusky_b200.plugin import RENIField, RENIFieldHeadNames

    g = golden("reni")
    f = RENIField(nb_init.init_reni_params(int(g["seed"])), device=dev)
    dirs, Z, sc = (torch.from_numpy(g[k]).to(dev) for k in ("dirs", "latents", "scale"))
    K, D = Z.shape[0], dirs.shape[0]
    # the reference's calling convention (neusky_model.py:470-493): one row per (camera, direction)
    rs = _ray_samples(None, dirs[None].expand(K, D, 3).reshape(-1, 3), None, None, cam=torch.arange(K, device=dev)[:, None].expand(K, D).reshape(-1, 1))
    lat = Z[:, None].expand(K, D, *Z.shape[1:]).reshape(K * D, *Z.shape[1:])
    out = f(rs, None, lat, sc[:, None].expand(K, D).reshape(-1))
    rad = f.unnormalise(out[RENIFieldHeadNames.RGB]).reshape(K, D, 3).cpu()
    ref = torch.from_numpy(g["radiance"])
    assert torch.allclose(rad, ref, rtol=1e-3, atol=1e-6)
    with pytest.raises(NotImplementedError):
        f(rs, torch.eye(3, device=dev)[None].expand(4, 3, 3), lat, None)


def test_sdf_albedo_field_forward(dev):
    from neusky_b200.plugin import FieldHeadNames, NeuSkyFieldHeadNames, SDFAlbedoField
    from oracle import neusky_oracle as O

    log2_T = 14
    p = nb_init.init_sdf_params(8, log2_T=log2_T, bias=0.4)
    R, S = 9, 17
    g = torch.Generator().manual_seed(5)
    o = torch.tensor([0.0, -0.8, 0.2]).expand(R, 3) + 0.02 * torch.randn(R, 3, generator=g)
    d = torch.nn.functional.normalize(-o + 0.3 * torch.randn(R, 3, generator=g), dim=-1)
    t = torch.sort(torch.rand(R, S + 1, generator=g) * 1.2 + 0.1, dim=1).values
    starts, ends = t[:, :-1, None], t[:, 1:, None]
    x = o[:, None] + d[:, None] * starts
    ref = O.sdf_field(x.reshape(-1, 3), p, O.hash_scalings(), log2_T)
    inv_s = float(torch.exp(torch.tensor(10 * 0.1)))
    ref_alpha = O.neus_alpha(ref["sdf"].reshape(R, S, 1), ref["gradient"].reshape(R, S, 3), d[:, None], ends - starts, inv_s)
    for impl, tol in (("simt", 1e-4), ("tc", 3e-3)):
        f = SDFAlbedoField(p, device=dev, log2_T=log2_T, impl=impl)
        rs = _ray_samples(o[:, None].expand(R, S, 3).to(dev), d[:, None].expand(R, S, 3).to(dev), starts.to(dev), ends.to(dev), cam=torch.zeros(R, S, 1, dtype=torch.long, device=dev))
        out = f(rs, return_alphas=True)
        assert out[FieldHeadNames.SDF].shape == (R, S, 1) and out[NeuSkyFieldHeadNames.ALBEDO].shape == (R, S, 3)
        assert float((out[FieldHeadNames.SDF].cpu().reshape(-1, 1) - ref["sdf"]).abs().max()) <= tol
        assert float((out[FieldHeadNames.ALPHA].cpu() - ref_alpha).abs().max()) <= 10 * tol
        assert float((out[NeuSkyFieldHeadNames.ALBEDO].cpu().reshape(-1, 3) - ref["albedo"]).abs().max()) <= 10 * tol
        sd = f.get_sdf_at_pos(x.reshape(-1, 3).to(dev))
        assert sd.shape == (R * S, 1) and float((sd.cpu() - ref["sdf"]).abs().max()) <= 1e-4
        rs.camera_indices = None
        with pytest.raises(AttributeError):
            f(rs)
    assert abs(float(f.deviation_network.get_variance()) - inv_s) < 1e-4


def test_lambertian_renderer_compact_vs_reference_golden(dev, golden):
    from neusky_b200.plugin import RGBLambertianRendererWithVisibility

    g = golden("lambert")
    t = {k: torch.from_numpy(g[k]).to(dev) for k in g.files}
    R = t["albedo"].shape[0]
    rgb = RGBLambertianRendererWithVisibility()(t["albedo"], t["normals"], t["dirs"], t["light"], t["visibility"], t["bg"], t["weights"],
                                                camera_rows=torch.arange(R, dtype=torch.int32, device=dev))
    assert torch.allclose(rgb.cpu(), t["rgb"].cpu(), rtol=1e-4, atol=1e-5), (rgb.cpu() - t["rgb"].cpu()).abs().max()
