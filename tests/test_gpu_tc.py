"""Tensor-core (tcgen05, fp16 operands / fp32 accumulate) shading path vs the exact fp32 CUDA path and
the CPU oracle.  Tolerances for this path are stated here, separately from the fp32 path (BASELINE.json
north_star), and kept at ~3x what the B200 measures (gpurun_out/test_errors.jsonl via conftest.log_err;
profiles/r02_test_errors.jsonl is the committed copy):

  default kernel `tc2` (CTA pairs; first trunk layer carried as an fp16 hi/lo split):
    random-init DDF as the reference initialises it (gain 1): visibility |err| <= 1e-3 max  (measured 4.4e-4) -- north_star's bar
    x8 stress gain on the DDF output layer (the goldens): visibility <= 1e-2 max / 3e-4 mean (measured 3.3e-3 / 8e-5),
    DDF distance <= 2.5e-3 (7e-4), shaded linear RGB <= 1.5e-3 relative (4.4e-4)
  single-CTA variant `tc` (no hi/lo split; kept as the bring-up kernel): 2e-2 / 5e-4 / 4e-3 / 3e-3."""
import pytest
import torch

from conftest import log_err
from neusky_b200 import init as nb_init

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from neusky_b200 import _lib

    _lib.load()
    return torch.device("cuda:0")


def _scene(R, D, S, seed):
    g = torch.Generator().manual_seed(seed)
    pts = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1) * torch.rand(R, 1, generator=g) ** (1 / 3) * 0.95
    normals = torch.nn.functional.normalize(torch.randn(R, S, 3, generator=g), dim=-1)
    wa = torch.rand(R, S, 3, generator=g) / S
    dirs = torch.nn.functional.normalize(torch.randn(D, 3, generator=g), dim=-1)
    dirs[: max(1, D // 2), 2] = dirs[: max(1, D // 2), 2].abs() + 1e-3  # at least half the set is in the upper hemisphere
    dirs = torch.nn.functional.normalize(dirs, dim=-1)
    radiance = torch.exp(torch.randn(1, D, 3, generator=g))
    return pts, normals, wa, dirs, radiance


@pytest.mark.parametrize("impl", ["tc", "tc2"])
@pytest.mark.parametrize("R,D,S,gain", [(1, 1, 1, 8.0), (3, 50, 2, 8.0), (64, 162, 1, 8.0), (700, 642, 1, 1.0), (257, 300, 3, 8.0)])
def test_tc_vs_simt(dev, R, D, S, gain, impl):
    from neusky_b200.render import SkyShader

    p = nb_init.init_ddf_params(21, final_gain=gain)
    pts, normals, wa, dirs, radiance = _scene(R, D, S, R * 7 + D)
    sh = SkyShader(p, None, device=dev)
    sh.set_directions(dirs)
    args = (pts.to(dev), normals.to(dev), wa.to(dev), radiance.to(dev))
    ref = sh.shade(*args, want_vis=True, want_ddf=True, impl="simt")
    out = sh.shade(*args, want_vis=True, want_ddf=True, impl=impl)
    torch.cuda.synchronize()
    assert torch.allclose(out["termination_dist"], ref["termination_dist"], rtol=1e-5, atol=2e-6)
    e_ddf = (out["expected_termination_dist"] - ref["expected_termination_dist"]).abs()
    e_vis = (out["visibility"] - ref["visibility"]).abs()
    denom = ref["rgb_lin"].abs().clamp_min(1e-3)
    rel = ((out["rgb_lin"] - ref["rgb_lin"]).abs() / denom).max()
    log_err(f"tc_vs_simt[{impl},{R},{D},{S},{gain}]", ddf_max=e_ddf.max(), vis_max=e_vis.max(), vis_mean=e_vis.mean(), rgb_rel=rel)
    t_ddf, t_vmax, t_vmean, t_rel = (2.5e-3, 1e-2, 3e-4, 1.5e-3) if impl == "tc2" else (4e-3, 2e-2, 5e-4, 3e-3)
    if impl == "tc2" and gain == 1.0:
        t_vmax = 1e-3          # north_star's visibility bar on the reference's own initialisation
    assert float(e_ddf.max()) <= t_ddf, float(e_ddf.max())
    assert float(e_vis.max()) <= t_vmax and float(e_vis.mean()) <= t_vmean, (float(e_vis.max()), float(e_vis.mean()))
    assert float(rel) <= t_rel, float(rel)


def test_tc_vs_reference_golden(dev, golden):
    """The tensor-core path against visibility produced by the reference's own compute_visibility."""
    from neusky_b200 import ops
    from neusky_b200.render import SkyShader

    g = golden("visibility")
    p = nb_init.init_ddf_params(int(g["seed"]), final_gain=float(g["final_gain"]))
    sh = SkyShader(p, None, device=dev)
    o, d, p2p, dirs = (torch.from_numpy(g[k]).to(dev) for k in ("origins", "ray_dirs", "p2p", "dirs"))
    pts = ops.surface_points(o, d, p2p, 1.0)
    sh.set_directions(dirs)
    R, D = pts.shape[0], dirs.shape[0]
    out = sh.shade(pts, torch.ones(R, 1, 3, device=dev) / 3**0.5, torch.ones(R, 1, 3, device=dev), torch.ones(1, D, 3, device=dev),
                   want_vis=True, want_ddf=True, threshold=float(g["threshold"]), sigmoid_scale=float(g["sigmoid_scale"]))
    ref = torch.from_numpy(g["visibility"])
    e = (out["visibility"].cpu() - ref).abs()
    e_ddf = (out["expected_termination_dist"].cpu() - torch.from_numpy(g["expected_termination_dist"])).abs()
    log_err("tc_vs_reference_golden", vis_max=e.max(), vis_mean=e.mean(), ddf_max=e_ddf.max())
    assert float(e.max()) <= 1e-2 and float(e.mean()) <= 1.5e-4, (float(e.max()), float(e.mean()))      # measured 3.0e-3 / 4.2e-5 (x8 stress gain)
    assert float(e_ddf.max()) <= 2e-3                                                                       # measured 6.6e-4


@pytest.mark.parametrize("impl", ["tc", "tc2"])
def test_tc_multi_tile_persistent_and_no_vis_buffer(dev, impl):
    """More tiles than SMs (persistent loop, ring phases wrap many times); vis tensor not requested."""
    from neusky_b200.render import SkyShader

    p = nb_init.init_ddf_params(5, final_gain=8.0)
    R, D = 2048, 642
    pts, normals, wa, dirs, radiance = _scene(R, D, 1, 99)
    sh = SkyShader(p, None, device=dev)
    sh.set_directions(dirs)
    args = (pts.to(dev), normals.to(dev), wa.to(dev), radiance.to(dev))
    ref = sh.shade(*args, impl="simt")
    out = sh.shade(*args, impl=impl)
    out2 = sh.shade(*args, impl=impl)
    torch.cuda.synchronize()
    assert "visibility" not in out
    rel = ((out["rgb_lin"] - ref["rgb_lin"]).abs() / ref["rgb_lin"].abs().clamp_min(1e-3)).max()
    log_err(f"tc_multi_tile[{impl}]", rgb_rel=rel)
    assert float(rel) <= (1e-3 if impl == "tc2" else 3e-3), float(rel)      # measured 3.4e-4 / 7.2e-4
    # run-to-run: only the fp32 atomic summation order may differ
    assert torch.allclose(out["rgb_lin"], out2["rgb_lin"], rtol=1e-5, atol=1e-7)
