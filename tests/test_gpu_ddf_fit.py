"""DDF fitting pass (neusky_b200/ddf_fit.py) on the GPU against the reference's own outputs
(tests/golden/ddf_fit.npz) and against fp64 autograd through oracle/ddf_fit_oracle.py.

Tolerances: split=3 (3xTF32, fp32-accurate) -- DDF distances within 3e-4 of the reference's fp32 output, losses within 2e-4
relative of the fp64 oracle, gradients within 1e-2 norm-wise of fp64 autograd or no further from it than twice the fp32
oracle's own distance (the FiLM frequencies 15 f + 30 amplify fp32 rounding; see tests/test_gpu_train.py).
"""
import numpy as np
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


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_ddf_rows_forward_vs_reference_golden(dev, golden):
    """DDFModel.get_outputs of the reference on the sampled rays, the multi-view rows and the sky rows (full-size 2^19 table)."""
    from neusky_b200 import ddf_fit as F
    from neusky_b200 import train as T
    from oracle import neusky_oracle as O

    g = golden("ddf_fit")
    t = {k: torch.from_numpy(g[k]) for k in g.files if g[k].dtype == np.float32}
    ddf_p = {k: v.to(dev) for k, v in nb_init.init_ddf_params(int(g["seed"]), final_gain=float(g["final_gain"])).items()}
    cfg = T.DDFConfig(scalings=O.hash_scalings().to(dev), log2_T=19, split=3)
    that = T.ddf_termination(cfg, t["origins"].to(dev), t["directions"].to(dev), ddf_p["position_encoding.hash_table"],
                             ddf_p["ddf.final_layer.weight"], ddf_p["ddf.final_layer.bias"], T.ddf_param_list(ddf_p))
    assert float((that.cpu() - t["out_expected_termination_dist"]).abs().max()) <= 3e-4
    exit_pts = F.ray_sphere_exit(t["sky_origins"], t["sky_directions"], 1.0)
    sky = T.ddf_termination(cfg, exit_pts.to(dev), (-t["sky_directions"]).to(dev), ddf_p["position_encoding.hash_table"],
                            ddf_p["ddf.final_layer.weight"], ddf_p["ddf.final_layer.bias"], T.ddf_param_list(ddf_p))
    assert float((sky.cpu() - t["out_sky_ray_expected_termination_dist"]).abs().max()) <= 3e-4


def _setup(log2_T, S, dev, split=3):
    from neusky_b200 import train as T

    g = torch.Generator().manual_seed(91)
    sdf_p = nb_init.init_sdf_params(3, log2_T=log2_T)
    sdf_p["encoding.hash_table"] = (torch.rand(sdf_p["encoding.hash_table"].shape, generator=g) * 2 - 1) * 0.05
    for l in range(3):
        sdf_p[f"glin{l}.weight_v"] = sdf_p[f"glin{l}.weight_v"] + 0.02 * torch.randn(sdf_p[f"glin{l}.weight_v"].shape, generator=g)
    # geometric init is a sphere of radius ~0.1: enlarge it so that a good share of the vMF rays hit the surface
    sdf_p["glin2.bias"] = sdf_p["glin2.bias"].clone()
    sdf_p["glin2.bias"][0] -= 0.35
    sdf_p["deviation_network.variance"] = torch.tensor(0.25)
    ddf_p = nb_init.init_ddf_params(5, final_gain=8.0, log2_T=log2_T, table_scale=0.1)
    reni_p = nb_init.init_reni_params(8)
    step = T.NeuSkyTrainStep(sdf_p, ddf_p, reni_p, num_cameras=2, device=dev, log2_T=log2_T, num_samples=S, split_geo=split, split=split, threshold_init=0.4)
    return sdf_p, ddf_p, step


def test_ground_truth_render_vs_oracle(dev):
    from neusky_b200 import ddf_fit as F
    from oracle import ddf_fit_oracle as FO

    log2_T, S = 14, 24
    sdf_p, _, step = _setup(log2_T, S, dev)
    sampler = F.VMFDDFSampler(F.DDFSamplerConfig(num_samples_on_sphere=4, num_rays_per_sample=32), device=dev)
    torch.manual_seed(7)
    o, d = sampler()
    fit = F.DDFFit(step, sampler=sampler)
    with torch.no_grad():
        data = fit.generate_ddf_ground_truth(o, d, 0.5)
    ref = FO.generate_ddf_ground_truth(o.cpu().double(), d.cpu().double(), {k: v.double() for k, v in sdf_p.items()}, S, log2_T, 0.5)
    acc_r = ref["accumulations"].detach()
    assert 0.1 < float((acc_r > 0.5).double().mean()) < 0.95, "test case should mix hits and misses"
    assert float((data["accumulations"].cpu().double() - acc_r).abs().max()) <= 2e-4
    assert float((data["termination_dist"].cpu().double() - ref["termination_dist"].detach()).abs().max()) <= 5e-4
    assert float((data["normals"].cpu().double() - ref["normals"].detach()).abs().max()) <= 2e-3
    decided = (acc_r - 0.5).abs() > 1e-3
    assert torch.equal(data["mask"].cpu().double()[decided], ref["mask"][decided])


@pytest.mark.parametrize("stop_sdf_gradients", [False, True])
def test_fit_pass_losses_and_gradients_vs_oracle_autograd(dev, stop_sdf_gradients):
    """Sampler -> ground truth -> DDF (rays + multi-view + sky rows) -> sdf at termination -> losses -> backward into the DDF,
    its hash table and (stop_sdf_gradients=False, the shipped setting neusky_config.py:45) the SDF field through the
    ground-truth render, the multi-view directions and the termination points."""
    from neusky_b200 import ddf_fit as F
    from oracle import ddf_fit_oracle as FO

    log2_T, S = 14, 16
    sdf_p, ddf_p, step = _setup(log2_T, S, dev)
    sampler = F.VMFDDFSampler(F.DDFSamplerConfig(num_samples_on_sphere=3, num_rays_per_sample=16), device=dev)
    torch.manual_seed(11)
    o, d = sampler()
    N = o.shape[0]
    g = torch.Generator().manual_seed(12)
    sky_o = torch.tensor([0.0, -0.6, 0.1]).expand(10, 3) + 0.1 * torch.randn(10, 3, generator=g)
    sky_d = torch.nn.functional.normalize(torch.randn(10, 3, generator=g) + torch.tensor([0.0, 0.0, 1.0]), dim=-1)
    mv = F.random_points_on_unit_sphere(N)
    mv[:, 2] = mv[:, 2].abs()

    def run_oracle(dt):
        sp = {k: v.to(dt).requires_grad_(True) for k, v in sdf_p.items()}
        dp = {k: v.to(dt).requires_grad_(True) for k, v in ddf_p.items()}
        data = FO.generate_ddf_ground_truth(o.cpu().to(dt), d.cpu().to(dt), sp, S, log2_T, 0.0)
        if stop_sdf_gradients:                                                # neusky_pipeline.py:505-513
            data = {k: v.detach() for k, v in data.items()}
        data["sky_origins"], data["sky_directions"] = sky_o.to(dt), sky_d.to(dt)
        out = FO.ddf_get_outputs(data, dp, sp, log2_T, mv.to(dt), stop_gradients=stop_sdf_gradients)
        L = FO.ddf_loss_dict(out, data)
        sum(L.values()).backward()
        return sp, dp, out, L

    sp, dp, out_r, L_r = run_oracle(torch.float64)
    sp32, dp32, _, _ = run_oracle(torch.float32)

    fit = F.DDFFit(step, sampler=sampler, stop_sdf_gradients=stop_sdf_gradients)
    if stop_sdf_gradients:
        with torch.no_grad():
            data = fit.generate_ddf_ground_truth(o, d, 0.0)
    else:
        data = fit.generate_ddf_ground_truth(o, d, 0.0)
    data["sky_origins"], data["sky_directions"] = sky_o.to(dev), sky_d.to(dev)
    out = fit.get_outputs(data, stop_gradients=stop_sdf_gradients, multi_view_points=mv)
    L = fit.get_loss_dict(out, data)
    sum(L.values()).backward()

    assert set(L) == set(L_r)
    for k in L_r:
        assert abs(float(L[k]) - float(L_r[k])) <= 2e-4 * max(1.0, abs(float(L_r[k]))), f"{k}: {float(L[k])} vs {float(L_r[k])}"
    for k in ("expected_termination_dist", "multi_view_expected_termination_dist", "sky_ray_expected_termination_dist", "sdf_at_termination"):
        assert float((out[k].detach().cpu().double() - out_r[k].detach()).abs().max()) <= 5e-4, k
    m = fit.get_metrics_dict(out, data)
    assert np.isfinite(float(m["depth_psnr"]))

    worst = {}
    for grp, ref, ref32 in (("ddf", dp, dp32), ("sdf", sp, sp32)):
        for k, v in step.group(grp).items():
            if ref[k].grad is None or float(ref[k].grad.abs().max()) == 0.0:
                assert v.grad is None or float(v.grad.abs().max()) <= 1e-12, f"{grp}.{k} should carry no gradient"
                continue
            assert v.grad is not None, f"{grp}.{k}"
            r64, r32 = _rel(v.grad, ref[k].grad), _rel(v.grad, ref32[k].grad)
            cond = _rel(ref32[k].grad, ref[k].grad)
            worst[f"{grp}.{k}"] = (r64, r32, cond)
    if stop_sdf_gradients:
        assert not any(k.startswith("sdf.") for k in worst), "stop_sdf_gradients must cut every path into the SDF field"
    else:
        assert any(k.startswith("sdf.") for k in worst)
    net_cond = {grp: max([c for k, (_, _, c) in worst.items() if k.startswith(grp + ".")] or [0.0]) for grp in ("sdf", "ddf")}
    bad = {k: v for k, v in worst.items() if not (v[0] <= 1e-2 or v[1] <= 1e-2 or v[0] <= 2.0 * net_cond[k.split(".")[0]])}
    assert not bad, f"gradient mismatch {{name: (ours vs fp64, ours vs fp32 oracle, fp32 oracle vs fp64)}} = {bad}"


def test_fit_pass_end_to_end_default_config(dev):
    """The shipped configuration (8 x 128 vMF rays, 256 sky rays) runs as one pass and produces finite losses and gradients
    for every DDF parameter; the three DDF batches go through the network as one 2304-row batch."""
    from neusky_b200 import _lib
    from neusky_b200 import ddf_fit as F

    log2_T, S = 14, 16
    _, _, step = _setup(log2_T, S, dev, split=1)
    fit = F.DDFFit(step)
    g = torch.Generator().manual_seed(5)
    sky_o = (torch.tensor([0.0, -0.6, 0.1]).expand(256, 3) + 0.1 * torch.randn(256, 3, generator=g)).to(dev)
    sky_d = torch.nn.functional.normalize(torch.randn(256, 3, generator=g) + torch.tensor([0.0, 0.0, 1.0]), dim=-1).to(dev)
    torch.manual_seed(3)
    n0 = _lib.launches
    loss, L, out, batch = fit(sky_o, sky_d)
    loss.backward()
    torch.cuda.synchronize()
    assert _lib.launches > n0
    assert batch["origins"].shape == (1024, 3) and out["expected_termination_dist"].shape == (1024,)
    assert out["multi_view_expected_termination_dist"].shape == (1024,) and out["sky_ray_expected_termination_dist"].shape == (256,)
    assert all(torch.isfinite(v) for v in L.values()) and set(L) == {"depth_l1_loss", "sdf_l2_loss", "multi_view_loss", "sky_ray_loss"}
    for k, v in step.group("ddf").items():
        assert v.grad is not None and bool(torch.isfinite(v.grad).all()), k


def test_fit_pass_uses_the_steps_proposal_sampler(dev):
    """With proposal networks on the training step, the ground-truth render of the fitting pass places its samples with the
    model's own sampler (neusky_model.py:1343) and the whole iteration (main pass + fitting pass) still backpropagates."""
    from neusky_b200 import ddf_fit as F
    from neusky_b200 import train as T
    from oracle import neusky_oracle as O

    log2_T, S = 14, 16
    g = torch.Generator().manual_seed(5)
    sdf_p = nb_init.init_sdf_params(3, log2_T=log2_T)
    sdf_p["glin2.bias"] = sdf_p["glin2.bias"].clone()
    sdf_p["glin2.bias"][0] -= 0.35
    ddf_p = nb_init.init_ddf_params(5, final_gain=8.0, log2_T=log2_T, table_scale=0.1)
    prop = [nb_init.init_proposal_params(1, table_scale=1.0, density_bias=1.0), nb_init.init_proposal_params(2, table_scale=1.0, density_bias=2.0)]
    step = T.NeuSkyTrainStep(sdf_p, ddf_p, nb_init.init_reni_params(8), num_cameras=2, device=dev, log2_T=log2_T, num_samples=S, split_geo=1, split=1,
                             threshold_init=0.4, proposal_params=prop, num_proposal_samples_per_ray=(32, 24))
    step.set_directions(O.icosphere_directions(100))
    R = 32
    o = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1) * 0.7
    d = torch.nn.functional.normalize(-o + 0.1 * torch.randn(R, 3, generator=g), dim=-1)
    batch = {"origins": o, "directions": d, "dnorm": torch.ones(R, 1), "cam": torch.randint(0, 2, (R,), generator=g), "image": torch.rand(R, 3, generator=g),
             "fg": torch.ones(R), "ground": torch.zeros(R), "sky": torch.zeros(R)}
    loss, L, _ = step({k: v.to(dev) for k, v in batch.items()})
    fit = F.DDFFit(step, sampler=F.VMFDDFSampler(F.DDFSamplerConfig(num_samples_on_sphere=2, num_rays_per_sample=16), device=dev))
    sky_o = (torch.tensor([0.0, -0.6, 0.1]).expand(8, 3) + 0.05 * torch.randn(8, 3, generator=g)).to(dev)
    sky_d = torch.nn.functional.normalize(torch.randn(8, 3, generator=g) + torch.tensor([0.0, 0.0, 1.0]), dim=-1).to(dev)
    lfit, Lf, _, data = fit(sky_o, sky_d)
    (loss + lfit).backward()
    assert "interlevel_loss" in L and all(bool(torch.isfinite(v)) for v in {**L, **Lf}.values())
    assert data["termination_dist"].shape == (32, 1)
    for grp in ("sdf", "ddf"):
        assert all(v.grad is not None and bool(torch.isfinite(v.grad).all()) for k, v in step.group(grp).items() if k != "embedding_appearance.embedding.weight")
