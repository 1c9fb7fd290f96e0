"""DDF fitting pass, CPU side: the host ray samplers must reproduce the reference's samplers bit for bit under the same
torch seed, and the oracle restatement of DDFModel.get_outputs (training) / get_loss_dict must match the reference's own
methods (fixtures from tests/golden/make_golden.py::golden_ddf_fit).  CPU only."""
import hashlib

import numpy as np
import torch

from neusky_b200 import init as nb_init
from neusky_b200 import ddf_fit as F
from oracle import ddf_fit_oracle as FO


def _sha(params):
    h = hashlib.sha256()
    for k in sorted(params):
        h.update(k.encode())
        h.update(params[k].detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_vmf_sampler_bit_exact_vs_reference(golden):
    g = golden("ddf_fit")
    s = F.VMFDDFSampler(F.DDFSamplerConfig(num_samples_on_sphere=8, num_rays_per_sample=128, only_sample_upper_hemisphere=True, concentration=20.0))
    torch.manual_seed(int(g["vmf_seed"]))
    o, d = s()
    assert o.shape == (1024, 3) and d.shape == (1024, 3) and s.num_rays == 1024
    assert np.array_equal(_bits(o.numpy()), _bits(g["vmf_origins"]))
    assert np.array_equal(_bits(d.numpy()), _bits(g["vmf_directions"]))
    # sampler invariants (ddf_sampler.py:255-266): origins on the upper unit sphere, directions unit and inward facing
    assert float(o[:, 2].min()) >= 0 and torch.allclose(o.norm(dim=-1), torch.ones(1024), atol=1e-6)
    assert torch.allclose(d.norm(dim=-1), torch.ones(1024), atol=1e-5) and float((d * -o).sum(-1).min()) >= 0


def test_uniform_sampler_bit_exact_vs_reference(golden):
    g = golden("ddf_fit")
    s = F.UniformDDFSampler(F.DDFSamplerConfig(num_samples_on_sphere=4, num_rays_per_sample=16, only_sample_upper_hemisphere=True))
    torch.manual_seed(int(g["uniform_seed"]))
    o, d = s()
    assert np.array_equal(_bits(o.numpy()), _bits(g["uniform_origins"]))
    assert np.array_equal(_bits(d.numpy()), _bits(g["uniform_directions"]))


def test_sampler_explicit_positions_and_counts():
    s = F.VMFDDFSampler(F.DDFSamplerConfig(num_samples_on_sphere=2, num_rays_per_sample=5, only_sample_upper_hemisphere=False, concentration=5.0))
    torch.manual_seed(1)
    pos = torch.tensor([[0.0, 0.0, -1.0], [1.0, 0.0, 0.0], [0.0, 0.6, 0.8]])
    o, d = s(num_directions=7, positions=pos.clone())
    assert o.shape == (21, 3) and torch.equal(o.reshape(3, 7, 3)[:, 0], pos)       # lower-hemisphere point kept: no flip requested
    assert float((d.reshape(3, 7, 3) * -pos[:, None]).sum(-1).min()) >= 0


def _golden_batch(g):
    t = {k: torch.from_numpy(g[k]) for k in g.files if g[k].dtype == np.float32}
    batch = {"origins": t["origins"], "directions": t["directions"], "termination_dist": t["termination_dist"], "mask": t["mask"],
             "sky_origins": t["sky_origins"], "sky_directions": t["sky_directions"]}
    torch.manual_seed(int(g["multi_view_seed"]))
    mv = F.random_points_on_unit_sphere(batch["origins"].shape[0])
    mv[:, 2] = torch.abs(mv[:, 2])
    return t, batch, mv


def test_oracle_ddf_outputs_and_losses_vs_reference(golden):
    g = golden("ddf_fit")
    ddf_p = nb_init.init_ddf_params(int(g["seed"]), final_gain=float(g["final_gain"]))
    assert _sha(ddf_p) == str(g["weights_sha256"]), "seeded RNG stream drifted: regenerate goldens"
    sdf_p = nb_init.init_sdf_params(int(g["sdf_seed"]), log2_T=int(g["sdf_log2_T"]))
    t, batch, mv = _golden_batch(g)
    with torch.no_grad():
        out = FO.ddf_get_outputs(batch, ddf_p, sdf_p, int(g["sdf_log2_T"]), mv, log2_T_ddf=19)
        L = FO.ddf_loss_dict(out, batch)
    for k in ("expected_termination_dist", "distance_weight", "sdf_at_termination", "multi_view_expected_termination_dist",
              "sky_ray_termination_dist", "sky_ray_expected_termination_dist"):
        assert torch.allclose(out[k], t["out_" + k], rtol=1e-4, atol=2e-5), (k, float((out[k] - t["out_" + k]).abs().max()))
    assert set(L) == {"depth_l1_loss", "sdf_l2_loss", "multi_view_loss", "sky_ray_loss"}
    for k, v in L.items():
        assert abs(float(v) - float(g["loss_" + k])) <= 1e-5 * max(1.0, abs(float(g["loss_" + k]))), (k, float(v), float(g["loss_" + k]))


def test_host_loss_dict_matches_reference_on_golden_outputs(golden):
    """neusky_b200.ddf_fit.DDFFit.get_loss_dict is plain torch on [N]-sized tensors: run it on the reference's outputs on CPU."""
    g = golden("ddf_fit")
    t, batch, _ = _golden_batch(g)
    out = {k[4:]: v for k, v in t.items() if k.startswith("out_")}
    fit = F.DDFFit.__new__(F.DDFFit)
    fit.config, fit.ddf_radius = F.DDFModelConfig(), 1.0
    L = fit.get_loss_dict(out, batch)
    for k, v in L.items():
        assert abs(float(v) - float(g["loss_" + k])) <= 1e-6 * max(1.0, abs(float(g["loss_" + k]))), k
    m = fit.get_metrics_dict(out, batch)
    assert np.isfinite(float(m["depth_psnr"]))


def test_draw_host_consumes_the_cpu_generator_like_the_in_line_pass():
    """DDFFit.draw_host (the host half of a fitting pass, hoisted out of the captured iteration: neusky_b200/graphed.py) must draw
    exactly what the in-line pass draws, in the same order on torch's CPU generator: the sampler's rays (ddf_sampler.py), then the
    multi-view sphere points with |z| (ddf_model.py:279-284) -- so a graphed training run sees the reference's random sequence."""
    fit = F.DDFFit.__new__(F.DDFFit)                      # host logic only: no NeuSkyTrainStep (that needs the CUDA library's device)
    fit.config, fit.training = F.DDFModelConfig(), True
    fit.sampler = F.VMFDDFSampler(F.DDFSamplerConfig(), ddf_sphere_radius=1.0, device="cpu")
    torch.manual_seed(11)
    o, d, mv = fit.draw_host()
    torch.manual_seed(11)
    o2, d2 = fit.sampler()
    mv2 = F.random_points_on_unit_sphere(o2.shape[0])
    mv2[:, 2] = torch.abs(mv2[:, 2])
    assert torch.equal(o, o2) and torch.equal(d, d2) and torch.equal(mv, mv2)
    assert o.device.type == "cpu" and o.shape == (1024, 3) and bool((mv[:, 2] >= 0).all())
    assert str(fit.sampler.device) == "cpu"
    # without the multi-view loss (or in eval mode) no extra draw is consumed
    fit.training = False
    torch.manual_seed(11)
    _, _, mv3 = fit.draw_host()
    nxt = torch.rand(1)
    torch.manual_seed(11)
    fit.sampler()
    assert mv3 is None and torch.equal(nxt, torch.rand(1))
