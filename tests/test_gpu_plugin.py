"""The drop-in nn.Module surface (neusky_b200/fields.py, models.py, plugin.py) driven the way the reference drives its own
classes -- same constructors, same call arguments (expanded [R*S, D, 3] tensors included), state loaded with load_state_dict --
against fixtures produced by the reference's code (tests/golden, tests/golden/make_golden.py) and the CPU oracle."""
import pytest
import torch

from neusky_b200 import init as nb_init
from neusky_b200.rays import Frustums, RayBundle, RaySamples

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from neusky_b200 import _lib

    _lib.load()
    return torch.device("cuda:0")


def _ray_samples(origins, directions, starts, ends, cam=None):
    return RaySamples(frustums=Frustums(origins=origins, directions=directions, starts=starts, ends=ends), deltas=None if ends is None else ends - starts, camera_indices=cam)


def _ddf_model(params, dev):
    from neusky_b200.models import DDFModelConfig

    m = DDFModelConfig().setup(ddf_radius=1.0)
    m.field.load_state_dict(params, strict=True)
    return m.to(dev)


def _reni_field(params, dev):
    from neusky_b200.fields import RENIFieldConfig

    f = RENIFieldConfig().setup(num_train_data=None, num_eval_data=None, normalisations={"min_max": None, "log_domain": True})
    missing, unexpected = f.load_state_dict(params, strict=False)
    assert set(missing) <= {"min_max", "log_domain"} and not unexpected
    return f.to(dev)


def _sdf_field(params, dev, log2_T, impl):
    from neusky_b200.fields import SDFAlbedoFieldConfig

    f = SDFAlbedoFieldConfig(log2_hashmap_size=log2_T, impl=impl).setup(aabb=torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_images=3)
    missing, unexpected = f.load_state_dict(params, strict=False)
    assert set(missing) == {"aabb", "embedding_appearance.embedding.weight"} and not unexpected
    return f.to(dev)


# ------------------------------------------------------------------------------------------------ fields
def test_ddf_model_forward_vs_reference_golden(dev, golden):
    """DDFModel.get_outputs exactly as tests/golden/make_golden.py calls the reference's (ddf_model.py:183-219)."""
    g = golden("ddf_model")
    m = _ddf_model(nb_init.init_ddf_params(int(g["seed"]), final_gain=float(g["final_gain"])), dev).eval()
    q, d = torch.from_numpy(g["positions"]).to(dev), torch.from_numpy(g["directions"]).to(dev)
    with torch.no_grad():
        out = m(RayBundle(origins=q, directions=d, pixel_area=torch.ones(q.shape[0], device=dev)), None, None, True)
    ref = torch.from_numpy(g["expected_termination_dist"])
    assert out["expected_termination_dist"].shape == ref.shape
    assert float((out["expected_termination_dist"].cpu() - ref).abs().max()) <= 3e-4      # 3xTF32 rows, fp32-accurate
    # the (H, W) viewer layout of :188-190, :366-368
    out2 = m(RayBundle(origins=q.reshape(16, 32, 3), directions=d.reshape(16, 32, 3)), None, None, True)
    assert out2["expected_termination_dist"].shape == (16, 32, 1, 1)
    # field-level call with directions already in the local frame (directional_distance_field.py:308-315)
    from neusky_b200.fields import NeuSkyFieldHeadNames

    M = m.get_localised_transforms(q)
    rs = _ray_samples(q, torch.einsum("ijl,ij->il", M, d), None, None)
    t = m.field(rs)[NeuSkyFieldHeadNames.TERMINATION_DISTANCE]
    assert torch.allclose(t, out["expected_termination_dist"], atol=1e-6)


def test_ddf_field_gradients_flow_to_reference_named_parameters(dev):
    m = _ddf_model(nb_init.init_ddf_params(3, final_gain=8.0), dev).train()
    g = torch.Generator().manual_seed(1)
    q = torch.nn.functional.normalize(torch.randn(256, 3, generator=g), dim=-1).to(dev)
    d = torch.nn.functional.normalize(-q + 0.3 * torch.randn(256, 3, generator=g).to(dev), dim=-1)
    out = m(RayBundle(origins=q, directions=d), None, None, True)
    out["expected_termination_dist"].sum().backward()
    got = {n for n, p in m.named_parameters() if p.grad is not None and float(p.grad.abs().sum()) > 0}
    assert {"field.position_encoding.hash_table", "field.ddf.final_layer.weight", "field.ddf.net.0.layer.weight", "field.ddf.mapping_network.network.10.weight"} <= got


def test_reni_field_forward_vs_reference_golden(dev, golden):
    from neusky_b200.fields import RENIFieldHeadNames

    g = golden("reni")
    f = _reni_field(nb_init.init_reni_params(int(g["seed"])), dev).eval()
    dirs, Z, sc, rot = (torch.from_numpy(g[k]).to(dev) for k in ("dirs", "latents", "scale", "rotation"))
    K, D = Z.shape[0], dirs.shape[0]
    # the reference's calling convention (neusky_model.py:470-493): one row per (camera, direction)
    rs = _ray_samples(None, dirs[None].expand(K, D, 3).reshape(-1, 3), None, None, cam=torch.arange(K, device=dev)[:, None].expand(K, D).reshape(-1, 1))
    lat = Z[:, None].expand(K, D, *Z.shape[1:]).reshape(K * D, *Z.shape[1:])
    for key, R in (("radiance", None), ("radiance_rot", rot)):
        out = f(rs, R, lat, sc[:, None].expand(K, D).reshape(-1))
        assert out[RENIFieldHeadNames.MU] is None and out[RENIFieldHeadNames.LOG_VAR] is None
        log_rgb = out[RENIFieldHeadNames.RGB]
        rad = f.unnormalise(log_rgb).reshape(K, D, 3).cpu()
        assert torch.allclose(rad, torch.from_numpy(g[key]), rtol=1e-3, atol=1e-6), key
        # the log-domain value is returned directly (no exp -> log round trip)
        assert torch.allclose(log_rgb.reshape(K, D, 3).cpu(), torch.log(torch.from_numpy(g[key])), rtol=0, atol=1e-3)
    # without camera indices: codes are grouped by value
    rs2 = _ray_samples(None, rs.frustums.directions, None, None, cam=None)
    out2 = f(rs2, None, lat.contiguous(), sc[:, None].expand(K, D).reshape(-1))
    assert torch.allclose(out2[RENIFieldHeadNames.RGB], f(rs, None, lat, sc[:, None].expand(K, D).reshape(-1))[RENIFieldHeadNames.RGB], atol=1e-6)
    with pytest.raises(NotImplementedError):
        f(rs, torch.eye(3, device=dev)[None].expand(4, 3, 3), lat, None)
    # latent gradients through the module call (decoder frozen)
    latg = Z.clone().requires_grad_(True)
    o = f(rs, None, latg[:, None].expand(K, D, *Z.shape[1:]).reshape(K * D, *Z.shape[1:]), None)[RENIFieldHeadNames.RGB]
    o.sum().backward()
    assert latg.grad is not None and float(latg.grad.abs().sum()) > 0


def test_sdf_albedo_field_forward(dev):
    from neusky_b200.fields import FieldHeadNames, NeuSkyFieldHeadNames
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
        f = _sdf_field(p, dev, log2_T, impl).eval()
        rs = _ray_samples(o[:, None].expand(R, S, 3).to(dev), d[:, None].expand(R, S, 3).to(dev), starts.to(dev), ends.to(dev), cam=torch.zeros(R, S, 1, dtype=torch.long, device=dev))
        with torch.no_grad():
            out = f(rs, return_alphas=True)
            sd = f.get_sdf_at_pos(x.reshape(-1, 3).to(dev))
            geo = f.forward_geonetwork(x.reshape(-1, 3).to(dev))
        assert out[FieldHeadNames.SDF].shape == (R, S, 1) and out[NeuSkyFieldHeadNames.ALBEDO].shape == (R, S, 3)
        assert float((out[FieldHeadNames.SDF].cpu().reshape(-1, 1) - ref["sdf"]).abs().max()) <= tol
        assert float((out[FieldHeadNames.ALPHA].cpu() - ref_alpha).abs().max()) <= 10 * tol
        assert float((out[NeuSkyFieldHeadNames.ALBEDO].cpu().reshape(-1, 3) - ref["albedo"]).abs().max()) <= 10 * tol
        assert sd.shape == (R * S, 1) and float((sd.cpu() - ref["sdf"]).abs().max()) <= 1e-4
        assert geo.shape == (R * S, 257) and torch.equal(geo[:, :1], sd)
        rs.camera_indices = None
        with pytest.raises(AttributeError):
            f(rs)
    assert abs(float(f.deviation_network.get_variance()) - inv_s) < 1e-4
    # training mode: the differentiable path, gradients under the reference's parameter names (incl. the weight_norm pair)
    f.train()
    rs.camera_indices = torch.zeros(R, S, 1, dtype=torch.long, device=dev)
    out = f(rs, return_alphas=True)
    assert float((out[FieldHeadNames.SDF].detach().cpu().reshape(-1, 1) - ref["sdf"]).abs().max()) <= 1e-4
    with torch.no_grad():      # the geometric init zeroes the first layer's hash-feature columns (no table gradient at step 0): perturb them
        f.glin0.weight_v[:, 39:] += 0.05 * torch.randn(256, 32, generator=g).to(dev)
    out = f(rs, return_alphas=True)
    (out[FieldHeadNames.ALPHA].sum() + out[NeuSkyFieldHeadNames.ALBEDO].sum()).backward()
    for n in ("glin0.weight_v", "glin0.weight_g", "glin2.bias", "clin2.weight_v", "encoding.hash_table", "deviation_network.variance"):
        gr = f.get_parameter(n).grad
        assert gr is not None and float(gr.abs().sum()) > 0, n


# ------------------------------------------------------------------------------------------------ renderer
def test_lambertian_renderer_reference_call_vs_reference_golden(dev, golden):
    """The EXPANDED argument layout of renderers.py:132-176, exactly as make_golden.py calls the reference; and the compact forms."""
    from neusky_b200.plugin import RGBLambertianRendererWithVisibility

    g = golden("lambert")
    t = {k: torch.from_numpy(g[k]).to(dev) for k in g.files}
    R, S, D = t["albedo"].shape[0], t["albedo"].shape[1], t["dirs"].shape[0]
    ren = RGBLambertianRendererWithVisibility().eval()
    rgb = ren(albedos=t["albedo"], normals=t["normals"], light_directions=t["dirs"][None].expand(R * S, D, 3),
              light_colors=t["light"][:, None].expand(R, S, D, 3).reshape(R * S, D, 3),
              visibility=t["visibility"][:, None].expand(R, S, D).reshape(R * S, D, 1), background_illumination=t["bg"], weights=t["weights"])
    assert torch.allclose(rgb.cpu(), t["rgb"].cpu(), rtol=1e-4, atol=1e-5), (rgb.cpu() - t["rgb"].cpu()).abs().max()
    rgb2 = ren(t["albedo"], t["normals"], t["dirs"], t["light"], t["visibility"], t["bg"], t["weights"], camera_rows=torch.arange(R, dtype=torch.int32, device=dev))
    assert torch.allclose(rgb2, rgb, atol=1e-6)
    # one camera: stride-0 expanded colours are read as ONE table
    one = t["light"][:1].expand(R * S, D, 3)
    rgb3 = ren(t["albedo"], t["normals"], t["dirs"][None].expand(R * S, D, 3), one, None, t["bg"], t["weights"])
    rgb4 = ren(t["albedo"], t["normals"], t["dirs"], t["light"][:1], None, t["bg"], t["weights"])
    assert torch.allclose(rgb3, rgb4, atol=1e-6)


# ------------------------------------------------------------------------------------------------ model
def _build_model(dev, log2_T=15, num_train=3, num_eval=2, seed=0, k4_impl="simt", sdf_impl="simt", S=40, proposal=True):
    from neusky_b200 import models as M
    from neusky_b200.fields import DirectionalDistanceFieldConfig, SDFAlbedoFieldConfig

    sdf_p = nb_init.init_sdf_params(seed, log2_T=log2_T, bias=0.45)
    sdf_p["deviation_network.variance"] = torch.tensor(0.3)
    ddf_p = nb_init.init_ddf_params(seed + 1, final_gain=8.0)
    reni_p = nb_init.init_reni_params(seed + 2)
    ddf = M.DDFModelConfig(ddf_field=DirectionalDistanceFieldConfig()).setup(ddf_radius=1.0)
    ddf.field.load_state_dict(ddf_p, strict=True)
    cfg = M.NeuSkyFactoModelConfig(sdf_field=SDFAlbedoFieldConfig(log2_hashmap_size=log2_T, impl=sdf_impl), num_neus_samples_per_ray=S, k4_impl=k4_impl,
                                   illumination_sampler=M.IcosahedronSamplerConfig(num_directions=100, apply_random_rotation=False),
                                   eval_tile=64)
    m = cfg.setup(scene_box=M.SceneBox(torch.tensor([[-1.0, -1, -1], [1, 1, 1]])), num_train_data=num_train, num_val_data=num_eval, num_test_data=0,
                  visibility_field=ddf, test_mode="val")
    m.field.load_state_dict(sdf_p, strict=False)
    m.illumination_field.load_state_dict({**reni_p, "log_domain": torch.tensor(True)}, strict=False)      # what the shipped decoder checkpoint carries
    prop = [nb_init.init_proposal_params(seed + 3, table_scale=1.0, density_bias=1.0), nb_init.init_proposal_params(seed + 4, table_scale=1.0, density_bias=2.0)]
    for net, pp in zip(m.proposal_networks, prop):
        net.load_state_dict(pp, strict=True)
    g = torch.Generator().manual_seed(seed + 9)
    with torch.no_grad():
        m.eval_illumination_latents.copy_(torch.randn(num_eval, 100, 3, generator=g))
        m.train_illumination_latents.copy_(torch.randn(num_train, 100, 3, generator=g))
        m.eval_scale.copy_(0.2 * torch.randn(num_eval, generator=g))
        m.visibility_threshold.fill_(0.1)
    return m.to(dev), dict(sdf_p=sdf_p, ddf_p=ddf_p, reni_p=reni_p, prop=prop, log2_T=log2_T, S=S)


def test_model_eval_forward_mixed_cameras_vs_oracle(dev):
    """NeuSkyFactoModel.forward on a bundle whose rays belong to TWO cameras (neusky_model.py:461): each half equals the oracle's
    single-camera render with that camera's latent code and scale; output keys follow :881-931."""
    from oracle import neusky_oracle as O

    m, p = _build_model(dev)
    m.eval()
    H = W = 12
    o, d, dn = O.pinhole_rays(H, W, float(W), float(W), W / 2, H / 2, O.look_at_camera((0.0, -0.9, 0.25)))
    n = o.shape[0]
    cam = torch.cat([torch.zeros(n // 2, 1, dtype=torch.long), torch.ones(n - n // 2, 1, dtype=torch.long)])
    rb = RayBundle(origins=o.to(dev), directions=d.to(dev), camera_indices=cam.to(dev), metadata={"directions_norm": dn.to(dev)})
    with torch.no_grad():
        out = m(rb)
    for k in ("rgb", "albedo", "accumulation", "depth", "p2p_dist", "normal", "weights", "hdr_background_colours", "directions_norm", "sdf_at_termination",
              "normal_vis", "prop_depth_0", "prop_depth_1", "visibility_batch"):
        assert k in out, k
    assert out["sdf_at_termination"] is None and out["weights"].shape == (n, p["S"], 1) and out["prop_depth_0"].shape == (n, 1)
    vb = out["visibility_batch"]
    assert vb["mask"].shape == vb["termination_dist"].shape and bool((vb["mask"] == 1).all()) and vb["sdf_at_termination"] is None
    assert torch.allclose(out["normal_vis"], (out["normal"] + 1) / 2)
    dirs = O.icosphere_directions(100)
    inv_s = float(torch.exp(torch.tensor(3.0)))
    Z, sc = m.eval_illumination_latents.detach().cpu(), m.eval_scale.detach().cpu()
    with torch.no_grad():
        # one oracle render of the WHOLE bundle per camera (same placement and depth clip range as the model's single pass), then pick that camera's rays
        refs = [O.render_rays(o, d, dn, p["S"], p["sdf_p"], p["ddf_p"], p["reni_p"], Z[c], sc[c], dirs, inv_s, log2_T=p["log2_T"], proposal_nets=p["prop"],
                              chunk=n) for c in range(2)]
    for c, sl in ((0, slice(0, n // 2)), (1, slice(n // 2, n))):
        for k in ("rgb", "albedo", "normal", "accumulation"):
            e = float((out[k][sl].cpu() - refs[c][k][sl]).abs().max())
            assert e <= 1e-3, (c, k, e)
        e = float((out["depth"][sl].cpu() - refs[c]["depth"][sl]).abs().max())
        assert e <= 1e-3 * float(refs[c]["depth"].abs().max()), (c, e)
    assert float((refs[0]["rgb"] - refs[1]["rgb"]).abs().max()) > 1e-2      # the two cameras' illuminations differ: the test can see a mix-up


def test_model_camera_bundle_chunk_clip_matches_reference_loop(dev):
    """get_outputs_for_camera_ray_bundle renders in chunks of config.eval_tile rays, each a forward() of its own, so -- like the
    reference's 256-ray loop (neusky_model.py:1413-1437) -- the depth clip range is per chunk: equals the oracle's clip_per_chunk mode."""
    from oracle import neusky_oracle as O

    m, p = _build_model(dev, proposal=True)
    H, W = 8, 16
    o, d, dn = O.pinhole_rays(H, W, float(W), float(W), W / 2, H / 2, O.look_at_camera((0.0, -0.9, 0.25)))
    rb = RayBundle(origins=o.reshape(H, W, 3).to(dev), directions=d.reshape(H, W, 3).to(dev), camera_indices=torch.ones(H, W, 1, dtype=torch.long, device=dev),
                   metadata={"directions_norm": dn.reshape(H, W, 1).to(dev)})
    m.train()
    out = m.get_outputs_for_camera_ray_bundle(rb)
    assert m.training and out["rgb"].shape == (H, W, 3) and out["depth"].shape == (H, W, 1) and "visibility_batch" not in out
    with torch.no_grad():
        ref = O.render_rays(o, d, dn, p["S"], p["sdf_p"], p["ddf_p"], p["reni_p"], m.eval_illumination_latents[1].detach().cpu(), m.eval_scale[1].detach().cpu(),
                            O.icosphere_directions(100), float(torch.exp(torch.tensor(3.0))), log2_T=p["log2_T"], proposal_nets=p["prop"], chunk=64, clip_per_chunk=True)
    assert float((out["rgb"].reshape(-1, 3).cpu() - ref["rgb"]).abs().max()) <= 1e-3
    assert float((out["p2p_dist"].reshape(-1, 1).cpu() - ref["p2p_dist"]).abs().max()) <= 1e-4 * float(ref["p2p_dist"].abs().max()) + 1e-5


def test_model_compute_visibility_reference_call_vs_golden(dev, golden):
    """NeuSkyFactoModel.compute_visibility with the reference's arguments (the call make_golden.py makes on the reference)."""
    from neusky_b200 import models as M

    g = golden("visibility")
    ddf = M.DDFModelConfig().setup(ddf_radius=1.0)
    ddf.field.load_state_dict(nb_init.init_ddf_params(int(g["seed"]), final_gain=float(g["final_gain"])), strict=True)
    o, d, p2p, dirs = (torch.from_numpy(g[k]).to(dev) for k in ("origins", "ray_dirs", "p2p", "dirs"))
    R, S, D = o.shape[0], 3, dirs.shape[0]
    rs = _ray_samples(o[:, None].expand(R, S, 3).contiguous(), d[:, None].expand(R, S, 3).contiguous(), torch.zeros(R, S, 1, device=dev), torch.ones(R, S, 1, device=dev))
    illum = dirs[None].expand(R * S, D, 3)
    for impl, tol in (("simt", 5e-4), ("tc2", 1e-2)):
        cfg = M.NeuSkyFactoModelConfig(k4_impl=impl)
        cfg.sdf_field.log2_hashmap_size = 12
        m = cfg.setup(scene_box=M.SceneBox(torch.tensor([[-1.0, -1, -1], [1, 1, 1]])), num_train_data=1, num_val_data=1, num_test_data=0, visibility_field=ddf,
                      test_mode="val").to(dev).eval()
        vd = m.compute_visibility(rs, p2p, illum, float(g["threshold"]), float(g["sigmoid_scale"]), compute_shadow_map=True)
        assert vd["visibility"].shape == (R * S, D, 1)
        vis = vd["visibility"].reshape(R, S, D)
        assert torch.equal(vis[:, 0], vis[:, 2])
        assert float((vis[:, 0].cpu() - torch.from_numpy(g["visibility"])).abs().max()) <= tol, impl
        vb = vd["visibility_batch"]
        assert torch.allclose(vb["termination_dist"].cpu(), torch.from_numpy(g["termination_dist"]), rtol=1e-5, atol=2e-6)
        assert vb["mask"].shape == vb["termination_dist"].shape and bool((vb["mask"] == 1).all())        # torch.ones_like(termination_dist), :1768-1771
        assert vd["difference"].shape == vd["expected_termination_dist"].shape
        ref_diff = torch.clamp(torch.from_numpy(g["termination_dist"]), max=2.0) - torch.from_numpy(g["expected_termination_dist"])
        assert float((vd["difference"].cpu() - ref_diff).abs().max()) <= (5e-4 if impl == "simt" else 4e-3)


def test_model_training_forward_backward_on_reference_named_parameters(dev):
    """Training-mode forward (ray bundle + targets) -> the reference's loss names -> backward: gradients land on the module
    parameters under the reference's names, in the reference's optimizer groups; the state_dict keeps the reference's layout."""
    from oracle import neusky_oracle as O

    m, p = _build_model(dev, S=16)
    m.train()
    R = 96
    g = torch.Generator().manual_seed(4)
    o = torch.tensor([0.0, -0.9, 0.25]).expand(R, 3) + 0.02 * torch.randn(R, 3, generator=g)
    d = torch.nn.functional.normalize(-o + 0.4 * torch.randn(R, 3, generator=g), dim=-1)
    rb = RayBundle(origins=o.to(dev), directions=d.to(dev), camera_indices=torch.randint(0, 3, (R, 1), generator=g).to(dev), metadata={"directions_norm": torch.ones(R, 1, device=dev)})
    batch = {"image": torch.rand(R, 3, generator=g).to(dev), "fg": (torch.rand(R, generator=g) > 0.3).float().to(dev),
             "ground": (torch.rand(R, generator=g) > 0.7).float().to(dev), "sky": (torch.rand(R, generator=g) > 0.8).float().to(dev)}
    keys_before = set(m.state_dict().keys())
    out = m(rb, batch=batch)
    losses = m.get_loss_dict(out, batch)
    assert {"rgb_l1_loss", "eikonal_loss", "fg_mask_loss", "sdf_level_set_visibility_loss", "sky_pixel_loss", "hashgrid_density_loss", "ground_plane_loss",
            "visibility_sigmoid_loss", "interlevel_loss"} <= set(losses)
    for k in ("rgb", "eik_grad", "weights", "accumulation", "depth", "p2p_dist", "normal", "normal_vis", "sdf_at_termination", "hdr_background_colours", "directions_norm"):
        assert k in out, k
    sum(losses.values()).backward()
    groups = m.get_param_groups()
    groups.update(m.visibility_field.get_param_groups())
    for name in ("fields", "proposal_networks", "illumination_field", "visibility_sigmoid", "ddf_field"):
        assert any(q.grad is not None and float(q.grad.abs().sum()) > 0 for q in groups[name]), name
    assert m.field.glin1.weight_v.grad is not None and m.visibility_field.field.ddf.net[2].layer.weight.grad is not None
    assert all(q.grad is None for q in m.illumination_field.parameters())      # frozen decoder
    assert set(m.state_dict().keys()) == keys_before
    # no_grad forward in training mode leaves every .grad untouched (the interlevel loss is an autograd node, not a side effect)
    snap = {n: q.grad.clone() for n, q in m.named_parameters() if q.grad is not None}
    rb_eval = RayBundle(origins=rb.origins, directions=rb.directions, camera_indices=rb.camera_indices % 2, metadata=rb.metadata)   # 2 eval images
    with torch.no_grad():
        m.eval()
        m(rb_eval)
        m.train()
    for n, q in m.named_parameters():
        if n in snap:
            assert torch.equal(q.grad, snap[n]), n


def test_model_graphed_iteration_trains_the_models_own_parameters(dev):
    """NeuSkyFactoModel.graphed_iteration: the whole iteration replayed as a CUDA graph (neusky_b200/graphed.py) over the model's own
    nn.Parameter objects -- after one eager and two replayed iterations the module's parameters have moved, the loss is finite and the
    state_dict still has the reference's layout."""
    from neusky_b200.parallel import GradBucketReducer

    m, p = _build_model(dev, S=16)
    m.train()
    ts = m.train_step()
    params = [q for q in ts.parameters() if q.requires_grad]
    red = GradBucketReducer(params, big_bytes=64 << 10)
    it = m.graphed_iteration(red, torch.optim.SGD(params, lr=1e-4), eager_warmup=1)
    R = 96
    g = torch.Generator().manual_seed(4)
    o = torch.tensor([0.0, -0.9, 0.25]).expand(R, 3) + 0.02 * torch.randn(R, 3, generator=g)
    d = torch.nn.functional.normalize(-o + 0.4 * torch.randn(R, 3, generator=g), dim=-1)
    batch = {"origins": o.contiguous(), "directions": d.contiguous(), "dnorm": torch.ones(R, 1), "cam": torch.randint(0, 3, (R,), generator=g).to(torch.int32),
             "image": torch.rand(R, 3, generator=g), "fg": (torch.rand(R, generator=g) > 0.3).float(), "ground": (torch.rand(R, generator=g) > 0.7).float(),
             "sky": (torch.rand(R, generator=g) > 0.8).float()}
    keys_before = set(m.state_dict().keys())
    w0 = m.field.glin1.weight_v.detach().clone()
    t0 = m.visibility_field.field.ddf.net[2].layer.weight.detach().clone()
    losses = []
    for _ in range(3):
        gp, gd = torch.rand(27, 3, generator=g) * 2 - 1, torch.nn.functional.normalize(torch.randn(27, 3, generator=g), dim=-1)
        losses.append(float(it(batch, m._illumination_directions().cpu(), gp, gd)))
    assert it.captures == 1 and it.eager_steps == 1 and it.replays == 2
    assert all(l == l and abs(l) < 1e6 for l in losses), losses
    assert not torch.equal(w0, m.field.glin1.weight_v) and not torch.equal(t0, m.visibility_field.field.ddf.net[2].layer.weight)
    assert set(m.state_dict().keys()) == keys_before
