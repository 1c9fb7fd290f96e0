"""Generate the golden fixtures in this directory by running the REFERENCE'S OWN CODE.

Run in the build container only (``python tests/golden/make_golden.py``): it imports
``neusky`` / ``reni`` from /root/reference through ``oracle/ref_shim`` (stand-ins for the
absent nerfstudio / tinycudann / nerfacc packages).  /root/reference does not exist on the
GPU box, so the resulting ``*.npz`` files are committed and tests only read those.

Weights are NOT stored: every fixture records the seed, and ``neusky_b200.init`` regenerates
the identical tensors (a checksum of the weights is stored to catch RNG drift).
"""
from __future__ import annotations

import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

ref_shim.install()

from neusky_b200 import init as nb_init  # noqa: E402


def checksum(params) -> str:
    h = hashlib.sha256()
    for k in sorted(params):
        h.update(k.encode())
        h.update(params[k].detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: " + ", ".join(f"{k}{tuple(np.shape(v))}" for k, v in out.items()))


# ------------------------------------------------------------------------------ icosphere
def golden_icosphere():
    from reni.model_components.illumination_samplers import IcosahedronSampler, IcosahedronSamplerConfig

    out = {}
    for n in (100, 256, 512):
        s = IcosahedronSampler(IcosahedronSamplerConfig(num_directions=n))
        out[f"dirs_{n}"] = s.directions
    save("icosphere", **out)


# ------------------------------------------------------------------------------ DDF + visibility
DDF_SEED = 1234
DDF_FINAL_GAIN = 8.0


def build_reference_ddf():
    from neusky.fields.directional_distance_field import DirectionalDistanceField, DirectionalDistanceFieldConfig

    cfg = DirectionalDistanceFieldConfig(
        ddf_type="ddf", position_encoding_type="hash", direction_encoding_type="nerf", conditioning="FiLM",
        termination_output_activation="sigmoid", probability_of_hit_output_activation="sigmoid",
        hidden_layers=5, hidden_features=256, mapping_layers=5, mapping_features=256,
        num_attention_heads=8, num_attention_layers=6, predict_probability_of_hit=False,
    )  # neusky/configs/neusky_config.py:162-177
    field = DirectionalDistanceField(cfg, ddf_radius=1.0)
    params = nb_init.init_ddf_params(DDF_SEED, final_gain=DDF_FINAL_GAIN)
    missing, unexpected = field.load_state_dict(params, strict=True), None
    field.eval()
    return field, params


def golden_ddf_and_visibility():
    from nerfstudio.cameras.rays import RayBundle
    from neusky.models.ddf_model import DDFModel
    from neusky.models.neusky_model import NeuSkyFactoModel

    field, params = build_reference_ddf()
    g = torch.Generator().manual_seed(77)

    # --- DDFModel.get_outputs on sphere points / inward directions (ddf_model.py:183-219)
    ddf_self = types.SimpleNamespace(
        field=field, training=False,
        config=types.SimpleNamespace(
            compute_normals=False, include_depth_loss_scene_center_weight=True,
            loss_inclusions={"sdf_l1_loss": False, "sdf_l2_loss": True, "multi_view_loss": True, "sky_ray_loss": True},
        ),
    )
    ddf_self.get_localised_transforms = lambda pos: DDFModel.get_localised_transforms(ddf_self, pos)
    N = 512
    q = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1)
    q[:, 2] = q[:, 2].abs()
    tgt = torch.randn(N, 3, generator=g) * 0.3
    d = torch.nn.functional.normalize(tgt - q, dim=-1)
    with torch.no_grad():
        out = DDFModel.get_outputs(ddf_self, RayBundle(origins=q, directions=d, pixel_area=torch.ones(N)), None, None, True)
    save("ddf_model", seed=DDF_SEED, final_gain=DDF_FINAL_GAIN, weights_sha256=checksum(params),
         positions=q, directions=d, expected_termination_dist=out["expected_termination_dist"])

    # --- NeuSkyFactoModel.compute_visibility (neusky_model.py:1624-1778) incl. outside-sphere rays
    from reni.model_components.illumination_samplers import IcosahedronSampler, IcosahedronSamplerConfig

    dirs = IcosahedronSampler(IcosahedronSamplerConfig(num_directions=100)).directions  # D=162
    R, S, D = 48, 3, dirs.shape[0]
    origins = torch.tensor([0.0, -0.9, 0.25]).expand(R, 3).clone() + 0.05 * torch.randn(R, 3, generator=g)
    rdirs = torch.nn.functional.normalize(-origins + 0.6 * torch.randn(R, 3, generator=g), dim=-1)
    p2p = 0.2 + 1.2 * torch.rand(R, 1, generator=g)
    p2p[:6] = 2.5  # these land outside the DDF sphere -> the "hack" branch (:1674-1683)
    model_self = types.SimpleNamespace(
        ddf_radius=1.0,
        config=types.SimpleNamespace(only_upperhemisphere_visibility=True, lower_hermisphere_visibility=True,
                                     sdf_to_visibility_stop_gradients="depth"),
        visibility_field=lambda rb, batch, neusky, stop_gradients: DDFModel.get_outputs(ddf_self, rb, batch, None, stop_gradients),
    )
    model_self.ray_sphere_intersection = lambda p, d_, r: NeuSkyFactoModel.ray_sphere_intersection(model_self, p, d_, r)
    from nerfstudio.cameras.rays import Frustums, RaySamples

    rs = RaySamples(frustums=Frustums(origins=origins[:, None].expand(R, S, 3).contiguous(),
                                      directions=rdirs[:, None].expand(R, S, 3).contiguous(),
                                      starts=torch.zeros(R, S, 1), ends=torch.ones(R, S, 1), pixel_area=torch.ones(R, S, 1)))
    illum = dirs[None].expand(R * S, D, 3)
    thr, scale = 0.1, 25.0
    with torch.no_grad():
        vd = NeuSkyFactoModel.compute_visibility(model_self, rs, p2p.clone(), illum, thr, scale)
    vis = vd["visibility"].reshape(R, S, D)
    assert torch.equal(vis[:, 0], vis[:, 1])
    save("visibility", seed=DDF_SEED, final_gain=DDF_FINAL_GAIN, weights_sha256=checksum(params),
         origins=origins, ray_dirs=rdirs, p2p=p2p, dirs=dirs, threshold=thr, sigmoid_scale=scale,
         visibility=vis[:, 0], expected_termination_dist=vd["expected_termination_dist"],
         termination_dist=vd["visibility_batch"]["termination_dist"])


# ------------------------------------------------------------------------------ RENI++
RENI_SEED = 4321


def golden_reni():
    from reni.illumination_fields.reni_illumination_field import RENIField, RENIFieldConfig
    from reni.field_components.field_heads import RENIFieldHeadNames
    from nerfstudio.cameras.rays import Frustums, RaySamples

    cfg = RENIFieldConfig(
        conditioning="Attention", invariant_function="VN", equivariance="SO2", axis_of_invariance="z",
        positional_encoding="NeRF", encoded_input="Directions", latent_dim=100, hidden_features=128, hidden_layers=9,
        mapping_layers=5, mapping_features=128, num_attention_heads=8, num_attention_layers=6,
        output_activation="None", last_layer_linear=True, fixed_decoder=True, trainable_scale=True,
    )  # neusky/configs/neusky_config.py:78-96
    field = RENIField(cfg, num_train_data=None, num_eval_data=None, normalisations={"min_max": None, "log_domain": True})
    params = nb_init.init_reni_params(RENI_SEED)
    sd = field.state_dict()
    for k, v in params.items():
        assert k in sd and sd[k].shape == v.shape, (k, v.shape, sd.get(k, torch.zeros(0)).shape)
    field.load_state_dict({**sd, **params}, strict=True)
    field.eval()
    g = torch.Generator().manual_seed(99)
    K, D = 3, 40
    dirs = torch.nn.functional.normalize(torch.randn(D, 3, generator=g), dim=-1)
    Z = torch.randn(K, 100, 3, generator=g)
    scale = 0.3 * torch.randn(K, generator=g)
    rot = torch.tensor([[0.8, -0.6, 0.0], [0.6, 0.8, 0.0], [0.0, 0.0, 1.0]])
    outs = {}
    for tag, R in (("", None), ("_rot", rot)):
        rows = []
        for k in range(K):
            rs = RaySamples(frustums=Frustums(origins=torch.zeros(D, 3), directions=dirs, starts=torch.zeros(D), ends=torch.ones(D), pixel_area=torch.ones(D)),
                            camera_indices=torch.full((D,), k))
            with torch.no_grad():
                o = field.forward(rs, rotation=R, latent_codes=Z[k : k + 1].expand(D, -1, -1), scale=scale[k : k + 1].expand(D))
            rows.append(field.unnormalise(o[RENIFieldHeadNames.RGB]))
        outs["radiance" + tag] = torch.stack(rows, 0)
    save("reni", seed=RENI_SEED, weights_sha256=checksum(params), dirs=dirs, latents=Z, scale=scale, rotation=rot, **outs)


# ------------------------------------------------------------------------------ Lambertian renderer
def golden_lambert():
    from neusky.model_components.renderers import RGBLambertianRendererWithVisibility

    g = torch.Generator().manual_seed(5)
    R, S, D = 16, 4, 37
    albedo = torch.rand(R, S, 3, generator=g)
    normals = torch.nn.functional.normalize(torch.randn(R, S, 3, generator=g), dim=-1)
    dirs = torch.nn.functional.normalize(torch.randn(D, 3, generator=g), dim=-1)
    light = torch.exp(torch.randn(R, D, 3, generator=g))
    vis = torch.rand(R, D, generator=g)
    bg = torch.rand(R, 3, generator=g)
    w = torch.rand(R, S, 1, generator=g) / S
    ren = RGBLambertianRendererWithVisibility()
    ren.eval()
    rgb = ren(
        albedos=albedo, normals=normals,
        light_directions=dirs[None].expand(R * S, D, 3),
        light_colors=light[:, None].expand(R, S, D, 3).reshape(R * S, D, 3),
        visibility=vis[:, None].expand(R, S, D).reshape(R * S, D, 1),
        background_illumination=bg, weights=w,
    )
    save("lambert", albedo=albedo, normals=normals, dirs=dirs, light=light, visibility=vis, bg=bg, weights=w, rgb=rgb)


# ------------------------------------------------------------------------------ DDF fitting pass (samplers, DDFModel training outputs + losses)
DDF_FIT_SDF_SEED = 3
DDF_FIT_LOG2_T_SDF = 14


def golden_ddf_fit():
    from neusky.model_components.ddf_sampler import UniformDDFSampler, UniformDDFSamplerConfig, VMFDDFSampler, VMFDDFSamplerConfig
    from neusky.models.ddf_model import DDFModel
    from nerfstudio.cameras.rays import RayBundle
    from oracle import neusky_oracle as O

    # --- samplers under a fixed torch CPU seed (ddf_sampler.py:119-286; NeuSky config neusky_config.py:207-212)
    vmf = VMFDDFSampler(VMFDDFSamplerConfig(num_samples_on_sphere=8, num_rays_per_sample=128, only_sample_upper_hemisphere=True, concentration=20.0))
    torch.manual_seed(2024)
    rb_v = vmf()
    uni = UniformDDFSampler(UniformDDFSamplerConfig(num_samples_on_sphere=4, num_rays_per_sample=16, only_sample_upper_hemisphere=True))
    torch.manual_seed(2025)
    rb_u = uni()

    # --- DDFModel.get_outputs (training) + get_loss_dict with the NeuSky loss configuration (neusky_config.py:178-205)
    field, params = build_reference_ddf()
    inclusions = {"depth_l1_loss": True, "depth_l2_loss": False, "sdf_l1_loss": False, "sdf_l2_loss": True, "prob_hit_loss": False,
                  "normal_loss": False, "multi_view_loss": True, "sky_ray_loss": True}
    coefficients = {"depth_l1_loss": 1.0, "depth_l2_loss": 0.0, "sdf_l1_loss": 1.0, "sdf_l2_loss": 0.01, "prob_hit_loss": 0.01,
                    "normal_loss": 1.0, "multi_view_loss": 0.01, "sky_ray_loss": 1.0}
    ddf_self = types.SimpleNamespace(
        field=field, training=True, ddf_radius=1.0,
        config=types.SimpleNamespace(compute_normals=False, include_depth_loss_scene_center_weight=True, scene_center_weight_exp=3.0,
                                     scene_center_weight_include_z=False, mask_to_circumference=False, inverse_depth_weight=False,
                                     loss_inclusions=inclusions, loss_coefficients=coefficients),
        depth_l1_loss=torch.nn.L1Loss(reduction="none"), sdf_l2_loss=torch.nn.MSELoss(), sky_ray_loss=torch.nn.L1Loss(),
    )
    ddf_self.get_localised_transforms = lambda pos: DDFModel.get_localised_transforms(ddf_self, pos)
    sdf_p = nb_init.init_sdf_params(DDF_FIT_SDF_SEED, log2_T=DDF_FIT_LOG2_T_SDF)
    sca = O.hash_scalings()
    neusky = types.SimpleNamespace(field=types.SimpleNamespace(
        get_sdf_at_pos=lambda x: O.sdf_geo_network(x, sdf_p, sca, DDF_FIT_LOG2_T_SDF)[:, :1]))   # nerfstudio SDFField: restated, unpinned
    g = torch.Generator().manual_seed(31)
    N = 256
    origins, directions = rb_v.origins[:N].clone(), rb_v.directions[:N].clone()
    term = 0.3 + 1.5 * torch.rand(N, 1, generator=g)
    mask = (torch.rand(N, 1, generator=g) > 0.25).float()
    n_sky = 48
    sky_o = torch.tensor([0.0, -0.6, 0.1]).expand(n_sky, 3) + 0.1 * torch.randn(n_sky, 3, generator=g)
    sky_d = torch.nn.functional.normalize(torch.randn(n_sky, 3, generator=g) + torch.tensor([0.0, 0.0, 1.0]), dim=-1)
    batch = {"termination_dist": term, "mask": mask, "sky_ray_bundle": RayBundle(origins=sky_o, directions=sky_d, pixel_area=torch.ones(n_sky, 1))}
    MV_SEED = 77
    torch.manual_seed(MV_SEED)          # the multi-view points are the first draw inside get_outputs (ddf_model.py:289)
    with torch.no_grad():
        out = DDFModel.get_outputs(ddf_self, RayBundle(origins=origins, directions=directions, pixel_area=torch.ones(N, 1)), batch, neusky, False)
        losses = DDFModel.get_loss_dict(ddf_self, out, batch)
    save("ddf_fit", seed=DDF_SEED, final_gain=DDF_FINAL_GAIN, weights_sha256=checksum(params), sdf_seed=DDF_FIT_SDF_SEED, sdf_log2_T=DDF_FIT_LOG2_T_SDF,
         vmf_seed=2024, vmf_origins=rb_v.origins, vmf_directions=rb_v.directions, uniform_seed=2025, uniform_origins=rb_u.origins, uniform_directions=rb_u.directions,
         origins=origins, directions=directions, termination_dist=term, mask=mask, sky_origins=sky_o, sky_directions=sky_d, multi_view_seed=MV_SEED,
         **{"out_" + k: v for k, v in out.items()}, **{"loss_" + k: v for k, v in losses.items()})


# ------------------------------------------------------------------------------ light-sum shaders (values + autograd gradients)
def golden_shaders():
    from reni.model_components.shaders import BlinnPhongShader, LambertianShader
    from neusky.model_components.renderers import RGBBlinnPhongRendererWithVisibility

    g = torch.Generator().manual_seed(61)
    N, M, K = 70, 37, 3
    leaf = lambda t: t.clone().requires_grad_(True)
    albedo = leaf(torch.rand(N, 3, generator=g))
    normals = leaf(torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1))
    dirs = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
    table = leaf(torch.exp(0.5 * torch.randn(K, M, 3, generator=g)))
    cam = torch.randint(0, K, (N,), generator=g)
    specular = leaf(torch.rand(N, 3, generator=g))
    shininess = leaf(1.5 + 30.0 * torch.rand(N, generator=g))
    view = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1)
    cot = torch.randn(N, 3, generator=g)
    cot2 = torch.randn(N, 3, generator=g)
    out = dict(albedo=albedo, normals=normals, dirs=dirs, table=table, cam=cam, specular=specular, shininess=shininess, view=view, cot=cot, cot2=cot2)
    leaves = dict(albedo=albedo, normals=normals, table=table, specular=specular, shininess=shininess)

    def grads(prefix, loss, names):
        gs = torch.autograd.grad(loss, [leaves[n] for n in names], allow_unused=True, retain_graph=True)
        for n, gr in zip(names, gs):
            out[f"{prefix}_d_{n}"] = torch.zeros_like(leaves[n]) if gr is None else gr

    ld, lc = dirs[None].expand(N, M, 3), table[cam]
    s0, rgb0 = LambertianShader.forward(albedo, normals, ld, lc, detach_normals=False)
    out["lambert_sum"], out["lambert_rgb"] = s0, rgb0
    grads("lambert", (s0 * cot).sum() + (rgb0 * cot2).sum(), ["albedo", "normals", "table"])

    bp = BlinnPhongShader.forward(albedo, normals, ld, lc, specular, shininess, view, detach_normals=False, normalize_directions=False)
    out["blinn_phong"] = bp
    grads("blinn_phong", (bp * cot).sum(), ["albedo", "normals", "table", "specular", "shininess"])
    bp_n = BlinnPhongShader.forward(albedo, normals, 1.7 * dirs[None], table[:1], specular, shininess, view, normalize_directions=True)
    out["blinn_phong_normalized_broadcast"] = bp_n

    # --- NeuSky's renderer: R rays x S samples, visibility per ray repeated over the samples (neusky_model.py:1755-1759)
    R, S = 10, 7
    assert R * S == N
    vis_ray = leaf(torch.rand(R, M, generator=g))
    weights = leaf(torch.rand(R, S, 1, generator=g) / S)
    bg = torch.rand(R, 3, generator=g)
    c2w = torch.randn(R, 3, 4, generator=g)
    c2w_s = c2w[:, None].expand(R, S, 3, 4).contiguous()
    leaves.update(vis_ray=vis_ray, weights=weights)
    ren = RGBBlinnPhongRendererWithVisibility()
    ren.train()
    rgb = ren(albedos=albedo.view(R, S, 3), normals=normals.view(R, S, 3), light_directions=ld, light_colors=lc,
              visibility=vis_ray[:, None].expand(R, S, M).reshape(N, M, 1), background_illumination=bg, weights=weights,
              shininess=shininess.view(R, S, 1), c2w_matrices=c2w_s)
    out.update(ren_vis=vis_ray, ren_weights=weights, ren_bg=bg, ren_c2w=c2w, ren_rgb=rgb, ren_cot=cot[:R])
    grads("ren", (rgb * cot[:R]).sum(), ["albedo", "normals", "table", "shininess", "vis_ray", "weights"])
    save("shaders", **out)


# ------------------------------------------------------------------------------ NeuSky-specific training losses
def golden_losses():
    from neusky.model_components.losses import RENISkyPixelLoss
    from neusky.utils.utils import linear_to_sRGB

    g = torch.Generator().manual_seed(71)
    R = 97
    hdr = torch.exp(torch.randn(R, 3, generator=g)).requires_grad_(True)           # hdr_background_colours
    image = torch.rand(R, 3, generator=g)
    sky = (torch.rand(R, generator=g) > 0.6).float()
    loss = RENISkyPixelLoss(alpha=0.1)(linear_to_sRGB(hdr), image, sky[:, None].expand(R, 3))    # neusky_model.py:1005-1012
    (grad,) = torch.autograd.grad(loss, hdr)
    save("losses", hdr=hdr, image=image, sky=sky, sky_pixel_loss=loss, sky_pixel_loss_d_hdr=grad)


# ------------------------------------------------------------------------------ state_dict layout of the reference's modules (drop-in boundary)
def golden_state_dicts():
    """Names, shapes and requires_grad flags of the reference's own modules (the drop-in modules of neusky_b200/fields.py must
    register exactly these; tests/test_dropin_state_dict.py).  The hash encodings are tcnn.Encoding in the reference (key
    `params`); under the shim they are the torch hash grid (key `hash_table`) -- recorded as the shim builds them.
    SDFAlbedoField cannot be built here (its parent class nerfstudio.fields.sdf_field.SDFField is absent)."""
    import json
    from reni.illumination_fields.reni_illumination_field import RENIField, RENIFieldConfig

    def layout(m):
        req = {n: bool(p.requires_grad) for n, p in m.named_parameters()}
        return {k: {"shape": list(v.shape), "requires_grad": req.get(k)} for k, v in m.state_dict().items()}

    field, _ = build_reference_ddf()
    cfg = RENIFieldConfig(
        conditioning="Attention", invariant_function="VN", equivariance="SO2", axis_of_invariance="z",
        positional_encoding="NeRF", encoded_input="Directions", latent_dim=100, hidden_features=128, hidden_layers=9,
        mapping_layers=5, mapping_features=128, num_attention_heads=8, num_attention_layers=6,
        output_activation="None", last_layer_linear=True, fixed_decoder=True, trainable_scale=True,
    )
    out = {
        "DirectionalDistanceField": layout(field),
        "RENIField": layout(RENIField(cfg, num_train_data=None, num_eval_data=None)),
        "RENIField_7_3": layout(RENIField(cfg, num_train_data=7, num_eval_data=3, normalisations={"min_max": None, "log_domain": True})),
    }
    path = os.path.join(HERE, "state_dict_keys.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"wrote {path}: " + ", ".join(f"{k} ({len(v)} entries)" for k, v in out.items()))


if __name__ == "__main__":
    only = sys.argv[1:]
    torch.manual_seed(0)
    for fn in (golden_icosphere, golden_lambert, golden_reni, golden_ddf_and_visibility, golden_ddf_fit, golden_shaders, golden_losses, golden_state_dicts):
        if not only or fn.__name__ in only:
            fn()
