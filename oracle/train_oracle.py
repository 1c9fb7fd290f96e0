"""TEST INFRASTRUCTURE -- CPU restatement (torch autograd) of one NeuSky training iteration's forward and losses.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
It strings the per-function oracle (oracle/neusky_oracle.py) together the way the reference's training path does:

  NeuSkyFactoModel.get_outputs (training)            neusky/models/neusky_model.py:738-931
    sample_and_forward_field                         :553-736   (sample placement: uniform, see note)
    SDFAlbedoField.get_outputs                       neusky/fields/sdf_albedo_field.py:211-269
    compute_visibility, depth detached               :596-632, 1624-1778  (sdf_to_visibility_stop_gradients="depth",
                                                     neusky/configs/neusky_config.py:156)
    DDFModel.get_outputs sdf_at_termination branch   neusky/models/ddf_model.py:241-251
    RGBLambertianRendererWithVisibility              neusky/model_components/renderers.py:60-176
  NeuSkyFactoModel.get_loss_dict (training branch)   :935-1031 with the coefficients of neusky_config.py:132-146

Note (parity unpinned items): the proposal sampler / interlevel loss are outside the hot path built so far;
sample placement is the deterministic uniform placement of oracle.uniform_samples, and the interlevel loss is absent
on both sides.  The losses whose definitions live in nerfstudio (monosdf_normal_loss) are restated from memory.
Gradients come from torch autograd, which is what the reference uses.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import neusky_oracle as O

Tensor = torch.Tensor

LOSS_COEFFICIENTS = {  # neusky/configs/neusky_config.py:132-146
    "rgb_l1_loss": 1.0, "eikonal_loss": 0.1, "fg_mask_loss": 1.0, "sdf_level_set_visibility_loss": 1.0, "sky_pixel_loss": 1.0,
    "hashgrid_density_loss": 1e-4, "ground_plane_loss": 0.1, "visibility_sigmoid_loss": 0.01,
}


def monosdf_normal_loss(normal_pred: Tensor, normal_gt: Tensor) -> Tensor:
    """nerfstudio.model_components.losses.monosdf_normal_loss [NS-mem]: L1 + (1 - cos) on normalised vectors."""
    normal_gt = torch.nn.functional.normalize(normal_gt, p=2, dim=-1)
    normal_pred = torch.nn.functional.normalize(normal_pred, p=2, dim=-1)
    l1 = torch.abs(normal_pred - normal_gt).sum(dim=-1).mean()
    cos = (1.0 - torch.sum(normal_pred * normal_gt, dim=-1)).mean()
    return l1 + cos


def sky_pixel_loss(inputs: Tensor, targets: Tensor, mask: Tensor, alpha: float = 0.1) -> Tensor:
    """neusky/model_components/losses.py:44-58 (RENISkyPixelLoss)."""
    inputs, targets = inputs * mask, targets * mask
    mse = torch.nn.functional.mse_loss(inputs, targets)
    sim = torch.nn.functional.cosine_similarity(inputs, targets, dim=1, eps=1e-20)
    return mse + alpha * (1 - sim.mean())


def training_forward(batch: Dict[str, Tensor], sdf_p, ddf_p, reni_p, latents: Tensor, scales: Tensor, threshold: Tensor, dirs: Tensor, S: int,
                     log2_T: int, ddf_radius: float = 1.0, sigmoid_scale: float = 25.0, cos_anneal_ratio: float = 1.0,
                     grid_positions: Optional[Tensor] = None, grid_dirs: Optional[Tensor] = None, grid_gap: float = 0.2,
                     sample_edges: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """batch: origins, directions [R,3], dnorm [R,1], cam [R] int64, image [R,3], fg/ground/sky masks [R].
    sample_edges [R,S+1]: euclidean bin edges from the proposal-network sampler (neusky_model.py:561; placement is detached
    there too) instead of the uniform placement; the sampler itself is covered by oracle/sampler_oracle.py.
    Returns the outputs dict (rgb, eik_grad, weights, normal, accumulation, hdr_background_colours, sdf_at_termination,
    visibility, expected_termination_dist, grid_density) of neusky_model.py:881-931 in training mode."""
    o, d, dn, cam = batch["origins"], batch["directions"], batch["dnorm"], batch["cam"]
    R = o.shape[0]
    dt = o.dtype
    sca = O.hash_scalings().to(dt)
    near, far = O.sphere_collider(o, d, radius=1.0, training=True)
    if sample_edges is not None:
        starts, ends = sample_edges[:, :-1, None].to(dt), sample_edges[:, 1:, None].to(dt)
    else:
        starts, ends = O.uniform_samples(near, far, S)
    x = (o[:, None, :] + d[:, None, :] * starts).reshape(-1, 3).detach().requires_grad_(True)
    h = O.sdf_geo_network(x, sdf_p, sca, log2_T)                                         # sdf_albedo_field.py:233
    sdf, geo = h[:, :1], h[:, 1:]
    grad = torch.autograd.grad(sdf, x, torch.ones_like(sdf), create_graph=True, retain_graph=True)[0]   # :235-238
    alb = O.sdf_colour_network(x, geo, sdf_p)                                           # :241
    sdf, grad, alb = sdf.reshape(R, S, 1), grad.reshape(R, S, 3), alb.reshape(R, S, 3)
    inv_s = torch.exp(sdf_p["deviation_network.variance"] * 10.0).clip(1e-6, 1e6)       # LearnedVariance [NS-mem A.4]
    alpha = O.neus_alpha(sdf, grad, d[:, None, :], ends - starts, inv_s, cos_anneal_ratio)   # :266
    w, T = O.weights_from_alphas(alpha)                                                 # neusky_model.py:565
    acc = w.sum(-2)
    normals = torch.nn.functional.normalize(grad, p=2, dim=-1)                           # :251
    p2p = O.render_depth_expected(w, starts, ends)                                      # :591
    normal = (w * normals).sum(-2)                                                      # :806-813
    radiance = O.reni_radiance_table(dirs, latents, scales, reni_p)                     # [K,D,3]  :460-518
    bg_all = O.reni_radiance_table(d, latents, scales, reni_p)                          # [K,R,3]  :535-549 (per-ray camera below)
    bg = bg_all[cam, torch.arange(R)]
    pts = O.surface_points(o, d, p2p.detach(), ddf_radius)                              # :608-611 depth detached, :1667-1683
    v = O.compute_visibility(pts, dirs, ddf_p, sca, log2_T, ddf_radius, threshold, sigmoid_scale)
    mask = v["mask"]
    d_sel = dirs[mask]
    Dp = d_sel.shape[0]
    pos = pts[:, None, :].expand(R, Dp, 3).reshape(-1, 3)
    dd = d_sel[None].expand(R, Dp, 3).reshape(-1, 3)
    q = O.ray_sphere_intersection(pos, dd, ddf_radius)
    term_pts = q + (-dd) * v["expected_termination_dist"][:, None]                      # ddf_model.py:243
    sdf_term = O.sdf_geo_network(term_pts, sdf_p, sca, log2_T)[:, :1]                   # :250 (stop_gradients False)
    rgb = O.lambertian_render(alb, normals, dirs, radiance[cam], v["visibility"], bg, w, training=True)
    out = {"rgb": rgb, "eik_grad": grad, "weights": w, "normal": normal, "accumulation": acc, "hdr_background_colours": bg, "p2p_dist": p2p,
           "sdf_at_termination": sdf_term, "visibility": v["visibility"], "expected_termination_dist": v["expected_termination_dist"]}
    if grid_positions is not None:                                                      # neusky_model.py:675-734
        gx = grid_positions.detach().requires_grad_(True)
        gs = O.sdf_geo_network(gx, sdf_p, sca, log2_T)[:, :1]
        gg = torch.autograd.grad(gs, gx, torch.ones_like(gs), create_graph=True, retain_graph=True)[0]
        out["grid_density"] = O.neus_alpha(gs, gg, grid_dirs, torch.full_like(gs, grid_gap), inv_s, cos_anneal_ratio)
    return out


def training_losses(out: Dict[str, Tensor], batch: Dict[str, Tensor], threshold: Tensor, target_min_bias: float = 0.1) -> Dict[str, Tensor]:
    """neusky_model.py:935-1031 (training branch), scaled by LOSS_COEFFICIENTS (:1066)."""
    image, fg, ground, sky = batch["image"], batch["fg"], batch["ground"], batch["sky"]
    keep = (1.0 - sky.to(image.dtype))[:, None]
    L: Dict[str, Tensor] = {}
    L["rgb_l1_loss"] = torch.nn.functional.l1_loss(image * keep, out["rgb"] * keep)                            # :945-952
    L["eikonal_loss"] = ((out["eik_grad"].norm(2, dim=-1) - 1) ** 2).mean()                                    # :960-962
    ws = out["weights"].sum(dim=1).clip(1e-3, 1.0 - 1e-3)                                                      # :966-969
    L["fg_mask_loss"] = torch.nn.functional.binary_cross_entropy(ws, fg.to(ws.dtype)[:, None])
    gm = ground.to(image.dtype)[:, None]
    up = torch.tensor([0.0, 0.0, 1.0], dtype=image.dtype).expand_as(out["normal"])
    L["ground_plane_loss"] = monosdf_normal_loss(out["normal"] * gm, up * gm)                                  # :998-1003
    srgb_bg = O.linear_to_srgb(out["hdr_background_colours"])
    L["sky_pixel_loss"] = sky_pixel_loss(srgb_bg, image, sky.to(image.dtype)[:, None].expand_as(srgb_bg))      # :1005-1012
    L["visibility_sigmoid_loss"] = torch.nn.functional.mse_loss(threshold, torch.tensor(target_min_bias, dtype=threshold.dtype))   # :1014-1033
    L["sdf_level_set_visibility_loss"] = torch.nn.functional.mse_loss(out["sdf_at_termination"], torch.zeros_like(out["sdf_at_termination"]))
    if "grid_density" in out:
        L["hashgrid_density_loss"] = torch.nn.functional.l1_loss(out["grid_density"], torch.zeros_like(out["grid_density"]))
    return {k: v * LOSS_COEFFICIENTS[k] for k, v in L.items()}
