"""The reference's plugin surface for the hot path in one namespace: drop-in ``nn.Module`` classes with the
reference's constructor / call signatures, ``state_dict`` names and optimizer-group names, the C-ABI kernels behind them.

    fields   SDFAlbedoField, DirectionalDistanceField, RENIField (+ their *Config dataclasses)      neusky_b200/fields.py
    models   NeuSkyFactoModel, DDFModel (+ configs)                                                 neusky_b200/models.py
    renderer RGBLambertianRendererWithVisibility                                                    below
    samplers IcosahedronSampler, EquirectangularSampler                                             neusky_b200/samplers.py
    shaders  LambertianShader, BlinnPhongShader, RGBBlinnPhongRendererWithVisibility                neusky_b200/shaders.py
    method   neusky_b200.neusky_config.NeuSkyB200 (`ns-train neusky-b200`)                           neusky_b200/neusky_config.py

nerfstudio itself is not a dependency: ray containers are duck-typed (neusky_b200.rays mirrors the attribute names), output
dictionaries are keyed by enums whose ``.value`` strings equal nerfstudio's FieldHeadNames / the reference's
NeuSkyFieldHeadNames / RENIFieldHeadNames (``fields._rekey`` converts).
What stays in the reference unchanged: data parsers / managers, pipelines, optimizers, schedulers, viewer, metrics.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from . import ops
from .fields import (DirectionalDistanceField, DirectionalDistanceFieldConfig, FieldHeadNames, LearnedVariance, NeuSkyFieldHeadNames, RENIField,  # noqa: F401
                     RENIFieldConfig, RENIFieldHeadNames, SDFAlbedoField, SDFAlbedoFieldConfig)
from .models import DDFModel, DDFModelConfig, NeuSkyFactoModel, NeuSkyFactoModelConfig, SceneBox  # noqa: F401
from .rays import Frustums, RayBundle, RaySamples  # noqa: F401
from .render import SkyShader

Tensor = torch.Tensor


class RGBLambertianRendererWithVisibility(nn.Module):
    """neusky/model_components/renderers.py:60-176 with the reference's arguments:

        albedos, normals [R,S,3]; light_directions, light_colors [R*S, D, 3]; visibility [R*S, D, 1] | None;
        background_illumination [R,3]; weights [R,S,1]   ->   rgb [R,3]

    The reference's three big tensors are redundant by construction -- every row of ``light_directions`` is the same direction
    set (neusky_model.py:520-525), ``light_colors`` is constant over a ray's samples (:512-518) and so is ``visibility``
    (:1755-1759) -- so they are read through views: row 0 of the directions, every S-th row of colours and visibility (no copy
    when the caller passes the stride-0 expanded tensors ``NeuSkyFactoModel.sample_illumination`` / ``compute_visibility`` return).
    Compact forms are accepted too: light_directions [D,3], light_colors [K,D,3] (+ ``camera_rows`` [R] int32) or [R,D,3],
    visibility [R,D]."""

    def forward(self, albedos: Tensor, normals: Tensor, light_directions: Tensor, light_colors: Tensor, visibility: Optional[Tensor],
                background_illumination: Tensor, weights: Tensor, ray_indices=None, num_rays=None, camera_rows: Optional[Tensor] = None) -> Tensor:
        if ray_indices is not None:
            raise NotImplementedError("packed samples are never used on the NeuSky path (ray_indices is None, neusky_model.py:797-805)")
        R, S = albedos.shape[0], albedos.shape[1]
        dirs = light_directions[0] if light_directions.dim() == 3 else light_directions                       # renderers.py:93-98 broadcast
        D = dirs.shape[0]
        dirs = dirs.contiguous()
        if light_colors.dim() == 3 and light_colors.shape[0] == R * S and not (S == 1 and camera_rows is not None):
            lc = light_colors.reshape(R, S, D, 3)[:, 0] if S > 1 else light_colors                            # per-ray table [R,D,3]
            if R > 0 and lc.stride(0) == 0:                                                                    # one camera, expanded view
                lc, camera_rows = lc[:1], None
            else:
                camera_rows = torch.arange(R, dtype=torch.int32, device=albedos.device)
        else:
            lc = light_colors.reshape(-1, D, 3)
            if lc.shape[0] == R and camera_rows is None and R != 1:
                camera_rows = torch.arange(R, dtype=torch.int32, device=albedos.device)
        lc = lc.contiguous()
        if visibility is None:
            vis = torch.ones((R, D), device=albedos.device)
        elif visibility.dim() == 3 and visibility.shape[0] == R * S:
            vis = visibility.reshape(R, S, D)[:, 0].contiguous()
        else:
            vis = visibility.reshape(R, D).contiguous()
        w = weights.reshape(R, S, 1)
        wa = (w * albedos).contiguous()
        normals = normals.contiguous()
        sel = torch.arange(D, dtype=torch.int32, device=albedos.device)
        inv_count, _ = ops.lambert_prep(normals, wa, dirs, torch.ones(D, dtype=torch.uint8, device=albedos.device), lc, camera_rows, 1.0)
        lin = ops.lambert_relight(normals, wa, inv_count, dirs, sel, lc, vis, camera_rows, 1.0)
        return ops.shade_finalize(lin, background_illumination.contiguous(), w.sum(1).reshape(R), training=self.training)   # eval clamp :173-174


class NeuSkyVisibility:
    """The visibility step of NeuSkyFactoModel (neusky/models/neusky_model.py:1624-1778) as a stand-alone object over a DDF
    state_dict (the model-level entry point is ``NeuSkyFactoModel.compute_visibility``).  ``ddf_params`` = state_dict of the DDF
    field (DirectionalDistanceField)."""

    def __init__(self, ddf_params: Dict[str, Tensor], device="cuda", ddf_radius: float = 1.0, log2_T: int = 19, impl: str = "tc2",
                 only_upperhemisphere_visibility: bool = True, lower_hemisphere_visibility: float = 1.0):
        self.shader = SkyShader(ddf_params, None, device=device, ddf_radius=ddf_radius, log2_T=log2_T, only_upper_hemisphere=only_upperhemisphere_visibility,
                                lower_hemisphere_visibility=lower_hemisphere_visibility, impl=impl)
        self.ddf_radius = ddf_radius

    def compute_visibility(self, ray_samples, depth: Tensor, illumination_directions: Tensor, threshold_distance: float, sigmoid_scale: float,
                           compute_shadow_map: bool = False) -> Dict[str, Tensor]:
        """ray_samples [R,S]; depth [R,1] (p2p distance, B.11); illumination_directions [R*S,D,3] or [D,3] (row 0 is used,
        :1648) -> {"visibility" [R*S,D,1], "expected_termination_dist" [R*D'], "visibility_batch": {...}}."""
        sh = self.shader
        dirs = illumination_directions[0] if illumination_directions.dim() == 3 else illumination_directions
        o = ray_samples.frustums.origins[:, 0].contiguous()
        d = ray_samples.frustums.directions[:, 0].contiguous()
        R, S = ray_samples.frustums.origins.shape[0], ray_samples.frustums.origins.shape[1]
        D = dirs.shape[0]
        sh.set_directions(dirs)
        pts = ops.surface_points(o, d, depth.reshape(R), sh.radius)
        dummy = torch.zeros((R, 1, 3), device=pts.device)
        rad = torch.zeros((1, D, 3), device=pts.device)
        out = sh.shade(pts, dummy, dummy, rad, want_vis=True, want_ddf=True, threshold=float(threshold_distance), sigmoid_scale=float(sigmoid_scale))
        vis = out["visibility"]                                                           # [R,D]
        term = out["termination_dist"]
        vd = {
            "visibility": vis[:, None, :].expand(R, S, D).reshape(R * S, D, 1),           # :1755-1759
            "expected_termination_dist": out["expected_termination_dist"],
            "visibility_batch": {"termination_dist": term, "mask": torch.ones_like(term), "sdf_at_termination": None},      # :1766-1776
        }
        if compute_shadow_map:
            vd["difference"] = torch.clamp(term, max=2.0 * sh.radius) - out["expected_termination_dist"]      # :1724-1727, :1764-1765
        return vd
