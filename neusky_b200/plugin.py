"""Host-side mirror of the reference's plugin surface for the hot path: the same class names, call
signatures, output keys and error behaviour as the nerfstudio Field / Model components the reference registers, with
the C-ABI kernels behind them.  nerfstudio itself is not a dependency: ray containers are duck-typed (anything with
``.frustums.origins/.directions/.starts/.ends``, ``.deltas``, ``.camera_indices`` works -- nerfstudio's RaySamples
does), and the output dictionaries are keyed by enums whose ``.value`` strings equal nerfstudio's FieldHeadNames /
the reference's NeuSkyFieldHeadNames / RENIFieldHeadNames, so a real plugin re-keys them with one dict comprehension.

What stays in the reference unchanged: configs, samplers' ProposalNetworkSampler, losses, pipelines, data.
"""
from __future__ import annotations

from enum import Enum
from typing import Dict, Optional

import torch

from . import ops, packing
from .init import hash_scalings
from .render import SkyShader

Tensor = torch.Tensor


class FieldHeadNames(Enum):            # nerfstudio.field_components.field_heads.FieldHeadNames (subset used on the path)
    SDF = "sdf"
    NORMALS = "normals"
    GRADIENT = "gradient"
    ALPHA = "alpha"


class NeuSkyFieldHeadNames(Enum):      # neusky/field_components/neusky_fieldheadnames.py:6-14
    ALBEDO = "albedo"
    SHININESS = "shininess"
    VISIBILITY = "visibility"
    TERMINATION_DISTANCE = "termination_distance"
    PROBABILITY_OF_HIT = "probability_of_hit"


class RENIFieldHeadNames(Enum):        # ns_reni/reni/field_components/field_heads.py
    RGB = "rgb"
    MU = "mu"
    LOG_VAR = "log_var"


class _LearnedVariance:
    """nerfstudio LearnedVariance [SURVEY A.4]: get_variance() = exp(10 * variance).clip(1e-6, 1e6)."""

    def __init__(self, variance: Tensor):
        self.variance = variance

    def get_variance(self) -> Tensor:
        return torch.exp(self.variance * 10.0).clip(1e-6, 1e6)


class SDFAlbedoField:
    """neusky/fields/sdf_albedo_field.py:80-282.  ``params`` is the reference state_dict of the field
    (``glin{l}.weight_v/.weight_g/.bias``, ``clin{l}.*``, ``encoding.hash_table``, ``deviation_network.variance``)."""

    def __init__(self, params: Dict[str, Tensor], device="cuda", log2_T: int = 19, num_levels: int = 16, impl: str = "tc"):
        self.device = torch.device(device)
        self.log2_T, self.impl = log2_T, impl
        self.scalings = hash_scalings(num_levels).to(self.device)
        self._cos_anneal_ratio = 1.0                                     # sdf_albedo_field.py:167
        self.load_params(params)

    def load_params(self, params: Dict[str, Tensor]) -> None:
        p = {k: v.detach().to(self.device) for k, v in params.items()}
        self.hash_table = p["encoding.hash_table"].to(torch.float32).contiguous()
        self.blob = packing.pack_sdf_tc(p) if self.impl == "tc" else packing.pack_sdf_simt(p)
        self._blob_simt = packing.pack_sdf_simt(p) if self.impl == "tc" else self.blob
        self.deviation_network = _LearnedVariance(p["deviation_network.variance"].to(torch.float32))

    def set_cos_anneal_ratio(self, anneal: float) -> None:
        self._cos_anneal_ratio = float(anneal)

    def get_sdf_at_pos(self, positions: Tensor) -> Tensor:
        """:169-174 -> [N,1] (exact fp32 kernel: the callers compare this value against thresholds)."""
        return ops.sdf_field(positions.reshape(-1, 3), self._blob_simt, self.hash_table, self.scalings, self.log2_T, want_grad=False, want_albedo=False, impl="simt")["sdf"]

    def get_alpha(self, ray_samples, sdf: Optional[Tensor] = None, gradients: Optional[Tensor] = None) -> Tensor:
        """nerfstudio SDFField.get_alpha [SURVEY A.5] (called at :266 and neusky_model.py:732)."""
        if sdf is None or gradients is None:
            x = ray_samples.frustums.origins + ray_samples.frustums.directions * ray_samples.frustums.starts
            f = ops.sdf_field(x, self.blob, self.hash_table, self.scalings, self.log2_T, want_albedo=(self.impl == "tc"), impl=self.impl)
            sdf, gradients = f["sdf"], f["gradient"]
        inv_s = self.deviation_network.get_variance()
        true_cos = (ray_samples.frustums.directions * gradients).sum(-1, keepdim=True)
        rho = self._cos_anneal_ratio
        iter_cos = -(torch.relu(-true_cos * 0.5 + 0.5) * (1.0 - rho) + torch.relu(-true_cos) * rho)
        nxt = sdf + iter_cos * ray_samples.deltas * 0.5
        prv = sdf - iter_cos * ray_samples.deltas * 0.5
        prev_cdf, next_cdf = torch.sigmoid(prv * inv_s), torch.sigmoid(nxt * inv_s)
        return ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).clip(0.0, 1.0)

    def forward(self, ray_samples, compute_normals: bool = False, return_alphas: bool = False) -> Dict[Enum, Tensor]:
        """:211-282.  Output shapes follow the reference: [*batch, 3] / [*batch, 1]."""
        if ray_samples.camera_indices is None:
            raise AttributeError("Camera indices are not provided.")       # :218-219
        x = ray_samples.frustums.origins + ray_samples.frustums.directions * ray_samples.frustums.starts   # get_start_positions (:225)
        f = ops.sdf_field(x, self.blob, self.hash_table, self.scalings, self.log2_T, impl=self.impl)
        out = {
            NeuSkyFieldHeadNames.ALBEDO: f["albedo"],
            FieldHeadNames.SDF: f["sdf"],
            FieldHeadNames.NORMALS: torch.nn.functional.normalize(f["gradient"], p=2, dim=-1),     # :251
            FieldHeadNames.GRADIENT: f["gradient"],
        }
        if return_alphas:
            out[FieldHeadNames.ALPHA] = self.get_alpha(ray_samples, f["sdf"], f["gradient"])       # :266
        return out

    __call__ = forward


class RENIField:
    """ns_reni/reni/illumination_fields/reni_illumination_field.py:90-593 (the decoder as NeuSky uses it: SO2
    invariance, attention conditioning, fixed decoder).  ``params`` = the reference state_dict of the field."""

    def __init__(self, params: Dict[str, Tensor], device="cuda", latent_dim: int = 100, hidden: int = 128, num_layers: int = 6, log_domain: bool = True):
        self.device = torch.device(device)
        self.latent_dim, self.hidden, self.num_layers, self.log_domain = latent_dim, hidden, num_layers, log_domain
        self.blob = packing.pack_reni({k: v.detach().to(self.device) for k, v in params.items()}, num_layers)

    def radiance_table(self, directions: Tensor, latent_codes: Tensor, scale: Optional[Tensor], rotation: Optional[Tensor] = None) -> Tensor:
        """[D,3] directions x [K,L,3] latent codes -> unnormalised HDR radiance [K,D,3]: what sample_illumination
        (neusky_model.py:445-551) needs, without expanding the latent code per (camera, direction) row."""
        return ops.reni_radiance_table(directions, latent_codes, scale, self.blob, rotation, self.hidden, self.num_layers, self.log_domain)

    def unnormalise(self, x: Tensor) -> Tensor:
        return torch.exp(x) if self.log_domain else x                  # base_spherical_field.py:143-154 (min_max unused: False buffer)

    def forward(self, ray_samples, rotation: Optional[Tensor] = None, latent_codes: Optional[Tensor] = None, scale: Optional[Tensor] = None) -> Dict[Enum, Tensor]:
        """:575-593.  directions [N,3]; latent_codes [N,L,3] / scale [N] as the reference passes them (one row per ray).
        Rows are grouped by ``ray_samples.camera_indices`` (the reference builds latent_codes by indexing with exactly
        those, neusky_model.py:481-493); returns the log-domain RGB like the reference (call unnormalise())."""
        if rotation is not None and rotation.dim() == 3:
            raise NotImplementedError("Batched rotation not implemented yet")          # :520-521
        d = ray_samples.frustums.directions.reshape(-1, 3)
        N = d.shape[0]
        if latent_codes is None:
            raise ValueError("RENIField.forward: latent_codes are required on this path")
        cam = ray_samples.camera_indices
        cam = torch.zeros(N, dtype=torch.long, device=d.device) if cam is None else cam.reshape(-1).to(torch.long)
        rgb = torch.empty((N, 3), device=d.device, dtype=torch.float32)
        for c in torch.unique(cam).tolist():
            m = cam == c
            first = int(torch.nonzero(m)[0])
            sc = None if scale is None else scale.reshape(-1)[first:first + 1].contiguous()
            tab = ops.reni_radiance_table(d[m].contiguous(), latent_codes[first:first + 1].contiguous(), sc, self.blob, rotation, self.hidden, self.num_layers, self.log_domain)
            # the kernel returns unnormalised radiance (exp of the log-domain output, scale added in the log domain,
            # :561-565); the reference's forward returns the log-domain value and leaves exp to unnormalise()
            rgb[m] = torch.log(tab[0]) if self.log_domain else tab[0]
        return {RENIFieldHeadNames.RGB: rgb, RENIFieldHeadNames.MU: None, RENIFieldHeadNames.LOG_VAR: None}

    __call__ = forward


class RGBLambertianRendererWithVisibility:
    """neusky/model_components/renderers.py:60-176 in compact form: instead of the reference's expanded
    light_directions / light_colors [R*S,D,3] and visibility [R*S,D,1], pass the direction set [D,3], the radiance
    table [K,D,3] (+ per-ray row) and the per-ray visibility [R,D] (it is per ray in the reference too:
    neusky_model.py:1755-1759 repeats it over samples)."""

    def forward(self, albedos: Tensor, normals: Tensor, light_directions: Tensor, light_colors: Tensor, visibility: Optional[Tensor],
                background_illumination: Tensor, weights: Tensor, ray_indices=None, num_rays=None, camera_rows: Optional[Tensor] = None) -> Tensor:
        if ray_indices is not None:
            raise NotImplementedError("packed samples are never used on the NeuSky path (ray_indices is None, neusky_model.py:797-805)")
        R, S = albedos.shape[0], albedos.shape[1]
        D = light_directions.shape[0]
        w = weights.reshape(R, S, 1)
        wa = (w * albedos).contiguous()
        sel = torch.arange(D, dtype=torch.int32, device=albedos.device)
        vis = torch.ones((R, D), device=albedos.device) if visibility is None else visibility.reshape(R, D).contiguous()
        inv_count, _ = ops.lambert_prep(normals, wa, light_directions, torch.ones(D, dtype=torch.uint8, device=albedos.device), light_colors.reshape(-1, D, 3), camera_rows, 1.0)
        lin = ops.lambert_relight(normals, wa, inv_count, light_directions, sel, light_colors.reshape(-1, D, 3), vis, camera_rows, 1.0)
        return ops.shade_finalize(lin, background_illumination, w.sum(1).reshape(R))

    __call__ = forward


class NeuSkyVisibility:
    """The visibility step of NeuSkyFactoModel (neusky/models/neusky_model.py:1624-1778) behind the reference's
    ``compute_visibility`` signature.  ``ddf_params`` = state_dict of the DDF field (DirectionalDistanceField)."""

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
        return {
            "visibility": vis[:, None, :].expand(R, S, D).reshape(R * S, D, 1),           # :1755-1759
            "expected_termination_dist": out["expected_termination_dist"],
            "visibility_batch": {"termination_dist": out["termination_dist"], "mask": sh.mask},
        }
