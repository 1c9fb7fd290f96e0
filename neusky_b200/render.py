"""Host-side orchestration of the shading hot path: RENI++ radiance table -> Lambert pre-pass ->
fused DDF visibility + cosine-weighted sum (K4) -> background blend + sRGB.

Mirrors what NeuSkyFactoModel.sample_illumination / compute_visibility / lambertian_renderer do
between them (neusky/models/neusky_model.py:445-551, 1624-1778, 797-805) without materialising the
[R*S, D, 3] direction / colour tensors or the [R*S, D] visibility tensor.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import ops, packing
from .init import hash_scalings

Tensor = torch.Tensor


class SkyShader:
    """Holds device-resident packed weights for the DDF visibility field and the RENI++ decoder."""

    def __init__(self, ddf_params: Dict[str, Tensor], reni_params: Optional[Dict[str, Tensor]], device="cuda", ddf_radius: float = 1.0,
                 log2_T: int = 19, num_levels: int = 16, only_upper_hemisphere: bool = True, lower_hemisphere_visibility: float = 1.0,
                 impl: str = "tc2"):
        self.device = torch.device(device)
        self.radius = float(ddf_radius)
        self.log2_T = log2_T
        self.only_upper = only_upper_hemisphere
        self.lower_vis = float(lower_hemisphere_visibility)
        self.impl = impl
        self.scalings = hash_scalings(num_levels).to(self.device)
        self.hash_table = ddf_params["position_encoding.hash_table"].to(self.device, torch.float32).contiguous()
        # imported tiny-cuda-nn position grid (tcnn_import): int32 [L,4] level table + tcnn's interpolation flag; None = nerfstudio torch grid
        self.grid_meta: Optional[Tensor] = None
        self.grid_smoothstep = True
        self.set_ddf_weights(ddf_params)
        self.reni_blob = packing.pack_reni(reni_params, device=self.device) if reni_params is not None else None
        self.reni_gemm = packing.pack_reni_gemm(reni_params, device=self.device) if reni_params is not None else None
        self.reni_fused = packing.pack_reni_fused(reni_params, device=self.device) if (reni_params is not None and reni_params["network.residual_projection.weight"].shape[1] == 510) else None
        self.reni_tc_min_rows = 8192    # below this the fp32 SIMT decode wins (one launch, exact fp32)
        # frame-sized row batches: "fused" = one tcgen05 kernel, fp16 operands (radiance within 7e-4 relative of fp32, measured: mean 1e-4);
        # "3xtf32" = the layer-wise fp32-accurate GEMM chain (6x slower)
        self.reni_rows_impl = "fused" if self.reni_fused is not None else "3xtf32"
        # visibility sigmoid: bias and scale.  Defaults = the values the method config trains towards / fixes (target_min_bias 0.1,
        # target_max_scale 25: neusky_config.py:122-123); a model passes its learnable `visibility_threshold` (initialised to
        # 2 * ddf_radius, neusky_model.py:234) explicitly or sets these attributes from the checkpoint
        self.threshold = 0.1
        self.sigmoid_scale = 25.0
        self.k4_events = None   # bench hook: when a list, (start, end) CUDA events are recorded around every K4 launch

    def set_ddf_weights(self, ddf_params: Dict[str, Tensor]) -> None:
        """Packed on the host, one upload per kernel variant, lazily: only the blob of the variant that is actually launched exists."""
        self._ddf_p = {k: v.detach() for k, v in ddf_params.items() if k.startswith("ddf.")}
        self._ddf_blobs: Dict[str, Tensor] = {}

    _DDF_PACKERS = {"simt": "pack_ddf_simt", "tc": "pack_ddf_tc", "tc2": "pack_ddf_tc2"}

    def ddf_blob(self, impl: str) -> Tensor:
        b = self._ddf_blobs.get(impl)
        if b is None:
            if impl not in self._DDF_PACKERS:
                raise ValueError(f"impl must be one of {sorted(self._DDF_PACKERS)}, got {impl!r}")
            b = self._ddf_blobs[impl] = getattr(packing, self._DDF_PACKERS[impl])(self._ddf_p, device=self.device)
        return b

    ddf_blob_simt = property(lambda self: self.ddf_blob("simt"))
    ddf_blob_tc = property(lambda self: self.ddf_blob("tc"))
    ddf_blob_tc2 = property(lambda self: self.ddf_blob("tc2"))

    # -- direction set -------------------------------------------------------------------------
    def set_directions(self, dirs: Tensor) -> None:
        """dirs [D,3] (unit).  Mask = upper hemisphere (neusky_model.py:1650-1657)."""
        self.dirs = dirs.to(self.device, torch.float32).contiguous()
        if self.only_upper:
            m = self.dirs[:, 2] > 0
        else:
            m = torch.ones(self.dirs.shape[0], dtype=torch.bool, device=self.device)
        self.mask = m
        self.mask_u8 = m.to(torch.uint8).contiguous()
        self.dirs_sel = self.dirs[m].contiguous()
        self.sel_index = torch.where(m, torch.cumsum(m.to(torch.int32), 0, dtype=torch.int32) - 1, torch.full_like(m, -1, dtype=torch.int32)).to(torch.int32).contiguous()

    def radiance_table(self, latents: Tensor, scale: Optional[Tensor], rotation: Optional[Tensor] = None) -> Tensor:
        return ops.reni_radiance_table(self.dirs, latents, scale, self.reni_blob, rotation)

    def radiance_rows(self, ray_directions: Tensor, latents: Tensor, scale: Optional[Tensor], rotation: Optional[Tensor] = None) -> Tensor:
        """HDR radiance [N,3] along N directions for ONE latent code [1,L,3] (per-ray background, neusky_model.py:535-549): the
        tensor-core GEMM chain for frame-sized batches, the fp32 SIMT decode for small ones."""
        if ray_directions.shape[0] >= self.reni_tc_min_rows:
            if self.reni_rows_impl == "fused":
                return ops.reni_rows_fused(ray_directions, latents, scale, self.reni_blob, self.reni_fused, rotation)
            return ops.reni_rows_tc(ray_directions, latents, scale, self.reni_blob, self.reni_gemm, rotation)
        return ops.reni_radiance_table(ray_directions, latents, scale, self.reni_blob, rotation)[0]

    # -- shading ---------------------------------------------------------------------------------
    def shade(self, points: Tensor, normals: Tensor, wa: Tensor, radiance: Tensor, cam: Optional[Tensor] = None,
              want_vis: bool = False, want_ddf: bool = False, threshold: Optional[float] = None, sigmoid_scale: Optional[float] = None,
              impl: Optional[str] = None) -> Dict[str, Tensor]:
        """points [R,3]; normals, wa [R,S,3]; radiance [K,D,3] -> linear radiance sum [R,3]
        (sum_s w_s * albedo_s * sum_j c_sj vis_rj L_j) and, on request, the per-pair tensors."""
        impl = impl or self.impl
        threshold = self.threshold if threshold is None else float(threshold)
        sigmoid_scale = self.sigmoid_scale if sigmoid_scale is None else float(sigmoid_scale)
        inv_count, rgb_lin = ops.lambert_prep(normals, wa, self.dirs, self.mask_u8, radiance, cam, self.lower_vis)
        rad_sel = radiance[:, self.mask].contiguous()
        blob = self.ddf_blob(impl)
        if self.k4_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        vis, ddf, term = ops.sky_shade(points, normals, wa, inv_count, self.dirs_sel, rad_sel, blob, self.hash_table, self.scalings,
                                       self.log2_T, self.radius, threshold, sigmoid_scale, rgb_lin, cam, want_vis, want_ddf, impl,
                                       grid_meta=self.grid_meta, smoothstep=self.grid_smoothstep)
        if self.k4_events is not None:
            ev[1].record()
            self.k4_events.append(ev)
        out = {"rgb_lin": rgb_lin, "inv_count": inv_count}
        if vis is not None:
            full = torch.full((points.shape[0], self.dirs.shape[0]), self.lower_vis, device=self.device)
            full[:, self.mask] = vis
            out["visibility"] = full
            out["visibility_sel"] = vis
        if ddf is not None:
            out["expected_termination_dist"], out["termination_dist"] = ddf, term
        return out

    # -- config-2 entry points (BASELINE.json: surface points x RENI++ directions) ------------------
    def shade_points(self, points: Tensor, normals: Tensor, albedo: Tensor, latents: Tensor, scale: Optional[Tensor] = None,
                     rotation: Optional[Tensor] = None, threshold: Optional[float] = None, sigmoid_scale: Optional[float] = None) -> Tensor:
        """Device-resident inputs: points/normals/albedo [N,3] (one sample per point, weight 1), latents [1,L,3]
        -> sRGB [N,3].  RENI++ decode -> Lambert pre-pass -> fused DDF visibility + cosine sum -> sRGB."""
        N = points.shape[0]
        radiance = self.radiance_table(latents, scale, rotation)
        out = self.shade(points, normals.reshape(N, 1, 3), albedo.reshape(N, 1, 3), radiance, threshold=threshold, sigmoid_scale=sigmoid_scale)
        ones = torch.ones(N, device=self.device)
        return ops.shade_finalize(out["rgb_lin"], torch.zeros(N, 3, device=self.device), ones)

    def shade_points_host(self, points_h: Tensor, normals_h: Tensor, albedo_h: Tensor, latents: Tensor, scale: Optional[Tensor] = None,
                          out_h: Optional[Tensor] = None, **kw) -> Tensor:
        """Same, from (pinned) HOST buffers to a (pinned) host result: the call a user with CPU-side
        G-buffers makes.  Copies ride the current stream; returns after the result has landed."""
        for n, t in (("points_h", points_h), ("normals_h", normals_h), ("albedo_h", albedo_h)):
            if t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or t.shape[1] != 3:
                raise ValueError(f"{n}: expected a host fp32 [N,3] tensor")
        pts, nrm, alb = (t.to(self.device, non_blocking=True) for t in (points_h, normals_h, albedo_h))
        rgb = self.shade_points(pts, nrm, alb, latents, scale, **kw)
        if out_h is None:
            out_h = torch.empty(rgb.shape, dtype=torch.float32, pin_memory=True)
        out_h.copy_(rgb, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return out_h


# ==================================================================================================
# Full per-ray eval render (BASELINE.json configs 1 and 3)
# ==================================================================================================
def pinhole_rays(H: int, W: int, fx: float, fy: float, cx: float, cy: float, c2w: Tensor, device) -> tuple:
    """Camera rays exactly as nerfstudio Cameras.generate_rays builds them for a perspective camera [SURVEY A.8]:
    pixel centres at +0.5, d_cam = ((x-cx)/fx, -(y-cy)/fy, -1).  Returns origins, unit directions [H*W,3] and
    directions_norm [H*W,1], row-major, on `device`."""
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=device) + 0.5, torch.arange(W, dtype=torch.float32, device=device) + 0.5, indexing="ij")
    d_cam = torch.stack([(xs - cx) / fx, -(ys - cy) / fy, -torch.ones_like(xs)], -1).reshape(-1, 3)
    c2w = c2w.to(device, torch.float32)
    d = d_cam @ c2w[:3, :3].T
    dn = d.norm(dim=-1, keepdim=True)
    return c2w[:3, 3].expand(H * W, 3).contiguous(), (d / dn).contiguous(), dn.contiguous()


def sphere_collider(origins: Tensor, directions: Tensor, radius: float = 1.0, near_plane: float = 0.05, training: bool = False):
    """nerfstudio SphereCollider [SURVEY A.6], set at neusky/models/neusky_model.py:440."""
    ox, oy, oz = origins[..., 0:1], origins[..., 1:2], origins[..., 2:3]
    dx, dy, dz = directions[..., 0:1], directions[..., 1:2], directions[..., 2:3]
    # explicit component arithmetic (no reductions): every op is a correctly rounded fp32 elementwise op, so the CPU
    # oracle and the GPU host mirror produce bit-identical near / far
    a = dx * dx + dy * dy + dz * dz
    b = 2 * (ox * dx + oy * dy + oz * dz)
    c = (ox * ox + oy * oy + oz * oz) - radius**2
    disc = b * b - 4 * a * c
    t0 = (-b - torch.sqrt(disc)) / (2 * a)
    t1 = (-b + torch.sqrt(disc)) / (2 * a)
    near = torch.clamp(t0, min=near_plane if training else 0.0)
    far = torch.maximum(t1, near + 1e-6)
    return torch.nan_to_num(near, nan=0.0), torch.nan_to_num(far, nan=0.0)


_UNIT_BINS: Dict[tuple, Tensor] = {}


def _unit_bins(S: int, like: Tensor) -> Tensor:
    """torch.linspace(0, 1, S + 1)[None] with the CPU kernel's rounding, built on the host ONCE per (S, dtype, device): a fresh
    host tensor per call is a pageable host->device copy, i.e. a host synchronisation inside every placement (and illegal in a
    CUDA-graph capture of the training iteration)."""
    key = (int(S), like.dtype, str(like.device))
    t = _UNIT_BINS.get(key)
    if t is None:
        t = _UNIT_BINS[key] = torch.linspace(0.0, 1.0, S + 1, dtype=like.dtype)[None].to(like.device)
    return t


def uniform_samples(near: Tensor, far: Tensor, S: int):
    """nerfstudio UniformSampler eval placement [SURVEY A.6]: same op sequence as the oracle so that starts / ends
    are bit-identical to the CPU reference."""
    bins = _unit_bins(S, near)                                                        # built on the host: CPU linspace rounding
    e = bins * far + (1 - bins) * near
    return e[:, :-1].contiguous(), e[:, 1:].contiguous()


class RayRenderer:
    """Eval render of a ray bundle: sample placement -> K2 (SDF/albedo field + analytic normals) -> K3 (NeuS alpha,
    transmittance, composites) -> RENI++ radiance -> K4 (DDF sky visibility + Lambertian sum) -> sRGB.
    Mirrors NeuSkyFactoModel.forward / get_outputs in eval mode (neusky/models/neusky_model.py:425-443, 553-931) for
    one camera (one latent code) per call; outputs use the reference's keys (:881-931)."""

    def __init__(self, sdf_params: Dict[str, Tensor], ddf_params: Dict[str, Tensor], reni_params: Dict[str, Tensor], device="cuda",
                 log2_T: int = 19, num_levels: int = 16, ddf_radius: float = 1.0, impl: str = "tc2", sdf_impl: str = "tc",
                 proposal_params: Optional[Sequence[Dict[str, Tensor]]] = None, proposal_max_res: Sequence[int] = (64, 256),
                 num_proposal_samples_per_ray: Sequence[int] = (256, 96), proposal_log2_T: int = 17, ddf_log2_T: Optional[int] = None):
        """``proposal_params``: state of the two HashMLPDensityFields -> sample placement by the proposal-network sampler
        (the shipped NeuS-facto configuration, neusky_model.py:561); None -> uniform placement."""
        self.device = torch.device(device)
        self.proposal_fields = None
        if proposal_params is not None:
            from . import proposal as _proposal

            self.proposal_fields = [_proposal.HashMLPDensityField(p, mr, log2_hashmap_size=proposal_log2_T, device=device) for p, mr in zip(proposal_params, proposal_max_res)]
            self._proposal_counts = tuple(num_proposal_samples_per_ray)
            self._proposal_samplers = {}
        self.log2_T = log2_T
        self.sdf_impl = sdf_impl
        # the DDF's position encoding is always 16 x 2^19 in the reference (directional_distance_field.py:139-145); `ddf_log2_T` lets reduced tests shrink it
        self.shader = SkyShader(ddf_params, reni_params, device=device, ddf_radius=ddf_radius, log2_T=log2_T if ddf_log2_T is None else ddf_log2_T,
                                num_levels=num_levels, impl=impl)
        self.scalings = self.shader.scalings
        self.sweep_events = None
        self.sdf_table = sdf_params["encoding.hash_table"].to(self.device, torch.float32).contiguous()
        self.sdf_grid_meta: Optional[Tensor] = None       # imported tiny-cuda-nn grid of the SDF field (see SkyShader.grid_meta)
        self.sdf_grid_smoothstep = True
        self.set_sdf_weights(sdf_params)

    def set_sdf_weights(self, sdf_params: Dict[str, Tensor]) -> None:
        p = {k: v for k, v in sdf_params.items() if k.startswith(("glin", "clin"))}
        self.sdf_blob = packing.pack_sdf_tc(p, device=self.device) if self.sdf_impl == "tc" else packing.pack_sdf_simt(p, device=self.device)
        var = sdf_params["deviation_network.variance"]
        self.inv_s = float(torch.exp(10.0 * var.detach().float().cpu()).clip(1e-6, 1e6))   # LearnedVariance.get_variance [SURVEY A.4]

    def set_directions(self, dirs: Tensor) -> None:
        self.shader.set_directions(dirs)

    def radiance_rows_cam(self, ray_directions: Tensor, cam: Tensor, latents: Tensor, scale: Optional[Tensor], rotation: Optional[Tensor] = None) -> Tensor:
        """Per-ray background radiance of a MIXED-camera batch: ray n is decoded with latent code cam[n] (neusky_model.py:535-549)."""
        sh = self.shader
        if ray_directions.shape[0] >= sh.reni_tc_min_rows:
            if sh.reni_rows_impl == "fused":
                return ops.reni_rows_fused(ray_directions, latents, scale, sh.reni_blob, sh.reni_fused, rotation, row_cam=cam)
            return ops.reni_rows_tc(ray_directions, latents, scale, sh.reni_blob, sh.reni_gemm, rotation, row_cam=cam)
        return ops.reni_radiance_rows(ray_directions, cam, latents, scale, sh.reni_blob, rotation)

    @torch.no_grad()
    def render(self, origins: Tensor, directions: Tensor, dnorm: Tensor, S: int, latent: Tensor, scale: Tensor, rotation: Optional[Tensor] = None,
               threshold: Optional[float] = None, sigmoid_scale: Optional[float] = None, cos_anneal_ratio: float = 1.0, want_vis: bool = False,
               steps_minmax: Optional[Tensor] = None, want_cache: bool = False, collapse_cache: bool = False,
               radiance: Optional[Tensor] = None, background: Optional[Tensor] = None, cam: Optional[Tensor] = None,
               want_visibility_batch: bool = False, want_prop_depth: bool = False, compact_cache: bool = False) -> Dict[str, Tensor]:
        """origins/directions [R,3], dnorm [R,1].  One camera: latent [L,3], scale scalar tensor.  Mixed-camera batch (what
        torch.unique(camera_indices) handles in the reference, neusky_model.py:461): latent [K,L,3], scale [K] and `cam` [R]
        int32 = the latent row of every ray.
        `radiance` [K,D,3] / `background` [R,3]: RENI++ decodes the caller already has (a tiled frame decodes once per frame,
        `illumination_for`, instead of once per tile).  Returns the reference's output keys (neusky_model.py:881-931)."""
        R = origins.shape[0]
        sh = self.shader
        near, far = sphere_collider(origins, directions)
        rs = weights_list = samples_list = None
        if self.proposal_fields is not None:
            from . import proposal as _proposal

            smp = self._proposal_samplers.get(S)
            if smp is None:
                smp = self._proposal_samplers[S] = _proposal.ProposalNetworkSampler(S, self._proposal_counts, len(self.proposal_fields))
            rs, weights_list, samples_list = smp.generate_ray_samples(origins, directions, near, far, self.proposal_fields)
            e = rs.euclidean_bins
            starts, ends = e[:, :-1].contiguous(), e[:, 1:].contiguous()
        else:
            starts, ends = uniform_samples(near, far, S)
        x = origins[:, None, :] + directions[:, None, :] * starts[..., None]          # get_start_positions
        f = ops.sdf_field(x, self.sdf_blob, self.sdf_table, self.scalings, self.log2_T, impl=self.sdf_impl, grid_meta=self.sdf_grid_meta,
                          smoothstep=self.sdf_grid_smoothstep)
        c = ops.neus_composite(f["sdf"], f["gradient"], f["albedo"], directions, starts, ends, ends - starts, dnorm, self.inv_s, cos_anneal_ratio, False,
                               steps_minmax=steps_minmax)
        if cam is not None:
            cam = cam.reshape(-1).to(torch.int32).contiguous()
        if radiance is None or background is None:
            Z = latent.reshape(-1, latent.shape[-2], 3).to(self.device, torch.float32)
            sc = scale.reshape(-1).to(self.device, torch.float32)
            if Z.shape[0] != 1 and cam is None:
                raise ValueError("render: several latent codes need `cam` (the latent row of every ray)")
            if radiance is None:
                radiance = sh.radiance_table(Z, sc, rotation)                            # [K,D,3]
            if background is None:                                                       # per-ray background (neusky_model.py:535-549)
                background = sh.radiance_rows(directions, Z, sc, rotation) if cam is None else self.radiance_rows_cam(directions, cam, Z, sc, rotation)
        bg = background
        pts = ops.surface_points(origins, directions, c["p2p_dist"], sh.radius)
        s = sh.shade(pts, c["normals"], c["wa"], radiance, cam=cam, want_vis=want_vis or want_cache, want_ddf=want_visibility_batch,
                     threshold=threshold, sigmoid_scale=sigmoid_scale)
        rgb = ops.shade_finalize(s["rgb_lin"], bg, c["accumulation"])
        out = {"rgb": rgb, "albedo": c["albedo"], "accumulation": c["accumulation"][:, None], "depth": c["depth"][:, None], "p2p_dist": c["p2p_dist"][:, None],
               "normal": c["normal"], "weights": c["weights"][..., None], "hdr_background_colours": bg, "directions_norm": dnorm,
               "sdf_at_termination": None,                                              # training-only branch (ddf_model.py:241-251)
               "normal_vis": (c["normal"] + 1.0) / 2.0,                                 # viewer output (neusky_model.py:919)
               "starts": starts, "ends": ends}
        if want_vis:
            out["visibility"] = s["visibility"]
        if want_visibility_batch:
            # neusky_model.py:1766-1776: the DDF loss batch of this render (mask = ones, no sdf_at_termination outside training)
            out["expected_termination_dist"] = s["expected_termination_dist"]
            out["visibility_batch"] = {"termination_dist": s["termination_dist"], "mask": torch.ones_like(s["termination_dist"]), "sdf_at_termination": None}
        if want_prop_depth and samples_list is not None:
            # neusky_model.py:910-917: expected depth of every proposal level (DepthRenderer "expected" on that level's own weights / bins)
            nf, ff = near.reshape(-1, 1), far.reshape(-1, 1)
            for i, (w, smp_i) in enumerate(zip(weights_list, samples_list)):
                e = smp_i.spacing_bins * ff + (1.0 - smp_i.spacing_bins) * nf
                steps = (e[:, :-1] + e[:, 1:]) * 0.5
                w2 = w.reshape(R, -1)
                d_i = (w2 * steps).sum(-1) / (w2.sum(-1) + 1e-10)
                out[f"prop_depth_{i}"] = torch.clip(d_i, steps.min(), steps.max())[:, None]
        if want_cache:
            # everything a new illumination needs (fixed geometry): per-sample shading inputs, per-ray visibility of the
            # DDF directions, accumulation.  The geometry-only outputs above stay valid for every latent code.
            if collapse_cache:
                # collapsed form: per (ray, direction) coefficients with the visibility folded in -- D x 3 floats
                # per ray, independent of S; a new illumination is one streaming pass over it
                # G over ALL directions (block-per-ray kernel), then fold the per-ray visibility in (lower hemisphere = lower_vis)
                H = ops.lambert_collapse_sel(c["normals"], c["wa"], s["inv_count"], sh.dirs) * s["visibility"][:, :, None]
                if compact_cache:
                    # compact form: fp16 rows only for rays that hit something, per-row scale; background rows only where 1 - accumulation > 0
                    acc = c["accumulation"]
                    rows = torch.nonzero(acc > 0).reshape(-1).to(torch.int32)
                    H16, hscale = ops.relight_pack_h16(H, rows)
                    out["relight_cache"] = {"H16": H16, "hscale": hscale, "rows": rows, "bg_rows": torch.nonzero(acc < 1).reshape(-1).to(torch.int32),
                                            "accumulation": acc, "directions": directions, "n_dirs": sh.dirs.shape[0]}
                else:
                    out["relight_cache"] = {"H": H, "accumulation": c["accumulation"], "directions": directions}
            else:
                out["relight_cache"] = {"normals": c["normals"], "wa": c["wa"], "inv_count": s["inv_count"], "visibility_sel": s["visibility_sel"],
                                        "accumulation": c["accumulation"], "directions": directions}
        return out

    @torch.no_grad()
    def shadow_map(self, origins: Tensor, directions: Tensor, p2p_dist: Tensor, accumulation: Tensor, azimuth_deg: float, elevation_deg: float,
                   threshold: float, sigmoid_scale: float, accumulation_mask_threshold: float = 0.0) -> Dict[str, Tensor]:
        """Viewer shadow map (neusky_model.py:632-672): visibility of ONE light direction (azimuth / elevation, z up) from the
        rendered surface points, and the raw `difference` = min(|p - q|, 2r) - DDF (:1724-1727, returned when
        compute_shadow_map=True, :1764-1765); both masked by accumulation > threshold."""
        import math

        az, el = math.radians(azimuth_deg), math.radians(elevation_deg)
        d = torch.tensor([[math.cos(az) * math.cos(el), math.sin(az) * math.cos(el), math.sin(el)]], device=self.device, dtype=torch.float32)
        sh = self.shader
        saved = (sh.dirs, sh.mask, sh.mask_u8, sh.dirs_sel, sh.sel_index) if hasattr(sh, "dirs") else None
        try:
            sh.set_directions(d)
            R = origins.shape[0]
            pts = ops.surface_points(origins, directions, p2p_dist.reshape(R), sh.radius)
            zero = torch.zeros((R, 1, 3), device=self.device)
            o = sh.shade(pts, zero, zero, torch.zeros((1, 1, 3), device=self.device), want_vis=True, want_ddf=True, threshold=threshold, sigmoid_scale=sigmoid_scale)
        finally:
            if saved is not None:
                sh.dirs, sh.mask, sh.mask_u8, sh.dirs_sel, sh.sel_index = saved
        m = (accumulation.reshape(R, 1) > accumulation_mask_threshold).to(torch.float32)
        vis = o["visibility"].reshape(R, 1, 1) * m[:, :, None]
        if "expected_termination_dist" in o:
            diff = torch.clamp(o["termination_dist"], max=2.0 * sh.radius) - o["expected_termination_dist"]
            diff = diff.reshape(R, -1) * m
        else:                                   # the direction points into the lower hemisphere: no DDF query, fully lit (:1742-1753)
            diff = torch.zeros((R, 0), device=self.device)
        return {"visibility": vis, "difference": diff}

    @torch.no_grad()
    def relight(self, cache: Dict[str, Tensor], latent: Tensor, scale: Tensor, rotation: Optional[Tensor] = None,
                radiance: Optional[Tensor] = None, background: Optional[Tensor] = None) -> Tensor:
        """sRGB [R,3] of the cached rays under another RENI++ latent code / rotation (BASELINE.json config 5): RENI++ decode
        of the direction set and of the per-ray background, then one Lambertian pass over the cached visibility --
        no SDF field, no compositing, no DDF.  Equals render(...)["rgb"] for the same latent.
        `radiance` [1,D,3] / `background` [R,3]: already decoded for this latent (a frame cut into tiles decodes the direction
        table and the background of all its rays once per latent instead of once per tile: `illumination_for`)."""
        sh = self.shader
        if radiance is None or background is None:
            Z = latent.reshape(1, -1, 3).to(self.device, torch.float32)
            sc = scale.reshape(1).to(self.device, torch.float32)
            if radiance is None:
                radiance = sh.radiance_table(Z, sc, rotation)
            if background is None:
                background = sh.radiance_rows(cache["directions"], Z, sc, rotation)
        bg = background
        if "H16" in cache:
            lin = ops.relight_h16_multi(cache["H16"], cache["hscale"], cache["rows"], cache["accumulation"].shape[0], cache["n_dirs"], radiance)[0]
        elif "H" in cache:
            lin = ops.relight_collapsed(cache["H"], radiance)
        else:
            lin = ops.lambert_relight(cache["normals"], cache["wa"], cache["inv_count"], sh.dirs, sh.sel_index, radiance, cache["visibility_sel"], None, sh.lower_vis)
        return ops.shade_finalize(lin, bg, cache["accumulation"])

    @torch.no_grad()
    def relight_many(self, cache: Dict[str, Tensor], radiance: Tensor, background: Tensor) -> Tensor:
        """sRGB [NL,R,3] of the cached rays (collapsed cache) under NL illuminations at once: radiance [NL,D,3] and background [NL,R,3]
        from `illumination_for`, one pass over the cache per four illuminations."""
        if "H" not in cache and "H16" not in cache:
            raise ValueError("relight_many needs the collapsed cache (render(..., want_cache=True, collapse_cache=True))")
        NL, R = radiance.shape[0], cache["accumulation"].shape[0]
        if "H16" in cache:
            lin = ops.relight_h16_multi(cache["H16"], cache["hscale"], cache["rows"], R, cache["n_dirs"], radiance)
        else:
            lin = ops.relight_collapsed_multi(cache["H"], radiance)
        rgb = ops.shade_finalize(lin.reshape(NL * R, 3), background.reshape(NL * R, 3).contiguous(), cache["accumulation"].repeat(NL))
        return rgb.reshape(NL, R, 3)

    @staticmethod
    def merge_caches(caches: Sequence[Dict[str, Tensor]]) -> Dict[str, Tensor]:
        """Compact caches of consecutive ray tiles -> one cache of the whole bundle (row indices shifted by the tile offsets), so that
        a sweep over a frame is a handful of launches instead of one set per tile."""
        if not caches or "H16" not in caches[0]:
            raise ValueError("merge_caches takes compact caches (render(..., want_cache=True, collapse_cache=True, compact_cache=True))")
        off, rows, bg = 0, [], []
        for c in caches:
            rows.append(c["rows"] + off)
            bg.append(c["bg_rows"] + off)
            off += c["accumulation"].shape[0]
        return {"H16": torch.cat([c["H16"] for c in caches], 0), "hscale": torch.cat([c["hscale"] for c in caches], 0), "rows": torch.cat(rows, 0),
                "bg_rows": torch.cat(bg, 0), "accumulation": torch.cat([c["accumulation"] for c in caches], 0),
                "directions": torch.cat([c["directions"] for c in caches], 0), "n_dirs": caches[0]["n_dirs"]}

    @torch.no_grad()
    def relight_sweep(self, cache: Dict[str, Tensor], latents: Tensor, scales: Tensor, rotation: Optional[Tensor] = None, group: int = 32) -> Tensor:
        """sRGB [NL, R, 3] of the cached rays (compact cache) under NL latent codes [NL, L, 3] / scales [NL] (BASELINE.json configs[4]).
        Per group of latent codes: ONE RENI++ table decode, ONE fused row decode of the background of the rays whose
        1 - accumulation is not zero (every other ray gets no background: it is multiplied by 0), ONE streaming pass over the cache
        (up to 32 codes per read), one finalize -- no per-tile, per-code Python loop."""
        if "H16" not in cache:
            raise ValueError("relight_sweep takes the compact cache")
        sh = self.shader
        NL, R = latents.shape[0], cache["accumulation"].shape[0]
        Z = latents.to(self.device, torch.float32).contiguous()
        sc = scales.reshape(-1).to(self.device, torch.float32).contiguous()
        bg_rows = cache["bg_rows"].long()
        Nb = bg_rows.shape[0]
        dirs_b = cache["directions"][bg_rows].contiguous()
        out = torch.empty((NL, R, 3), device=self.device, dtype=torch.float32)
        ev = self.sweep_events          # bench hook: when a dict, CUDA-event pairs per stage ("table", "pass", "background", "finalize")

        def timed(name, fn):
            if ev is None:
                return fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r_ = fn(); e1.record()
            ev.setdefault(name, []).append((e0, e1))
            return r_

        all_rows = Nb == R
        for a in range(0, NL, group):
            b = min(NL, a + group)
            g = b - a
            rad = timed("table", lambda: sh.radiance_table(Z[a:b], sc[a:b], rotation))                                # [g, D, 3]
            lin = timed("pass", lambda: ops.relight_h16_multi(cache["H16"], cache["hscale"], cache["rows"], R, cache["n_dirs"], rad))      # [g, R, 3]
            if Nb:
                row_cam = torch.arange(g, device=self.device, dtype=torch.int32).repeat_interleave(Nb)
                rows_rad = timed("background", lambda: self.radiance_rows_cam(dirs_b.repeat(g, 1), row_cam, Z[a:b], sc[a:b], rotation))    # [g * Nb, 3]
            if all_rows:
                bg = rows_rad.reshape(g, R, 3)
            else:
                bg = torch.zeros((g, R, 3), device=self.device, dtype=torch.float32)
                if Nb:
                    bg[:, bg_rows] = rows_rad.reshape(g, Nb, 3)
            out[a:b] = timed("finalize", lambda: ops.shade_finalize(lin.reshape(g * R, 3), bg.reshape(g * R, 3), cache["accumulation"].repeat(g))).reshape(g, R, 3)
        return out

    @torch.no_grad()
    def illumination_for(self, latent: Tensor, scale: Tensor, ray_directions: Tensor, rotation: Optional[Tensor] = None):
        """(radiance table [1,D,3], per-ray background [R,3]) of one latent code for a whole ray bundle: the two RENI++ decodes
        of `relight`, done once per latent."""
        Z = latent.reshape(1, -1, 3).to(self.device, torch.float32)
        sc = scale.reshape(1).to(self.device, torch.float32)
        return self.shader.radiance_table(Z, sc, rotation), self.shader.radiance_rows(ray_directions, Z, sc, rotation)


def global_steps_minmax(origins: Tensor, directions: Tensor, S: int) -> Tensor:
    """[min, max] of the sample mid-points over a whole ray bundle (what nerfstudio's DepthRenderer clips to when the
    bundle is rendered in one batch), from the first and last bins only -- same arithmetic as uniform_samples, so the
    values are bit-identical to a full placement.  Lets tiles / ranks clip to an image-global range."""
    near, far = sphere_collider(origins, directions)
    bins = _unit_bins(S, near)
    sel = bins[:, [0, 1, S - 1, S]]
    e = sel * far + (1 - sel) * near
    return torch.stack([((e[:, 0] + e[:, 1]) / 2).min(), ((e[:, 2] + e[:, 3]) / 2).max()]).contiguous()


def render_image(renderer: "RayRenderer", origins: Tensor, directions: Tensor, dnorm: Tensor, S: int, latent: Tensor, scale: Tensor,
                 tile: int = 16384, keys=("rgb", "albedo", "normal", "depth", "p2p_dist", "accumulation"), group=None, **kw) -> Dict[str, Tensor]:
    """Eval render of a full camera (BASELINE.json config 3): the H*W rays are cut into tiles of `tile` rays, tile i is
    rendered by rank i mod world (weights replicated, no data-path collective) and the per-ray outputs are gathered
    once at the end (neusky_b200/parallel.py).  Single process: a plain loop over tiles.  The reference renders the
    same image serially in 256-ray chunks on one GPU (neusky/models/neusky_model.py:1413-1437)."""
    from . import parallel

    n = origins.shape[0]
    # uniform placement: clip depth to the image-global range (known in closed form); proposal placement is data dependent, so each
    # tile clips to its own range -- as the reference does per 256-ray chunk (neusky_model.py:1413-1437)
    mm = global_steps_minmax(origins, directions, S) if renderer.proposal_fields is None else None

    def fn(idx: Tensor) -> Dict[str, Tensor]:
        outs = {k: [] for k in keys}
        rad = bg_all = None
        if idx.numel():          # the frame has one latent code: decode its direction table and the background of this rank's rays once
            rad, bg_all = renderer.illumination_for(latent, scale, directions[idx].contiguous(), kw.get("rotation"))
        for a in range(0, idx.shape[0], tile):
            sel = idx[a:a + tile]
            o = renderer.render(origins[sel].contiguous(), directions[sel].contiguous(), dnorm[sel].contiguous(), S, latent, scale, steps_minmax=mm,
                                radiance=rad, background=bg_all[a:a + tile], **kw)
            for k in keys:
                outs[k].append(o[k])
        if not idx.numel():
            w = {"rgb": 3, "albedo": 3, "normal": 3}
            return {k: torch.zeros((0, w.get(k, 1)), device=origins.device) for k in keys}
        return {k: torch.cat(v, 0) for k, v in outs.items()}

    return parallel.render_sharded(fn, n, tile, tuple(keys), device=origins.device, group=group)
