"""DDF fitting pass: every `ns-train neusky` iteration also fits the sky-visibility DDF to the current
SDF scene (fit_visibility_field=True, neusky/pipelines/neusky_pipeline.py:272-289, 493-515):

  visibility_train_sampler()            VMFDDFSampler: 8 points on the DDF sphere x 128 von-Mises-Fisher directions
                                        (neusky/model_components/ddf_sampler.py:193-286; config neusky_config.py:207-212)
  generate_ddf_ground_truth(rays)       render accumulation / expected depth / normals of those rays through the SDF field
                                        (neusky/models/neusky_model.py:1337-1367)
  DDFModel.get_outputs (training)       DDF on the sampled rays, on the multi-view batch, on the sky-ray batch, and the SDF at
                                        the predicted termination points (neusky/models/ddf_model.py:183-369)
  DDFModel.get_loss_dict                depth_l1 (scene-centre weighted), sdf_l2, multi_view, sky_ray (:407-493,
                                        coefficients neusky_config.py:178-205)

The samplers draw from torch's CPU generator in the reference (the rays are moved to the device afterwards), so they stay
host code here and reproduce the reference's ray placement bit for bit under the same seed.  Everything that touches a
network runs on the CUDA ops of this package (csrc/train_ops.cu `nsk_ddf_rows_fwd`, the tf32 tcgen05 GEMM chain of
train.py, the SDF field op and the warp-per-ray compositing kernel); the three DDF batches (1024 + 1024 + 256 rows at the
default config) go through the network as ONE row batch.  No torch fallback: the ops raise without the CUDA library.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch

from . import ops
from .train import NeuSkyTrainStep, ddf_param_list, ddf_termination, sdf_field, sdf_param_list

Tensor = torch.Tensor


# =====================================================================================================================
# host samplers (neusky/model_components/ddf_sampler.py)
# =====================================================================================================================
def random_points_on_unit_sphere(num_points: int) -> Tensor:
    """Uniform points on S^2 from two torch.rand draws (azimuth first, then the polar angle through acos), cartesian
    (ddf_sampler.py:72-85 == neusky/utils/utils.py:33-46, sph2cart :95-99)."""
    azimuth = 2 * torch.pi * torch.rand(num_points)
    polar = torch.acos(2 * torch.rand(num_points) - 1)
    sp = torch.sin(polar)
    return torch.stack((sp * torch.cos(azimuth), sp * torch.sin(azimuth), torch.cos(polar)), dim=1)


def _flip_to_upper(positions: Tensor) -> Tensor:
    below = positions[:, 2] < 0
    positions[below] = -positions[below]
    return positions


@dataclass
class DDFSamplerConfig:
    """ddf_sampler.py:40-50 with the NeuSky values of neusky_config.py:207-212."""
    num_samples_on_sphere: int = 8
    num_rays_per_sample: int = 128
    only_sample_upper_hemisphere: bool = True
    concentration: float = 20.0


class DDFSampler:
    """Host ray sampler for DDF fitting; `__call__` returns (origins [N,3], directions [N,3]) on `device` where the reference
    returns a RayBundle with the same two tensors (pixel_area = 1, camera_indices = 0, directions_norm = 1)."""

    def __init__(self, config: Optional[DDFSamplerConfig] = None, ddf_sphere_radius: float = 1.0, device="cpu"):
        self.config = config or DDFSamplerConfig()
        self.ddf_sphere_radius = ddf_sphere_radius
        self.device = torch.device(device)
        self.num_rays = self.config.num_samples_on_sphere * self.config.num_rays_per_sample

    def generate_ddf_samples(self, num_positions: int, num_directions: int, positions: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
        raise NotImplementedError

    def __call__(self, num_positions: Optional[int] = None, num_directions: Optional[int] = None, positions: Optional[Tensor] = None):
        n_pos = self.config.num_samples_on_sphere if num_positions is None else num_positions
        n_dir = self.config.num_rays_per_sample if num_directions is None else num_directions
        return self.generate_ddf_samples(n_pos, n_dir, positions)

    forward = __call__


class UniformDDFSampler(DDFSampler):
    """Uniform inward-facing directions per sphere point (ddf_sampler.py:119-180).  The reference tiles the positions with
    `positions.repeat(num_directions, 1)` (position index fastest) against directions laid out position-major; that pairing
    is reproduced as is."""

    def generate_ddf_samples(self, num_positions, num_directions, positions=None):
        if positions is None:
            positions = random_points_on_unit_sphere(num_positions)
        if self.config.only_sample_upper_hemisphere:
            positions = _flip_to_upper(positions)
        inward = -positions
        dirs = random_points_on_unit_sphere(num_directions * inward.shape[0]).reshape(inward.shape[0], num_directions, 3)
        wrong_side = torch.sum(inward.unsqueeze(1) * dirs, dim=2) < 0
        dirs[wrong_side] = -dirs[wrong_side]
        positions = positions * self.ddf_sphere_radius
        return positions.repeat(num_directions, 1).to(self.device), dirs.reshape(-1, 3).to(self.device)


class VMFDDFSampler(DDFSampler):
    """von-Mises-Fisher directions around the inward normal of each sphere point (ddf_sampler.py:193-286; Wood's rejection
    sampler for the cosine, :205-223)."""

    def _vmf_cosines(self, dim: int, kappa: float, n: int) -> Tensor:
        b = (dim - 1) / (2 * kappa + (4 * kappa ** 2 + (dim - 1) ** 2) ** 0.5)
        x0 = torch.tensor((1 - b) / (1 + b))
        c = kappa * x0 + (dim - 1) * torch.log(1 - x0 ** 2)
        kept, have = [], 0
        beta = torch.distributions.beta.Beta((dim - 1) / 2, (dim - 1) / 2)
        while have < n:
            m = min(n, int((n - have) * 1.5))
            z = beta.sample((m,))
            t = (1 - (1 + b) * z) / (1 - (1 - b) * z)
            score = kappa * t + (dim - 1) * torch.log(1 - x0 * t) - c
            ok = score >= -torch.exp(torch.ones(m))      # the reference's acceptance threshold is -e (not log u): kept as is
            kept.append(t[ok])
            have += len(kept[-1])
        return torch.cat(kept)[:n]

    def random_vmf(self, normals: Tensor, kappa: float, num_samples: int) -> Tensor:
        normals = normals / torch.norm(normals, dim=-1, keepdim=True)
        N, dim = normals.shape
        tang = torch.normal(0, 1, (N, num_samples, dim))
        tang = tang / torch.norm(tang, dim=-1, keepdim=True)
        tang = tang - (torch.einsum("nij,nj->ni", tang, normals))[..., None] * normals[:, None, :]
        tang = tang / torch.norm(tang, dim=-1, keepdim=True)
        cos = self._vmf_cosines(dim, kappa, N * num_samples).reshape(N, num_samples)
        sin = torch.sqrt(1 - cos ** 2)
        x = tang * sin[..., None] + cos[..., None] * normals[:, None, :]
        return x / torch.norm(x, dim=-1, keepdim=True)

    def generate_ddf_samples(self, num_positions, num_directions, positions=None):
        if positions is None:
            positions = random_points_on_unit_sphere(num_positions)
        if self.config.only_sample_upper_hemisphere:
            positions = _flip_to_upper(positions)
        dirs = self.random_vmf(-positions, self.config.concentration, num_directions)
        wrong_side = torch.einsum("nij,nj->ni", dirs, -positions) < 0
        dirs[wrong_side] = -dirs[wrong_side]
        positions = positions * self.ddf_sphere_radius
        origins = positions.unsqueeze(1).repeat(1, num_directions, 1).reshape(-1, 3)
        return origins.to(self.device), dirs.reshape(-1, 3).to(self.device)


# =====================================================================================================================
# DDFModel training path
# =====================================================================================================================
@dataclass
class DDFModelConfig:
    """neusky/models/ddf_model.py:53-86 with the values NeuSky trains with (neusky_config.py:162-206)."""
    include_depth_loss_scene_center_weight: bool = True
    scene_center_weight_exp: float = 3.0
    scene_center_weight_include_z: bool = False
    mask_to_circumference: bool = False
    inverse_depth_weight: bool = False
    loss_inclusions: Dict[str, bool] = field(default_factory=lambda: {
        "depth_l1_loss": True, "depth_l2_loss": False, "sdf_l1_loss": False, "sdf_l2_loss": True,
        "prob_hit_loss": False, "normal_loss": False, "multi_view_loss": True, "sky_ray_loss": True})
    loss_coefficients: Dict[str, float] = field(default_factory=lambda: {
        "depth_l1_loss": 1.0, "depth_l2_loss": 0.0, "sdf_l1_loss": 1.0, "sdf_l2_loss": 0.01,
        "prob_hit_loss": 0.01, "normal_loss": 1.0, "multi_view_loss": 0.01, "sky_ray_loss": 1.0})


def ray_sphere_exit(positions: Tensor, directions: Tensor, radius: float) -> Tensor:
    """Far intersection of unit-direction rays with the origin-centred sphere (neusky/utils/utils.py:68-93)."""
    b = 2 * (directions * positions).sum(-1)
    c = (positions * positions).sum(-1) - radius ** 2
    disc = b ** 2 - 4 * c
    t = torch.max((-b - torch.sqrt(disc)) / 2, (-b + torch.sqrt(disc)) / 2)
    return positions + t.unsqueeze(-1) * directions


class DDFFit:
    """Host mirror of the visibility-field half of NeuSkyPipeline.get_train_loss_dict (neusky_pipeline.py:272-289): owns no
    parameters -- the DDF and SDF parameters are those of the `NeuSkyTrainStep` it is given (the reference's
    `model.visibility_field` / `model.field`), so one optimizer step updates both passes' gradients together."""

    def __init__(self, step: NeuSkyTrainStep, config: Optional[DDFModelConfig] = None, sampler: Optional[DDFSampler] = None,
                 stop_sdf_gradients: bool = False, accumulation_mask_threshold: float = 0.0):
        self.step = step
        self.config = config or DDFModelConfig()
        self.ddf_radius = step.radius
        self.sampler = sampler or VMFDDFSampler(DDFSamplerConfig(), ddf_sphere_radius=step.radius, device=step.dev)
        self.stop_sdf_gradients = stop_sdf_gradients                     # neusky_config.py:45 (False)
        self.accumulation_mask_threshold = accumulation_mask_threshold    # neusky_config.py:214 (0.0)
        self.training = True

    # -- neusky_model.py:1337-1367 ---------------------------------------------------------------------------------
    def generate_ddf_ground_truth(self, origins: Tensor, directions: Tensor, mask_threshold: float = 0.5) -> Dict[str, Tensor]:
        """Render accumulation, expected depth (clamped to the sphere diameter) and normals of the rays through the SDF field:
        collider -> sample placement (the step's proposal sampler when it has one, else uniform) -> SDFAlbedoField(return_alphas) ->
        weights -> renderers.  Differentiable w.r.t. the SDF
        field (the reference only cuts this graph when stop_sdf_gradients is set, neusky_pipeline.py:505-513)."""
        from . import autograd as nba
        from .render import sphere_collider, uniform_samples

        st = self.step
        R, S = origins.shape[0], st.S
        sdf_p = st.group("sdf")
        near, far = sphere_collider(origins, directions, radius=1.0, training=True)
        if st.proposal_fields is not None:
            # the model's own sampler, as in the reference (:1343); its state for the interlevel loss was consumed in the main pass
            with torch.no_grad():
                rs, _, _ = st.proposal_sampler.generate_ray_samples(origins, directions, near, far, st.proposal_fields)
            e = rs.euclidean_bins
            starts, ends = e[:, :-1].contiguous(), e[:, 1:].contiguous()
        else:
            starts, ends = uniform_samples(near, far, S)
        x = (origins[:, None, :] + directions[:, None, :] * starts[:, :, None]).reshape(-1, 3)
        sdf, grad, _ = sdf_field(st.sdf_cfg, x, sdf_p["encoding.hash_table"], st.sdf_weights(), want_normals=True, want_albedo=False)
        inv_s = torch.exp(sdf_p["deviation_network.variance"] * 10.0).clip(1e-6, 1e6)
        dn = torch.ones(R, device=origins.device)
        alb0 = torch.zeros(R, S, 3, device=origins.device)
        _w, _wa, _n, acc, p2p_raw, normal, _a, _bgT = nba.neus_composite(sdf.reshape(R, S), grad.reshape(R, S, 3), alb0, inv_s, directions.contiguous(),
                                                                         starts, ends, ends - starts, dn, st.cos_anneal_ratio)
        mids = (starts + ends) * 0.5
        p2p = torch.clip(p2p_raw, mids.min(), mids.max())                       # DepthRenderer("expected") clip [NS-mem A.7]
        p2p = torch.clamp(p2p, max=2 * self.ddf_radius).reshape(-1, 1)          # :1350-1351
        accumulations = acc.reshape(-1, 1)
        return {"origins": origins, "directions": directions, "accumulations": accumulations,
                "mask": (accumulations > mask_threshold).float(), "termination_dist": p2p, "normals": normal.reshape(-1, 3)}

    # -- neusky_pipeline.py:493-515 --------------------------------------------------------------------------------
    def draw_host(self) -> Tuple[Tensor, Tensor, Optional[Tensor]]:
        """The HOST random draws of one fitting pass, in the reference's order on torch's CPU generator: the sampler's rays
        (ddf_sampler.py), then the multi-view sphere points (ddf_model.py:279-284) -> CPU tensors (origins [N,3], directions [N,3],
        multi_view_points [N,3] or None).  `forward(..., rays=, multi_view_points=)` takes them back (on the device): a captured
        iteration (graphed.py) cannot draw on the host, so it draws here, outside the graph, and copies into its static inputs."""
        dev, self.sampler.device = self.sampler.device, torch.device("cpu")
        try:
            origins, directions = self.sampler()
        finally:
            self.sampler.device = dev
        mv = None
        if self.config.loss_inclusions.get("multi_view_loss") and self.training:
            mv = random_points_on_unit_sphere(origins.shape[0])
            mv[:, 2] = torch.abs(mv[:, 2])
        return origins, directions, mv

    def generate_ddf_samples(self, sky_origins: Optional[Tensor] = None, sky_directions: Optional[Tensor] = None,
                             rays: Optional[Tuple[Tensor, Tensor]] = None) -> Dict[str, Tensor]:
        """Sampler -> ground truth (+ the sky-ray bundle the datamanager supplies, `get_sky_ray_bundle(256)`).  `rays`: pre-drawn
        (origins, directions) on the device instead of a fresh draw from the sampler (`draw_host`)."""
        origins, directions = self.sampler() if rays is None else rays
        if self.stop_sdf_gradients:
            with torch.no_grad():
                data = self.generate_ddf_ground_truth(origins, directions, self.accumulation_mask_threshold)
        else:
            data = self.generate_ddf_ground_truth(origins, directions, self.accumulation_mask_threshold)
        if sky_origins is not None:
            data["sky_origins"], data["sky_directions"] = sky_origins, sky_directions
        return data

    # -- ddf_model.py:183-369 --------------------------------------------------------------------------------------
    def get_outputs(self, batch: Dict[str, Tensor], stop_gradients: bool = True, multi_view_points: Optional[Tensor] = None) -> Dict[str, Tensor]:
        """Training-mode DDFModel.get_outputs on batch["origins"/"directions"].  `multi_view_points` [N,3] overrides the random
        sphere points of the multi-view batch (tests); by default they are drawn on the host like the reference does."""
        st, cfg = self.step, self.config
        inc = cfg.loss_inclusions
        positions, directions = batch["origins"].reshape(-1, 3), batch["directions"].reshape(-1, 3)
        N = positions.shape[0]
        rows_o, rows_d = [positions], [directions]
        out: Dict[str, Tensor] = {}

        n_mv = 0
        if inc.get("multi_view_loss") and self.training:
            # a random second viewpoint on the (upper) sphere per ground-truth termination point (:279-307)
            gt_points = positions + directions * batch["termination_dist"].repeat(1, 3)
            if multi_view_points is None:
                multi_view_points = random_points_on_unit_sphere(N)
                multi_view_points[:, 2] = torch.abs(multi_view_points[:, 2])
            pts = multi_view_points.to(gt_points)
            to_gt = gt_points - pts
            to_gt = to_gt / torch.norm(to_gt, dim=-1).unsqueeze(-1)
            rows_o.append(pts)
            rows_d.append(to_gt)
            n_mv = N
        n_sky = 0
        if inc.get("sky_ray_loss") and self.training:
            # camera rays that reach the sky: from their exit point on the sphere, looking back, the DDF must return the distance
            # to the camera (:324-363)
            cam_o, cam_d = batch["sky_origins"].reshape(-1, 3), batch["sky_directions"].reshape(-1, 3)
            exit_pts = ray_sphere_exit(cam_o, cam_d, self.ddf_radius)
            out["sky_ray_termination_dist"] = torch.norm(cam_o - exit_pts, dim=-1)
            rows_o.append(exit_pts)
            rows_d.append(-cam_d)
            n_sky = cam_o.shape[0]

        ddf_p = st.group("ddf")
        that = ddf_termination(st.ddf_cfg, torch.cat(rows_o).detach(), torch.cat(rows_d), ddf_p["position_encoding.hash_table"],
                               ddf_p["ddf.final_layer.weight"], ddf_p["ddf.final_layer.bias"], ddf_param_list(ddf_p))
        expected = that[:N]
        out["expected_termination_dist"] = expected
        if n_mv:
            out["multi_view_termintation_dist"] = batch["termination_dist"]              # (sic) key and value as in :321
            out["multi_view_expected_termination_dist"] = that[N:N + n_mv]
        if n_sky:
            out["sky_ray_expected_termination_dist"] = that[N + n_mv:N + n_mv + n_sky]

        if cfg.include_depth_loss_scene_center_weight and self.training:
            rad = torch.norm(positions if cfg.scene_center_weight_include_z else positions[..., :2], dim=-1) / self.ddf_radius
            out["distance_weight"] = 1.0 - rad ** cfg.scene_center_weight_exp            # :224-238

        if (inc.get("sdf_l1_loss") or inc.get("sdf_l2_loss")) and self.training:
            sdf_p = st.group("sdf")
            if stop_gradients:                                                           # :244-248
                with torch.no_grad():
                    term_pts = positions + directions * expected.unsqueeze(-1)
                    s, _, _ = sdf_field(st.sdf_cfg, term_pts, sdf_p["encoding.hash_table"], st.sdf_weights(), want_normals=False, want_albedo=False)
            else:
                term_pts = positions + directions * expected.unsqueeze(-1)
                s, _, _ = sdf_field(st.sdf_cfg, term_pts, sdf_p["encoding.hash_table"], st.sdf_weights(), want_normals=False, want_albedo=False)
            out["sdf_at_termination"] = s.reshape(-1, 1)
        return out

    # -- ddf_model.py:381-405 --------------------------------------------------------------------------------------
    def get_metrics_dict(self, outputs: Dict[str, Tensor], batch: Dict[str, Tensor]) -> Dict[str, Tensor]:
        """depth_psnr with data range (0, ddf_radius): torchmetrics PeakSignalNoiseRatio clamps both arguments into the range
        [from memory; torchmetrics is not in the image]."""
        m = batch["mask"]
        pred = (outputs["expected_termination_dist"].detach().unsqueeze(1) * m).clamp(0.0, self.ddf_radius)
        gt = (batch["termination_dist"].detach() * m).clamp(0.0, self.ddf_radius)
        mse = torch.mean((pred - gt) ** 2)
        return {"depth_psnr": 10.0 * torch.log10(self.ddf_radius ** 2 / mse)}

    # -- ddf_model.py:407-493 --------------------------------------------------------------------------------------
    def get_loss_dict(self, outputs: Dict[str, Tensor], batch: Dict[str, Tensor]) -> Dict[str, Tensor]:
        cfg = self.config
        inc = cfg.loss_inclusions
        L: Dict[str, Tensor] = {}
        mask = batch["mask"]
        if cfg.mask_to_circumference:
            expected = outputs["expected_termination_dist"].unsqueeze(1)
            gt = torch.where(mask == 0, torch.full_like(batch["termination_dist"], self.ddf_radius * 2), batch["termination_dist"])
        else:
            expected = outputs["expected_termination_dist"].unsqueeze(1) * mask
            gt = batch["termination_dist"] * mask
        inv_w = 1.0 / (gt + 1e-6) if cfg.inverse_depth_weight else 1.0
        for name, fn in (("depth_l1_loss", lambda a, b: (a - b).abs()), ("depth_l2_loss", lambda a, b: (a - b) ** 2)):
            if not inc.get(name):
                continue
            if cfg.include_depth_loss_scene_center_weight:
                L[name] = torch.mean(fn(expected, gt) * outputs["distance_weight"].unsqueeze(-1) * inv_w)
            else:
                L[name] = torch.mean(torch.mean(fn(expected, gt)) * inv_w)
        if inc.get("sdf_l1_loss"):
            L["sdf_l1_loss"] = (outputs["sdf_at_termination"] * mask).abs().mean()
        if inc.get("sdf_l2_loss"):
            L["sdf_l2_loss"] = ((outputs["sdf_at_termination"] * mask) ** 2).mean()
        if inc.get("multi_view_loss"):
            # [N] - [N,1] broadcasts to [N,N] in the reference (:478-483): every prediction is penalised against every ray's
            # ground-truth distance.  Reproduced as is.
            L["multi_view_loss"] = torch.mean(torch.relu(outputs["multi_view_expected_termination_dist"] - outputs["multi_view_termintation_dist"]) ** 2)
        if inc.get("sky_ray_loss"):
            L["sky_ray_loss"] = (outputs["sky_ray_expected_termination_dist"] - outputs["sky_ray_termination_dist"]).abs().mean()
        return {k: v * cfg.loss_coefficients[k] for k, v in L.items() if k in cfg.loss_coefficients}

    # -- neusky_pipeline.py:272-289 --------------------------------------------------------------------------------
    def forward(self, sky_origins: Optional[Tensor] = None, sky_directions: Optional[Tensor] = None,
                rays: Optional[Tuple[Tensor, Tensor]] = None, multi_view_points: Optional[Tensor] = None):
        """One fitting pass: (sum of losses, loss dict, outputs, batch).  Add the sum to the main step's loss before backward().
        `rays` / `multi_view_points`: the host draws of `draw_host()`, already on the device (no host work inside the pass)."""
        batch = self.generate_ddf_samples(sky_origins, sky_directions, rays=rays)
        outputs = self.get_outputs(batch, stop_gradients=self.stop_sdf_gradients, multi_view_points=multi_view_points)
        losses = self.get_loss_dict(outputs, batch)
        return sum(losses.values()), losses, outputs, batch

    __call__ = forward
