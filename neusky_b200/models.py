"""Drop-in model classes behind the reference's Model surface:

  DDFModel(config, ddf_radius, **kwargs)                          neusky/models/ddf_model.py:88-379
  NeuSkyFactoModel(config, scene_box, num_train_data, num_val_data, num_test_data, visibility_field, test_mode, **kwargs)
                                                                  neusky/models/neusky_model.py:172-931, 1360-1440, 1590-1778

Both are ``torch.nn.Module``s whose sub-modules and parameters carry the reference's names (``field.*``, ``proposal_networks.*``,
``illumination_field.*``, ``train_illumination_latents``, ``train_scale``, ``eval_illumination_latents``, ``eval_scale``,
``eval_rotation``, ``visibility_threshold``; ``field.*`` for the DDF model) and whose ``get_param_groups`` returns the optimizer
group names of the reference's method config (``fields``, ``proposal_networks``, ``illumination_field``, ``visibility_sigmoid``,
``ddf_field``: neusky_model.py:379-398, ddf_model.py:151-156, neusky_config.py:216-237).

Execution: ``forward(ray_bundle, batch=None, rotation=None, step=None)`` applies the sphere collider and
  * in eval mode renders through the fused kernels (neusky_b200/render.py: proposal sampling -> K2 -> K3 -> RENI++ -> K4 -> sRGB),
    mixed-camera batches included (one RENI++ table per distinct camera, per-ray camera rows into K4);
  * in training mode runs one differentiable iteration through neusky_b200/train.py (the same nn.Parameter objects, shared).
Out of scope: data managers, pipelines, optimizers / schedulers, viewer widgets, metrics, eval-latent fitting loop.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Dict, List, Literal, Optional, Tuple, Type, Union

import torch
from torch import nn
from torch.nn import Parameter

from . import ops, proposal as _proposal, samplers as _samplers
from .fields import (DirectionalDistanceField, DirectionalDistanceFieldConfig, InstantiateConfig, NeuSkyFieldHeadNames, RENIField, RENIFieldConfig,
                     RENIFieldHeadNames, SDFAlbedoField, SDFAlbedoFieldConfig)
from .rays import Frustums, RayBundle, RaySamples

Tensor = torch.Tensor


# =====================================================================================================================
# DDFModel
# =====================================================================================================================
@dataclass
class DDFModelConfig(InstantiateConfig):
    """ddf_model.py:54-85; defaults = the NeuSky method config (neusky_config.py:160-205)."""

    _target: Type = field(default_factory=lambda: DDFModel)
    ddf_field: DirectionalDistanceFieldConfig = field(default_factory=DirectionalDistanceFieldConfig)
    compute_normals: bool = False
    include_depth_loss_scene_center_weight: bool = True
    scene_center_weight_exp: float = 3.0
    scene_center_weight_include_z: bool = False
    mask_to_circumference: bool = False
    inverse_depth_weight: bool = False
    log_depth: bool = False
    loss_inclusions: Dict[str, bool] = field(default_factory=lambda: {
        "depth_l1_loss": True, "depth_l2_loss": False, "sdf_l1_loss": False, "sdf_l2_loss": True, "prob_hit_loss": False,
        "normal_loss": False, "multi_view_loss": True, "sky_ray_loss": True})
    loss_coefficients: Dict[str, float] = field(default_factory=lambda: {
        "depth_l1_loss": 1.0, "depth_l2_loss": 0.0, "sdf_l1_loss": 1.0, "sdf_l2_loss": 0.01, "prob_hit_loss": 0.01,
        "normal_loss": 1.0, "multi_view_loss": 0.01, "sky_ray_loss": 1.0})
    eval_num_rays_per_chunk: int = 1024


class DDFModel(nn.Module):
    """neusky/models/ddf_model.py:88-379 (the field, the local frame, get_outputs / forward, the optimizer group)."""

    def __init__(self, config: DDFModelConfig, ddf_radius: float, **kwargs) -> None:
        super().__init__()
        self.config = config
        self.ddf_radius = float(ddf_radius)
        self.kwargs = kwargs
        self.populate_modules()

    def populate_modules(self) -> None:
        self.field: DirectionalDistanceField = self.config.ddf_field.setup(ddf_radius=self.ddf_radius)        # :114

    def get_param_groups(self) -> Dict[str, List[Parameter]]:
        if self.field is None:
            raise ValueError("populate_fields() must be called before get_param_groups")                    # :153-154
        return {"ddf_field": list(self.field.parameters())}

    def get_localised_transforms(self, positions: Tensor) -> Tensor:
        """:158-181 -- columns (x, y, z) of the local frame at each sphere point: y = -q, x = norm(up x y), z = norm(y x x)."""
        up = torch.tensor([0.0, 0.0, 1.0]).type_as(positions).expand_as(positions)
        y = -positions
        x = torch.linalg.cross(up, y)
        x = x / x.norm(dim=-1, keepdim=True)
        z = torch.linalg.cross(y, x)
        z = z / z.norm(dim=-1, keepdim=True)
        return torch.stack((x, y, z), dim=-1)

    def get_outputs(self, ray_bundle, batch, neusky, stop_gradients: bool = True) -> Dict[str, Tensor]:
        """:183-369.  ray_bundle.origins on the DDF sphere, .directions in world space -> expected_termination_dist [N]
        (+ distance_weight, sdf_at_termination in training, as the reference)."""
        if self.field is None:
            raise ValueError("populate_fields() must be called before get_outputs")
        H = W = None
        if ray_bundle.origins.dim() in (3, 4):
            H, W = ray_bundle.origins.shape[:2]
        positions = ray_bundle.origins.reshape(-1, 3)
        directions = ray_bundle.directions.reshape(-1, 3)
        if torch.is_grad_enabled() and directions.requires_grad:
            # world directions that depend on another network (multi-view batch, :286-307): the differentiable row op carries d/d directions
            from . import train

            p = self.field.named_ddf_state()
            that = train.ddf_termination(self.field._cfg(), positions.contiguous(), directions.contiguous(), p["position_encoding.hash_table"],
                                         p["ddf.final_layer.weight"], p["ddf.final_layer.bias"], train.ddf_param_list(p))
        else:
            M = self.get_localised_transforms(positions)
            d_local = torch.einsum("ijl,ij->il", M, directions)                                              # :196-200
            rs = RaySamples(frustums=Frustums(origins=positions, directions=d_local, starts=torch.zeros_like(positions), ends=torch.zeros_like(positions),
                                              pixel_area=torch.ones_like(positions[..., 0])))
            that = self.field.forward(rs)[NeuSkyFieldHeadNames.TERMINATION_DISTANCE]                          # :217
        outputs: Dict[str, Tensor] = {"expected_termination_dist": that}
        c = self.config
        if c.include_depth_loss_scene_center_weight and self.training and batch is not None:
            dist = positions.norm(dim=-1) if c.scene_center_weight_include_z else positions[..., :2].norm(dim=-1)
            outputs["distance_weight"] = 1.0 - (dist / self.ddf_radius) ** c.scene_center_weight_exp          # :224-238
        if (c.loss_inclusions["sdf_l1_loss"] or c.loss_inclusions["sdf_l2_loss"]) and self.training:          # :241-254
            if neusky is not None:
                term_pts = positions + directions * that.unsqueeze(-1)
                if stop_gradients:
                    with torch.no_grad():
                        sdf_t = neusky.field.get_sdf_at_pos(term_pts).detach()
                else:
                    sdf_t = neusky.field.get_sdf_at_pos(term_pts)
                outputs["sdf_at_termination"] = sdf_t
            elif batch is not None and "sdf_at_termination" in batch:
                outputs["sdf_at_termination"] = batch["sdf_at_termination"]
        if H is not None:
            outputs = {k: v.reshape(H, W, 1, -1) for k, v in outputs.items()}
        return outputs

    def forward(self, ray_bundle, batch, neusky, stop_gradients: bool = True) -> Dict[str, Tensor]:
        """:371-379."""
        return self.get_outputs(ray_bundle, batch, neusky, stop_gradients=stop_gradients)


# =====================================================================================================================
# NeuSkyFactoModel
# =====================================================================================================================
@dataclass
class SceneBox:                        # nerfstudio.data.scene_box.SceneBox [NS-mem]: only .aabb is read
    aabb: Tensor = None


@dataclass
class IcosahedronSamplerConfig(InstantiateConfig):
    """ns_reni illumination_samplers.py:73-85."""

    _target: Type = field(default_factory=lambda: _samplers.IcosahedronSampler)
    num_directions: int = 512
    apply_random_rotation: bool = True
    remove_lower_hemisphere: bool = False

    def setup(self, **kwargs):
        return _samplers.IcosahedronSampler(self.num_directions, self.apply_random_rotation, self.remove_lower_hemisphere, **kwargs)


@dataclass
class NeuSkyFactoModelConfig(InstantiateConfig):
    """neusky_model.py:80-170 + the nerfstudio NeuS-facto fields the path reads [NS-mem A.6]; defaults = neusky_config.py:65-159."""

    _target: Type = field(default_factory=lambda: NeuSkyFactoModel)
    sdf_field: SDFAlbedoFieldConfig = field(default_factory=SDFAlbedoFieldConfig)
    illumination_field: RENIFieldConfig = field(default_factory=RENIFieldConfig)
    illumination_field_ckpt_path: Optional[str] = None          # decoder weights (.ckpt with "_model.field.*" keys, neusky_model.py:271-299)
    illumination_field_ckpt_step: int = 50000
    illumination_sampler: IcosahedronSamplerConfig = field(default_factory=IcosahedronSamplerConfig)
    only_upperhemisphere_visibility: bool = True
    lower_hermisphere_visibility: bool = True
    sdf_to_visibility_stop_gradients: Literal["none", "sdf", "depth", "both"] = "depth"
    fix_test_illumination_directions: bool = True
    use_visibility: bool = True
    fit_visibility_field: bool = True
    scene_contraction_order: Literal["Linf", "L2"] = "L2"
    collider_shape: Literal["sphere", "box"] = "sphere"
    render_ambient_light: bool = False
    eval_num_rays_per_chunk: int = 256
    # NeuS-facto sampling [NS-mem A.6]
    num_proposal_samples_per_ray: Tuple[int, ...] = (256, 96)
    num_neus_samples_per_ray: int = 48
    num_proposal_iterations: int = 2
    proposal_net_args_list: Tuple[Dict[str, int], ...] = ({"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 64},
                                                          {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 256})
    use_single_jitter: bool = True
    loss_inclusions: Dict[str, Any] = field(default_factory=lambda: {
        "rgb_l1_loss": True, "rgb_l2_loss": False, "cosine_colour_loss": False, "eikonal loss": True, "fg_mask_loss": True, "normal_loss": False,
        "depth_loss": False, "sdf_level_set_visibility_loss": True, "interlevel_loss": True, "sky_pixel_loss": {"enabled": True, "cosine_weight": 0.1},
        "hashgrid_density_loss": {"enabled": True, "grid_resolution": 10}, "ground_plane_loss": True,
        "visibility_sigmoid_loss": {"visibility_threshold_method": "learnable", "optimise_sigmoid_bias": True, "optimise_sigmoid_scale": False,
                                    "target_min_bias": 0.1, "target_max_scale": 25, "steps_until_min_bias": 50000}})
    # ours
    eval_tile: int = 16384              # rays per kernel pass in get_outputs_for_camera_ray_bundle (the reference's 256-ray chunks underfill a B200)
    k4_impl: str = "tc2"


class NeuSkyFactoModel(nn.Module):
    """The NeuSky model surface (neusky_model.py:172-931, 1360-1440, 1590-1778) over the fused kernels.  See the module docstring."""

    def __init__(self, config: NeuSkyFactoModelConfig, scene_box, num_train_data: int, num_val_data: int = 0, num_test_data: int = 0,
                 visibility_field: Optional[DDFModel] = None, test_mode: str = "val", **kwargs) -> None:
        super().__init__()
        self.config = config
        self.scene_box = scene_box
        self.num_train_data = num_train_data
        self.num_val_data, self.num_test_data, self.test_mode = num_val_data, num_test_data, test_mode
        self.num_eval_data = num_val_data if test_mode == "val" else num_test_data                          # :196
        self.fitting_eval_latents = False
        self.train_metadata, self.eval_metadata = kwargs.get("train_metadata"), kwargs.get("eval_metadata")
        self.kwargs = kwargs
        self.populate_modules()
        self.visibility_field = visibility_field                                                            # :203 (a DDFModel, trained by the pipeline)
        lv = config.loss_inclusions["visibility_sigmoid_loss"]
        self.visibility_threshold_method = lv["visibility_threshold_method"]
        self.sigmoid_scale = torch.tensor(float(lv["target_max_scale"]))                                     # :221-223
        if visibility_field is not None:
            self.ddf_radius = visibility_field.ddf_radius                                                   # :218
            if self.visibility_threshold_method == "learnable":
                if not (lv["optimise_sigmoid_bias"] or lv["optimise_sigmoid_scale"]):
                    raise AssertionError("Must optimise sigmoid bias or scale")                             # :227-230
                if lv["optimise_sigmoid_bias"]:
                    self.visibility_threshold = Parameter(torch.tensor(self.ddf_radius * 2.0))               # :234
                if lv["optimise_sigmoid_scale"]:
                    self.sigmoid_scale = Parameter(torch.tensor(1.0))
            elif self.visibility_threshold_method == "exponential_decay":
                self.visibility_threshold_start = torch.tensor(self.ddf_radius * 2.0)
                self.visibility_threshold_end = torch.tensor(float(lv["target_min_bias"]))
            elif self.visibility_threshold_method == "fixed":
                self.visibility_threshold = torch.tensor(float(lv["target_min_bias"]))
        # viewer state the reference snapshots per frame (:1389-1403); plain attributes here
        self.render_shadow_map_flag = False
        self.shadow_map_azimuth, self.shadow_map_elevation = 0.0, 45.0
        self.shadow_map_threshold, self.shadow_map_sigmoid_scale = 0.1, 25.0
        self.accumulation_mask_threshold = 0.0
        self._renderer = None
        self._renderer_key = None
        object.__setattr__(self, "_train_step", None)
        self._cos_anneal_ratio = 1.0

    # -- modules ---------------------------------------------------------------------------------------------------
    def populate_modules(self) -> None:
        """neusky_model.py:248-372 (+ NeuSFactoModel.populate_modules [NS-mem A.6]: field, proposal networks, proposal sampler)."""
        c = self.config
        aabb = self.scene_box.aabb if self.scene_box is not None else torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])
        self.field: SDFAlbedoField = c.sdf_field.setup(aabb=aabb, num_images=self.num_train_data, use_average_appearance_embedding=False, spatial_distortion=None)
        from . import init as nb_init

        nets = []
        for i in range(c.num_proposal_iterations):
            a = c.proposal_net_args_list[min(i, len(c.proposal_net_args_list) - 1)]
            p = nb_init.init_proposal_params(int(torch.randint(0, 2**31 - 1, (1,))), num_levels=a["num_levels"], log2_T=a["log2_hashmap_size"], hidden=a["hidden_dim"])
            nets.append(_proposal.HashMLPDensityField(p, max_res=a["max_res"], num_levels=a["num_levels"], log2_hashmap_size=a["log2_hashmap_size"], device="cpu"))
        self.proposal_networks = nn.ModuleList(nets)
        self.density_fns = [n.density_fn for n in self.proposal_networks]
        self.illumination_field: RENIField = c.illumination_field.setup(num_train_data=None, num_eval_data=None)      # :253-256: decoder only
        L = self.illumination_field.latent_dim
        self.eval_rotation = Parameter(torch.ones(max(self.num_eval_data, 0)))                                # :258
        self.train_illumination_latents = Parameter(torch.zeros((self.num_train_data, L, 3)))                 # :260-263
        self.train_scale = Parameter(torch.ones(self.num_train_data))
        self.eval_illumination_latents = Parameter(torch.zeros((max(self.num_eval_data, 0), L, 3)))
        self.eval_scale = Parameter(torch.ones(max(self.num_eval_data, 0)))
        if c.illumination_field_ckpt_path is not None:
            self.load_illumination_decoder(c.illumination_field_ckpt_path)
        self.illumination_sampler = c.illumination_sampler.setup()

    def load_illumination_decoder(self, ckpt_path) -> None:
        """:271-299 -- decoder weights from a ns_reni checkpoint: keys ``_model.field.*`` minus the per-image latent codes."""
        import os

        if not os.path.exists(str(ckpt_path)):
            raise ValueError(f"Could not find illumination field checkpoint at {ckpt_path}")                 # :283-284
        ckpt = torch.load(str(ckpt_path), weights_only=False, map_location="cpu")
        pre, ignore = "_model.field.", ("train_logvar", "eval_logvar", "train_mu", "eval_mu")
        sd = {k[len(pre):]: v for k, v in ckpt["pipeline"].items() if k.startswith(pre) and not any(i in k for i in ignore)}
        self.illumination_field.load_state_dict(sd, strict=False)

    def get_param_groups(self) -> Dict[str, List[Parameter]]:
        """:379-398."""
        g: Dict[str, List[Parameter]] = {"fields": list(self.field.parameters()), "proposal_networks": list(self.proposal_networks.parameters())}
        g["illumination_field"] = [self.train_illumination_latents, self.train_scale] if self.train_scale is not None else [self.train_illumination_latents]
        lv = self.config.loss_inclusions["visibility_sigmoid_loss"]
        if lv["visibility_threshold_method"] == "learnable":
            ps = []
            if lv["optimise_sigmoid_bias"]:
                ps.append(self.visibility_threshold)
            if lv["optimise_sigmoid_scale"]:
                ps.append(self.sigmoid_scale)
            g["visibility_sigmoid"] = ps
        return g

    def get_illumination_field(self) -> Tuple[Tensor, Tensor]:
        """:400-410 -- (latent codes [N_img, L, 3], scales [N_img]) of the current split."""
        if self.training and not self.fitting_eval_latents:
            return self.train_illumination_latents, self.train_scale
        return self.eval_illumination_latents, self.eval_scale

    def set_cos_anneal_ratio(self, anneal: float) -> None:
        self._cos_anneal_ratio = float(anneal)
        self.field.set_cos_anneal_ratio(anneal)

    @property
    def device(self) -> torch.device:
        return self.field.encoding.hash_table.device

    def _threshold(self, step: Optional[int]) -> Union[Tensor, float]:
        if self.visibility_threshold_method == "exponential_decay":
            return self.decay_threshold(step)
        return self.visibility_threshold

    def decay_threshold(self, step: Optional[int]) -> Tensor:
        """neusky_model.py:1572-1588: exponential decay from 2r to target_min_bias over steps_until_min_bias."""
        n = self.config.loss_inclusions["visibility_sigmoid_loss"]["steps_until_min_bias"]
        t = min(max(float(step or 0) / float(n), 0.0), 1.0)
        return self.visibility_threshold_start * (self.visibility_threshold_end / self.visibility_threshold_start) ** t

    # -- fused eval renderer over the modules' parameters ---------------------------------------------------------------
    def _get_renderer(self):
        from .fields import _versions
        from .render import RayRenderer

        if self.visibility_field is None:
            raise ValueError("NeuSkyFactoModel: use_visibility needs a visibility_field (DDFModel)")
        mods = [self.field, self.visibility_field.field, self.illumination_field, self.proposal_networks]
        key = _versions([p for m in mods for p in m.parameters()])
        if self._renderer is None or self._renderer_key != key:
            sdf_p = {n: p.detach() for n, p in self.field.named_parameters()}
            ddf_p = {n: p.detach() for n, p in self.visibility_field.field.named_parameters()}
            c = self.config
            r = RayRenderer(sdf_p, ddf_p, self.illumination_field.decoder_state(), device=self.device, log2_T=self.field.encoding.log2_T,
                            ddf_radius=self.visibility_field.ddf_radius, impl=c.k4_impl, sdf_impl=c.sdf_field.impl,
                            proposal_params=[{n: p.detach() for n, p in net.named_parameters()} for net in self.proposal_networks],
                            proposal_max_res=[net.max_res for net in self.proposal_networks], num_proposal_samples_per_ray=c.num_proposal_samples_per_ray,
                            proposal_log2_T=self.proposal_networks[0].log2_T, ddf_log2_T=self.visibility_field.field.position_encoding.log2_T)
            ga, gd = self.field.encoding.grid_args(), self.visibility_field.field.position_encoding.grid_args()      # imported tcnn grids, if any
            r.sdf_grid_meta, r.sdf_grid_smoothstep = ga["grid_meta"], ga["smoothstep"]
            r.shader.grid_meta, r.shader.grid_smoothstep = gd["grid_meta"], gd["smoothstep"]
            r.shader.only_upper = c.only_upperhemisphere_visibility
            r.shader.lower_vis = 1.0 if c.lower_hermisphere_visibility else 0.0
            self._renderer, self._renderer_key = r, key
        return self._renderer

    # -- reference API: forward ---------------------------------------------------------------------------------------
    def forward(self, ray_bundle, batch: Optional[Dict] = None, rotation: Optional[Tensor] = None, step: Optional[int] = None) -> Dict[str, Tensor]:
        """:425-443.  The collider is the unit sphere (collider_shape="sphere", :214); it is applied inside the render path."""
        if self.config.collider_shape != "sphere":
            raise NotImplementedError("only collider_shape='sphere' (the NeuSky method config) is implemented")
        return self.get_outputs(ray_bundle, batch=batch, rotation=rotation, step=step)

    def _illumination_directions(self) -> Tensor:
        if not self.training and self.config.fix_test_illumination_directions:
            smp = self.illumination_sampler(apply_random_rotation=False)                                     # :449-452
        else:
            smp = self.illumination_sampler()
        return smp.frustums.directions.to(self.device, torch.float32)

    def get_outputs(self, ray_bundle, batch: Optional[Dict] = None, rotation: Optional[Tensor] = None, step: Optional[int] = None) -> Dict[str, Any]:
        """:738-931."""
        if self.training and torch.is_grad_enabled():
            return self._train_outputs(ray_bundle, batch, rotation, step)
        o = ray_bundle.origins.reshape(-1, 3).contiguous()
        d = ray_bundle.directions.reshape(-1, 3).contiguous()
        dn = ray_bundle.metadata["directions_norm"].reshape(-1, 1).contiguous()
        cam_idx = ray_bundle.camera_indices.reshape(-1).long()
        r = self._get_renderer()
        r.set_directions(self._illumination_directions())
        lat, sc = self.get_illumination_field()
        uniq, inv = torch.unique(cam_idx, return_inverse=True)                                                # :461
        Z, s_k = lat.detach()[uniq], sc.detach()[uniq]
        rot = rotation
        if rot is not None and rot.dim() == 3:
            raise NotImplementedError("Batched rotation not implemented yet")                                # reni_illumination_field.py:520-521
        thr = float(self._threshold(step))
        out = r.render(o, d, dn, self.config.num_neus_samples_per_ray, Z, s_k, rotation=rot, threshold=thr, sigmoid_scale=float(self.sigmoid_scale),
                       cos_anneal_ratio=self._cos_anneal_ratio, cam=inv.to(torch.int32) if uniq.shape[0] > 1 else None,
                       want_visibility_batch=True, want_prop_depth=True)
        if self.render_shadow_map_flag:
            sm = r.shadow_map(o, d, out["p2p_dist"], out["accumulation"], self.shadow_map_azimuth, self.shadow_map_elevation, self.shadow_map_threshold,
                              self.shadow_map_sigmoid_scale, self.accumulation_mask_threshold)
            out["shadow_map"] = sm["visibility"]                                                             # :899-901
            out["shadow_map_difference"] = sm["difference"]
        out.pop("starts", None), out.pop("ends", None)
        return out

    def _train_outputs(self, ray_bundle, batch, rotation, step) -> Dict[str, Any]:
        """Training forward + the tensors get_loss_dict needs, through neusky_b200/train.py on the SAME nn.Parameter objects."""
        ts = self.train_step()
        ts.cos_anneal_ratio = self._cos_anneal_ratio
        ts.set_directions(self._illumination_directions())
        if rotation is not None:
            raise NotImplementedError("rotation is an eval-time option (fit / relight); the training forward takes none, like ns-train")
        return self._train_outputs_with(ts, ray_bundle, batch)

    def train_step(self):
        """The training-path engine (neusky_b200/train.py:NeuSkyTrainStep) over THIS model's nn.Parameter objects (built once per device)."""
        from . import train

        if self.visibility_field is None:
            raise ValueError("NeuSkyFactoModel: training needs a visibility_field (DDFModel)")
        if self._train_step is None or self._train_step.dev != self.device:
            c = self.config
            sdf_p = dict(self.field.named_parameters())
            sdf_p = {k: v for k, v in sdf_p.items() if k.startswith(("glin", "clin", "encoding.", "deviation_network."))}
            ts = train.NeuSkyTrainStep(
                sdf_p, dict(self.visibility_field.field.named_parameters()), self.illumination_field.decoder_state(), num_cameras=self.num_train_data,
                device=self.device, log2_T=self.field.encoding.log2_T, num_samples=c.num_neus_samples_per_ray, ddf_radius=self.visibility_field.ddf_radius,
                sigmoid_scale=float(self.sigmoid_scale), only_upper_hemisphere=c.only_upperhemisphere_visibility,
                lower_hemisphere_visibility=1.0 if c.lower_hermisphere_visibility else 0.0, num_proposal_samples_per_ray=c.num_proposal_samples_per_ray,
                share_params=True, latents=self.train_illumination_latents, scale=self.train_scale, visibility_threshold=self.visibility_threshold,
                proposal_fields=list(self.proposal_networks), ddf_log2_T=self.visibility_field.field.position_encoding.log2_T)
            object.__setattr__(self, "_train_step", ts)      # NOT a registered sub-module: its parameters are ours already (state_dict stays the reference's)
        return self._train_step

    def graphed_iteration(self, reducer, optimizer, fit=None, **kwargs):
        """One whole training iteration of this model (zero-fill, forward, optional DDF fitting pass, backward, gradient all-reduce,
        optimizer step) replayed as a CUDA graph: neusky_b200/graphed.py.  `reducer`: parallel.GradBucketReducer over
        `list(model.train_step().parameters())`; `optimizer`: anything with `.step()` over the same parameters.  For loops that own
        the iteration (nerfstudio's Trainer calls forward / backward / step separately and stays on the eager path)."""
        from .graphed import GraphedTrainIteration

        ts = self.train_step()
        ts.cos_anneal_ratio = self._cos_anneal_ratio
        return GraphedTrainIteration(ts, reducer, optimizer, fit=fit, **kwargs)

    def _train_outputs_with(self, ts, ray_bundle, batch) -> Dict[str, Any]:
        b = {"origins": ray_bundle.origins.reshape(-1, 3).contiguous(), "directions": ray_bundle.directions.reshape(-1, 3).contiguous(),
             "dnorm": ray_bundle.metadata["directions_norm"].reshape(-1, 1).contiguous(), "cam": ray_bundle.camera_indices.reshape(-1)}
        grid = None
        lh = self.config.loss_inclusions["hashgrid_density_loss"]
        if lh["enabled"]:                                                                                    # :674-734
            n = lh["grid_resolution"]
            ts.grid_resolution = n
            lo, hi = self.scene_box.aabb[0].to(self.device), self.scene_box.aabb[1].to(self.device)
            ax = [torch.linspace(float(lo[i]), float(hi[i]), n, device=self.device) for i in range(3)]
            pos = torch.stack(torch.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
            gap = (hi - lo) / n
            pos = pos + torch.rand_like(pos) * gap - gap / 2
            dirs = torch.nn.functional.normalize(torch.randn_like(pos), dim=-1)
            grid = (pos.contiguous(), dirs.contiguous())
        if batch is not None:
            b.update({k: batch[k] for k in ("image", "fg", "ground", "sky") if k in batch})
        have_targets = all(k in b for k in ("image", "fg", "ground", "sky"))
        if not have_targets:      # forward only: placeholder targets, the loss dict is not returned
            R = b["origins"].shape[0]
            z = torch.zeros(R, device=self.device)
            b.update({"image": torch.zeros(R, 3, device=self.device), "fg": z, "ground": z, "sky": z})
        loss, losses, out = ts(b, grid_positions=None if grid is None else grid[0], grid_dirs=None if grid is None else grid[1])
        out["weights"] = out["weights"][..., None]
        out["accumulation"] = out["accumulation"][:, None]
        out["p2p_dist"] = out["p2p_dist"][:, None]
        out["depth"] = out["p2p_dist"] / b["dnorm"]
        out["directions_norm"] = b["dnorm"]
        out["normal_vis"] = (out["normal"] + 1.0) / 2.0
        out["visibility_batch"] = {"termination_dist": None, "mask": None, "sdf_at_termination": out["sdf_at_termination"]}
        if have_targets:
            out["loss_dict"] = losses
        return out

    def get_loss_dict(self, outputs: Dict[str, Any], batch: Dict[str, Any], metrics_dict=None) -> Dict[str, Tensor]:
        """:935-1031 (training branch): computed by the training forward on the same batch (``outputs["loss_dict"]``)."""
        if "loss_dict" not in outputs:
            raise ValueError("get_loss_dict: run forward(ray_bundle, batch=...) in training mode with the target tensors (image, fg, ground, sky)")
        return outputs["loss_dict"]

    # -- reference API: the pieces other components call ---------------------------------------------------------------
    def sample_illumination(self, ray_samples, rotation: Optional[Tensor] = None):
        """:445-551 -> (hdr_illumination_colours [R*S, D, 3], illumination_directions [R*S, D, 3], hdr_background_colours [R, 3]).
        The two big tensors are returned as stride-0 EXPANDED VIEWS of the [K, D, 3] radiance table / the [D, 3] direction set
        whenever the batch has one camera (nothing is materialised); the fused path never calls this."""
        cam = ray_samples.camera_indices.reshape(ray_samples.frustums.origins.shape[0], -1)
        R, S = cam.shape[0], ray_samples.frustums.origins.shape[1] if ray_samples.frustums.origins.dim() == 3 else 1
        lat, sc = self.get_illumination_field()
        dirs = self._illumination_directions()
        D = dirs.shape[0]
        uniq, inv = torch.unique(cam[:, 0], return_inverse=True)
        table = self.illumination_field.radiance_table(dirs, lat[uniq].contiguous(), sc[uniq].contiguous(), rotation)       # [K,D,3], unnormalised
        if uniq.shape[0] == 1:
            colours = table.expand(R * S, D, 3)
        else:
            colours = table[inv][:, None].expand(R, S, D, 3).reshape(R * S, D, 3)
        directions = dirs[None].expand(R * S, D, 3)
        d0 = ray_samples.frustums.directions.reshape(R, -1, 3)[:, 0].contiguous()
        rs0 = RaySamples(frustums=Frustums(origins=None, directions=d0), camera_indices=cam[:, 0])
        bg = self.illumination_field.forward(rs0, rotation=rotation, latent_codes=lat[cam[:, 0].long()], scale=sc[cam[:, 0].long()])[RENIFieldHeadNames.RGB]
        return colours, directions, self.illumination_field.unnormalise(bg)

    def ray_sphere_intersection(self, positions: Tensor, directions: Tensor, radius: float) -> Tensor:
        """:1590-1622."""
        d = directions / directions.norm(dim=-1, keepdim=True)
        b = 2.0 * (d * positions).sum(-1)
        c = (positions * positions).sum(-1) - radius**2
        disc = torch.clamp(b * b - 4.0 * c, min=0.0)
        t = torch.maximum((-b - torch.sqrt(disc)) / 2.0, (-b + torch.sqrt(disc)) / 2.0)
        return positions + t[:, None] * d

    @torch.no_grad()
    def compute_visibility(self, ray_samples, depth: Tensor, illumination_directions: Tensor, threshold_distance, sigmoid_scale,
                           compute_shadow_map: bool = False) -> Dict[str, Any]:
        """:1624-1778 with the reference's arguments: ray_samples [R,S], depth [R,1] (the p2p distance),
        illumination_directions [R*S, D, 3] (row 0 is read, :1648) or [D,3] -> {"visibility" [R*S, D, 1] (a stride-0 expanded view over
        the samples, :1755-1759), "expected_termination_dist" [R*D'], "visibility_batch": {...}, "difference" if compute_shadow_map}.
        Forward only (the fused K4 kernel); the differentiable visibility lives in the training forward."""
        from .render import SkyShader

        dirs = illumination_directions[0] if illumination_directions.dim() == 3 else illumination_directions
        fo = ray_samples.frustums.origins
        R, S = fo.shape[0], (fo.shape[1] if fo.dim() == 3 else 1)
        o = fo.reshape(R, S, 3)[:, 0].contiguous()
        d = ray_samples.frustums.directions.reshape(R, S, 3)[:, 0].contiguous()
        D = dirs.shape[0]
        sh: SkyShader = self._get_renderer().shader
        saved = (sh.dirs, sh.mask, sh.mask_u8, sh.dirs_sel, sh.sel_index) if hasattr(sh, "dirs") else None
        try:
            sh.set_directions(dirs)
            pts = ops.surface_points(o, d, depth.reshape(R).contiguous(), sh.radius)
            zero = torch.zeros((R, 1, 3), device=pts.device)
            out = sh.shade(pts, zero, zero, torch.zeros((1, D, 3), device=pts.device), want_vis=True, want_ddf=True,
                           threshold=float(threshold_distance), sigmoid_scale=float(sigmoid_scale))
        finally:
            if saved is not None:
                sh.dirs, sh.mask, sh.mask_u8, sh.dirs_sel, sh.sel_index = saved
        term = out["termination_dist"]
        vd: Dict[str, Any] = {
            "visibility": out["visibility"][:, None, :].expand(R, S, D).reshape(R * S, D, 1),
            "expected_termination_dist": out["expected_termination_dist"],
            "visibility_batch": {"termination_dist": term, "mask": torch.ones_like(term), "sdf_at_termination": None},      # :1766-1776
        }
        if compute_shadow_map:
            vd["difference"] = torch.clamp(term, max=2.0 * sh.radius) - out["expected_termination_dist"]     # :1724-1727, :1764-1765
        return vd

    @torch.no_grad()
    def get_outputs_for_camera_ray_bundle(self, camera_ray_bundle, show_progress: bool = False, rotation: Optional[Tensor] = None, to_cpu: bool = False,
                                          step: Optional[int] = None) -> Dict[str, Tensor]:
        """:1360-1440 -- a whole camera in chunks (ours: config.eval_tile rays per pass instead of the reference's 256), outputs
        concatenated and reshaped to [H, W, C].  Non-tensor outputs are dropped like the reference does (:1420-1423)."""
        H, W = camera_ray_bundle.origins.shape[:2]
        n = H * W
        flat = RayBundle(origins=camera_ray_bundle.origins.reshape(n, 3), directions=camera_ray_bundle.directions.reshape(n, 3),
                         camera_indices=camera_ray_bundle.camera_indices.reshape(n, -1),
                         metadata={k: v.reshape(n, -1) for k, v in camera_ray_bundle.metadata.items()})
        lists: Dict[str, List[Tensor]] = {}
        was_training = self.training
        self.eval()
        try:
            for a in range(0, n, self.config.eval_tile):
                out = self.forward(flat.get_row_major_sliced_ray_bundle(a, a + self.config.eval_tile), rotation=rotation, step=step)
                for k, v in out.items():
                    if torch.is_tensor(v):
                        lists.setdefault(k, []).append(v.cpu() if to_cpu else v)
        finally:
            self.train(was_training)
        return {k: torch.cat(v).view(H, W, -1) for k, v in lists.items() if sum(t.shape[0] for t in v) == n}
