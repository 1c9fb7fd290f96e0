"""Drop-in ``torch.nn.Module`` fields with the reference's constructor / call signatures and ``state_dict`` names
(SURVEY.md 8b), the C-ABI kernels behind them:

  SDFAlbedoField(config, aabb, num_images, use_average_appearance_embedding=False, spatial_distortion=None)
        neusky/fields/sdf_albedo_field.py:80-282 (+ nerfstudio SDFField.forward_geonetwork / get_alpha)
  DirectionalDistanceField(config, ddf_radius=1.0)           neusky/fields/directional_distance_field.py:96-315
  RENIField(config, num_train_data=None, num_eval_data=None, normalisations=None)
        ns_reni/reni/illumination_fields/reni_illumination_field.py:90-593 (+ base_spherical_field.py:143-154)

Each ``*Config`` is a dataclass with the reference's field names and a ``_target`` / ``setup(**kwargs)`` pair like
nerfstudio's ``InstantiateConfig``, so ``config.sdf_field.setup(aabb=..., num_images=...)`` builds ours when the method
config points ``_target`` here (neusky_b200/neusky_config.py).  Parameters are ``nn.Parameter``s under the names the
reference's modules register, so ``ours.load_state_dict(ref.state_dict(), strict=True)`` round-trips
(tests/test_dropin_state_dict.py builds the reference modules through oracle/ref_shim and checks exactly that).

Two execution paths per module, chosen per call:
  * autograd off (eval, viewer, ``torch.no_grad()``): the fused forward kernels (K2 tcgen05 / SIMT, RENI decode, ...);
  * autograd on and a parameter requires grad: the differentiable layer-wise ops of neusky_b200/train.py.
There is no CPU path: CPU tensors raise.  Kernel-side weight blobs are re-packed (on the host, one upload) when a
parameter's version counter changes, i.e. after an optimizer step or a ``load_state_dict``.
"""
from __future__ import annotations

import contextlib
import math
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Dict, List, Literal, Optional, Tuple, Type, Union

import torch
from torch import nn

from . import init as nb_init
from . import ops, packing
from .init import hash_scalings

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------- output keys
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


def _rekey(out: Dict[Enum, Tensor], *enums) -> Dict[Any, Tensor]:
    """Re-key an output dict to another package's enums of equal ``.value`` (nerfstudio's FieldHeadNames when installed)."""
    table = {m.value: m for e in enums for m in e}
    return {table.get(k.value, k): v for k, v in out.items()}


# ----------------------------------------------------------------------------------------------- configs
@dataclass
class InstantiateConfig:               # nerfstudio.configs.base_config.InstantiateConfig [NS-mem]
    _target: Type = None

    def setup(self, **kwargs) -> Any:
        return self._target(self, **kwargs)


@dataclass
class SDFAlbedoFieldConfig(InstantiateConfig):
    """nerfstudio SDFFieldConfig [NS-mem A.4] with the NeuSky overrides as defaults (neusky/configs/neusky_config.py:66-77) +
    ``predict_shininess`` (sdf_albedo_field.py:71-77).  ``impl``: eval kernel, "tc" (tcgen05 fp16 operands) or "simt" (fp32)."""

    _target: Type = field(default_factory=lambda: SDFAlbedoField)
    num_layers: int = 2
    hidden_dim: int = 256
    geo_feat_dim: int = 256
    num_layers_color: int = 2
    hidden_dim_color: int = 256
    appearance_embedding_dim: int = 32
    use_appearance_embedding: bool = False
    bias: float = 0.1
    geometric_init: bool = True
    inside_outside: bool = False
    weight_norm: bool = True
    use_grid_feature: bool = True
    divide_factor: float = 2.0
    beta_init: float = 0.1
    encoding_type: Literal["hash", "periodic", "tensorf_vm"] = "hash"
    num_levels: int = 16
    max_res: int = 2048
    base_res: int = 16
    log2_hashmap_size: int = 19
    features_per_level: int = 2
    use_hash: bool = True
    smoothstep: bool = True
    predict_shininess: bool = False
    impl: str = "tc"


@dataclass
class DirectionalDistanceFieldConfig(InstantiateConfig):
    """directional_distance_field.py:47-93; defaults = the NeuSky method config (neusky_config.py:162-177)."""

    _target: Type = field(default_factory=lambda: DirectionalDistanceField)
    position_encoding_type: Literal["hash", "nerf", "sh", "icosphere_hash", "none"] = "hash"
    direction_encoding_type: Literal["hash", "nerf", "sh", "icosphere_hash", "none"] = "nerf"
    conditioning: Literal["FiLM", "Concat", "Attention"] = "FiLM"
    termination_output_activation: Literal["sigmoid", "tanh", "relu"] = "sigmoid"
    probability_of_hit_output_activation: Literal["sigmoid", "tanh", "relu"] = "sigmoid"
    hidden_layers: int = 5
    hidden_features: int = 256
    mapping_layers: int = 5
    mapping_features: int = 256
    num_attention_heads: int = 8
    num_attention_layers: int = 6
    out_features: int = 3
    last_layer_linear: bool = True
    first_omega_0: float = 30.0
    hidden_omega_0: float = 30.0
    predict_probability_of_hit: bool = False
    ddf_type: Literal["ddf", "pddf"] = "ddf"
    num_dirac_components: int = 2
    eta_T: float = 1.0
    epsilon_s: float = 1e-5
    split: int = 3          # ours: GEMM precision of the row-wise path (1 = tf32, 3 = 3xTF32, fp32-accurate)


@dataclass
class RENIFieldConfig(InstantiateConfig):
    """reni_illumination_field.py:39-87; defaults = the NeuSky method config (neusky_config.py:78-96)."""

    _target: Type = field(default_factory=lambda: RENIField)
    conditioning: Literal["FiLM", "Concat", "Attention"] = "Attention"
    invariant_function: Literal["GramMatrix", "VN"] = "VN"
    equivariance: Literal["None", "SO2", "SO3"] = "SO2"
    axis_of_invariance: Literal["x", "y", "z"] = "z"
    positional_encoding: Literal["None", "NeRF"] = "NeRF"
    encoded_input: Literal["None", "Directions", "Conditioning", "Both"] = "Directions"
    latent_dim: int = 100
    hidden_layers: int = 9
    hidden_features: int = 128
    mapping_layers: int = 5
    mapping_features: int = 128
    num_attention_heads: int = 8
    num_attention_layers: int = 6
    out_features: int = 3
    last_layer_linear: bool = True
    output_activation: Literal["sigmoid", "tanh", "relu", "exp", "None"] = "None"
    first_omega_0: float = 30.0
    hidden_omega_0: float = 30.0
    fixed_decoder: bool = True
    trainable_scale: Union[bool, Literal["train", "eval", "both"]] = True
    old_implementation: bool = False
    view_train_latents: bool = False


# ----------------------------------------------------------------------------------------------- small building blocks
class _HashGrid(nn.Module):
    """The multiresolution hash table as ONE fp32 parameter ``hash_table`` [L*T, F] (nerfstudio HashEncoding layout).
    The reference builds a ``tcnn.Encoding`` here (fp16 ``params``); reference checkpoints go through
    neusky_b200/tcnn_import.py (see ``_load_from_state_dict`` of the owning fields)."""

    def __init__(self, num_levels: int = 16, log2_hashmap_size: int = 19, features_per_level: int = 2, base_res: int = 16, max_res: int = 2048):
        super().__init__()
        self.num_levels, self.log2_T, self.features = num_levels, log2_hashmap_size, features_per_level
        self.n_output_dims = num_levels * features_per_level
        self.base_res, self.max_res = base_res, max_res
        self.hash_table = nn.Parameter((torch.rand((num_levels << log2_hashmap_size, features_per_level)) * 2 - 1) * 1e-3)
        self.register_buffer("scalings", hash_scalings(num_levels, base_res, max_res), persistent=False)
        self.tcnn_levels = None         # set by tcnn_import when the table came from a tiny-cuda-nn checkpoint
        self.tcnn_smoothstep = True
        self._tcnn_meta = None

    def require_native(self, who: str) -> None:
        """Training paths only: the differentiable kernels (hash-encode backward, train.py) implement the nerfstudio torch grid."""
        if self.tcnn_levels is not None:
            raise NotImplementedError(f"{who}: this hash grid was imported from a tiny-cuda-nn checkpoint; it renders through the eval kernels "
                                      "(K1 / K2 / K4 take the imported grid), but training from it is not supported (DESIGN.md section 6).")

    def grid_args(self) -> Dict[str, Any]:
        """kwargs for ops.sdf_field / ops.sky_shade: the per-level table of an imported tiny-cuda-nn grid (None = nerfstudio torch grid)."""
        if self.tcnn_levels is None:
            return {"grid_meta": None, "smoothstep": True}
        from .tcnn_import import tcnn_level_meta

        if self._tcnn_meta is None or self._tcnn_meta.device != self.hash_table.device:
            self._tcnn_meta = tcnn_level_meta(self.tcnn_levels, self.hash_table.device)
        return {"grid_meta": self._tcnn_meta, "smoothstep": bool(self.tcnn_smoothstep)}

    def forward(self, x: Tensor) -> Tensor:
        from . import autograd as nba

        if self.tcnn_levels is not None:
            from .tcnn_import import tcnn_level_meta

            if self._tcnn_meta is None or self._tcnn_meta.device != self.hash_table.device:
                self._tcnn_meta = tcnn_level_meta(self.tcnn_levels, self.hash_table.device)
            return ops.hash_encode_tcnn(x, self.hash_table.detach(), self._tcnn_meta, self.log2_T, self.tcnn_smoothstep)
        if torch.is_grad_enabled() and (self.hash_table.requires_grad or x.requires_grad):
            return nba.hash_encode(x.reshape(-1, 3).contiguous(), self.hash_table, self.scalings, self.log2_T).reshape(*x.shape[:-1], -1)
        return ops.hash_encode(x, self.hash_table.detach(), self.scalings, self.log2_T)


class _WNLinear(nn.Module):
    """``nn.utils.weight_norm(nn.Linear(in, out))`` parameter layout: ``weight_g`` [out,1], ``weight_v`` [out,in], ``bias`` [out]."""

    def __init__(self, in_f: int, out_f: int):
        super().__init__()
        self.weight_g = nn.Parameter(torch.ones(out_f, 1))
        self.weight_v = nn.Parameter(torch.zeros(out_f, in_f))
        self.bias = nn.Parameter(torch.zeros(out_f))

    @property
    def weight(self) -> Tensor:
        return self.weight_v * (self.weight_g / self.weight_v.norm(dim=1, keepdim=True))


class LearnedVariance(nn.Module):
    """nerfstudio LearnedVariance [SURVEY A.4]: ``variance`` parameter; get_variance() = exp(10 * variance).clip(1e-6, 1e6)."""

    def __init__(self, init_val: float):
        super().__init__()
        self.register_parameter("variance", nn.Parameter(init_val * torch.ones(1), requires_grad=True))

    def forward(self, x: Tensor) -> Tensor:
        return torch.ones([len(x), 1], device=x.device) * torch.exp(self.variance * 10.0)

    def get_variance(self) -> Tensor:
        return torch.exp(self.variance * 10.0).clip(1e-6, 1e6)


class _Embedding(nn.Module):
    """nerfstudio Embedding: ``embedding`` = nn.Embedding(in_dim, out_dim) (the appearance embedding the reference creates but
    never reads on this path, sdf_albedo_field.py:110)."""

    def __init__(self, in_dim: int, out_dim: int):
        super().__init__()
        self.embedding = nn.Embedding(in_dim, out_dim)

    def forward(self, idx: Tensor) -> Tensor:
        return self.embedding(idx)


def _versions(params) -> tuple:
    return tuple((p.data_ptr(), p._version) for p in params)


def _need_autograd(params) -> bool:
    return torch.is_grad_enabled() and any(p.requires_grad for p in params)


def _load_scalar_compat(state_dict, key: str, like: Tensor) -> None:
    """Accept a 0-dim tensor where the module holds a 1-element one (and the reverse)."""
    v = state_dict.get(key)
    if isinstance(v, torch.Tensor) and v.numel() == like.numel() and v.shape != like.shape:
        state_dict[key] = v.reshape(like.shape)


# =====================================================================================================================
# SDFAlbedoField
# =====================================================================================================================
class SDFAlbedoField(nn.Module):
    """neusky/fields/sdf_albedo_field.py:80-282.  NeuSky shape only (71 -> 256 -> 256 -> 257 softplus(beta=100) geometry network,
    295 -> 256 -> 256 -> 3 colour network, hash grid 16 x 2^19 x 2): other configurations raise NotImplementedError."""

    def __init__(self, config: SDFAlbedoFieldConfig, aabb: Tensor, num_images: int, use_average_appearance_embedding: bool = False,
                 spatial_distortion=None) -> None:
        super().__init__()
        c = config
        if (c.num_layers, c.hidden_dim, c.geo_feat_dim, c.num_layers_color, c.hidden_dim_color) != (2, 256, 256, 2, 256) or c.encoding_type != "hash" \
                or not c.use_grid_feature or not c.weight_norm or c.num_levels != 16 or c.features_per_level != 2 or not c.use_hash:
            raise NotImplementedError("SDFAlbedoField: the kernels implement the NeuSky field shape (neusky_config.py:66-77): 2 x 256 geometry layers, "
                                      "256 geometry features, 2 x 256 colour layers, 16-level x 2-feature hash grid, weight_norm")
        if c.predict_shininess:
            raise NotImplementedError("predict_shininess=True (Blinn-Phong branch) is served by neusky_b200.shaders, not by this field")
        self.config = c
        self.aabb = nn.Parameter(torch.as_tensor(aabb, dtype=torch.float32).clone(), requires_grad=False)      # :104
        self.spatial_distortion = spatial_distortion
        self.num_images = num_images
        self.embedding_appearance = _Embedding(num_images, c.appearance_embedding_dim)                      # :110 (unused on the path)
        self.use_average_appearance_embedding = use_average_appearance_embedding
        self.use_grid_feature, self.divide_factor = c.use_grid_feature, c.divide_factor
        self.encoding = _HashGrid(c.num_levels, c.log2_hashmap_size, c.features_per_level, c.base_res, c.max_res)   # :117-130
        self.encoding.tcnn_smoothstep = bool(c.smoothstep)      # nerfstudio SDFField passes "interpolation": "Smoothstep" to tcnn when config.smoothstep
        in_dim = 3 + 36 + self.encoding.n_output_dims
        dims = [in_dim] + [c.hidden_dim] * c.num_layers + [1 + c.geo_feat_dim]
        self.num_layers = len(dims)
        for l in range(self.num_layers - 1):
            setattr(self, f"glin{l}", _WNLinear(dims[l], dims[l + 1]))
        self.deviation_network = LearnedVariance(init_val=c.beta_init)                                      # :146
        cdims = [3 + 36 + c.geo_feat_dim] + [c.hidden_dim_color] * c.num_layers_color + [3]
        self.num_layers_color = len(cdims)
        for l in range(self.num_layers_color - 1):
            setattr(self, f"clin{l}", _WNLinear(cdims[l], cdims[l + 1]))
        self._cos_anneal_ratio = 1.0
        self._blob_key: Dict[str, tuple] = {}
        self._blobs: Dict[str, Tensor] = {}
        self._w_key = None
        self.reset_parameters()

    # -- initialisation: geometric init of nerfstudio SDFField.initialize_geo_layers [A.4] + torch Linear default for the colour net
    def reset_parameters(self) -> None:
        c = self.config
        seed = int(torch.randint(0, 2**31 - 1, (1,)))
        p = nb_init.init_sdf_params(seed, c.hidden_dim, c.geo_feat_dim, c.num_layers, c.num_layers_color, c.bias, c.inside_outside, c.num_levels,
                                    c.log2_hashmap_size, c.features_per_level, c.beta_init)
        with torch.no_grad():
            for k, v in p.items():
                tgt = self.get_parameter(k)
                tgt.copy_(v.reshape(tgt.shape))

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        _load_scalar_compat(state_dict, prefix + "deviation_network.variance", self.deviation_network.variance)
        from .tcnn_import import convert_tcnn_state_dict_entry

        convert_tcnn_state_dict_entry(state_dict, prefix + "encoding.", self.encoding)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    # -- kernel-side views of the parameters ------------------------------------------------------------------------
    def _mlp_params(self) -> List[Tensor]:
        return [p for n, p in self.named_parameters() if n.startswith(("glin", "clin"))]

    def _state(self) -> Dict[str, Tensor]:
        return {n: p.detach() for n, p in self.named_parameters() if n.startswith(("glin", "clin"))}

    def _blob(self, impl: str) -> Tensor:
        key = _versions(self._mlp_params())
        if self._blob_key.get(impl) != key:
            dev = self.encoding.hash_table.device
            self._blobs[impl] = (packing.pack_sdf_tc if impl == "tc" else packing.pack_sdf_simt)(self._state(), device=dev)
            self._blob_key[impl] = key
        return self._blobs[impl]

    def _train_weights(self):
        from . import train

        return train.sdf_param_list({n: p for n, p in self.named_parameters() if n.startswith(("glin", "clin"))})

    def _train_cfg(self):
        from . import train

        return train.SDFConfig(scalings=self.encoding.scalings, log2_T=self.encoding.log2_T)

    # -- reference API ---------------------------------------------------------------------------------------------
    def set_cos_anneal_ratio(self, anneal: float) -> None:
        self._cos_anneal_ratio = float(anneal)

    def forward_geonetwork(self, inputs: Tensor) -> Tensor:
        """x [N,3] -> [N, 1 + geo_feat_dim] (sdf | geometry feature), nerfstudio SDFField.forward_geonetwork [A.4]: exact fp32 kernel."""
        x = inputs.reshape(-1, 3)
        f = ops.sdf_field(x, self._blob("simt"), self.encoding.hash_table.detach(), self.encoding.scalings, self.encoding.log2_T,
                          want_grad=False, want_albedo=False, want_geo=True, impl="simt", **self.encoding.grid_args())
        return torch.cat([f["sdf"], f["geo"]], dim=-1)

    def get_sdf_at_pos(self, positions: Tensor) -> Tensor:
        """:169-174 -> [N,1].  Differentiable (w.r.t. positions to first order, the hash table and the weights) when autograd is on:
        DDFModel's sdf_at_termination branch trains through it (ddf_model.py:241-251)."""
        x = positions.reshape(-1, 3)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            from . import train

            self.encoding.require_native("SDFAlbedoField.get_sdf_at_pos (differentiable)")
            sdf, _, _ = train.sdf_field(self._train_cfg(), x.contiguous(), self.encoding.hash_table, self._train_weights(), want_normals=False, want_albedo=False)
            return sdf[:, None]
        return ops.sdf_field(x, self._blob("simt"), self.encoding.hash_table.detach(), self.encoding.scalings, self.encoding.log2_T,
                             want_grad=False, want_albedo=False, impl="simt", **self.encoding.grid_args())["sdf"]

    def _field(self, x: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """x [...,3] -> (sdf [...,1], gradient [...,3], albedo [...,3])."""
        lead = x.shape[:-1]
        if _need_autograd(list(self.parameters())):
            from . import train

            self.encoding.require_native("SDFAlbedoField (training)")
            sdf, grad, alb = train.sdf_field(self._train_cfg(), x.reshape(-1, 3).contiguous(), self.encoding.hash_table, self._train_weights())
            return sdf.reshape(*lead, 1), grad.reshape(*lead, 3), alb.reshape(*lead, 3)
        impl = self.config.impl
        f = ops.sdf_field(x, self._blob(impl), self.encoding.hash_table.detach(), self.encoding.scalings, self.encoding.log2_T, impl=impl,
                          **self.encoding.grid_args())
        return f["sdf"], f["gradient"], f["albedo"]

    def get_alpha(self, ray_samples, sdf: Optional[Tensor] = None, gradients: Optional[Tensor] = None) -> Tensor:
        """nerfstudio SDFField.get_alpha [SURVEY A.5] (called at :266 and, without sdf, at neusky_model.py:732)."""
        if sdf is None or gradients is None:
            x = ray_samples.frustums.origins + ray_samples.frustums.directions * ray_samples.frustums.starts
            sdf, gradients, _ = self._field(x)
        inv_s = self.deviation_network.get_variance()
        true_cos = (ray_samples.frustums.directions * gradients).sum(-1, keepdim=True)
        rho = self._cos_anneal_ratio
        iter_cos = -(torch.relu(-true_cos * 0.5 + 0.5) * (1.0 - rho) + torch.relu(-true_cos) * rho)
        nxt = sdf + iter_cos * ray_samples.deltas * 0.5
        prv = sdf - iter_cos * ray_samples.deltas * 0.5
        prev_cdf, next_cdf = torch.sigmoid(prv * inv_s), torch.sigmoid(nxt * inv_s)
        return ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).clip(0.0, 1.0)

    def get_outputs(self, ray_samples, density_embedding: Optional[Tensor] = None, return_alphas: bool = False) -> Dict[Enum, Tensor]:
        """:211-269.  Output shapes follow the reference: [*batch, 3] / [*batch, 1]; the gradient is analytic (one reverse pass
        inside the kernel) instead of torch.autograd.grad (:235-238)."""
        if ray_samples.camera_indices is None:
            raise AttributeError("Camera indices are not provided.")       # :218-219
        x = ray_samples.frustums.origins + ray_samples.frustums.directions * ray_samples.frustums.starts   # get_start_positions (:225)
        sdf, grad, alb = self._field(x)
        out = {
            NeuSkyFieldHeadNames.ALBEDO: alb,
            FieldHeadNames.SDF: sdf,
            FieldHeadNames.NORMALS: torch.nn.functional.normalize(grad, p=2, dim=-1),     # :251
            FieldHeadNames.GRADIENT: grad,
        }
        if return_alphas:
            out[FieldHeadNames.ALPHA] = self.get_alpha(ray_samples, sdf, grad)             # :266
        return out

    def forward(self, ray_samples, compute_normals: bool = False, return_alphas: bool = False) -> Dict[Enum, Tensor]:
        """:271-282."""
        return self.get_outputs(ray_samples, return_alphas=return_alphas)


# =====================================================================================================================
# DirectionalDistanceField
# =====================================================================================================================
class _FiLMLayer(nn.Module):           # film_siren.py:14-43 (parameter container; the arithmetic runs in the kernels)
    def __init__(self, in_f: int, out_f: int):
        super().__init__()
        self.layer = nn.Linear(in_f, out_f)


class _MappingNetwork(nn.Module):      # film_siren.py:45-72: Linear / LeakyReLU(0.2) x layers, Linear -> 2 * trunk width * trunk layers
    def __init__(self, in_f: int, layers: int, features: int, out_f: int):
        super().__init__()
        mods: List[nn.Module] = []
        for i in range(layers):
            mods += [nn.Linear(in_f if i == 0 else features, features), nn.LeakyReLU(0.2, inplace=True)]
        mods.append(nn.Linear(features, out_f))
        self.network = nn.Sequential(*mods)


class _FiLMSiren(nn.Module):           # film_siren.py:75-156
    def __init__(self, in_dim: int, hidden_layers: int, hidden_features: int, map_in: int, map_layers: int, map_features: int, out_dim: int):
        super().__init__()
        self.net = nn.ModuleList([_FiLMLayer(in_dim if l == 0 else hidden_features, hidden_features) for l in range(hidden_layers)])
        self.final_layer = nn.Linear(hidden_features, out_dim)
        self.mapping_network = _MappingNetwork(map_in, map_layers, map_features, hidden_layers * hidden_features * 2)


def ddf_direction_features(d: Tensor) -> Tensor:
    """[d (3) | sin(2 pi d f) (6, index dim*2+f, f in {1,4}) | the same + pi/2 (6) | 0] = 16 columns: the trunk input rows
    (directional_distance_field.py:270-271 with NeRFEncoding(3, 2 freqs, no input)), zero-padded to the MMA K step."""
    ang = (2.0 * torch.pi * d)[:, :, None] * d.new_tensor([1.0, 4.0])
    ang = ang.reshape(-1, 6)
    return torch.cat([d, torch.sin(ang), torch.sin(ang + torch.pi / 2.0), torch.zeros_like(d[:, :1])], dim=-1)


class _DDFFieldRows(torch.autograd.Function):
    """DirectionalDistanceField.get_outputs on rows whose directions are ALREADY in the local frame (the field-level call,
    directional_distance_field.py:261-306): cond = [q | hash(q)], x = [d | PE(d)], FiLM-SIREN, sigmoid * 2r."""

    @staticmethod
    def forward(ctx, cfg, origins, dirs_local, table, w_final, b_final, *mlp):
        from . import train

        n = origins.shape[0]
        feat = ops.hash_encode(origins, table, cfg.scalings, cfg.log2_T)                       # [n,32]
        cond = torch.cat([origins, feat, origins.new_zeros((n, 40 - 3 - feat.shape[1]))], dim=-1).contiguous()
        xin = ddf_direction_features(dirs_local).contiguous()
        term = origins.new_zeros(n)
        thr = origins.new_zeros(())
        that, _vis, (film, Wm, Wt, hs, zs, acts) = train._ddf_forward_core(cfg, cond, xin, term, thr, w_final, b_final, mlp)
        ctx.cfg, ctx.b_shape, ctx.n_act = cfg, b_final.shape, (len(hs), len(zs))
        ctx.save_for_backward(cond, xin, origins, term, film, that, thr, w_final, *Wm, *Wt, *hs, *zs, *acts)
        return that

    @staticmethod
    def backward(ctx, d_that):
        from . import train

        sv = ctx.saved_tensors
        cond, xin, origins, term, film, that, thr, w_final = sv[:8]
        o = 8
        Wm = sv[o:o + train.DDF_MAP_LAYERS]; o += train.DDF_MAP_LAYERS
        Wt = sv[o:o + train.DDF_TRUNK_LAYERS]; o += train.DDF_TRUNK_LAYERS
        hs = sv[o:o + ctx.n_act[0]]; o += ctx.n_act[0]
        zs = sv[o:o + ctx.n_act[1]]; o += ctx.n_act[1]
        acts = sv[o:o + ctx.n_act[1]]
        _, d_table, d_wf, d_bf, grads_mlp, _ = train._ddf_backward_core(ctx.cfg, cond, xin, origins, term, film, that, thr, w_final, Wm, Wt, hs, zs, acts,
                                                                        None, d_that.contiguous(), ctx.needs_input_grad[3], False, ctx.b_shape)
        return (None, None, None, d_table, d_wf, d_bf, *grads_mlp)


class DirectionalDistanceField(nn.Module):
    """neusky/fields/directional_distance_field.py:96-315, NeuSky configuration only (hash position encoding, NeRF direction
    encoding, FiLM conditioning, 5 x 256 trunk and mapping layers, "ddf" head)."""

    def __init__(self, config: DirectionalDistanceFieldConfig, ddf_radius: float = 1.0) -> None:
        super().__init__()
        c = config
        if c.position_encoding_type == "icosphere_hash" or c.direction_encoding_type == "icosphere_hash":
            raise NotImplementedError("Icosphere hash encoding not implemented yet")           # :177-181
        if (c.position_encoding_type, c.direction_encoding_type, c.conditioning, c.ddf_type) != ("hash", "nerf", "FiLM", "ddf") or c.predict_probability_of_hit \
                or (c.hidden_layers, c.hidden_features, c.mapping_layers, c.mapping_features) != (5, 256, 5, 256) or c.termination_output_activation != "sigmoid":
            raise NotImplementedError("DirectionalDistanceField: the kernels implement the NeuSky DDF (neusky_config.py:162-177): hash position encoding, "
                                      "NeRF direction encoding, FiLM conditioning, 5 x 256 trunk / mapping layers, sigmoid 'ddf' head")
        self.config = c
        self.ddf_radius = float(ddf_radius)
        self.position_encoding = _HashGrid(16, 19, 2, 16, 2048)                                 # :138-156
        self.position_encoding.tcnn_smoothstep = False          # no "interpolation" key in the reference's encoding_config: tcnn's default, linear
        self.direction_encoding = None                                                          # NeRFEncoding has no parameters (:189-192)
        self.num_depth_components = c.num_dirac_components
        self.num_weight_components = c.num_dirac_components - 1
        self.ddf = _FiLMSiren(3 + 12, c.hidden_layers, c.hidden_features, 3 + self.position_encoding.n_output_dims, c.mapping_layers, c.mapping_features, 1)
        self.termination_output_activation = torch.sigmoid
        self.probability_of_hit_output_activation = torch.sigmoid
        self.reset_parameters()

    def reset_parameters(self) -> None:
        """FiLMSiren's own initialisers (film_siren.py:22-43, 60-62, 123, 135-136) + the hash table's U(-1,1) * 1e-3."""
        p = nb_init.init_ddf_params(int(torch.randint(0, 2**31 - 1, (1,))))
        with torch.no_grad():
            for k, v in p.items():
                self.get_parameter(k).copy_(v)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        from .tcnn_import import convert_tcnn_state_dict_entry

        convert_tcnn_state_dict_entry(state_dict, prefix + "position_encoding.", self.position_encoding)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def named_ddf_state(self) -> Dict[str, Tensor]:
        """The parameters under the reference's state_dict names (what packing.pack_ddf_* / train.ddf_param_list take)."""
        return dict(self.named_parameters())

    def _cfg(self):
        from . import train

        return train.DDFConfig(scalings=self.position_encoding.scalings, log2_T=self.position_encoding.log2_T, radius=self.ddf_radius, split=self.config.split)

    def get_density(self, ray_samples):
        raise NotImplementedError                                                                # :255-256

    def get_outputs(self, ray_samples) -> Dict[Enum, Tensor]:
        """:261-306.  ray_samples.frustums.origins [N,3] (on the DDF sphere), .directions [N,3] (in the local frame of the origin,
        as DDFModel passes them) -> {TERMINATION_DISTANCE: [N]}."""
        from . import train

        self.position_encoding.require_native("DirectionalDistanceField")
        o = ray_samples.frustums.origins.reshape(-1, 3).contiguous()
        d = ray_samples.frustums.directions.reshape(-1, 3).contiguous()
        p = self.named_ddf_state()
        args = (self._cfg(), o, d, p["position_encoding.hash_table"], p["ddf.final_layer.weight"], p["ddf.final_layer.bias"], *train.ddf_param_list(p))
        if _need_autograd(list(self.parameters())):
            that = _DDFFieldRows.apply(*args)
        else:
            with torch.no_grad():
                that = _DDFFieldRows.apply(*args)
        return {NeuSkyFieldHeadNames.TERMINATION_DISTANCE: that}

    def forward(self, ray_samples) -> Dict[Enum, Tensor]:
        """:308-315."""
        return self.get_outputs(ray_samples)


# =====================================================================================================================
# RENIField
# =====================================================================================================================
class _VNLinear(nn.Module):            # vn_layers.py:191-216: weight [dim_out, dim_in] (randn)
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(dim_out, dim_in))


class _VNReLU(nn.Module):              # vn_layers.py:219-246: W, U [dim, dim] (randn)
    def __init__(self, dim: int):
        super().__init__()
        self.W = nn.Parameter(torch.randn(dim, dim))
        self.U = nn.Parameter(torch.randn(dim, dim))


class _VNInvariant(nn.Module):         # vn_layers.py:404-419: mlp = Sequential(VNLinear(dim, dim_coor), VNReLU(dim_coor), Rearrange)
    def __init__(self, dim: int, dim_coor: int):
        super().__init__()
        self.mlp = nn.Sequential(_VNLinear(dim, dim_coor), _VNReLU(dim_coor), nn.Identity())


class _MHA(nn.Module):                 # transformer_decoder.py:21-71 (query / key never influence the output: softmax over ONE key)
    def __init__(self, q_dim: int, kv_dim: int, hidden: int):
        super().__init__()
        self.query = nn.Linear(q_dim, hidden)
        self.key = nn.Linear(kv_dim, hidden)
        self.value = nn.Linear(kv_dim, hidden)
        self.fc_out = nn.Linear(hidden, hidden)


class _AttentionLayer(nn.Module):      # transformer_decoder.py:74-97
    def __init__(self, kv_dim: int, hidden: int):
        super().__init__()
        self.mha = _MHA(hidden, kv_dim, hidden)
        self.norm1 = nn.LayerNorm(hidden)
        self.norm2 = nn.LayerNorm(hidden)
        self.fc = nn.Sequential(nn.Linear(hidden, hidden), nn.ReLU(), nn.Linear(hidden, hidden))


class _Decoder(nn.Module):             # transformer_decoder.py:100-155
    def __init__(self, in_dim: int, cond_dim: int, hidden: int, num_layers: int, out_dim: int):
        super().__init__()
        self.residual_projection = nn.Linear(in_dim, hidden)
        self.layers = nn.ModuleList([_AttentionLayer(cond_dim, hidden) for _ in range(num_layers)])
        self.fc = nn.Linear(hidden, out_dim)


class _ReniLogRows(torch.autograd.Function):
    """Log-domain RENI++ output of N rows with K distinct (latent, scale) codes: RENIField.get_outputs (:493-573).  Backward
    into the latent codes and scales (decoder frozen), through nsk_reni_decode_bwd."""

    @staticmethod
    def forward(ctx, dirs, row_cam, latents, scale, packed, packed_bwd, rotation, mode: int):
        out = ops.reni_radiance_rows(dirs, row_cam, latents, scale, packed, rotation=rotation, log_domain=mode)
        e = dirs.new_zeros(0)
        ctx.save_for_backward(dirs, row_cam, latents, scale if scale is not None else e, packed, packed_bwd if packed_bwd is not None else e,
                              rotation if rotation is not None else e, out)
        ctx.flags = (scale is not None, rotation is not None, mode)
        return out

    @staticmethod
    def backward(ctx, g):
        dirs, row_cam, latents, scale, packed, packed_bwd, rotation, out = ctx.saved_tensors
        has_scale, has_rot, mode = ctx.flags
        d_lat = torch.zeros_like(latents)
        d_scale = torch.zeros_like(scale) if has_scale else None
        # mode 2 returned the raw log value o: d o / d o = 1, which the log-domain backward computes as g * out with out = 1
        o = torch.ones_like(out) if mode == 2 else out
        ops.reni_decode_bwd(dirs, row_cam, latents, scale if has_scale else None, packed, packed_bwd, o, g.contiguous(), d_lat, d_scale,
                            rotation=rotation if has_rot else None, log_domain=(mode != 0))
        return None, None, d_lat, d_scale, None, None, None, None


class RENIField(nn.Module):
    """ns_reni/reni/illumination_fields/reni_illumination_field.py:90-593 in the configuration NeuSky uses (neusky_config.py:78-96):
    SO2 equivariance about z, VN invariant layers, attention conditioning (one key/value token => a conditioned MLP),
    NeRF positional encoding of the directional input, linear output.  Other configurations raise NotImplementedError."""

    def __init__(self, config: RENIFieldConfig, num_train_data: Optional[int] = None, num_eval_data: Optional[int] = None,
                 normalisations: Optional[Dict[str, Any]] = None) -> None:
        super().__init__()
        c = config
        if (c.conditioning, c.invariant_function, c.equivariance, c.axis_of_invariance, c.positional_encoding, c.encoded_input) != \
                ("Attention", "VN", "SO2", "z", "NeRF", "Directions") or c.output_activation != "None" or c.old_implementation or c.out_features != 3:
            raise NotImplementedError("RENIField: the kernels implement the RENI++ decoder NeuSky ships (neusky_config.py:78-96): Attention conditioning, VN "
                                      "invariance, SO2 about z, NeRF encoding of the directions, linear output")
        self.config = c
        # BaseRENIField (base_spherical_field.py:60-95): counts, normalisation buffers, fixed_decoder
        self.num_train_data, self.num_eval_data = num_train_data, num_eval_data
        self.normalisations = normalisations
        self.register_buffer("min_max", torch.tensor(False))
        self.register_buffer("log_domain", torch.tensor(False))          # the shipped checkpoint carries True and overwrites it on load
        if normalisations is not None:
            if normalisations.get("min_max") is not None:
                self.min_max.data = torch.tensor(normalisations["min_max"])
            if normalisations.get("log_domain") is not None:
                self.log_domain.data = torch.tensor(normalisations["log_domain"])
        self._ld_key, self._ld_val = None, False
        self.fixed_decoder = c.fixed_decoder
        self.equivariance, self.conditioning = c.equivariance, c.conditioning
        self.latent_dim, self.hidden_layers, self.hidden_features = c.latent_dim, c.hidden_layers, c.hidden_features
        self.mapping_layers, self.mapping_features, self.out_features = c.mapping_layers, c.mapping_features, c.out_features
        self.last_layer_linear, self.output_activation = c.last_layer_linear, c.output_activation
        self.axis_of_invariance = ["x", "y", "z"].index(c.axis_of_invariance)
        if num_train_data is not None:
            self.train_mu = nn.Parameter(torch.zeros(num_train_data, c.latent_dim, 3))                       # :117-124, init_latent_codes "train": zeros
            self.train_logvar = nn.Parameter(torch.zeros(num_train_data, c.latent_dim, 3))
            if c.trainable_scale in [True, "train", "both"]:
                self.train_scale = nn.Parameter(torch.ones(num_train_data))
        if num_eval_data is not None:
            self.eval_mu = nn.Parameter(torch.zeros(num_eval_data, c.latent_dim, 3))
            self.eval_logvar = nn.Parameter(torch.zeros(num_eval_data, c.latent_dim, 3), requires_grad=False)
            if c.trainable_scale in [True, "eval", "both"]:
                self.eval_scale = nn.Parameter(torch.ones(num_eval_data))
        self.vn_proj_in = nn.Sequential(nn.Identity(), _VNLinear(1, 1))                                      # :135-137
        self.vn_invar = _VNInvariant(dim=1, dim_coor=2)                                                      # :138-139
        d_in = (c.latent_dim + 2) * 5                                                                        # NeRF PE, 2 freqs, include_input (:345-348, 486)
        self.network = _Decoder(d_in, c.latent_dim * 3, c.hidden_features, c.num_attention_layers, c.out_features)     # :398-407
        if self.fixed_decoder:                                                                               # :145-155
            for p in list(self.network.parameters()) + list(self.vn_proj_in.parameters()) + list(self.vn_invar.parameters()):
                p.requires_grad = False
        self._blob_key = None
        self._blob_fwd = self._blob_bwd = None
        self._gemm_w = None

    # -- packed decoder -----------------------------------------------------------------------------------------------
    def _decoder_params(self) -> List[Tensor]:
        return list(self.network.parameters()) + list(self.vn_proj_in.parameters()) + list(self.vn_invar.parameters())

    def decoder_state(self) -> Dict[str, Tensor]:
        return {n: p.detach() for n, p in self.named_parameters() if n.startswith(("network.", "vn_"))}

    def _packed(self):
        key = _versions(self._decoder_params())
        if self._blob_key != key:
            dev = self.network.fc.weight.device
            st = self.decoder_state()
            L = self.config.num_attention_layers
            self._blob_fwd = packing.pack_reni(st, L, device=dev)
            self._blob_bwd = packing.pack_reni_bwd(st, L, device=dev)
            self._gemm_w = None
            self._blob_key = key
        return self._blob_fwd, self._blob_bwd

    @contextlib.contextmanager
    def hold_decoder_fixed(self):
        """:157-196 -- freeze the decoder (and train_scale) inside the block, restore the previous requires_grad flags after."""
        ps = self._decoder_params()
        prev = [p.requires_grad for p in ps]
        for p in ps:
            p.requires_grad = False
        ts = getattr(self, "train_scale", None) if self.config.trainable_scale in [True, "train", "both"] and self.num_train_data is not None else None
        prev_ts = None if ts is None else ts.requires_grad
        if ts is not None:
            ts.requires_grad = False
        prev_fixed = self.fixed_decoder
        self.fixed_decoder = True
        try:
            yield
        finally:
            for p, r in zip(ps, prev):
                p.requires_grad_(r)
            if ts is not None:
                ts.requires_grad_(prev_ts)
            self.fixed_decoder = prev_fixed

    def _is_log_domain(self) -> bool:
        """Host copy of the ``log_domain`` buffer, refreshed only when the buffer changes (load_state_dict): no sync per call."""
        key = (self.log_domain.data_ptr(), self.log_domain._version)
        if self._ld_key != key:
            self._ld_key, self._ld_val = key, bool(self.log_domain)
        return self._ld_val

    # -- reference API -----------------------------------------------------------------------------------------------
    def unnormalise(self, x: Tensor) -> Tensor:
        """base_spherical_field.py:143-154: undo min-max, then exp in the log domain."""
        if not self.min_max.dtype == torch.bool:
            lo, hi = self.min_max
            x = 0.5 * (x + 1) * (hi - lo) + lo
        return torch.exp(x) if self._is_log_domain() else x

    def select_scale(self) -> Optional[Tensor]:
        name = "train_scale" if (self.training or self.config.view_train_latents) else "eval_scale"
        return getattr(self, name, None)

    def sample_latent(self, idx: Tensor):
        """base_spherical_field / reni_illumination_field: the field's own per-image codes (unused by NeuSky, which keeps its latents
        on the model, neusky_model.py:261-269).  Returns (latent, mu, log_var) with latent = mu (eval) or mu + eps * std (train)."""
        if self.training and not self.fixed_decoder:
            mu, log_var = self.train_mu[idx], self.train_logvar[idx]
            return mu + torch.randn_like(mu) * torch.exp(0.5 * log_var), mu, log_var
        name = "train" if (self.training or self.config.view_train_latents) else "eval"
        mu, log_var = getattr(self, name + "_mu")[idx], getattr(self, name + "_logvar")[idx]
        return mu, mu, log_var

    def radiance_table(self, directions: Tensor, latent_codes: Tensor, scale: Optional[Tensor], rotation: Optional[Tensor] = None) -> Tensor:
        """[D,3] directions x [K,L,3] latent codes -> unnormalised HDR radiance [K,D,3]: what sample_illumination (neusky_model.py:445-551)
        needs, without expanding the latent code per (camera, direction) row.  Differentiable w.r.t. latent_codes / scale."""
        from . import autograd as nba

        fwd, bwd = self._packed()
        if torch.is_grad_enabled() and (latent_codes.requires_grad or (scale is not None and scale.requires_grad)):
            return nba.reni_radiance(directions, latent_codes, scale, fwd, bwd, rotation=rotation, log_domain=self._is_log_domain())
        return ops.reni_radiance_table(directions, latent_codes.detach(), None if scale is None else scale.detach(), fwd, rotation,
                                       self.hidden_features, self.config.num_attention_layers, self._is_log_domain())

    def get_outputs(self, ray_samples, rotation: Optional[Tensor] = None, latent_codes: Optional[Tensor] = None, scale: Optional[Tensor] = None) -> Dict[Enum, Tensor]:
        """:493-573.  directions [N,3]; latent_codes [N,L,3] / scale [N] as the reference passes them (one row per ray, built by
        indexing with ``ray_samples.camera_indices``, neusky_model.py:481-493): rows are grouped by camera index, each distinct
        code is conditioned ONCE (the reference repeats that per row) and all rows go through one launch.  Returns the model
        output in the model's own domain (log-HDR for the shipped decoder), like the reference; call ``unnormalise()``."""
        if rotation is not None and rotation.dim() == 3:
            raise NotImplementedError("Batched rotation not implemented yet")          # :520-521
        d = ray_samples.frustums.directions.reshape(-1, 3).contiguous()
        N = d.shape[0]
        cam = getattr(ray_samples, "camera_indices", None)
        mu = log_var = None
        if latent_codes is None:                                                       # :505-513: the field's own codes
            if cam is None:
                raise ValueError("RENIField: camera_indices are needed to select the field's own latent codes")
            latent_codes, mu, log_var = self.sample_latent(cam.reshape(-1).long())
            if scale is None and self.select_scale() is not None:
                scale = self.select_scale()[cam.reshape(-1).long()]
        if latent_codes.shape[0] != N:
            raise ValueError(f"latent_codes: expected one row per ray ([{N}, L, 3]), got {tuple(latent_codes.shape)}")
        if cam is not None:
            uniq, inv = torch.unique(cam.reshape(-1), return_inverse=True)             # as neusky_model.py:461
            K = int(uniq.shape[0])
        elif N > 0 and latent_codes.stride(0) == 0:                                    # one code broadcast over the rows
            K, inv = 1, torch.zeros(N, dtype=torch.long, device=d.device)
        else:
            _, inv = torch.unique(latent_codes.reshape(N, -1), dim=0, return_inverse=True)
            K = int(inv.max()) + 1 if N > 0 else 0
        first = torch.full((K,), N, dtype=torch.long, device=d.device).scatter_reduce(0, inv, torch.arange(N, device=d.device), "amin")
        Z = latent_codes[first].contiguous()
        sc = None if scale is None else scale.reshape(-1)[first].contiguous()
        fwd, bwd = self._packed()
        mode = 2 if self._is_log_domain() else 0
        rgb = _ReniLogRows.apply(d, inv.to(torch.int32).contiguous(), Z, sc, fwd, bwd, rotation, mode)
        return {RENIFieldHeadNames.RGB: rgb, RENIFieldHeadNames.MU: mu, RENIFieldHeadNames.LOG_VAR: log_var}

    def forward(self, ray_samples, rotation: Optional[Tensor] = None, latent_codes: Optional[Tensor] = None, scale: Optional[Tensor] = None) -> Dict[Enum, Tensor]:
        """:575-593."""
        return self.get_outputs(ray_samples=ray_samples, rotation=rotation, latent_codes=latent_codes, scale=scale)
