"""Seeded random initialisers for every weight set on the hot path.

Distributions follow the reference's own initialisers (cited per function); the random
stream is ours (a ``torch.Generator`` on CPU), so a seed reproduces the same weights on any
box with the same torch build.  Parameter names are the reference's ``state_dict`` names so
reference checkpoints load into the drop-in modules unchanged.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

Tensor = torch.Tensor


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _uniform(g, shape, bound) -> Tensor:
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def _linear_bias(g, out_f, in_f) -> Tensor:
    # torch.nn.Linear default: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    return _uniform(g, (out_f,), 1.0 / math.sqrt(in_f))


def _linear_default(g, out_f, in_f):
    # torch.nn.Linear default weight init: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), +)
    return _uniform(g, (out_f, in_f), 1.0 / math.sqrt(in_f)), _linear_bias(g, out_f, in_f)


def hash_scalings(num_levels: int = 16, min_res: int = 16, max_res: int = 2048) -> Tensor:
    """Per-level grid scale exactly as nerfstudio's HashEncoding computes it (float32 pow,
    hence 2047 at the top level) [NS-mem, SURVEY A.3]."""
    levels = torch.arange(num_levels)
    growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1.0
    return torch.floor(min_res * growth**levels).to(torch.float32)


def init_hash_table(seed: int, num_levels: int = 16, log2_T: int = 19, features: int = 2, scale: float = 1e-3) -> Tensor:
    """nerfstudio HashEncoding init: (rand*2-1)*1e-3, one table [L*T, F] [NS-mem A.3]."""
    return _uniform(_gen(seed), ((1 << log2_T) * num_levels, features), scale)


def init_ddf_params(
    seed: int,
    hidden: int = 256,
    layers: int = 5,
    map_hidden: int = 256,
    map_layers: int = 5,
    num_levels: int = 16,
    log2_T: int = 19,
    features: int = 2,
    final_gain: float = 1.0,
    table_scale: float = 1e-3,
) -> Dict[str, Tensor]:
    """DirectionalDistanceField weights (neusky/fields/directional_distance_field.py:220-243
    with neusky/configs/neusky_config.py:162-177) initialised as FiLMSiren does
    (ns_reni/reni/field_components/film_siren.py:22-43, 60-62, 123, 135-136)."""
    g = _gen(seed)
    p: Dict[str, Tensor] = {}
    in_map = 3 + num_levels * features
    in_dir = 3 + 12
    gain = math.sqrt(2.0 / (1 + 0.2**2))  # kaiming_normal_(a=0.2, fan_in, leaky_relu)
    dims = [in_map] + [map_hidden] * map_layers + [layers * hidden * 2]
    for i in range(len(dims) - 1):
        W = torch.randn((dims[i + 1], dims[i]), generator=g) * (gain / math.sqrt(dims[i]))
        if i == len(dims) - 2:
            W = W * 0.25  # film_siren.py:61-62
        p[f"ddf.mapping_network.network.{2 * i}.weight"] = W
        p[f"ddf.mapping_network.network.{2 * i}.bias"] = _linear_bias(g, dims[i + 1], dims[i])
    for l in range(layers):
        fan = in_dir if l == 0 else hidden
        bound = 1.0 / fan if l == 0 else math.sqrt(6.0 / fan) / 25.0  # :38-42, :28-33
        p[f"ddf.net.{l}.layer.weight"] = _uniform(g, (hidden, fan), bound)
        p[f"ddf.net.{l}.layer.bias"] = _linear_bias(g, hidden, fan)
    p["ddf.final_layer.weight"] = _uniform(g, (1, hidden), math.sqrt(6.0 / hidden) / 25.0) * final_gain
    p["ddf.final_layer.bias"] = _linear_bias(g, 1, hidden)
    p["position_encoding.hash_table"] = _uniform(g, ((1 << log2_T) * num_levels, features), table_scale)
    return p


def init_reni_params(seed: int, latent_dim: int = 100, hidden: int = 128, num_layers: int = 6) -> Dict[str, Tensor]:
    """RENI++ decoder weights (ns_reni/reni/illumination_fields/reni_illumination_field.py:135-145,
    398-407; ns_reni/reni/field_components/transformer_decoder.py:21-133; vn_layers.py:191-232):
    torch-default Linear/LayerNorm init, randn VN weights.  Query/key projections are created
    (checkpoint compatibility) but never influence the output."""
    g = _gen(seed)
    p: Dict[str, Tensor] = {}
    d_in = (latent_dim + 2) * 5  # NeRF PE, 2 freqs, include_input
    c_in = latent_dim * 3
    p["vn_proj_in.1.weight"] = torch.randn((1, 1), generator=g)
    p["vn_invar.mlp.0.weight"] = torch.randn((2, 1), generator=g)
    p["vn_invar.mlp.1.W"] = torch.randn((2, 2), generator=g)
    p["vn_invar.mlp.1.U"] = torch.randn((2, 2), generator=g)
    p["network.residual_projection.weight"], p["network.residual_projection.bias"] = _linear_default(g, hidden, d_in)
    for i in range(num_layers):
        pre = f"network.layers.{i}."
        p[pre + "mha.query.weight"], p[pre + "mha.query.bias"] = _linear_default(g, hidden, hidden)
        p[pre + "mha.key.weight"], p[pre + "mha.key.bias"] = _linear_default(g, hidden, c_in)
        p[pre + "mha.value.weight"], p[pre + "mha.value.bias"] = _linear_default(g, hidden, c_in)
        p[pre + "mha.fc_out.weight"], p[pre + "mha.fc_out.bias"] = _linear_default(g, hidden, hidden)
        for n in ("norm1", "norm2"):
            p[pre + n + ".weight"] = torch.ones(hidden)
            p[pre + n + ".bias"] = torch.zeros(hidden)
        p[pre + "fc.0.weight"], p[pre + "fc.0.bias"] = _linear_default(g, hidden, hidden)
        p[pre + "fc.2.weight"], p[pre + "fc.2.bias"] = _linear_default(g, hidden, hidden)
    p["network.fc.weight"], p["network.fc.bias"] = _linear_default(g, 3, hidden)
    return p


def init_sdf_params(
    seed: int,
    hidden: int = 256,
    geo_feat: int = 256,
    num_layers: int = 2,
    num_layers_color: int = 2,
    bias: float = 0.1,
    inside_outside: bool = False,
    num_levels: int = 16,
    log2_T: int = 19,
    features: int = 2,
    beta_init: float = 0.1,
) -> Dict[str, Tensor]:
    """SDFAlbedoField weights: geometric init of nerfstudio SDFField.initialize_geo_layers
    [NS-mem, SURVEY A.4] with the NeuSky overrides (neusky/configs/neusky_config.py:66-77);
    colour network per neusky/fields/sdf_albedo_field.py:148-161.  weight_norm is stored
    as (weight_g, weight_v) like nn.utils.weight_norm."""
    g = _gen(seed)
    p: Dict[str, Tensor] = {}
    in_dim = 3 + 36 + num_levels * features
    dims = [in_dim] + [hidden] * num_layers + [1 + geo_feat]
    n = len(dims) - 1
    for l in range(n):
        out_dim, fan = dims[l + 1], dims[l]
        if l == n - 1:
            mean = math.sqrt(math.pi) / math.sqrt(fan)
            W = torch.randn((out_dim, fan), generator=g) * 1e-4 + (mean if not inside_outside else -mean)
            b = torch.full((out_dim,), -bias if not inside_outside else bias)
        elif l == 0:
            W = torch.zeros((out_dim, fan))
            W[:, :3] = torch.randn((out_dim, 3), generator=g) * (math.sqrt(2) / math.sqrt(out_dim))
            b = torch.zeros(out_dim)
        else:
            W = torch.randn((out_dim, fan), generator=g) * (math.sqrt(2) / math.sqrt(out_dim))
            b = torch.zeros(out_dim)
        p[f"glin{l}.weight_v"] = W
        p[f"glin{l}.weight_g"] = W.norm(dim=1, keepdim=True)
        p[f"glin{l}.bias"] = b
    cdims = [3 + 36 + geo_feat] + [hidden] * num_layers_color + [3]
    for l in range(len(cdims) - 1):
        W, b = _linear_default(g, cdims[l + 1], cdims[l])
        p[f"clin{l}.weight_v"] = W
        p[f"clin{l}.weight_g"] = W.norm(dim=1, keepdim=True)
        p[f"clin{l}.bias"] = b
    p["encoding.hash_table"] = _uniform(g, ((1 << log2_T) * num_levels, features), 1e-3)
    p["deviation_network.variance"] = torch.tensor(beta_init)
    return p


def init_proposal_params(seed: int, num_levels: int = 5, log2_T: int = 17, hidden: int = 16, table_scale: float = 1e-3, density_bias: float = 0.0) -> Dict[str, Tensor]:
    """Random-init state of one nerfstudio HashMLPDensityField (proposal network): hash table U(-1,1)*table_scale
    (nerfstudio: 1e-3), torch Linear default init for Linear(2L,16) and Linear(16,1).  A larger ``table_scale`` / ``density_bias``
    gives a non-trivial density (benchmarks and tests use that so the resampled placement is not uniform)."""
    g = _gen(seed)
    p = {"encoding.hash_table": _uniform(g, ((1 << log2_T) * num_levels, 2), table_scale)}
    for i, (fin, fout) in enumerate(((2 * num_levels, hidden), (hidden, 1))):
        p[f"mlp.{i}.weight"], p[f"mlp.{i}.bias"] = _linear_default(g, fout, fin)
    p["mlp.1.bias"] = p["mlp.1.bias"] + density_bias
    return p
