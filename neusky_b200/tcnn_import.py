"""Import of tiny-cuda-nn ("tcnn") hash-grid parameters (second half).

The reference builds its SDF and DDF position encodings with ``tcnn.Encoding`` (neusky/fields/sdf_albedo_field.py:117-130,
neusky/fields/directional_distance_field.py:139-156), so a trained reference checkpoint stores them as ONE flat fp16
vector ``<encoding>.params``.  tcnn's grid differs from the nerfstudio torch hash grid our kernels implement (SURVEY A.3
"tcnn differences"; restated here FROM MEMORY of tiny-cuda-nn's ``grid.h`` -- tiny-cuda-nn is not in this image, so this
file is pinned by self-consistency tests only, tests/test_tcnn_import.py, and marked UNPINNED against tcnn itself):

  * per level l: ``scale_l = 2^(l * log2(per_level_scale)) * base_resolution - 1``, ``res_l = ceil(scale_l) + 1``;
  * level l owns ``n_l = min(round_up(res_l^3, 8), 2^log2_hashmap_size)`` entries, stored back to back (NOT T per level);
  * a level with ``res_l^3 <= n_l`` is DENSE: index = x + y res + z res^2; otherwise the coherent prime hash
    ``x ^ (y * 2654435761) ^ (z * 805459861)``; both taken modulo ``n_l``;
  * the grid position is ``x * scale_l + 0.5`` (corners at floor / floor + 1, not floor / ceil), x in [0,1];
  * interpolation weights are linear or smoothstep ``w^2 (3 - 2 w)`` (nerfstudio's SDFFieldConfig.smoothstep = True);
  * parameters are fp16, features interleaved ([entry][feature]).

``tcnn_levels`` computes the level geometry, ``tcnn_params_to_table`` re-lays the flat vector out as the fp32 ``[L * T, F]``
table our fields hold (level l at rows ``[l T, l T + n_l)``, the rest zero) and ``nsk_hash_encode_tcnn_fwd``
(``ops.hash_encode_tcnn``) evaluates it with tcnn's indexing, offset and interpolation.  The fused field kernels take the same
per-level table through their ``*_fwd_ex`` entry points (``GridMode`` in csrc/nsk_common.cuh): K2 (``nsk_sdf_field_tc_fwd_ex`` /
``_simt_fwd_ex``, incl. the smoothstep factor of the analytic normal) and K4 (``nsk_sky_shade_tc2_fwd_ex`` / ``_simt_fwd_ex``), so an
imported checkpoint renders through the same eval path; training from an imported grid is not supported.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List

import numpy as np
import torch

Tensor = torch.Tensor

PRIMES = (1, 2654435761, 805459861)


@dataclass(frozen=True)
class TcnnLevel:
    scale: float          # float32 value of the grid scale
    resolution: int
    size: int             # entries owned by the level (its modulus)
    offset: int           # first entry of the level in tcnn's flat parameter vector
    dense: bool


def tcnn_levels(num_levels: int = 16, base_res: int = 16, max_res: int = 2048, log2_hashmap_size: int = 19) -> List[TcnnLevel]:
    """Level geometry as tcnn's GridEncodingTemplated constructor computes it; per_level_scale as the reference passes it
    (``growth_factor = exp((ln max_res - ln base_res) / (num_levels - 1))``, sdf_albedo_field.py:115)."""
    growth = float(np.exp((np.log(max_res) - np.log(base_res)) / (num_levels - 1))) if num_levels > 1 else 1.0
    log2_pls = np.float32(np.log2(np.float32(growth)))
    out, offset = [], 0
    for l in range(num_levels):
        scale = np.float32(np.exp2(np.float32(l) * log2_pls) * np.float32(base_res) - np.float32(1.0))
        res = int(math.ceil(float(scale))) + 1
        n = res ** 3
        n = (n + 7) // 8 * 8
        size = min(n, 1 << log2_hashmap_size)
        out.append(TcnnLevel(float(scale), res, size, offset, res ** 3 <= size))
        offset += size
    return out


def tcnn_num_params(levels: List[TcnnLevel], features: int = 2) -> int:
    return (levels[-1].offset + levels[-1].size) * features


def tcnn_params_to_table(params: Tensor, levels: List[TcnnLevel], log2_hashmap_size: int = 19, features: int = 2) -> Tensor:
    """Flat tcnn ``params`` (fp16 or fp32, ``[sum_l n_l * F]``) -> fp32 ``[L * T, F]`` table in our level-major layout."""
    T = 1 << log2_hashmap_size
    flat = params.detach().reshape(-1).to(torch.float32)
    need = tcnn_num_params(levels, features)
    if flat.numel() != need:
        raise ValueError(f"tcnn params: expected {need} values for this grid configuration, got {flat.numel()}")
    table = torch.zeros((len(levels) * T, features), dtype=torch.float32, device=flat.device)
    for l, lv in enumerate(levels):
        table[l * T:l * T + lv.size] = flat[lv.offset * features:(lv.offset + lv.size) * features].reshape(lv.size, features)
    return table


def tcnn_level_meta(levels: List[TcnnLevel], device) -> Tensor:
    """int32 [L,4] = (float bits of scale, resolution, size, dense) -- the per-level table nsk_hash_encode_tcnn_fwd reads."""
    rows = [[int(np.float32(lv.scale).view(np.int32)), lv.resolution, lv.size, int(lv.dense)] for lv in levels]
    return torch.tensor(rows, dtype=torch.int32, device=device)


def tcnn_grid_encode_torch(x: Tensor, table: Tensor, levels: List[TcnnLevel], log2_hashmap_size: int = 19, smoothstep: bool = True) -> Tensor:
    """Plain-torch statement of tcnn's grid forward on our table layout (test yardstick; x in [0,1], [n,3] -> [n, L*F])."""
    T = 1 << log2_hashmap_size
    outs = []
    for l, lv in enumerate(levels):
        # tcnn: fmaf(scale, x, 0.5f) -- one rounding; the fp64 product of two fp32 values is exact, so this is the same number
        pos = (x.to(torch.float64) * float(np.float32(lv.scale)) + 0.5).to(torch.float32)
        g = torch.floor(pos)
        w = pos - g
        if smoothstep:
            w = w * w * (3.0 - 2.0 * w)
        g = g.to(torch.int64)
        acc = 0
        for c in range(8):
            bit = [(c >> d) & 1 for d in range(3)]
            gc = torch.stack([g[:, d] + bit[d] for d in range(3)], -1)
            wc = torch.ones_like(w[:, 0])
            for d in range(3):
                wc = wc * (w[:, d] if bit[d] else (1.0 - w[:, d]))
            u = gc & 0xFFFFFFFF                      # grid coordinates as tcnn holds them: uint32 (negative positions wrap around)
            if lv.dense:
                idx = (u[:, 0] + ((u[:, 1] * lv.resolution) & 0xFFFFFFFF) + ((u[:, 2] * (lv.resolution * lv.resolution)) & 0xFFFFFFFF)) & 0xFFFFFFFF
            else:
                idx = (u[:, 0] * PRIMES[0]) ^ ((u[:, 1] * PRIMES[1]) & 0xFFFFFFFF) ^ ((u[:, 2] * PRIMES[2]) & 0xFFFFFFFF)
            idx = idx % lv.size
            acc = acc + wc[:, None] * table[l * T + idx]
        outs.append(acc)
    return torch.cat(outs, -1)


def convert_tcnn_state_dict_entry(state_dict: Dict[str, Tensor], prefix: str, grid) -> None:
    """``_load_from_state_dict`` hook of the fields: a reference checkpoint carries ``<prefix>params`` (tcnn) where our module
    holds ``<prefix>hash_table``.  The flat vector is re-laid out as our table and the grid module is switched to tcnn
    semantics (``grid.tcnn_levels``).  The fused field kernels (K2, K4) evaluate the nerfstudio torch-grid semantics only, so a
    module carrying an imported tcnn grid serves the stand-alone encode (``_HashGrid.forward`` -> nsk_hash_encode_tcnn_fwd) and
    raises in the fused paths -- see DESIGN.md section 6."""
    key = prefix + "params"
    if key not in state_dict:
        return
    levels = tcnn_levels(grid.num_levels, grid.base_res, grid.max_res, grid.log2_T)
    state_dict[prefix + "hash_table"] = tcnn_params_to_table(state_dict.pop(key), levels, grid.log2_T, grid.features)
    grid.tcnn_levels = levels
