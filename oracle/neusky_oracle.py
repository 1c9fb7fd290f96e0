"""CPU oracle for the NeuSky per-ray render-and-shade hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import it.  The product path (``neusky_b200``) never does and fails loudly
when its CUDA library is missing.

It is a plain-torch (CPU, fp32 or fp64) restatement of the reference algorithm.  Every
function cites the reference ``file:line`` it follows (paths relative to the
reference root; ``[NS-mem]`` marks nerfstudio behaviour restated from memory because
nerfstudio is an un-vendored, un-pinned dependency -- SURVEY.md Appendix A).

Parity status
-------------
* Pieces whose source IS in the reference tree (FiLM-SIREN DDF network, RENI++ decoder,
  VN invariant layers, local DDF frame, ray/sphere exit, visibility sigmoid, Lambertian
  renderer, icosphere directions, sRGB) are pinned: ``tests/golden/make_golden.py``
  imports the reference's own modules from /root/reference and the fixtures under
  ``tests/golden/`` hold their outputs; ``tests/test_oracle_golden.py`` checks this file
  against them.
* Pieces that live in nerfstudio / tiny-cuda-nn (hash grid, NeRF encoding, SDFField
  geo network, NeuS alpha, samplers, renderers) are restated from memory:
  **parity unpinned** for those (the reference has no tests or golden vectors).

Parameters are plain ``dict[str, Tensor]`` using the reference's ``state_dict`` names.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------
# Hash grid  -- nerfstudio HashEncoding.pytorch_fwd semantics [NS-mem, SURVEY A.3]
# substituted for tcnn.Encoding at neusky/fields/sdf_albedo_field.py:117-130 and
# neusky/fields/directional_distance_field.py:139-156
# --------------------------------------------------------------------------------------

HASH_PRIMES = (1, 2654435761, 805459861)


def hash_scalings(num_levels: int = 16, min_res: int = 16, max_res: int = 2048) -> Tensor:
    """Per-level scale, computed exactly the way nerfstudio does (float32 ``pow``):
    ``floor(min_res * growth ** arange(L))``.  Note the top level comes out as 2047, not
    2048, because the pow is evaluated in float32.  Kernels take this array as an input."""
    levels = torch.arange(num_levels)
    growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1.0
    return torch.floor(min_res * growth**levels).to(torch.float32)


def hash_corner_indices(x: Tensor, scalings: Tensor, log2_T: int) -> Tuple[Tensor, Tensor]:
    """Integer work of the hash grid.  x [N,3] float32 -> (idx [N,L,8] int64 including the
    per-level offset l*T, offset [N,L,3] float32).  Corner order 0..7 =
    (c,c,c),(c,f,c),(f,f,c),(f,c,c),(c,c,f),(c,f,f),(f,f,f),(f,c,f)  [SURVEY A.3]."""
    T = 1 << log2_T
    L = scalings.shape[0]
    scaled = x[..., None, :] * scalings.view(-1, 1).to(x.dtype)  # [N,L,3]
    sc = torch.ceil(scaled).to(torch.int32)
    sf = torch.floor(scaled).to(torch.int32)
    offset = scaled - sf
    primes = torch.tensor(HASH_PRIMES, dtype=torch.int64)

    def h(cx, cy, cz):
        v = torch.stack([cx, cy, cz], dim=-1).to(torch.int64) * primes  # int64 products
        r = torch.bitwise_xor(torch.bitwise_xor(v[..., 0], v[..., 1]), v[..., 2])
        r = torch.remainder(r, T)  # non-negative
        return r + (torch.arange(L, dtype=torch.int64) * T)

    cx, cy, cz = sc[..., 0], sc[..., 1], sc[..., 2]
    fx, fy, fz = sf[..., 0], sf[..., 1], sf[..., 2]
    idx = torch.stack(
        [h(cx, cy, cz), h(cx, fy, cz), h(fx, fy, cz), h(fx, cy, cz), h(cx, cy, fz), h(cx, fy, fz), h(fx, fy, fz), h(fx, cy, fz)],
        dim=-1,
    )
    return idx, offset


def hash_encode(x: Tensor, table: Tensor, scalings: Tensor, log2_T: int) -> Tensor:
    """x [N,3] -> [N, L*F] (level-major), linear interpolation in the nerfstudio order."""
    idx, o = hash_corner_indices(x.to(torch.float32), scalings, log2_T)
    o = o.to(table.dtype)
    f = [table[idx[..., c]] for c in range(8)]  # each [N,L,F]
    ox, oy, oz = o[..., 0:1], o[..., 1:2], o[..., 2:3]
    f03 = f[0] * ox + f[3] * (1 - ox)
    f12 = f[1] * ox + f[2] * (1 - ox)
    f56 = f[5] * ox + f[6] * (1 - ox)
    f47 = f[4] * ox + f[7] * (1 - ox)
    f0312 = f03 * oy + f12 * (1 - oy)
    f4756 = f47 * oy + f56 * (1 - oy)
    enc = f0312 * oz + f4756 * (1 - oz)
    return enc.flatten(-2, -1)


# --------------------------------------------------------------------------------------
# NeRF positional encoding [NS-mem, SURVEY A.2]
# --------------------------------------------------------------------------------------


def nerf_encode(x: Tensor, num_frequencies: int, min_freq_exp: float, max_freq_exp: float, include_input: bool) -> Tensor:
    freqs = 2 ** torch.linspace(min_freq_exp, max_freq_exp, num_frequencies, dtype=x.dtype)
    s = (2 * torch.pi * x)[..., None] * freqs
    s = s.reshape(*s.shape[:-2], -1)
    enc = torch.sin(torch.cat([s, s + torch.pi / 2.0], dim=-1))
    if include_input:
        enc = torch.cat([enc, x], dim=-1)
    return enc


# --------------------------------------------------------------------------------------
# DDF: local frame + FiLM-SIREN field
# --------------------------------------------------------------------------------------


def ddf_local_directions(positions: Tensor, directions: Tensor) -> Tensor:
    """neusky/models/ddf_model.py:158-181 (get_localised_transforms) and :196-200."""
    up = torch.tensor([0.0, 0.0, 1.0], dtype=positions.dtype).expand_as(positions)
    y = -positions
    x_l = torch.linalg.cross(up, y)
    x_l = x_l / x_l.norm(dim=-1, keepdim=True)
    z_l = torch.linalg.cross(y, x_l)
    z_l = z_l / z_l.norm(dim=-1, keepdim=True)
    rot = torch.stack((x_l, y, z_l), dim=-1)  # columns are the local axes
    return torch.einsum("ijl,ij->il", rot, directions)


def film_siren(x: Tensor, cond: Tensor, p: Dict[str, Tensor], prefix: str = "ddf.") -> Tensor:
    """ns_reni/reni/field_components/film_siren.py:45-156 (FiLMSiren with
    outermost_linear=True, no output activation)."""
    h = cond
    i = 0
    while f"{prefix}mapping_network.network.{i}.weight" in p:
        W = p[f"{prefix}mapping_network.network.{i}.weight"]
        b = p[f"{prefix}mapping_network.network.{i}.bias"]
        h = h @ W.T + b
        if f"{prefix}mapping_network.network.{i + 2}.weight" in p:
            h = torch.nn.functional.leaky_relu(h, 0.2)  # film_siren.py:53
        i += 2
    half = h.shape[-1] // 2
    freq, phase = h[..., :half], h[..., half:]  # film_siren.py:66-67
    freq = freq * 15 + 30  # film_siren.py:140
    l = 0
    hid = p[f"{prefix}net.0.layer.weight"].shape[0]
    while f"{prefix}net.{l}.layer.weight" in p:
        W = p[f"{prefix}net.{l}.layer.weight"]
        b = p[f"{prefix}net.{l}.layer.bias"]
        x = x @ W.T + b
        x = torch.sin(freq[..., l * hid : (l + 1) * hid] * x + phase[..., l * hid : (l + 1) * hid])  # :81
        l += 1
    return x @ p[f"{prefix}final_layer.weight"].T + p[f"{prefix}final_layer.bias"]  # :147


def ddf_field(q: Tensor, d_local: Tensor, p: Dict[str, Tensor], scalings: Tensor, log2_T: int, ddf_radius: float) -> Tensor:
    """neusky/fields/directional_distance_field.py:261-306 with position_encoding_type="hash",
    direction_encoding_type="nerf", conditioning="FiLM", sigmoid termination
    (neusky/configs/neusky_config.py:162-177).  Returns expected termination distance [N]."""
    cond = torch.cat([q, hash_encode(q, p["position_encoding.hash_table"], scalings, log2_T).to(q.dtype)], dim=-1)  # :268
    x = torch.cat([d_local, nerf_encode(d_local, 2, 0.0, 2.0, False)], dim=-1)  # :271, :188-191
    out = film_siren(x, cond, p)  # :276
    return torch.sigmoid(out[..., 0]) * (2 * ddf_radius)  # :297-299


def ddf_model(positions: Tensor, directions: Tensor, p, scalings, log2_T, ddf_radius) -> Tensor:
    """neusky/models/ddf_model.py:183-219: localise directions, run the field."""
    return ddf_field(positions, ddf_local_directions(positions, directions), p, scalings, log2_T, ddf_radius)


# --------------------------------------------------------------------------------------
# Visibility  (neusky/models/neusky_model.py:1590-1778)
# --------------------------------------------------------------------------------------


def ray_sphere_intersection(positions: Tensor, directions: Tensor, radius: float) -> Tensor:
    """neusky/models/neusky_model.py:1590-1622 (normalises directions, clamps discriminant)."""
    directions = directions / torch.norm(directions, dim=-1, keepdim=True)
    b = 2 * (directions * positions).sum(-1)
    c = (positions * positions).sum(-1) - radius**2
    disc = torch.clamp(b**2 - 4 * c, min=0.0)
    t0 = (-b - torch.sqrt(disc)) / 2
    t1 = (-b + torch.sqrt(disc)) / 2
    t = torch.max(t0, t1)
    return positions + t.unsqueeze(-1) * directions


def surface_points(origins: Tensor, ray_dirs: Tensor, p2p: Tensor, ddf_radius: float) -> Tensor:
    """neusky_model.py:1667-1683 incl. the element-wise outside-sphere "hack" (SURVEY B.1).
    origins, ray_dirs [R,3]; p2p [R,1] -> [R,3]."""
    pos = origins + ray_dirs * p2p
    inside = pos.norm(dim=-1) < ddf_radius
    if (~inside).any():
        pos = pos.clone()
        pos[~inside] = ray_sphere_intersection(origins[~inside], ray_dirs[~inside], ddf_radius) * 0.01 * -ray_dirs[~inside]
    return pos


def compute_visibility(
    points: Tensor,
    dirs: Tensor,
    p: Dict[str, Tensor],
    scalings: Tensor,
    log2_T: int,
    ddf_radius: float,
    threshold: float,
    sigmoid_scale: float,
    only_upper: bool = True,
    lower_vis: float = 1.0,
    chunk: int = 65536,
) -> Dict[str, Tensor]:
    """neusky_model.py:1624-1778 for already-computed surface points [R,3] and light
    directions [D,3].  Returns visibility [R,D], expected_termination_dist [R*D'],
    termination_dist [R*D'] (distance point -> sphere exit), mask [D] bool."""
    R = points.shape[0]
    D = dirs.shape[0]
    mask = (dirs[:, 2] > 0) if only_upper else torch.ones(D, dtype=torch.bool)  # :1650-1657
    d_sel = dirs[mask]
    Dp = d_sel.shape[0]
    pos = points[:, None, :].expand(R, Dp, 3).reshape(-1, 3)  # :1685-1690
    dd = d_sel[None].expand(R, Dp, 3).reshape(-1, 3)
    outs_ddf, outs_term = [], []
    for s in range(0, pos.shape[0], chunk):
        pc, dc = pos[s : s + chunk], dd[s : s + chunk]
        q = ray_sphere_intersection(pc, dc, ddf_radius)  # :1693
        outs_term.append(torch.norm(q - pc, dim=-1))  # :1697
        outs_ddf.append(ddf_model(q, -dc, p, scalings, log2_T, ddf_radius))  # :1702-1718
    term = torch.cat(outs_term)
    ddf = torch.cat(outs_ddf)
    gt = torch.clamp(term, max=ddf_radius * 2.0)  # :1724-1727
    diff = gt - ddf  # :1730
    vis_sel = 1.0 - torch.sigmoid(sigmoid_scale * (diff - threshold))  # :1739-1740
    vis = torch.full((R, D), float(lower_vis), dtype=points.dtype)  # :1745-1753
    vis[:, mask] = vis_sel.reshape(R, Dp)
    return {
        "visibility": vis,
        "expected_termination_dist": ddf,
        "termination_dist": term,
        "difference": diff,
        "mask": mask,
    }


# --------------------------------------------------------------------------------------
# RENI++ illumination field
# --------------------------------------------------------------------------------------


def vn_invariant_so2(z_xy: Tensor, p: Dict[str, Tensor]) -> Tensor:
    """vn_proj_in (VNLinear(1,1)) + VNInvariant(dim=1, dim_coor=2)
    ns_reni/reni/field_components/vn_layers.py:191-216, 218-246, 404-419 as used at
    ns_reni/reni/illumination_fields/reni_illumination_field.py:138-142, 224-225.
    z_xy [B,L,2] -> [B,L,2]."""
    x = z_xy.unsqueeze(-2)  # '... c -> ... 1 c'  [B,L,1,2]
    x = torch.einsum("...ic,oi->...oc", x, p["vn_proj_in.1.weight"])  # [B,L,1,2]
    # VNInvariant.mlp = VNLinear(1, 2) -> VNReLU(2) -> rearrange '... d e -> ... e d'
    y = torch.einsum("...ic,oi->...oc", x, p["vn_invar.mlp.0.weight"])  # [B,L,2,2]
    q = torch.einsum("...ic,oi->...oc", y, p["vn_invar.mlp.1.W"])
    k = torch.einsum("...ic,oi->...oc", y, p["vn_invar.mlp.1.U"])
    qk = (q * k).sum(-1, keepdim=True)
    k_norm = torch.sqrt((k**2).sum(dim=-1, keepdim=True).clamp(min=1e-6))
    q_proj = q - (q * (k / k_norm)).sum(-1, keepdim=True) * k
    y = torch.where(qk >= 0.0, q, q_proj)  # [B,L,d=2,e=2]
    y = y.transpose(-1, -2)  # '... d e -> ... e d'
    return torch.einsum("bndi,bnio->bno", x, y)  # [B,L,2]


def reni_inputs(dirs: Tensor, Z: Tensor, p: Dict[str, Tensor], rotation: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """reni_illumination_field.py:517-519 (rotation on the latent), :219-246 (SO2 about z,
    VN invariant function), :481-491 + :345-348 (NeRF PE on the directional input).
    dirs [N,3], Z [N,L,3] -> (directional input [N, 5*(L+2)], conditioning [N, 3L])."""
    if rotation is not None:
        Z = torch.matmul(Z, rotation)
    z_xy = torch.stack((Z[:, :, 0], Z[:, :, 1]), -1)
    d_xy = torch.stack((dirs[:, 0], dirs[:, 1]), -1).unsqueeze(1)
    z_inv = vn_invariant_so2(z_xy, p)
    z_z = Z[:, :, 2].unsqueeze(-1)
    inner = (z_xy * d_xy).sum(-1)
    d_norm = torch.sqrt(dirs[:, 0] ** 2 + dirs[:, 1] ** 2).unsqueeze(-1)
    d_z = dirs[:, 2].unsqueeze(-1)
    directional = torch.cat((inner, d_z, d_norm), 1)
    cond = torch.cat((z_inv, z_z), dim=-1).flatten(1)
    return nerf_encode(directional, 2, 0.0, 2.0, True), cond


def reni_decoder(x: Tensor, cond: Tensor, p: Dict[str, Tensor], num_layers: int = 6, eps: float = 1e-5) -> Tensor:
    """ns_reni/reni/field_components/transformer_decoder.py:21-155.  The attention has one
    key/value token, so softmax == 1 and mha(q,c,c) == fc_out(value(c));
    the query/key projections do not influence the output."""
    ln = torch.nn.functional.layer_norm
    x = x @ p["network.residual_projection.weight"].T + p["network.residual_projection.bias"]
    H = x.shape[-1]
    for i in range(num_layers):
        pre = f"network.layers.{i}."
        v = cond @ p[pre + "mha.value.weight"].T + p[pre + "mha.value.bias"]
        a = v @ p[pre + "mha.fc_out.weight"].T + p[pre + "mha.fc_out.bias"]
        o1 = ln(a + x, (H,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], eps)
        f = torch.relu(o1 @ p[pre + "fc.0.weight"].T + p[pre + "fc.0.bias"])
        f = f @ p[pre + "fc.2.weight"].T + p[pre + "fc.2.bias"]
        x = ln(f + o1, (H,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], eps)
    return x @ p["network.fc.weight"].T + p["network.fc.bias"]


def reni_field(dirs: Tensor, Z: Tensor, scale: Optional[Tensor], p: Dict[str, Tensor], rotation: Optional[Tensor] = None, log_domain: bool = True) -> Tensor:
    """RENIField.get_outputs (reni_illumination_field.py:493-573) with the NeuSky config
    (neusky/configs/neusky_config.py:78-96): Attention conditioning, VN invariance, SO2 about
    z, PE on directions, output_activation None.  Returns the *normalised* (log-HDR) RGB."""
    x, cond = reni_inputs(dirs, Z, p, rotation)
    out = reni_decoder(x, cond, p)
    if scale is not None:
        s = torch.exp(scale)  # :562
        out = out + torch.log(s.unsqueeze(1)) if log_domain else out * s.unsqueeze(1)  # :564-567
    return out


def reni_unnormalise(x: Tensor, log_domain: bool = True, min_max: Optional[Tuple[float, float]] = None) -> Tensor:
    """ns_reni/reni/illumination_fields/base_spherical_field.py:143-154."""
    if min_max is not None:
        x = 0.5 * (x + 1) * (min_max[1] - min_max[0]) + min_max[0]
    return torch.exp(x) if log_domain else x


def reni_radiance_table(dirs: Tensor, Zs: Tensor, scales: Tensor, p, rotation=None, log_domain=True) -> Tensor:
    """neusky_model.py:460-510: HDR radiance for K latent codes x D directions -> [K,D,3]."""
    K, D = Zs.shape[0], dirs.shape[0]
    out = []
    for k in range(K):
        Z = Zs[k : k + 1].expand(D, -1, -1)
        sc = scales[k : k + 1].expand(D)
        out.append(reni_unnormalise(reni_field(dirs, Z, sc, p, rotation, log_domain), log_domain))
    return torch.stack(out, 0)


# --------------------------------------------------------------------------------------
# Lambertian shading with visibility (neusky/model_components/renderers.py:60-176)
# --------------------------------------------------------------------------------------


def linear_to_srgb(c: Tensor) -> Tensor:
    """neusky/utils/utils.py:11-31."""
    c = torch.where(c <= 0.0031308, 12.92 * c, 1.055 * torch.pow(torch.abs(c), 1 / 2.4) - 0.055)
    return torch.clamp(c, 0.0, 1.0)


def lambertian_radiance(albedo: Tensor, normals: Tensor, dirs: Tensor, light: Tensor, vis: Optional[Tensor]) -> Tensor:
    """renderers.py:89-113 per sample.  albedo/normals [N,3]; dirs [D,3]; light [N,D,3] or
    [D,3]; vis [N,D] or None -> radiance [N,3]."""
    dot = (normals @ dirs.T).clamp(0.0, 1.0)  # :93-98
    count = (dot > 0).to(dot.dtype).sum(1, keepdim=True)  # :101
    count = torch.where(count > 0, count, torch.ones_like(count))  # :104
    dot = dot / count  # :106
    if vis is not None:
        dot = dot * vis  # :110
    if light.dim() == 2:
        return albedo * (dot @ light)
    return albedo * torch.einsum("bj,bji->bi", dot, light)  # :113


def lambertian_render(albedo, normals, dirs, light, vis, bg, weights, training: bool = False) -> Tensor:
    """renderers.py:60-176.  albedo/normals [R,S,3]; light [R,D,3] (per ray) or [D,3]; vis
    [R,D] per ray (applied to every sample, neusky_model.py:1755-1759); bg [R,3];
    weights [R,S,1] -> sRGB [R,3]."""
    R, S = albedo.shape[:2]
    lightN = light if light.dim() == 2 else light[:, None].expand(R, S, -1, 3).reshape(R * S, -1, 3)
    visN = None if vis is None else vis[:, None].expand(R, S, -1).reshape(R * S, -1)
    rad = lambertian_radiance(albedo.reshape(-1, 3), normals.reshape(-1, 3), dirs, lightN, visN).reshape(R, S, 3)
    comp = (weights * rad).sum(-2)  # :122
    acc = weights.sum(-2)  # :123
    comp = comp + bg * (1.0 - acc)  # :127
    comp = linear_to_srgb(comp)  # :128
    if not training:
        comp = comp.clamp(0.0, 1.0)  # :173-174
    return comp


# --------------------------------------------------------------------------------------
# SDF / albedo field and NeuS compositing
# (neusky/fields/sdf_albedo_field.py + nerfstudio SDFField [NS-mem A.4, A.5, A.7])
# --------------------------------------------------------------------------------------


def scene_contraction_linf(x: Tensor) -> Tensor:
    """nerfstudio SceneContraction(order=inf) [NS-mem A.4]; the field keeps L-inf."""
    mag = torch.linalg.norm(x, ord=float("inf"), dim=-1)[..., None]
    return torch.where(mag < 1, x, (2 - (1 / mag)) * (x / mag))


def _wn(p, name):
    """weight_norm fold: W = g * v / ||v||_row (nn.utils.weight_norm, dim=0)."""
    if name + ".weight" in p:
        return p[name + ".weight"]
    v, g = p[name + ".weight_v"], p[name + ".weight_g"]
    return v * (g / v.norm(dim=1, keepdim=True))


def softplus100(x: Tensor) -> Tensor:
    return torch.nn.functional.softplus(x, beta=100)


def sdf_geo_network(x: Tensor, p, scalings, log2_T) -> Tensor:
    """SDFField.forward_geonetwork [NS-mem A.4], called at sdf_albedo_field.py:172,233.
    x [N,3] -> [N, 1+geo_feat]."""
    pos = (scene_contraction_linf(x) + 2.0) / 4.0
    feat = hash_encode(pos, p["encoding.hash_table"], scalings, log2_T).to(x.dtype)
    pe = nerf_encode(x, 6, 0.0, 5.0, False)  # sdf_albedo_field.py:133-135
    h = torch.cat((x, pe, feat), dim=-1)
    n = 0
    while f"glin{n}.bias" in p:
        n += 1
    for l in range(n):
        h = h @ _wn(p, f"glin{l}").T + p[f"glin{l}.bias"]
        if l < n - 1:
            h = softplus100(h)
    return h


def sdf_colour_network(x: Tensor, geo: Tensor, p) -> Tensor:
    """sdf_albedo_field.py:185-209."""
    h = torch.cat([x, nerf_encode(x, 6, 0.0, 5.0, False), geo], dim=-1)
    n = 0
    while f"clin{n}.bias" in p:
        n += 1
    for l in range(n):
        h = h @ _wn(p, f"clin{l}").T + p[f"clin{l}.bias"]
        if l < n - 1:
            h = torch.relu(h)
    return torch.sigmoid(h)


def sdf_field(x: Tensor, p, scalings, log2_T) -> Dict[str, Tensor]:
    """sdf_albedo_field.py:211-269 without alpha: sdf, gradient (autograd), normals, albedo."""
    x = x.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        h = sdf_geo_network(x, p, scalings, log2_T)
        sdf, geo = h[:, :1], h[:, 1:]
        grad = torch.autograd.grad(sdf, x, torch.ones_like(sdf), create_graph=False, retain_graph=True)[0]
    albedo = sdf_colour_network(x, geo, p)
    normals = torch.nn.functional.normalize(grad, p=2, dim=-1)
    return {"sdf": sdf.detach(), "gradient": grad.detach(), "normals": normals.detach(), "albedo": albedo.detach(), "geo": geo.detach()}


def neus_alpha(sdf, grad, ray_dirs, deltas, inv_s: float, cos_anneal_ratio: float = 1.0) -> Tensor:
    """SDFField.get_alpha [NS-mem A.5], called at sdf_albedo_field.py:266.  All [R,S,.]."""
    true_cos = (ray_dirs * grad).sum(-1, keepdim=True)
    iter_cos = -(torch.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio) + torch.relu(-true_cos) * cos_anneal_ratio)
    nxt = sdf + iter_cos * deltas * 0.5
    prv = sdf - iter_cos * deltas * 0.5
    prev_cdf = torch.sigmoid(prv * inv_s)
    next_cdf = torch.sigmoid(nxt * inv_s)
    return ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).clip(0.0, 1.0)


def weights_from_alphas(alpha: Tensor) -> Tuple[Tensor, Tensor]:
    """RaySamples.get_weights_and_transmittance_from_alphas [NS-mem A.5] (neusky_model.py:565)."""
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-7], 1), 1)
    return alpha * T[:, :-1], T


def render_depth_expected(weights: Tensor, starts: Tensor, ends: Tensor) -> Tensor:
    """DepthRenderer('expected') [NS-mem A.7] (neusky_model.py:591): global clip to steps range."""
    steps = (starts + ends) / 2
    d = (weights * steps).sum(-2) / (weights.sum(-2) + 1e-10)
    return torch.clip(d, steps.min(), steps.max())


def neus_composite(sdf, grad, albedo, radiance, ray_dirs, starts, ends, deltas, bg, dnorm, inv_s, cos_anneal_ratio=1.0, training=False):
    """Everything K3 fuses: alpha (A.5), weights/transmittance, accumulation, expected depth,
    normal (sum w n), albedo (white background), shaded rgb (renderers.py:122-128).
    sdf [R,S,1], grad/albedo/radiance [R,S,3], ray_dirs [R,3], starts/ends/deltas [R,S,1],
    bg [R,3], dnorm [R,1]."""
    alpha = neus_alpha(sdf, grad, ray_dirs[:, None, :], deltas, inv_s, cos_anneal_ratio)
    w, T = weights_from_alphas(alpha)
    acc = w.sum(-2)
    normals = torch.nn.functional.normalize(grad, p=2, dim=-1)
    steps = (starts + ends) / 2
    p2p_raw = (w * steps).sum(-2) / (acc + 1e-10)
    p2p = torch.clip(p2p_raw, steps.min(), steps.max())
    rgb = (w * radiance).sum(-2) + bg * (1.0 - acc)
    rgb = linear_to_srgb(rgb)
    if not training:
        rgb = rgb.clamp(0.0, 1.0)
    alb = (w * albedo).sum(-2) + (1.0 - acc)
    if not training:
        alb = alb.clamp(0.0, 1.0)
    return {
        "alpha": alpha, "weights": w, "transmittance": T, "bg_transmittance": T[:, -1], "accumulation": acc,
        "p2p_dist": p2p, "p2p_raw": p2p_raw, "depth": p2p / dnorm, "normal": (w * normals).sum(-2), "albedo": alb, "rgb": rgb,
    }


# --------------------------------------------------------------------------------------
# Illumination direction sets
# --------------------------------------------------------------------------------------


def icosphere_directions(num_directions: int) -> Tensor:
    """ns_reni/reni/model_components/illumination_samplers.py:87-326 (geodesic icosphere with
    the smallest subdivision frequency giving >= num_directions vertices), float64 numpy ->
    float32.  Vertex ORDER follows the reference construction: 12 icosahedron vertices, then
    edge-interior vertices edge by edge, then face-interior vertices face by face."""
    from oracle._icosphere import icosphere

    v, _ = icosphere(nr_verts=num_directions)
    return torch.from_numpy(v).float()


def equirect_directions(width: int) -> Tensor:
    """EquirectangularSampler (illumination_samplers.py:373-432) through nerfstudio
    Cameras.generate_rays for CameraType.EQUIRECTANGULAR [NS-mem A.8]; z-up."""
    H, W = width // 2, width
    fx = fy = float(H)
    cx, cy = float(W // 2), float(H // 2)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32) + 0.5, torch.arange(W, dtype=torch.float32) + 0.5, indexing="ij")
    u = (xs - cx) / fx
    v = -(ys - cy) / fy
    theta = -torch.pi * u
    phi = torch.pi * (0.5 - v)
    d_cam = torch.stack([-torch.sin(theta) * torch.sin(phi), torch.cos(phi), -torch.cos(theta) * torch.sin(phi)], -1)
    c2w = torch.tensor([[1.0, 0, 0], [0, 0, 1.0], [0, 1.0, 0]])
    d = (d_cam.reshape(-1, 3) @ c2w.T)
    return d / d.norm(dim=-1, keepdim=True)


# --------------------------------------------------------------------------------------
# End-to-end shading of surface points (BASELINE.json config 2)
# --------------------------------------------------------------------------------------


def shade_points(points, normals, albedo, dirs, radiance, ddf_p, scalings, log2_T, ddf_radius, threshold, sigmoid_scale, only_upper=True):
    """Config 2: one sample per point with weight 1.  Returns linear radiance [N,3] and
    the visibility dict.  radiance [D,3] is the HDR table for the single latent code."""
    v = compute_visibility(points, dirs, ddf_p, scalings, log2_T, ddf_radius, threshold, sigmoid_scale, only_upper)
    rad = lambertian_radiance(albedo, normals, dirs, radiance, v["visibility"])
    return rad, v


# --------------------------------------------------------------------------------------
# End-to-end eval render of a ray bundle (BASELINE.json configs 1 and 3)
# NeuSkyFactoModel.forward -> get_outputs (neusky/models/neusky_model.py:425-443, 738-931), eval mode
# --------------------------------------------------------------------------------------


def pinhole_rays(H: int, W: int, fx: float, fy: float, cx: float, cy: float, c2w: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """nerfstudio Cameras.generate_rays, perspective [NS-mem A.8]: pixel centres at +0.5,
    d_cam = ((x-cx)/fx, -(y-cy)/fy, -1), d_world = R d_cam, directions_norm = |d_world|.
    Returns origins [H*W,3], unit directions [H*W,3], directions_norm [H*W,1] (row-major)."""
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32) + 0.5, torch.arange(W, dtype=torch.float32) + 0.5, indexing="ij")
    d_cam = torch.stack([(xs - cx) / fx, -(ys - cy) / fy, -torch.ones_like(xs)], -1).reshape(-1, 3)
    d = d_cam @ c2w[:3, :3].T
    dn = d.norm(dim=-1, keepdim=True)
    return c2w[:3, 3].expand(H * W, 3).contiguous(), d / dn, dn


def look_at_camera(eye, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)) -> Tensor:
    """c2w [3,4] of a camera at `eye` looking at `target` (OpenGL convention: -z forward, +y up), z-up world."""
    eye, target, up = (torch.tensor(v, dtype=torch.float32) for v in (eye, target, up))
    f = torch.nn.functional.normalize(target - eye, dim=0)
    r = torch.nn.functional.normalize(torch.linalg.cross(f, up), dim=0)
    u = torch.linalg.cross(r, f)
    return torch.cat([torch.stack([r, u, -f], 1), eye[:, None]], 1)


def sphere_collider(origins: Tensor, directions: Tensor, radius: float = 1.0, near_plane: float = 0.05, training: bool = False) -> Tuple[Tensor, Tensor]:
    """nerfstudio SphereCollider (centre 0) [NS-mem A.6] set at neusky_model.py:440."""
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


def uniform_samples(near: Tensor, far: Tensor, S: int) -> Tuple[Tensor, Tensor]:
    """nerfstudio UniformSampler, eval placement (no jitter) [NS-mem A.6]:
    bins = linspace(0,1,S+1); euclid = bins*far + (1-bins)*near.  Returns starts, ends [R,S,1]."""
    bins = torch.linspace(0.0, 1.0, S + 1, dtype=near.dtype, device=near.device)[None]
    e = bins * far + (1 - bins) * near
    return e[:, :-1, None], e[:, 1:, None]


def render_rays(origins, directions, dnorm, S, sdf_p, ddf_p, reni_p, latent, scale, dirs, inv_s, log2_T=19,
                ddf_radius=1.0, threshold=0.1, sigmoid_scale=25.0, rotation=None, chunk=256, proposal_nets=None, proposal_log2_T=17,
                clip_per_chunk=False, steps_minmax=None):
    """Eval render of R rays of ONE camera.  Sample placement: uniform, or -- with ``proposal_nets`` = the state of the two
    HashMLPDensityFields -- the proposal-network sampler (neusky_model.py:561; oracle/sampler_oracle.py).  Returns the outputs
    dict of neusky_model.py:881-931 (rgb, albedo, accumulation, depth, p2p_dist, normal) plus visibility [R,D].
    ``clip_per_chunk``: clip the expected depth to each chunk's own sample range, as the reference's chunked eval loop does."""
    sca = hash_scalings()
    radiance = reni_radiance_table(dirs, latent[None], scale.reshape(1), reni_p, rotation)[0]  # [D,3] (:488-518)
    out = {k: [] for k in ("rgb", "albedo", "accumulation", "depth", "p2p_dist", "normal", "visibility", "weights")}
    R = origins.shape[0]
    near, far = sphere_collider(origins, directions)
    if proposal_nets is not None:
        from . import sampler_oracle as SO

        e, _, _, _ = SO.proposal_sample(origins, directions, near, far, proposal_nets, num_final=S, log2_T=proposal_log2_T)
        e = torch.from_numpy(e)
        starts_all, ends_all = e[:, :-1, None], e[:, 1:, None]
    else:
        starts_all, ends_all = uniform_samples(near, far, S)
    mids = (starts_all + ends_all) / 2
    smin, smax = mids.min(), mids.max()
    if steps_minmax is not None:      # a sample of a larger bundle: clip to the range of the WHOLE bundle the sample was drawn from
        smin, smax = torch.as_tensor(steps_minmax[0], dtype=mids.dtype), torch.as_tensor(steps_minmax[1], dtype=mids.dtype)
    for s in range(0, R, chunk):
        o, d, dn = origins[s:s + chunk], directions[s:s + chunk], dnorm[s:s + chunk]
        starts, ends = starts_all[s:s + chunk], ends_all[s:s + chunk]
        r = o.shape[0]
        x = o[:, None, :] + d[:, None, :] * starts  # get_start_positions (sdf_albedo_field.py:225)
        f = sdf_field(x.reshape(-1, 3), sdf_p, sca, log2_T)
        sdf, grad, alb = f["sdf"].reshape(r, S, 1), f["gradient"].reshape(r, S, 3), f["albedo"].reshape(r, S, 3)
        bg = reni_radiance_table(d, latent[None], scale.reshape(1), reni_p, rotation)[0]  # :535-549
        c = neus_composite(sdf, grad, alb, torch.zeros(r, S, 3), d, starts, ends, ends - starts, bg, dn, inv_s, 1.0, False)
        if clip_per_chunk:
            # the reference's eval loop calls forward() once per `chunk` rays (neusky_model.py:1413-1437, eval_num_rays_per_chunk = 256), so its
            # DepthRenderer clips to the sample range of THAT chunk; the default here (and a one-pass render of the bundle) clips to the bundle's
            cm = (starts + ends) / 2
            smin, smax = cm.min(), cm.max()
        p2p = torch.clip(c["p2p_raw"], smin, smax)   # DepthRenderer clips to the batch-global range [A.7]
        pts = surface_points(o, d, p2p, ddf_radius)
        v = compute_visibility(pts, dirs, ddf_p, sca, log2_T, ddf_radius, threshold, sigmoid_scale)
        normals = torch.nn.functional.normalize(grad, dim=-1)
        rgb = lambertian_render(alb, normals, dirs, radiance, v["visibility"], bg, c["weights"], False)
        for k, t in (("rgb", rgb), ("albedo", c["albedo"]), ("accumulation", c["accumulation"]), ("depth", p2p / dn), ("p2p_dist", p2p),
                     ("normal", c["normal"]), ("visibility", v["visibility"]), ("weights", c["weights"])):
            out[k].append(t)
    return {k: torch.cat(v, 0) for k, v in out.items()}
