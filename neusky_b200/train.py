"""Training-step path (BASELINE config 4): differentiable CUDA ops for the parts of the reference's autograd graph
that carry the cost of a NeuSky training iteration.

The eval path runs the fused forward-only kernels (sky_shade_tc2.cu, sdf_field_tc.cu).  Training needs every layer's
activations again in the backward pass, so this path runs the networks layer by layer on the tcgen05 tf32 GEMM
(csrc/gemm_tf32.cu) with the pointwise stages in its epilogue or in the one-pass kernels of csrc/train_ops.cu, and
keeps activations in HBM (fp32) between forward and backward -- what torch autograd does for the reference, minus the
[R*S, D, 3] light tensors and with the normals' double backward written out analytically.

  ddf_visibility(...)   DDFModel.get_outputs + DirectionalDistanceField + FiLMSiren + the visibility sigmoid of
                        NeuSkyFactoModel.compute_visibility (neusky/models/neusky_model.py:1685-1740), differentiable
                        w.r.t. every DDF parameter, the DDF hash table and the learnable visibility threshold.

`split` selects the GEMM precision: 1 = tf32 (10-bit mantissa operands, fp32 accumulate), 3 = 3xTF32 (fp32-accurate).
There is no torch fallback: every op raises without the CUDA library.
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops

Tensor = torch.Tensor

DDF_MAP_LAYERS = 6      # 5 x (Linear + LeakyReLU 0.2) + Linear -> 2 * 5 * 256   (film_siren.py:45-62)
DDF_TRUNK_LAYERS = 5    # FiLM layers                                           (film_siren.py:107-113)


@dataclass(frozen=True)
class DDFConfig:
    scalings: Tensor
    log2_T: int = 19
    radius: float = 1.0
    sigmoid_scale: float = 25.0
    split: int = 1
    split_bwd: Optional[int] = None      # GEMM precision of the backward contractions (dX, dW); None = same as `split`


def _pad_cols(W: Tensor, mult: int = 8) -> Tensor:
    """Zero-pad the input dimension of a [out, in] weight to a multiple of the tf32 MMA K step."""
    k = W.shape[1]
    kp = (k + mult - 1) // mult * mult
    if kp == k:
        return W.contiguous()
    out = W.new_zeros((W.shape[0], kp))
    out[:, :k] = W
    return out


def ddf_param_list(p: Dict[str, Tensor], prefix: str = "ddf.") -> List[Tensor]:
    """Flatten the reference's DDF state-dict names (directional_distance_field.py:220-243 -> FiLMSiren) into the
    positional order `ddf_visibility` takes: mapping W0,b0..W5,b5, trunk W0,b0..W4,b4."""
    out: List[Tensor] = []
    for i in range(DDF_MAP_LAYERS):
        out += [p[f"{prefix}mapping_network.network.{2 * i}.weight"], p[f"{prefix}mapping_network.network.{2 * i}.bias"]]
    for l in range(DDF_TRUNK_LAYERS):
        out += [p[f"{prefix}net.{l}.layer.weight"], p[f"{prefix}net.{l}.layer.bias"]]
    return out


def _ddf_forward_core(cfg: DDFConfig, cond: Tensor, xin: Tensor, term: Tensor, threshold: Tensor, w_final: Tensor, b_final: Tensor, mlp: Sequence[Tensor]):
    """Mapping network + FiLM-SIREN trunk + head on prepared rows (film_siren.py:45-156, directional_distance_field.py:261-306).
    Returns (that, vis, tensors the backward needs)."""
    if len(mlp) != 2 * (DDF_MAP_LAYERS + DDF_TRUNK_LAYERS):
        raise ValueError(f"DDF network: expected {2 * (DDF_MAP_LAYERS + DDF_TRUNK_LAYERS)} MLP tensors, got {len(mlp)}")
    sp = cfg.split
    Wm, bm = mlp[0:2 * DDF_MAP_LAYERS:2], mlp[1:2 * DDF_MAP_LAYERS:2]
    Wt, bt = mlp[2 * DDF_MAP_LAYERS::2], mlp[2 * DDF_MAP_LAYERS + 1::2]
    h, hs = cond, []
    for i in range(DDF_MAP_LAYERS - 1):
        h = ops.gemm_nt(h, _pad_cols(Wm[i]), bias=bm[i], act="leaky", split=sp)
        hs.append(h)
    film = ops.gemm_nt(h, Wm[-1].contiguous(), bias=bm[-1], split=sp)          # [N, 2560]
    a, zs, acts = xin, [], []
    for l in range(DDF_TRUNK_LAYERS):
        z = ops.gemm_nt(a, _pad_cols(Wt[l]), bias=bt[l], split=sp)
        a = ops.film_sin(z, film, l)
        zs.append(z)
        acts.append(a)
    that, vis = ops.ddf_head(a, w_final, b_final, term, cfg.radius, threshold, cfg.sigmoid_scale)
    return that, vis, (film, Wm, Wt, hs, zs, acts)


def _ddf_backward_core(cfg: DDFConfig, cond, xin, q, term, film, that, threshold, w_final, Wm, Wt, hs, zs, acts, d_vis, d_that,
                       need_table: bool, need_xin: bool, b_shape):
    """Backward of `_ddf_forward_core`: gradients for threshold, hash table, final layer, every mapping / trunk weight and
    (optionally) the trunk input rows xin."""
    sp = cfg.split if cfg.split_bwd is None else cfg.split_bwd
    dev = cond.device
    zeros = lambda *s: torch.zeros(s, device=dev, dtype=torch.float32)

    d_wf, d_bf, d_thr = zeros(256), zeros(1), zeros(1)
    da = ops.ddf_head_bwd(acts[-1], w_final, that, term, None if d_vis is None else d_vis.contiguous(), None if d_that is None else d_that.contiguous(),
                          cfg.radius, threshold, cfg.sigmoid_scale, d_wf, d_bf, d_thr)
    dfilm = torch.empty_like(film)
    dWt: List[Optional[Tensor]] = [None] * DDF_TRUNK_LAYERS
    dbt: List[Optional[Tensor]] = [None] * DDF_TRUNK_LAYERS
    d_xin = None
    db_film = zeros(film.shape[1])          # column sums of dfilm (bias gradient of the last mapping layer): filled block by block below
    for l in reversed(range(DDF_TRUNK_LAYERS)):
        dbt[l] = zeros(256)
        dz = ops.film_sin_bwd(da, zs[l], film, l, dfilm, sum_dz=dbt[l], sum_dfilm=db_film)      # bias column sums in the same pass
        a_prev = acts[l - 1] if l > 0 else xin
        g = zeros(256, a_prev.shape[1])
        ops.gemm_tn(dz, a_prev, g, split=sp)
        dWt[l] = g[:, :Wt[l].shape[1]]
        if l > 0:
            da = ops.gemm_nt(dz, Wt[l].t().contiguous(), split=sp)
        elif need_xin:
            d_xin = ops.gemm_nt(dz, _pad_cols(Wt[0]).t().contiguous(), split=sp)          # [N,16]
    dWm: List[Optional[Tensor]] = [None] * DDF_MAP_LAYERS
    dbm: List[Optional[Tensor]] = [None] * DDF_MAP_LAYERS
    g = zeros(*Wm[-1].shape)
    ops.gemm_tn(dfilm, hs[-1], g, split=sp)
    dWm[-1] = g
    dbm[-1] = db_film
    dz = ops.gemm_nt(dfilm, Wm[-1].t().contiguous(), aux=hs[-1], dact="leaky", split=sp)
    del dfilm
    d_table = None
    for i in reversed(range(DDF_MAP_LAYERS - 1)):
        h_prev = hs[i - 1] if i > 0 else cond
        g = zeros(256, h_prev.shape[1])
        ops.gemm_tn(dz, h_prev, g, split=sp)
        dWm[i] = g[:, :Wm[i].shape[1]]
        dbm[i] = ops.colsum(dz, zeros(256))
        if i > 0:
            dz = ops.gemm_nt(dz, Wm[i].t().contiguous(), aux=hs[i - 1], dact="leaky", split=sp)
        elif need_table:
            L2 = 2 * cfg.scalings.numel()
            dhash = ops.gemm_nt(dz, Wm[0][:, 3:3 + L2].t().contiguous(), split=sp)     # [N, 32]: d cond[:, 3:35]
            d_table = ops.hash_encode_bwd(q, cfg.scalings, cfg.log2_T, dhash)
    grads_mlp: List[Optional[Tensor]] = []
    for i in range(DDF_MAP_LAYERS):
        grads_mlp += [dWm[i], dbm[i]]
    for l in range(DDF_TRUNK_LAYERS):
        grads_mlp += [dWt[l], dbt[l]]
    return d_thr.reshape(threshold.shape), d_table, d_wf.reshape(w_final.shape), d_bf.reshape(b_shape), grads_mlp, d_xin


class _DDFVisibility(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg: DDFConfig, points: Tensor, dirs_sel: Tensor, threshold: Tensor, table: Tensor, w_final: Tensor, b_final: Tensor, *mlp: Tensor):
        R, D = points.shape[0], dirs_sel.shape[0]
        cond, xin, q, term = ops.ddf_pairs(points, dirs_sel, table, cfg.scalings, cfg.log2_T, cfg.radius)
        that, vis, (film, Wm, Wt, hs, zs, acts) = _ddf_forward_core(cfg, cond, xin, term, threshold, w_final, b_final, mlp)
        ctx.cfg = cfg
        ctx.b_shape = b_final.shape
        ctx.n_act = (len(hs), len(zs))
        ctx.save_for_backward(cond, xin, q, term, film, that, threshold, w_final, *Wm, *Wt, *hs, *zs, *acts)
        ctx.mark_non_differentiable(q, term)
        return vis.view(R, D), that, q, term

    @staticmethod
    def backward(ctx, d_vis, d_that, _dq, _dterm):
        sv = ctx.saved_tensors
        cond, xin, q, term, film, that, threshold, w_final = sv[:8]
        o = 8
        Wm = sv[o:o + DDF_MAP_LAYERS]; o += DDF_MAP_LAYERS
        Wt = sv[o:o + DDF_TRUNK_LAYERS]; o += DDF_TRUNK_LAYERS
        hs = sv[o:o + ctx.n_act[0]]; o += ctx.n_act[0]
        zs = sv[o:o + ctx.n_act[1]]; o += ctx.n_act[1]
        acts = sv[o:o + ctx.n_act[1]]
        d_thr, d_table, d_wf, d_bf, grads_mlp, _ = _ddf_backward_core(ctx.cfg, cond, xin, q, term, film, that, threshold, w_final, Wm, Wt, hs, zs, acts,
                                                                      d_vis, d_that, ctx.needs_input_grad[4], False, ctx.b_shape)
        return (None, None, None, d_thr, d_table, d_wf, d_bf, *grads_mlp)


def ddf_visibility(cfg: DDFConfig, points: Tensor, dirs_sel: Tensor, threshold: Tensor, table: Tensor, w_final: Tensor, b_final: Tensor,
                   mlp: Sequence[Tensor]) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """points [R,3] (detached surface points), dirs_sel [D',3] -> (visibility [R,D'], expected termination distance [R*D'],
    sphere points q [R*D',3], termination_dist [R*D']).  Differentiable w.r.t. threshold, the hash table, the final layer and
    every mapping / trunk weight (`mlp` in `ddf_param_list` order)."""
    return _DDFVisibility.apply(cfg, points, dirs_sel, threshold, table, w_final, b_final, *mlp)


def ddf_row_features_torch(origins: Tensor, directions: Tensor) -> Tensor:
    """Local-frame direction and its NeRF encoding for DDF rows in torch ops ([N,15]): only used to carry the cotangent of
    the trunk input back to `directions` when those depend on another network (multi-view batch with
    stop_sdf_gradients=False, ddf_model.py:286-307).  ddf_model.py:158-200, directional_distance_field.py:270-271."""
    y = -origins
    x = torch.stack([-y[:, 1], y[:, 0], torch.zeros_like(y[:, 0])], dim=-1)
    x = x / x.norm(dim=-1, keepdim=True)
    z = torch.linalg.cross(y, x)
    z = z / z.norm(dim=-1, keepdim=True)
    dl = torch.stack([(x * directions).sum(-1), (y * directions).sum(-1), (z * directions).sum(-1)], dim=-1)
    a1 = 2.0 * torch.pi * dl
    ang = torch.stack([a1 * 1.0, a1 * 4.0], dim=-1)                                          # [N,3,2]; no host-built constant tensor (a pageable h2d copy = a sync)
    ang = ang.reshape(-1, 6)
    return torch.cat([dl, torch.sin(ang), torch.sin(ang + torch.pi / 2.0)], dim=-1)


class _DDFRows(torch.autograd.Function):
    """DDFModel.get_outputs -> DirectionalDistanceField.forward on rows (origin on the sphere, world direction):
    expected termination distance [N] (ddf_model.py:193-219)."""

    @staticmethod
    def forward(ctx, cfg: DDFConfig, origins: Tensor, directions: Tensor, table: Tensor, w_final: Tensor, b_final: Tensor, *mlp: Tensor):
        origins, directions = origins.contiguous(), directions.contiguous()
        cond, xin = ops.ddf_rows(origins, directions, table, cfg.scalings, cfg.log2_T)
        term = torch.zeros(origins.shape[0], device=origins.device, dtype=torch.float32)        # visibility head unused here
        thr = torch.zeros((), device=origins.device, dtype=torch.float32)
        that, _vis, (film, Wm, Wt, hs, zs, acts) = _ddf_forward_core(cfg, cond, xin, term, thr, w_final, b_final, mlp)
        ctx.cfg = cfg
        ctx.b_shape = b_final.shape
        ctx.n_act = (len(hs), len(zs))
        ctx.save_for_backward(cond, xin, origins, directions, term, film, that, thr, w_final, *Wm, *Wt, *hs, *zs, *acts)
        return that

    @staticmethod
    def backward(ctx, d_that):
        sv = ctx.saved_tensors
        cond, xin, origins, directions, term, film, that, thr, w_final = sv[:9]
        o = 9
        Wm = sv[o:o + DDF_MAP_LAYERS]; o += DDF_MAP_LAYERS
        Wt = sv[o:o + DDF_TRUNK_LAYERS]; o += DDF_TRUNK_LAYERS
        hs = sv[o:o + ctx.n_act[0]]; o += ctx.n_act[0]
        zs = sv[o:o + ctx.n_act[1]]; o += ctx.n_act[1]
        acts = sv[o:o + ctx.n_act[1]]
        need_dir = ctx.needs_input_grad[2]
        _, d_table, d_wf, d_bf, grads_mlp, d_xin = _ddf_backward_core(ctx.cfg, cond, xin, origins, term, film, that, thr, w_final, Wm, Wt, hs, zs, acts,
                                                                      None, d_that, ctx.needs_input_grad[3], need_dir, ctx.b_shape)
        d_dir = None
        if need_dir:
            with torch.enable_grad():
                dd = directions.detach().requires_grad_(True)
                feat = ddf_row_features_torch(origins.detach(), dd)
                (d_dir,) = torch.autograd.grad(feat, dd, d_xin[:, :15])
        return (None, None, d_dir, d_table, d_wf, d_bf, *grads_mlp)


def ddf_termination(cfg: DDFConfig, origins: Tensor, directions: Tensor, table: Tensor, w_final: Tensor, b_final: Tensor, mlp: Sequence[Tensor]) -> Tensor:
    """Row-wise DDF: origins [N,3] on the sphere, world directions [N,3] -> expected termination distance [N].
    Differentiable w.r.t. the DDF hash table, final layer, every mapping / trunk weight, and `directions`."""
    return _DDFRows.apply(cfg, origins, directions, table, w_final, b_final, *mlp)


# =====================================================================================================================
# SDF / albedo field (neusky/fields/sdf_albedo_field.py:211-269; nerfstudio SDFField.forward_geonetwork)
# =====================================================================================================================
@dataclass(frozen=True)
class SDFConfig:
    scalings: Tensor
    log2_T: int = 19
    split_geo: int = 3       # the sdf feeds the NeuS logistic CDF scaled by inv_s: keep the geometry network fp32-accurate
    split_colour: int = 1


def sdf_param_list(p: Dict[str, Tensor]) -> List[Tensor]:
    """weight_norm-folded weights in the positional order `sdf_field` takes: glin0..2 (W, b), clin0..2 (W, b).  The fold
    W = g v / |v| (nn.utils.weight_norm, dim=0) is done here with torch ops so autograd carries the gradient back to
    weight_g / weight_v; it touches 6 small matrices per step."""
    out: List[Tensor] = []
    for name in ("glin0", "glin1", "glin2", "clin0", "clin1", "clin2"):
        if name + ".weight" in p:
            W = p[name + ".weight"]
        else:
            v, g = p[name + ".weight_v"], p[name + ".weight_g"]
            W = v * (g / v.norm(dim=1, keepdim=True))
        out += [W, p[name + ".bias"]]
    return out


class _SDFField(torch.autograd.Function):
    """sdf [n], grad_x sdf [n,3] (analytic reverse pass instead of torch.autograd.grad, sdf_albedo_field.py:235-238) and
    albedo [n,3]; backward for cotangents on all three, including the double backward through the reverse pass that the
    eikonal loss and the rendered normals need (the reference gets it from create_graph=True)."""

    @staticmethod
    def forward(ctx, cfg: SDFConfig, want_normals: bool, want_albedo: bool, x: Tensor, table: Tensor, W0, b0, W1, b1, W2, b2, C0, c0, C1, c1, C2, c2):
        sg, sc = cfg.split_geo, cfg.split_colour
        n, dev = x.shape[0], x.device
        x = x.contiguous()
        Hc = torch.empty((n, 296), device=dev, dtype=torch.float32) if want_albedo else None
        H0, pos, J = ops.sdf_inputs(x, table, cfg.scalings, cfg.log2_T, tail=None if Hc is None else Hc[:, 256:296])
        W0p = _pad_cols(W0)                                        # [256, 72]
        A1 = ops.gemm_nt(H0, W0p, bias=b0, act="softplus100", split=sg)
        A2 = ops.gemm_nt(A1, W1.contiguous(), bias=b1, act="softplus100", split=sg)
        w2s = W2[0].contiguous()
        sdf = ops.rowdot256(A2, w2s, b2[0:1])
        empty = x.new_zeros(0)
        Ca1 = Ca2 = alb = C0p = empty
        if want_albedo:
            ops.gemm_nt(A2, W2[1:].contiguous(), bias=b2[1:].contiguous(), out=Hc[:, :256], split=sg)
            # colour-net input reordered to (geo | x | PE | 0) so the geometry feature lands 16-byte aligned; C0's columns follow
            C0p = torch.cat([C0[:, 39:295], C0[:, :39], C0.new_zeros((C0.shape[0], 1))], dim=1).contiguous()
            Ca1 = ops.gemm_nt(Hc, C0p, bias=c0, act="relu", split=sc)
            Ca2 = ops.gemm_nt(Ca1, C1.contiguous(), bias=c1, act="relu", split=sc)
            alb = ops.gemm_nt(Ca2, C2.contiguous(), bias=c2, act="sigmoid", split=sc)
        G2 = P1 = D1 = G0 = grad = empty
        if want_normals:
            G2 = ops.ew256("sp_chain", n, a=A2, w=w2s)
            P1 = ops.gemm_nt(G2, W1.t().contiguous(), split=sg)
            D1 = ops.ew256("mul_dsp", n, a=P1, b=A1)
            G0 = ops.gemm_nt(D1, W0p.t().contiguous(), split=sg)    # [n, 72] = d sdf / d H0
            gpos = ops.hash_encode_grad_x(pos, table, cfg.scalings, cfg.log2_T, G0[:, 39:71].contiguous())
            grad = ops.sdf_grad_assemble(x, G0, gpos, J)
        ctx.cfg, ctx.want = cfg, (want_normals, want_albedo)
        ctx.save_for_backward(x, table, pos, J, H0, A1, A2, W0p, W1, W2, C0p, C1, C2, Hc if Hc is not None else empty, Ca1, Ca2, alb, G2, P1, D1, G0)
        return sdf, grad, alb

    @staticmethod
    def backward(ctx, g_sdf, g_grad, g_alb):
        cfg: SDFConfig = ctx.cfg
        sg, sc = cfg.split_geo, cfg.split_colour
        want_normals, want_albedo = ctx.want
        x, table, pos, J, H0, A1, A2, W0p, W1, W2, C0p, C1, C2, Hc, Ca1, Ca2, alb, G2, P1, D1, G0 = ctx.saved_tensors
        n, dev = x.shape[0], x.device
        zeros = lambda *s: torch.zeros(s, device=dev, dtype=torch.float32)
        w2s = W2[0].contiguous()
        need_table = ctx.needs_input_grad[4]
        d_table = zeros(*table.shape) if need_table else None
        dW0p, dW1, dW2 = zeros(*W0p.shape), zeros(*W1.shape), zeros(*W2.shape)
        db0, db1, db2 = zeros(256), zeros(256), zeros(W2.shape[0])
        dC0 = dc0 = dC1 = dc1 = dC2 = dc2 = None

        # ---- colour network -> d geo feature -> d A2
        dA2 = None
        if want_albedo and g_alb is not None:
            dz3 = zeros(n, 8)
            dz3[:, :3] = g_alb * alb * (1.0 - alb)                                   # sigmoid'
            dC2 = ops.gemm_tn(dz3[:, :3], Ca2, zeros(3, 256), split=sc)
            dc2 = ops.colsum(dz3[:, :3], zeros(3))
            C2t = zeros(256, 8)
            C2t[:, :3] = C2.t()
            dz2 = ops.gemm_nt(dz3, C2t, aux=Ca2, dact="relu", split=sc)              # [n,256]
            dC1 = ops.gemm_tn(dz2, Ca1, zeros(256, 256), split=sc)
            dc1 = ops.colsum(dz2, zeros(256))
            dz1 = ops.gemm_nt(dz2, C1.t().contiguous(), aux=Ca1, dact="relu", split=sc)
            dC0p = ops.gemm_tn(dz1, Hc, zeros(256, 296), split=sc)
            dc0 = ops.colsum(dz1, zeros(256))
            dC0 = torch.cat([dC0p[:, 256:295], dC0p[:, :256]], dim=1)                # back to the reference's (x | PE | geo) order
            dgeo = ops.gemm_nt(dz1, C0p[:, :256].t().contiguous(), split=sc)         # d Hc[:, :256]
            ops.gemm_tn(dgeo, A2, dW2[1:], split=sg)
            ops.colsum(dgeo, db2[1:])
            dA2 = ops.gemm_nt(dgeo, W2[1:].t().contiguous(), split=sg)
        # ---- sdf head
        if g_sdf is not None:
            gs = g_sdf.contiguous()
            dA2 = ops.ew256("outer_add", n, a=dA2, w=w2s, s=gs)
            ops.colsum_w(A2, gs, dW2[0])
            db2[0:1] += gs.sum().reshape(1)
        # ---- double backward through the reverse pass (cotangent on grad_x sdf)
        dD1 = dG2 = None
        if want_normals and g_grad is not None:
            dG0, cpos = ops.sdf_grad_assemble_bwd(x, g_grad.contiguous(), J)
            d_g, d_t2 = ops.hash_encode_grad_x_bwd(pos, table, cfg.scalings, cfg.log2_T, G0[:, 39:71].contiguous(), cpos, True, need_table)
            dG0[:, 39:71] = d_g
            if need_table:
                d_table = d_t2                                                        # freshly zero-filled by the op: reuse as the accumulator
            dD1 = ops.gemm_nt(dG0, W0p, split=sg)                                     # [n,256] = dG0 . W0p^T
            ops.gemm_tn(D1, dG0, dW0p, split=sg)
            dP1 = ops.ew256("mul_dsp", n, a=dD1, b=A1)
            dG2 = ops.gemm_nt(dP1, W1.contiguous(), split=sg)                         # dG2[i] = sum_j dP1[j] W1[i,j]
            ops.gemm_tn(G2, dP1, dW1, split=sg)
            ops.colsum(ops.ew256("mul_dsp", n, a=dG2, b=A2), dW2[0])                  # G2 = w2s * s'(A2)
        # ---- geometry network
        if dA2 is None and dG2 is None:
            dZ2 = None
        else:
            dZ2 = ops.ew256("sp_bwd2_w", n, a=dA2, b=A2, c=dG2, w=w2s)
        d_x = None
        if dZ2 is not None:
            ops.gemm_tn(dZ2, A1, dW1, split=sg)
            ops.colsum(dZ2, db1)
            dA1 = ops.gemm_nt(dZ2, W1.t().contiguous(), split=sg)
            dZ1 = ops.ew256("sp_bwd2", n, a=dA1, b=A1, c=dD1, d=P1 if dD1 is not None else None)
        elif dD1 is not None:
            dZ1 = ops.ew256("sp_bwd2", n, a=None, b=A1, c=dD1, d=P1)
        else:
            dZ1 = None
        if dZ1 is not None:
            ops.gemm_tn(dZ1, H0, dW0p, split=sg)
            ops.colsum(dZ1, db0)
            if need_table or ctx.needs_input_grad[3]:
                dH0 = ops.gemm_nt(dZ1, W0p.t().contiguous(), split=sg)                # [n,72]
                dH0h = dH0[:, 39:71].contiguous()
                if need_table:
                    ops.hash_encode_bwd(pos, cfg.scalings, cfg.log2_T, dH0h, d_table)
                if ctx.needs_input_grad[3]:
                    # first-order input gradient (positions that depend on other networks, e.g. sdf_at_termination); the
                    # second-order d(grad_x)/dx term is not propagated (the reference's sample positions are detached)
                    d_x = ops.sdf_grad_assemble(x, dH0, ops.hash_encode_grad_x(pos, table, cfg.scalings, cfg.log2_T, dH0h), J)
        dW0 = dW0p[:, :71]
        return (None, None, None, d_x, d_table, dW0, db0, dW1, db1, dW2, db2, dC0, dc0, dC1, dc1, dC2, dc2)


def sdf_field(cfg: SDFConfig, x: Tensor, table: Tensor, weights: Sequence[Tensor], want_normals: bool = True, want_albedo: bool = True) -> Tuple[Tensor, Tensor, Tensor]:
    """x [n,3] -> (sdf [n], gradient [n,3], albedo [n,3]); `weights` in `sdf_param_list` order.  Differentiable w.r.t. the
    hash table and all weights (twice through the gradient output), and to first order w.r.t. x."""
    return _SDFField.apply(cfg, want_normals, want_albedo, x, table, *weights)


# =====================================================================================================================
# One training iteration: NeuSkyFactoModel.get_outputs (training) + get_loss_dict   (neusky_model.py:553-1068)
# =====================================================================================================================
LOSS_COEFFICIENTS = {  # neusky/configs/neusky_config.py:132-146
    "rgb_l1_loss": 1.0, "eikonal_loss": 0.1, "fg_mask_loss": 1.0, "sdf_level_set_visibility_loss": 1.0, "sky_pixel_loss": 1.0,
    "hashgrid_density_loss": 1e-4, "ground_plane_loss": 0.1, "visibility_sigmoid_loss": 0.01, "interlevel_loss": 1.0,
}


def _monosdf_normal_loss(normal_pred: Tensor, normal_gt: Tensor) -> Tensor:
    normal_gt = torch.nn.functional.normalize(normal_gt, p=2, dim=-1)
    normal_pred = torch.nn.functional.normalize(normal_pred, p=2, dim=-1)
    return torch.abs(normal_pred - normal_gt).sum(dim=-1).mean() + (1.0 - torch.sum(normal_pred * normal_gt, dim=-1)).mean()


def sky_pixel_loss(inputs: Tensor, targets: Tensor, mask: Tensor, alpha: float = 0.1) -> Tensor:
    """RENISkyPixelLoss (neusky/model_components/losses.py:44-58): MSE + alpha * (1 - mean cosine similarity) on masked pixels;
    alpha = cosine_weight of neusky_config.py (0.1)."""
    a, b = inputs * mask, targets * mask
    return torch.nn.functional.mse_loss(a, b) + alpha * (1 - torch.nn.functional.cosine_similarity(a, b, dim=1, eps=1e-20).mean())


def _linear_to_srgb(c: Tensor) -> Tensor:
    c = torch.where(c <= 0.0031308, 12.92 * c, 1.055 * torch.pow(torch.abs(c), 1 / 2.4) - 0.055)
    return torch.clamp(c, 0.0, 1.0)


def _neus_alpha(sdf, grad, ray_dirs, deltas, inv_s, rho: float):
    """SDFField.get_alpha on a handful of grid samples (hashgrid_density_loss, neusky_model.py:675-734): torch glue."""
    true_cos = (ray_dirs * grad).sum(-1, keepdim=True)
    iter_cos = -(torch.relu(-true_cos * 0.5 + 0.5) * (1.0 - rho) + torch.relu(-true_cos) * rho)
    nxt, prv = sdf + iter_cos * deltas * 0.5, sdf - iter_cos * deltas * 0.5
    p, q = torch.sigmoid(prv * inv_s), torch.sigmoid(nxt * inv_s)
    return ((p - q + 1e-5) / (p + 1e-5)).clip(0.0, 1.0)


class NeuSkyTrainStep(torch.nn.Module):
    """Host mirror of one `ns-train neusky` iteration for a ray batch: the reference's NeuSkyFactoModel.forward in training
    mode plus get_loss_dict, with every heavy stage on the CUDA ops of this package.  Parameters carry the reference's
    state-dict names (dots replaced by '__' for nn.Module registration) and its optimizer groups
    (neusky_model.py:379-398): fields, ddf_field, illumination latents, visibility_sigmoid.

    The RENI++ decoder is frozen (fixed_decoder=True in the reference, neusky_config.py:94); the per-image latent codes and
    scales feeding it are trained through csrc/reni_decode_bwd.cu."""

    def __init__(self, sdf_params: Dict[str, Tensor], ddf_params: Dict[str, Tensor], reni_params: Dict[str, Tensor], num_cameras: int, device="cuda",
                 log2_T: int = 19, num_levels: int = 16, num_samples: int = 48, ddf_radius: float = 1.0, sigmoid_scale: float = 25.0,
                 split_geo: int = 3, split: int = 3, ddf_split_bwd: Optional[int] = None, threshold_init: Optional[float] = None, only_upper_hemisphere: bool = True,
                 lower_hemisphere_visibility: float = 1.0, proposal_params: Optional[Sequence[Dict[str, Tensor]]] = None,
                 proposal_max_res: Sequence[int] = (64, 256), num_proposal_samples_per_ray: Sequence[int] = (256, 96), proposal_log2_T: int = 17,
                 share_params: bool = False, latents: Optional[torch.nn.Parameter] = None, scale: Optional[torch.nn.Parameter] = None,
                 visibility_threshold: Optional[torch.nn.Parameter] = None, proposal_fields: Optional[Sequence] = None, ddf_log2_T: Optional[int] = None):
        """``proposal_params``: state of the two HashMLPDensityFields -> the shipped NeuS-facto sample placement (proposal-network
        sampler, neusky_model.py:561) and its interlevel loss (:987-988, coefficient 1.0); None -> uniform placement.
        ``share_params``: register the ``nn.Parameter`` objects passed in ``sdf_params`` / ``ddf_params`` themselves instead of
        copies (the drop-in model, neusky_b200/models.py, owns the parameters under the reference's module names and runs its
        training forward through this class); ``latents`` / ``scale`` / ``visibility_threshold`` / ``proposal_fields`` likewise.
        ``ddf_split_bwd``: GEMM precision of the DDF's backward contractions only (1 = single-pass tf32 dX / dW with the forward
        still at ``split``); None keeps them at ``split``."""
        super().__init__()
        from . import packing
        from .init import hash_scalings

        self.dev = torch.device(device)
        self.S, self.radius, self.log2_T = num_samples, float(ddf_radius), log2_T
        self.only_upper, self.lower_vis = only_upper_hemisphere, float(lower_hemisphere_visibility)
        self.scalings = hash_scalings(num_levels).to(self.dev)
        self._names: Dict[str, Dict[str, str]] = {"sdf": {}, "ddf": {}}
        for grp, params in (("sdf", sdf_params), ("ddf", ddf_params)):
            for k, v in params.items():
                reg = f"{grp}__{k.replace('.', '__')}"
                if share_params:
                    if not isinstance(v, torch.nn.Parameter) or v.device != self.dev or v.dtype != torch.float32:
                        raise ValueError(f"share_params: {grp}.{k} must be an fp32 nn.Parameter on {self.dev}")
                    self.register_parameter(reg, v)
                else:
                    self.register_parameter(reg, torch.nn.Parameter(v.detach().to(self.dev, torch.float32).clone().contiguous()))
                self._names[grp][k] = reg
        self.latents = latents if latents is not None else torch.nn.Parameter(torch.zeros(num_cameras, 100, 3, device=self.dev))      # neusky_model.py:261-269 (zero init)
        self.scale = scale if scale is not None else torch.nn.Parameter(torch.zeros(num_cameras, device=self.dev))
        thr0 = 2.0 * self.radius if threshold_init is None else threshold_init                    # :234
        self.visibility_threshold = visibility_threshold if visibility_threshold is not None else torch.nn.Parameter(torch.tensor(float(thr0), device=self.dev))
        self.reni_blob = packing.pack_reni(reni_params, device=self.dev)
        self.reni_blob_bwd = packing.pack_reni_bwd(reni_params, device=self.dev)
        self.sdf_cfg = SDFConfig(scalings=self.scalings, log2_T=log2_T, split_geo=split_geo, split_colour=split)
        self.ddf_cfg = DDFConfig(scalings=self.scalings, log2_T=log2_T if ddf_log2_T is None else ddf_log2_T, radius=self.radius, sigmoid_scale=sigmoid_scale, split=split, split_bwd=ddf_split_bwd)
        self.cos_anneal_ratio = 1.0
        self.aux_stream: Optional[torch.cuda.Stream] = None      # side stream for the RENI++ branch of forward (None = in line)
        self.grid_resolution = 10                                                                  # neusky_config.py:127
        self.proposal_fields, self.proposal_sampler, self.proposal_anneal = None, None, 1.0
        if proposal_params is not None or proposal_fields is not None:
            from . import proposal as _proposal
            self.proposal_fields = list(proposal_fields) if proposal_fields is not None else [
                _proposal.HashMLPDensityField(pp, max_res=mr, log2_hashmap_size=proposal_log2_T, device=self.dev) for pp, mr in zip(proposal_params, proposal_max_res)]
            for i, f in enumerate(self.proposal_fields):          # the fields' own Parameter objects, registered here for optimizers / reducers
                for k, v in f.params.items():
                    self.register_parameter(f"prop{i}__{k.replace('.', '__')}", v)
            self.proposal_sampler = _proposal.ProposalNetworkSampler(num_samples, num_proposal_samples_per_ray, len(self.proposal_fields))
            self.proposal_sampler.training = True

    # -- parameter access under the reference's names -------------------------------------------------------
    def group(self, grp: str) -> Dict[str, Tensor]:
        return {k: getattr(self, reg) for k, reg in self._names[grp].items()}

    def sdf_weights(self) -> List[Tensor]:
        """`sdf_param_list` of the current SDF parameters (weight-norm fold), computed once per optimizer state: the main pass,
        the sdf_at_termination branch, the density probe and the DDF fitting pass all evaluate the same field in one iteration,
        and every fold is ~25 small torch ops forward plus their backward.  Keyed on the parameters' version counters, so an
        optimizer step (in-place update) invalidates it."""
        sdf_p = self.group("sdf")
        key = tuple(v._version for v in sdf_p.values()) + (torch.is_grad_enabled(),)
        if getattr(self, "_sdf_w_key", None) != key:
            self._sdf_w_key, self._sdf_w = key, sdf_param_list(sdf_p)
            for w in self._sdf_w:
                if w.requires_grad and not w.is_leaf:      # the fold's graph is freed by the first backward through it: drop the cache then
                    w.register_hook(self._drop_sdf_weights)
                    break
        return self._sdf_w

    def _drop_sdf_weights(self, grad):
        self._sdf_w_key, self._sdf_w = None, None          # also releases the fold's autograd nodes (they pin the parameters' AccumulateGrad nodes)
        return None

    def get_param_groups(self) -> Dict[str, List[torch.nn.Parameter]]:
        g = {"fields": list(self.group("sdf").values()), "ddf_field": list(self.group("ddf").values()),
             "illumination_field": [self.latents, self.scale], "visibility_sigmoid": [self.visibility_threshold]}
        if self.proposal_fields is not None:
            g["proposal_networks"] = [p for f in self.proposal_fields for p in f.parameters()]
        return g

    def set_directions(self, dirs: Tensor) -> None:
        """Illumination directions of this iteration [D,3] (IcosahedronSampler with its random rotation, neusky_model.py:452-456)."""
        d0, mask_u8, dirs_sel, sel = self.compact_directions(dirs)
        nb = d0.device.type == "cpu"
        self.dirs = d0.to(self.dev, non_blocking=nb)
        self.mask_u8 = mask_u8.to(self.dev, non_blocking=nb)
        self.dirs_sel = dirs_sel.to(self.dev, non_blocking=nb)
        self.sel_index = sel.to(self.dev, non_blocking=nb)

    def compact_directions(self, dirs: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        """dirs [D,3] -> (dirs fp32 [D,3], upper-hemisphere mask uint8 [D], the D' selected directions [D',3], index of every direction
        in that selection or -1, int32 [D]) on the device `dirs` lives on (the host for the icosphere sampler: no sync)."""
        d0 = dirs.to(torch.float32).contiguous()          # mask / compaction on the device the sampler produced them on (host: no sync)
        m = (d0[:, 2] > 0) if self.only_upper else torch.ones(d0.shape[0], dtype=torch.bool, device=d0.device)
        sel = torch.where(m, torch.cumsum(m.to(torch.int32), 0, dtype=torch.int32) - 1, torch.full_like(m, -1, dtype=torch.int32)).to(torch.int32)
        return d0, m.to(torch.uint8).contiguous(), d0[m].contiguous(), sel.contiguous()

    # -- forward ---------------------------------------------------------------------------------------------
    def forward(self, batch: Dict[str, Tensor], grid_positions: Optional[Tensor] = None, grid_dirs: Optional[Tensor] = None) -> Tuple[Tensor, Dict[str, Tensor], Dict[str, Tensor]]:
        from . import autograd as nba
        from .render import sphere_collider, uniform_samples

        o, d, dn = batch["origins"], batch["directions"], batch["dnorm"]
        cam = batch["cam"].to(torch.int32).contiguous()
        R, S = o.shape[0], self.S
        sdf_p, ddf_p = self.group("sdf"), self.group("ddf")
        sdf_w = self.sdf_weights()
        table = sdf_p["encoding.hash_table"]

        # RENI++ radiance of every (camera, light direction) pair and along each camera ray; differentiable w.r.t. the per-image
        # latent codes and scales, decoder frozen (neusky_model.py:261-269, 488-504, 535-549; neusky_config.py:94).  Independent of the
        # geometry until the shading sum, and small (K x D + R rows on fp32 SIMT kernels: one wave of 128 blocks for the per-ray rows):
        # with `aux_stream` set (graphed.GraphedTrainIteration) it runs as a parallel branch next to the SDF field instead of in line.
        inv_s = torch.exp(sdf_p["deviation_network.variance"] * 10.0).clip(1e-6, 1e6)
        aux, cur = self.aux_stream, None
        if aux is not None:
            cur = torch.cuda.current_stream()
            aux.wait_stream(cur)
        grid_density = None
        with torch.cuda.stream(aux) if aux is not None else contextlib.nullcontext():
            radiance = nba.reni_radiance(self.dirs, self.latents, self.scale, self.reni_blob, self.reni_blob_bwd)                 # [K,D,3]
            bg = nba.reni_radiance(d.contiguous(), self.latents, self.scale, self.reni_blob, self.reni_blob_bwd, row_cam=cam)     # [R,3]
            if grid_positions is not None:
                # hashgrid density probe (neusky_model.py:675-734): 1000 grid points through the geometry network, ~30 tiny kernels forward
                # + backward -- the same branch
                gs, gg, _ = sdf_field(self.sdf_cfg, grid_positions, table, sdf_w, want_normals=True, want_albedo=False)
                gap = 2.0 / self.grid_resolution
                grid_density = _neus_alpha(gs[:, None], gg, grid_dirs, torch.full_like(gs[:, None], gap), inv_s, self.cos_anneal_ratio)
        if aux is not None:
            for t in (inv_s, *sdf_w):
                t.record_stream(aux)

        near, far = sphere_collider(o, d, radius=1.0, training=True)
        rs = None
        if self.proposal_fields is not None:
            # proposal-network sampler (training mode: one jitter per ray and level); placement carries no gradient, the proposal
            # networks learn from the interlevel loss below
            for f in self.proposal_fields:
                f.refresh()
            self.proposal_sampler.set_anneal(self.proposal_anneal)
            with torch.no_grad():
                rs, _wl, _sl = self.proposal_sampler.generate_ray_samples(o, d, near, far, self.proposal_fields, jitters=batch.get("jitters"))
            e = rs.euclidean_bins
            starts, ends = e[:, :-1].contiguous(), e[:, 1:].contiguous()
        else:
            starts, ends = uniform_samples(near, far, S)                               # [R,S]
        x = (o[:, None, :] + d[:, None, :] * starts[:, :, None]).reshape(-1, 3)
        sdf, grad, alb = sdf_field(self.sdf_cfg, x, table, sdf_w)
        starts2, ends2 = starts.reshape(R, S), ends.reshape(R, S)
        weights, wa, normals, acc, p2p_raw, normal, _albedo, _bgT = nba.neus_composite(
            sdf.reshape(R, S), grad.reshape(R, S, 3), alb.reshape(R, S, 3), inv_s, d, starts2, ends2, ends2 - starts2, dn.reshape(R), self.cos_anneal_ratio)
        steps = (starts2 + ends2) * 0.5
        p2p = torch.clip(p2p_raw.detach(), steps.min(), steps.max())                   # DepthRenderer clip; detached (stop-gradients "depth")
        pts = ops.surface_points(o, d, p2p, self.radius)

        if aux is not None:                                        # join the RENI++ branch before its results are read on this stream
            cur.wait_stream(aux)
            for t in (radiance, bg) + (() if grid_density is None else (grid_density,)):
                t.record_stream(cur)

        vis, that, q, _term = ddf_visibility(self.ddf_cfg, pts, self.dirs_sel, self.visibility_threshold, ddf_p["position_encoding.hash_table"],
                                             ddf_p["ddf.final_layer.weight"], ddf_p["ddf.final_layer.bias"], ddf_param_list(ddf_p))
        Dp = self.dirs_sel.shape[0]
        term_pts = q + (-self.dirs_sel)[None].expand(R, Dp, 3).reshape(-1, 3) * that[:, None]          # ddf_model.py:243
        sdf_term, _, _ = sdf_field(self.sdf_cfg, term_pts, table, sdf_w, want_normals=False, want_albedo=False)

        inv_count, _ = ops.lambert_prep(normals.detach(), wa.detach(), self.dirs, self.mask_u8, radiance.detach(), cam, self.lower_vis)
        rgb_lin = nba.lambert_shade(normals, wa, radiance, vis, inv_count, self.dirs, self.sel_index, cam, self.lower_vis)
        rgb = nba.shade_finalize(rgb_lin, bg, acc)
        out = {"rgb": rgb, "eik_grad": grad.reshape(R, S, 3), "weights": weights, "normal": normal, "accumulation": acc, "hdr_background_colours": bg,
               "p2p_dist": p2p, "sdf_at_termination": sdf_term, "visibility_sel": vis, "expected_termination_dist": that}
        if rs is not None:
            # nerfstudio interlevel_loss with the fine NeuS weights appended (neusky_model.py:575-576, 987-988), through autograd into the
            # proposal networks (weights -> density -> MLP + hash table; the fine histogram is detached)
            out["interlevel_loss"] = self.proposal_sampler.interlevel_loss(weights.detach(), rs, o, d, near, far)
            out["starts"], out["ends"] = starts, ends
        if grid_density is not None:
            out["grid_density"] = grid_density
        losses = self.get_loss_dict(out, batch)
        return sum(losses.values()), losses, out

    def get_loss_dict(self, out: Dict[str, Tensor], batch: Dict[str, Tensor]) -> Dict[str, Tensor]:
        """neusky_model.py:935-1031, training branch; per-ray reductions on [R, .] tensors (torch)."""
        image, fg, ground, sky = batch["image"], batch["fg"], batch["ground"], batch["sky"]
        keep = (1.0 - sky.to(image.dtype))[:, None]
        L: Dict[str, Tensor] = {}
        L["rgb_l1_loss"] = torch.nn.functional.l1_loss(image * keep, out["rgb"] * keep)
        L["eikonal_loss"] = ((out["eik_grad"].norm(2, dim=-1) - 1) ** 2).mean()
        ws = out["accumulation"].clip(1e-3, 1.0 - 1e-3)
        L["fg_mask_loss"] = torch.nn.functional.binary_cross_entropy(ws, fg.to(ws.dtype))
        gm = ground.to(image.dtype)[:, None]
        up = torch.zeros_like(out["normal"])
        up[:, 2] = 1.0                                                                  # built on the device: no pageable h2d copy (a host sync)
        L["ground_plane_loss"] = _monosdf_normal_loss(out["normal"] * gm, up * gm)
        srgb_bg = _linear_to_srgb(out["hdr_background_colours"])
        L["sky_pixel_loss"] = sky_pixel_loss(srgb_bg, image, sky.to(image.dtype)[:, None].expand_as(srgb_bg))
        L["visibility_sigmoid_loss"] = (self.visibility_threshold - 0.1) ** 2
        L["sdf_level_set_visibility_loss"] = (out["sdf_at_termination"] ** 2).mean()
        if "grid_density" in out:
            L["hashgrid_density_loss"] = out["grid_density"].abs().mean()
        if "interlevel_loss" in out:
            L["interlevel_loss"] = out["interlevel_loss"]
        return {k: v * LOSS_COEFFICIENTS[k] for k, v in L.items()}
