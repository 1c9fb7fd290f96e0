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


class _DDFVisibility(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg: DDFConfig, points: Tensor, dirs_sel: Tensor, threshold: Tensor, table: Tensor, w_final: Tensor, b_final: Tensor, *mlp: Tensor):
        if len(mlp) != 2 * (DDF_MAP_LAYERS + DDF_TRUNK_LAYERS):
            raise ValueError(f"ddf_visibility: expected {2 * (DDF_MAP_LAYERS + DDF_TRUNK_LAYERS)} MLP tensors, got {len(mlp)}")
        sp = cfg.split
        Wm, bm = mlp[0:2 * DDF_MAP_LAYERS:2], mlp[1:2 * DDF_MAP_LAYERS:2]
        Wt, bt = mlp[2 * DDF_MAP_LAYERS::2], mlp[2 * DDF_MAP_LAYERS + 1::2]
        R, D = points.shape[0], dirs_sel.shape[0]
        cond, xin, q, term = ops.ddf_pairs(points, dirs_sel, table, cfg.scalings, cfg.log2_T, cfg.radius)
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
        ctx.cfg = cfg
        ctx.b_shape = b_final.shape
        ctx.n_act = (len(hs), len(zs))
        ctx.save_for_backward(cond, xin, q, term, film, that, threshold, w_final, *Wm, *Wt, *hs, *zs, *acts)
        ctx.mark_non_differentiable(q, term)
        return vis.view(R, D), that, q, term

    @staticmethod
    def backward(ctx, d_vis, d_that, _dq, _dterm):
        cfg: DDFConfig = ctx.cfg
        sp = cfg.split
        sv = ctx.saved_tensors
        cond, xin, q, term, film, that, threshold, w_final = sv[:8]
        o = 8
        Wm = sv[o:o + DDF_MAP_LAYERS]; o += DDF_MAP_LAYERS
        Wt = sv[o:o + DDF_TRUNK_LAYERS]; o += DDF_TRUNK_LAYERS
        hs = sv[o:o + ctx.n_act[0]]; o += ctx.n_act[0]
        zs = sv[o:o + ctx.n_act[1]]; o += ctx.n_act[1]
        acts = sv[o:o + ctx.n_act[1]]
        dev = cond.device
        zeros = lambda *s: torch.zeros(s, device=dev, dtype=torch.float32)

        d_wf, d_bf, d_thr = zeros(256), zeros(1), zeros(1)
        da = ops.ddf_head_bwd(acts[-1], w_final, that, term, None if d_vis is None else d_vis.contiguous(), None if d_that is None else d_that.contiguous(),
                              cfg.radius, threshold, cfg.sigmoid_scale, d_wf, d_bf, d_thr)
        dfilm = torch.empty_like(film)
        dWt: List[Optional[Tensor]] = [None] * DDF_TRUNK_LAYERS
        dbt: List[Optional[Tensor]] = [None] * DDF_TRUNK_LAYERS
        for l in reversed(range(DDF_TRUNK_LAYERS)):
            dz = ops.film_sin_bwd(da, zs[l], film, l, dfilm)
            a_prev = acts[l - 1] if l > 0 else xin
            g = zeros(256, a_prev.shape[1])
            ops.gemm_tn(dz, a_prev, g, split=sp)
            dWt[l] = g[:, :Wt[l].shape[1]]
            dbt[l] = ops.colsum(dz, zeros(256))
            if l > 0:
                da = ops.gemm_nt(dz, Wt[l].t().contiguous(), split=sp)
        dWm: List[Optional[Tensor]] = [None] * DDF_MAP_LAYERS
        dbm: List[Optional[Tensor]] = [None] * DDF_MAP_LAYERS
        g = zeros(*Wm[-1].shape)
        ops.gemm_tn(dfilm, hs[-1], g, split=sp)
        dWm[-1] = g
        dbm[-1] = ops.colsum(dfilm, zeros(film.shape[1]))
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
            elif ctx.needs_input_grad[4]:
                L2 = 2 * cfg.scalings.numel()
                dhash = ops.gemm_nt(dz, Wm[0][:, 3:3 + L2].t().contiguous(), split=sp)     # [N, 32]: d cond[:, 3:35]
                d_table = ops.hash_encode_bwd(q, cfg.scalings, cfg.log2_T, dhash)
        grads_mlp: List[Optional[Tensor]] = []
        for i in range(DDF_MAP_LAYERS):
            grads_mlp += [dWm[i], dbm[i]]
        for l in range(DDF_TRUNK_LAYERS):
            grads_mlp += [dWt[l], dbt[l]]
        return (None, None, None, d_thr.reshape(threshold.shape), d_table, d_wf.reshape(w_final.shape), d_bf.reshape(ctx.b_shape), *grads_mlp)


def ddf_visibility(cfg: DDFConfig, points: Tensor, dirs_sel: Tensor, threshold: Tensor, table: Tensor, w_final: Tensor, b_final: Tensor,
                   mlp: Sequence[Tensor]) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """points [R,3] (detached surface points), dirs_sel [D',3] -> (visibility [R,D'], expected termination distance [R*D'],
    sphere points q [R*D',3], termination_dist [R*D']).  Differentiable w.r.t. threshold, the hash table, the final layer and
    every mapping / trunk weight (`mlp` in `ddf_param_list` order)."""
    return _DDFVisibility.apply(cfg, points, dirs_sel, threshold, table, w_final, b_final, *mlp)
