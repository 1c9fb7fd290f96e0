"""torch.autograd bindings of the C-ABI ops that have hand-written backward kernels (the "torch custom-op layer" of
north_star for the training path).  Forward and backward both run the CUDA kernels; there is no torch fallback.

  hash_encode(x, table, ...)      d/d table, d/d x and the double backward d(d/d x)/d table  (nsk_hash_encode_bwd, _grad_x, _grad_x_bwd)
  neus_composite(sdf, grad, albedo, inv_s, ...)   d/d sdf, grad, albedo, inv_s  (nsk_neus_composite_bwd)
  lambert_shade(normals, wa, radiance, vis, ...)   d/d normals, wa, radiance, visibility  (nsk_lambert_relight_bwd)
  shade_finalize(rgb_lin, bg, acc)                 d/d rgb_lin, bg, acc  (nsk_shade_finalize_bwd)
  reni_radiance(dirs, latents, scale, ...)         d/d latents, scale with the decoder frozen  (nsk_reni_decode_bwd)
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import ops

Tensor = torch.Tensor


class _HashEncodeBwd(torch.autograd.Function):
    """The backward of the encode as a differentiable op: (x, table, g) -> (d x [n,3], d table [L*T,2]).  Its own backward
    is what makes `torch.autograd.grad(sdf, x, create_graph=True)` (the reference's normals, sdf_albedo_field.py:235-238)
    trainable: the normals depend on the table through the trilinear slopes."""

    @staticmethod
    def forward(ctx, x, table, g, scalings, log2_T: int):
        ctx.save_for_backward(x, table, g, scalings)
        ctx.log2_T = log2_T
        g = g.contiguous()
        return ops.hash_encode_grad_x(x, table, scalings, log2_T, g), ops.hash_encode_bwd(x, scalings, log2_T, g)

    @staticmethod
    def backward(ctx, c_gx, c_gtable):
        x, table, g, scalings = ctx.saved_tensors
        d_g = d_table = None
        if c_gx is not None:
            d_g, d_table = ops.hash_encode_grad_x_bwd(x, table, scalings, ctx.log2_T, g, c_gx.contiguous(), ctx.needs_input_grad[2], ctx.needs_input_grad[1])
        if c_gtable is not None and ctx.needs_input_grad[2]:
            extra = ops.hash_encode(x, c_gtable.contiguous(), scalings, ctx.log2_T)       # d g of the table scatter = encode with that "table"
            d_g = extra.reshape(g.shape) if d_g is None else d_g.reshape(g.shape) + extra.reshape(g.shape)
        return None, d_table, (None if d_g is None else d_g.reshape(g.shape)), None, None


class _HashEncode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, table: Tensor, scalings: Tensor, log2_T: int) -> Tensor:
        ctx.save_for_backward(x, table, scalings)
        ctx.log2_T = log2_T
        return ops.hash_encode(x, table, scalings, log2_T)

    @staticmethod
    def backward(ctx, g: Tensor):
        x, table, scalings = ctx.saved_tensors
        gx, gtable = _HashEncodeBwd.apply(x, table, g, scalings, ctx.log2_T)
        return (gx if ctx.needs_input_grad[0] else None), (gtable if ctx.needs_input_grad[1] else None), None, None


def hash_encode(x: Tensor, table: Tensor, scalings: Tensor, log2_T: int) -> Tensor:
    """Twice-differentiable hash-grid encode: d/d table and d/d x, and the derivative of d/d x wrt the table."""
    return _HashEncode.apply(x, table, scalings, log2_T)


class _NeusComposite(torch.autograd.Function):
    OUT = ("weights", "wa", "normals", "accumulation", "p2p_raw", "normal", "albedo", "bg_transmittance")

    @staticmethod
    def forward(ctx, sdf, grad, albedo, inv_s, ray_dirs, starts, ends, deltas, dnorm, cos_anneal_ratio: float):
        o = ops.neus_composite(sdf, grad, albedo, ray_dirs, starts, ends, deltas, dnorm, inv_s, cos_anneal_ratio, True)     # inv_s read on the device: no host sync
        ctx.save_for_backward(sdf, grad, albedo, inv_s, ray_dirs, starts, ends, deltas)
        ctx.rho = cos_anneal_ratio
        return tuple(o[k] for k in _NeusComposite.OUT)

    @staticmethod
    def backward(ctx, *gs):
        sdf, grad, albedo, inv_s, ray_dirs, starts, ends, deltas = ctx.saved_tensors
        g = {k: (None if v is None else v.contiguous()) for k, v in zip(_NeusComposite.OUT, gs)}
        d_sdf, d_grad, d_alb, d_inv = ops.neus_composite_bwd(sdf, grad, albedo, ray_dirs, starts, ends, deltas, inv_s, ctx.rho, g)
        return d_sdf.reshape(sdf.shape), d_grad, d_alb, d_inv.reshape(inv_s.shape), None, None, None, None, None, None


def neus_composite(sdf, grad, albedo, inv_s: Tensor, ray_dirs, starts, ends, deltas, dnorm, cos_anneal_ratio: float = 1.0) -> Tuple[Tensor, ...]:
    """Differentiable K3 (training mode).  Returns (weights [R,S], wa [R,S,3], normals [R,S,3], accumulation [R],
    p2p_raw [R], normal [R,3], albedo [R,3], bg_transmittance [R]); inv_s is a 0-dim / 1-element tensor."""
    return _NeusComposite.apply(sdf, grad, albedo, inv_s, ray_dirs, starts, ends, deltas, dnorm, cos_anneal_ratio)


class _LambertShade(torch.autograd.Function):
    @staticmethod
    def forward(ctx, normals, wa, radiance, vis_sel, inv_count, dirs, sel_index, cam, unocc: float):
        ctx.save_for_backward(normals, wa, radiance, vis_sel, inv_count, dirs, sel_index, cam if cam is not None else torch.empty(0))
        ctx.has_cam, ctx.unocc = cam is not None, unocc
        return ops.lambert_relight(normals, wa, inv_count, dirs, sel_index, radiance, vis_sel, cam, unocc)

    @staticmethod
    def backward(ctx, g):
        normals, wa, radiance, vis_sel, inv_count, dirs, sel_index, cam = ctx.saved_tensors
        cam = cam if ctx.has_cam else None
        d_wa, d_n, d_vis, d_rad = ops.lambert_relight_bwd(normals, wa, inv_count, dirs, sel_index, radiance, vis_sel, g.contiguous(), cam, ctx.unocc,
                                                          ctx.needs_input_grad[3], ctx.needs_input_grad[2])
        return d_n, d_wa, d_rad, d_vis, None, None, None, None, None


def lambert_shade(normals, wa, radiance, vis_sel, inv_count, dirs, sel_index, cam=None, unoccluded_vis: float = 1.0) -> Tensor:
    """Differentiable Lambertian sum with given per-ray visibility: linear rgb [R,3] (renderers.py:93-113)."""
    return _LambertShade.apply(normals, wa, radiance, vis_sel, inv_count, dirs, sel_index, cam, unoccluded_vis)


class _ShadeFinalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb_lin, bg, acc):
        ctx.save_for_backward(rgb_lin, bg, acc)
        return ops.shade_finalize(rgb_lin, bg, acc, training=True)

    @staticmethod
    def backward(ctx, g):
        rgb_lin, bg, acc = ctx.saved_tensors
        d_lin, d_bg, d_acc = ops.shade_finalize_bwd(rgb_lin, bg, acc, g.contiguous())
        return d_lin, d_bg, d_acc.reshape(acc.shape)


def shade_finalize(rgb_lin, bg, acc) -> Tensor:
    return _ShadeFinalize.apply(rgb_lin, bg, acc)


class _ReniRadiance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dirs, row_cam, latents, scale, packed, packed_bwd, rotation, log_domain: bool):
        if row_cam is None:
            out = ops.reni_radiance_table(dirs, latents, scale, packed, rotation=rotation, log_domain=log_domain)
        else:
            out = ops.reni_radiance_rows(dirs, row_cam, latents, scale, packed, rotation=rotation, log_domain=log_domain)
        e = torch.empty(0)
        ctx.save_for_backward(dirs, row_cam if row_cam is not None else e, latents, scale if scale is not None else e, packed, packed_bwd,
                              rotation if rotation is not None else e, out)
        ctx.flags = (row_cam is not None, scale is not None, rotation is not None, log_domain)
        return out

    @staticmethod
    def backward(ctx, g):
        dirs, row_cam, latents, scale, packed, packed_bwd, rotation, out = ctx.saved_tensors
        has_cam, has_scale, has_rot, log_domain = ctx.flags
        d_lat = torch.zeros_like(latents)
        d_scale = torch.zeros_like(scale) if has_scale else None
        ops.reni_decode_bwd(dirs, row_cam if has_cam else None, latents, scale if has_scale else None, packed, packed_bwd, out, g.contiguous(), d_lat, d_scale,
                            rotation=rotation if has_rot else None, log_domain=log_domain)
        return None, None, d_lat, d_scale, None, None, None, None


def reni_radiance(dirs: Tensor, latents: Tensor, scale, packed: Tensor, packed_bwd: Tensor, row_cam=None, rotation=None, log_domain: bool = True) -> Tensor:
    """RENI++ HDR radiance, differentiable w.r.t. the latent codes and scales (decoder frozen, as under the reference's
    hold_decoder_fixed): [K,D,3] for every (code, direction) pair, or [N,3] with one code per direction when row_cam [N] is given."""
    return _ReniRadiance.apply(dirs, row_cam, latents, scale, packed, packed_bwd, rotation, log_domain)
