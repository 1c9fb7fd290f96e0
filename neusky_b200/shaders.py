"""Reference-named shaders over csrc/shaders.cu.

  LambertianShader                      reni.model_components.shaders.LambertianShader      (ns_reni/reni/model_components/shaders.py:25-70)
  BlinnPhongShader                      reni.model_components.shaders.BlinnPhongShader      (ns_reni/reni/model_components/shaders.py:73-161)
  RGBBlinnPhongRendererWithVisibility   neusky.model_components.renderers.<same name>       (neusky/model_components/renderers.py:179-288)

Same call signatures and argument meaning as the reference classes.  The reference takes the lights as [N,M,3] tensors
(usually expand()-views of one [M,3] direction set and of the per-camera radiance rows); those are accepted as they are:
a stride-0 leading dimension is recognised and read once, a genuinely per-row tensor is read per row.  The compact form --
`light_directions` [M,3], `light_colors` [K,M,3] plus `light_index` [N] -- avoids materialising anything of size N x M and
is what the rest of this package passes.  All three are differentiable (albedo, normals, specular, shininess, light colours,
visibility) through one backward kernel.  CUDA only: the ops raise on CPU tensors.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops

Tensor = torch.Tensor


def _compact_lights(light_directions: Tensor, light_colors: Tensor, light_index: Optional[Tensor], N: int) -> Tuple[Tensor, Tensor, Optional[Tensor]]:
    """-> (dirs [M,3] | [N,M,3], radiance [K,M,3], cam [N] int32 | None)."""
    d = light_directions
    if d.dim() == 3 and (d.shape[0] == 1 or d.stride(0) == 0):
        d = d[0]
    c = light_colors
    if c.dim() == 2:
        c = c[None]
    cam = None
    if light_index is not None:
        cam = light_index.reshape(-1).to(torch.int32)
    elif c.shape[0] == 1 or c.stride(0) == 0:
        c = c[:1]
    elif c.shape[0] == N:                      # one light table per row, as the reference materialises it
        cam = torch.arange(N, device=c.device, dtype=torch.int32)
    else:
        raise ValueError(f"light_colors: expected [1|N, M, 3] or [K, M, 3] with light_index, got {tuple(c.shape)} for N = {N}")
    return d, c, cam


class _ShadeLights(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mode, normalize_dirs, albedo, normals, dirs, radiance, cam, specular, shininess, view_dirs, vis):
        out_a, out_b = ops.shade_lights(mode, albedo, normals, dirs, radiance, cam, specular, shininess, view_dirs, vis, normalize_dirs)
        ctx.mode, ctx.normalize_dirs = mode, normalize_dirs
        ctx.opt = (cam is not None, specular is not None, shininess is not None, view_dirs is not None, vis is not None)
        e = albedo.new_zeros(0)
        ctx.save_for_backward(albedo, normals, dirs, radiance, cam if cam is not None else e.to(torch.int32), specular if specular is not None else e,
                              shininess if shininess is not None else e, view_dirs if view_dirs is not None else e, vis if vis is not None else e)
        if mode == 0:
            return out_a, out_b
        return out_a, out_a.new_zeros(0)

    @staticmethod
    def backward(ctx, g_a, g_b):
        albedo, normals, dirs, radiance, cam, specular, shininess, view_dirs, vis = ctx.saved_tensors
        has_cam, has_spec, has_shin, has_view, has_vis = ctx.opt
        g_a = None if g_a is None else g_a.contiguous()
        g_b = None if (g_b is None or ctx.mode != 0) else g_b.contiguous()
        d_alb, d_nrm, d_spec, d_shin, d_rad, d_vis = ops.shade_lights_bwd(
            ctx.mode, albedo, normals, dirs, radiance, g_a, g_b, cam if has_cam else None, specular if has_spec else None, shininess if has_shin else None,
            view_dirs if has_view else None, vis if has_vis else None, ctx.normalize_dirs, want_radiance=ctx.needs_input_grad[5], want_vis=ctx.needs_input_grad[10])
        return None, None, d_alb, d_nrm, None, d_rad, None, d_spec, d_shin, None, d_vis


class LambertianShader(torch.nn.Module):
    """(textureless shading, shaded albedo) = (sum_j max(n.l_j, 0) L_j, albedo * that)."""

    @classmethod
    def forward(cls, albedo: Tensor, normals: Tensor, light_directions: Tensor, light_colors: Tensor, detach_normals: bool = True,
                light_index: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        if detach_normals:
            normals = normals.detach()
        lead = albedo.shape[:-1]
        a, n = albedo.reshape(-1, 3), normals.reshape(-1, 3)
        d, c, cam = _compact_lights(light_directions, light_colors, light_index, a.shape[0])
        s, rgb = _ShadeLights.apply(0, False, a, n, d, c, cam, None, None, None, None)
        return s.reshape(*lead, 3), rgb.reshape(*lead, 3)


class BlinnPhongShader(torch.nn.Module):
    """Diffuse + normalised Blinn-Phong specular lobe, clamped to >= 1e-3."""

    @classmethod
    def forward(cls, albedo: Tensor, normals: Tensor, light_directions: Tensor, light_colors: Tensor, specular: Tensor, shininess: Tensor,
                view_directions: Tensor, detach_normals: bool = False, normalize_directions: bool = False, light_index: Optional[Tensor] = None) -> Tensor:
        if detach_normals:
            normals = normals.detach()
        N = albedo.shape[0]
        if shininess.numel() != N:
            raise ValueError(f"shininess: expected {N} values, got {tuple(shininess.shape)}")
        d, c, cam = _compact_lights(light_directions, light_colors, light_index, N)
        out, _ = _ShadeLights.apply(1, bool(normalize_directions), albedo, normals, d, c, cam, specular, shininess.reshape(N), view_directions, None)
        return out


class BlinnPhongShaderChunked(BlinnPhongShader):
    """The reference chunks pixels to bound its [chunk, M, 3] temporaries (shaders.py:164-233); the kernel has none, so the
    chunked variant is the plain one (chunk_size is accepted and ignored)."""

    def __init__(self, chunk_size: int = 4096):
        super().__init__()
        self.chunk_size = chunk_size


class RGBBlinnPhongRendererWithVisibility(torch.nn.Module):
    """NeuSky's Blinn-Phong branch (predict_shininess=True): per-sample radiance with sky visibility, composited along the ray,
    background-blended and converted to sRGB."""

    @classmethod
    def render_and_combine_rgb(cls, albedos: Tensor, normals: Tensor, light_directions: Tensor, light_colors: Tensor, visibility: Optional[Tensor],
                               background_illumination: Tensor, weights: Tensor, shininess: Tensor, c2w_matrices: Tensor,
                               ray_indices: Optional[Tensor] = None, num_rays: Optional[int] = None, light_index: Optional[Tensor] = None) -> Tensor:
        if ray_indices is not None or num_rays is not None:
            raise NotImplementedError("packed (nerfacc) samples are not on NeuSky's path (SURVEY 2.1 N5/N6)")
        from .train import _linear_to_srgb

        a, n = albedos.reshape(-1, 3), normals.reshape(-1, 3)
        N = a.shape[0]
        # world-space view direction: c2w @ (0, 0, -1, 1) (renderers.py:205-210) = translation - third rotation column
        c2w = c2w_matrices.reshape(-1, 3, 4)
        view = (c2w[:, :, 3] - c2w[:, :, 2]).to(n.dtype).contiguous()
        d, c, cam = _compact_lights(light_directions, light_colors, light_index, N)
        vis = None
        if visibility is not None:
            v = visibility
            if v.dim() == 3:
                v = v[..., 0]
            vis = v                                  # [N, M] per sample (the reference's layout) or [R, M] per ray
        rad, _ = _ShadeLights.apply(2, False, a, n, d, c, cam, None, shininess.reshape(N), view, vis)
        radiance = rad.view(*weights.shape[:-1], 3)
        comp = torch.sum(weights * radiance, dim=-2)
        acc = torch.sum(weights, dim=-2)
        comp = comp + background_illumination.to(weights.device) * (1.0 - acc)
        return _linear_to_srgb(comp)

    def forward(self, albedos, normals, light_directions, light_colors, visibility, background_illumination, weights, shininess, c2w_matrices,
                ray_indices=None, num_rays=None, light_index=None) -> Tensor:
        rgb = self.render_and_combine_rgb(albedos, normals, light_directions, light_colors, visibility, background_illumination, weights, shininess,
                                          c2w_matrices, ray_indices, num_rays, light_index)
        if not self.training:
            rgb = torch.clamp(rgb, min=0.0, max=1.0)
        return rgb
