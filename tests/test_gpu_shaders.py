"""Light-sum shaders (csrc/shaders.cu, neusky_b200/shaders.py) against values AND autograd gradients produced by the reference's
own classes (tests/golden/shaders.npz <- reni.model_components.shaders, neusky.model_components.renderers).
Tolerance: fp32, 1e-5 relative on values, 2e-5 norm-wise on gradients (different summation order only)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from neusky_b200 import _lib

    _lib.load()
    return torch.device("cuda:0")


@pytest.fixture()
def G(golden, dev):
    g = golden("shaders")
    t = {k: torch.from_numpy(g[k]) for k in g.files}
    return t


def _close(a, b, rtol=1e-5, atol=1e-6):
    a, b = a.detach().cpu().double(), b.double()
    assert torch.allclose(a, b, rtol=rtol, atol=atol), float((a - b).abs().max())


def _rel(a, b):
    a, b = a.detach().cpu().double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _leaves(G, dev, names):
    return {n: G[n].to(dev).clone().requires_grad_(True) for n in names}


@pytest.mark.parametrize("layout", ["compact", "expanded", "materialised"])
def test_lambertian_shader_values_and_gradients(G, dev, layout):
    from neusky_b200.shaders import LambertianShader

    L = _leaves(G, dev, ["albedo", "normals", "table"])
    N, M = G["albedo"].shape[0], G["dirs"].shape[0]
    dirs, cam = G["dirs"].to(dev), G["cam"].to(dev)
    if layout == "compact":
        s, rgb = LambertianShader.forward(L["albedo"], L["normals"], dirs, L["table"], detach_normals=False, light_index=cam)
    elif layout == "expanded":       # the reference's calling convention: expand()-view directions, gathered colours
        s, rgb = LambertianShader.forward(L["albedo"], L["normals"], dirs[None].expand(N, M, 3), L["table"][cam], detach_normals=False)
    else:
        s, rgb = LambertianShader.forward(L["albedo"], L["normals"], dirs[None].expand(N, M, 3).contiguous(), L["table"][cam].contiguous(), detach_normals=False)
    _close(s, G["lambert_sum"])
    _close(rgb, G["lambert_rgb"])
    ((s * G["cot"].to(dev)).sum() + (rgb * G["cot2"].to(dev)).sum()).backward()
    for n in ("albedo", "normals", "table"):
        assert _rel(L[n].grad, G[f"lambert_d_{n}"]) <= 2e-5, n
    # detach_normals=True (the reference's default): no gradient reaches the normals
    L2 = _leaves(G, dev, ["albedo", "normals", "table"])
    s2, rgb2 = LambertianShader.forward(L2["albedo"], L2["normals"], dirs, L2["table"], light_index=cam)
    (s2.sum() + rgb2.sum()).backward()
    assert L2["normals"].grad is None and L2["albedo"].grad is not None


def test_blinn_phong_shader_values_and_gradients(G, dev):
    from neusky_b200.shaders import BlinnPhongShader, BlinnPhongShaderChunked

    names = ["albedo", "normals", "table", "specular", "shininess"]
    L = _leaves(G, dev, names)
    dirs, cam, view = G["dirs"].to(dev), G["cam"].to(dev), G["view"].to(dev)
    out = BlinnPhongShader.forward(L["albedo"], L["normals"], dirs, L["table"], L["specular"], L["shininess"], view, light_index=cam)
    _close(out, G["blinn_phong"], rtol=2e-5)
    (out * G["cot"].to(dev)).sum().backward()
    for n in names:
        assert _rel(L[n].grad, G[f"blinn_phong_d_{n}"]) <= 5e-5, n
    # (1, M, 3) broadcast lights with un-normalised directions
    with torch.no_grad():
        out_n = BlinnPhongShaderChunked(chunk_size=16).forward(L["albedo"], L["normals"], 1.7 * dirs[None], L["table"][:1], L["specular"], L["shininess"], view,
                                                               normalize_directions=True)
    _close(out_n, G["blinn_phong_normalized_broadcast"], rtol=2e-5)


def test_neusky_blinn_phong_renderer_with_visibility(G, dev):
    from neusky_b200 import ops
    from neusky_b200.shaders import RGBBlinnPhongRendererWithVisibility

    R, S, M = G["ren_vis"].shape[0], G["ren_weights"].shape[1], G["dirs"].shape[0]
    N = R * S
    L = _leaves(G, dev, ["albedo", "normals", "table", "shininess"])
    vis, w = G["ren_vis"].to(dev).requires_grad_(True), G["ren_weights"].to(dev).requires_grad_(True)
    dirs, cam = G["dirs"].to(dev), G["cam"].to(dev)
    c2w = G["ren_c2w"].to(dev)[:, None].expand(R, S, 3, 4)
    ren = RGBBlinnPhongRendererWithVisibility()
    ren.train()
    rgb = ren(albedos=L["albedo"].view(R, S, 3), normals=L["normals"].view(R, S, 3), light_directions=dirs, light_colors=L["table"], visibility=vis,
              background_illumination=G["ren_bg"].to(dev), weights=w, shininess=L["shininess"].view(R, S, 1), c2w_matrices=c2w, light_index=cam)
    _close(rgb, G["ren_rgb"], rtol=2e-5)
    (rgb * G["ren_cot"].to(dev)).sum().backward()
    for n in ("albedo", "normals", "table", "shininess"):
        assert _rel(L[n].grad, G[f"ren_d_{n}"]) <= 5e-5, n
    assert _rel(vis.grad, G["ren_d_vis_ray"]) <= 5e-5
    assert _rel(w.grad, G["ren_d_weights"]) <= 5e-5
    # the reference's own layout: [N, M, 1] visibility and [N, M, 3] lights, eval-mode clamp
    ren.eval()
    with torch.no_grad():
        rgb2 = ren(albedos=L["albedo"].view(R, S, 3), normals=L["normals"].view(R, S, 3), light_directions=dirs[None].expand(N, M, 3), light_colors=L["table"][cam],
                   visibility=vis[:, None].expand(R, S, M).reshape(N, M, 1), background_illumination=G["ren_bg"].to(dev), weights=w,
                   shininess=L["shininess"].view(R, S, 1), c2w_matrices=c2w.contiguous())
    _close(rgb2, G["ren_rgb"].clamp(0, 1), rtol=2e-5)
    # fused compositing in the forward kernel (eval path): rgb_lin[r] += w * radiance
    view = (c2w.reshape(-1, 3, 4)[:, :, 3] - c2w.reshape(-1, 3, 4)[:, :, 2]).contiguous()
    rgb_lin = torch.zeros(R, 3, device=dev)
    rad, _ = ops.shade_lights(2, L["albedo"].detach(), L["normals"].detach(), dirs, L["table"].detach(), cam.to(torch.int32), shininess=L["shininess"].detach(),
                              view_dirs=view, vis=vis.detach(), weights=w.detach().reshape(-1), rgb_lin=rgb_lin)
    ref_lin = (w.detach() * rad.view(R, S, 3)).sum(-2)
    assert torch.allclose(rgb_lin, ref_lin, rtol=1e-5, atol=1e-6)


def test_shader_input_validation(dev):
    from neusky_b200.shaders import BlinnPhongShader, LambertianShader, RGBBlinnPhongRendererWithVisibility

    a = torch.rand(4, 3)
    with pytest.raises(ValueError):      # CPU tensors: no fallback
        LambertianShader.forward(a, a, torch.rand(5, 3), torch.rand(1, 5, 3))
    ad = a.to(dev)
    with pytest.raises(ValueError):      # K light tables but no index and K != N
        LambertianShader.forward(ad, ad, torch.rand(5, 3, device=dev), torch.rand(3, 5, 3, device=dev))
    with pytest.raises(NotImplementedError):
        RGBBlinnPhongRendererWithVisibility.render_and_combine_rgb(ad.view(2, 2, 3), ad.view(2, 2, 3), torch.rand(5, 3, device=dev), torch.rand(1, 5, 3, device=dev), None,
                                                                  torch.rand(2, 3, device=dev), torch.rand(2, 2, 1, device=dev), torch.rand(2, 2, 1, device=dev),
                                                                  torch.rand(2, 2, 3, 4, device=dev), ray_indices=torch.zeros(4, device=dev), num_rays=2)
    with pytest.raises(ValueError):      # shininess of the wrong length
        BlinnPhongShader.forward(ad, ad, torch.rand(5, 3, device=dev), torch.rand(1, 5, 3, device=dev), ad, torch.rand(3, device=dev), ad)
