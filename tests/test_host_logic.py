"""Host-side logic that needs no GPU: weight packing for the tensor-core RENI++ chain and the light-argument normalisation of
the reference-named shaders."""
import pytest
import torch

from neusky_b200 import init as nb_init
from neusky_b200 import packing


def test_pack_reni_gemm_layout():
    p = nb_init.init_reni_params(4321)
    g = packing.pack_reni_gemm(p)
    assert g["res_w"].shape == (128, 512) and torch.equal(g["res_w"][:, :510], p["network.residual_projection.weight"])
    assert float(g["res_w"][:, 510:].abs().max()) == 0.0                      # K padding of the 510-wide decoder input
    for i in range(6):
        assert g[f"f0w{i}"].shape == (128, 128) and g[f"f2w{i}"].shape == (128, 128)
        assert torch.equal(g[f"n1w{i}"], p[f"network.layers.{i}.norm1.weight"])
    assert g["fc_w"].shape == (3, 128) and all(v.is_contiguous() for v in g.values())


def test_shader_light_arguments():
    from neusky_b200.shaders import _compact_lights

    N, M, K = 6, 5, 3
    dirs = torch.randn(M, 3)
    table = torch.rand(K, M, 3)
    cam = torch.tensor([0, 2, 1, 1, 0, 2])
    d, c, idx = _compact_lights(dirs[None].expand(N, M, 3), table[cam], None, N)      # the reference's expanded form
    assert d.shape == (M, 3) and c.shape == (N, M, 3) and torch.equal(idx, torch.arange(N, dtype=torch.int32))
    d, c, idx = _compact_lights(dirs, table, cam, N)                                   # compact form
    assert d.shape == (M, 3) and c.shape == (K, M, 3) and idx.dtype == torch.int32 and idx.tolist() == cam.tolist()
    d, c, idx = _compact_lights(dirs[None], table[:1], None, N)                        # (1, M, 3) broadcast lights
    assert c.shape == (1, M, 3) and idx is None
    with pytest.raises(ValueError):
        _compact_lights(dirs, table, None, N)                                          # K tables, no index, K != N


def test_ops_reject_cpu_tensors():
    from neusky_b200 import ops

    with pytest.raises(ValueError):
        ops.lambert_collapse_sel(torch.zeros(2, 4, 3), torch.zeros(2, 4, 3), torch.zeros(2, 4), torch.zeros(5, 3))
    with pytest.raises(ValueError):
        ops.relight_collapsed(torch.zeros(2, 5, 3), torch.zeros(1, 5, 3))


def test_sky_pixel_loss_vs_reference_golden(golden):
    """RENISkyPixelLoss(alpha=0.1) on sRGB(hdr background) vs the image under the sky mask: value and gradient from the
    reference's own class (tests/golden/make_golden.py::golden_losses); product formula and oracle restatement."""
    from neusky_b200 import train as T
    from oracle import train_oracle as TO
    from oracle import neusky_oracle as O

    g = golden("losses")
    hdr = torch.from_numpy(g["hdr"]).requires_grad_(True)
    image, sky = torch.from_numpy(g["image"]), torch.from_numpy(g["sky"])
    m = sky[:, None].expand_as(image)
    ours = T.sky_pixel_loss(T._linear_to_srgb(hdr), image, m)
    (grad,) = torch.autograd.grad(ours, hdr)
    assert abs(float(ours) - float(g["sky_pixel_loss"])) <= 1e-6
    assert torch.allclose(grad, torch.from_numpy(g["sky_pixel_loss_d_hdr"]), rtol=1e-5, atol=1e-8)
    ref2 = TO.sky_pixel_loss(O.linear_to_srgb(hdr.detach()), image, m)
    assert abs(float(ref2) - float(g["sky_pixel_loss"])) <= 1e-6
