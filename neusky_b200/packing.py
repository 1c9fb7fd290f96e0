"""Host-side packing of module parameters into the flat blobs the kernels read.

Layouts mirror ``simt_layout()`` (csrc/sky_shade_simt.cu), ``reni_layout()`` (csrc/reni_decode.cu)
and the tensor-core operand images of csrc/sky_shade_tc.cu.  Packing is done once per weight
update on the host/GPU with torch ops; the kernels never see nn.Parameters directly.
"""
from __future__ import annotations

from typing import Dict

import torch

Tensor = torch.Tensor
DDF_HID, DDF_LAYERS = 256, 5


def pack_ddf_simt(p: Dict[str, Tensor]) -> Tensor:
    """fp32, weights transposed to [K][N]: mapping 0..5, trunk 0..4, final (see simt_layout())."""
    parts = []
    for i in range(DDF_LAYERS + 1):
        parts += [p[f"ddf.mapping_network.network.{2 * i}.weight"].t().contiguous().flatten(), p[f"ddf.mapping_network.network.{2 * i}.bias"].flatten()]
    for l in range(DDF_LAYERS):
        parts += [p[f"ddf.net.{l}.layer.weight"].t().contiguous().flatten(), p[f"ddf.net.{l}.layer.bias"].flatten()]
    fb = torch.zeros(4, dtype=torch.float32, device=parts[0].device)
    fb[0] = p["ddf.final_layer.bias"].flatten()[0]
    parts += [p["ddf.final_layer.weight"].flatten(), fb]
    return torch.cat([x.to(torch.float32) for x in parts]).contiguous()


def pack_reni(p: Dict[str, Tensor], num_layers: int = 6) -> Tensor:
    """fp32 blob for nsk_reni_decode_fwd (see reni_layout()); linear weights transposed to [in][out]."""
    dev = p["network.fc.weight"].device
    vn = torch.zeros(16, dtype=torch.float32, device=dev)
    vn[0] = p["vn_proj_in.1.weight"].flatten()[0]
    vn[1:3] = p["vn_invar.mlp.0.weight"].flatten()
    vn[3:7] = p["vn_invar.mlp.1.W"].flatten()
    vn[7:11] = p["vn_invar.mlp.1.U"].flatten()
    parts = [vn, p["network.residual_projection.weight"].t().contiguous().flatten(), p["network.residual_projection.bias"]]
    for i in range(num_layers):
        pre = f"network.layers.{i}."
        parts += [
            p[pre + "mha.value.weight"].t().contiguous().flatten(), p[pre + "mha.value.bias"],
            p[pre + "mha.fc_out.weight"].t().contiguous().flatten(), p[pre + "mha.fc_out.bias"],
            p[pre + "norm1.weight"], p[pre + "norm1.bias"],
            p[pre + "fc.0.weight"].t().contiguous().flatten(), p[pre + "fc.0.bias"],
            p[pre + "fc.2.weight"].t().contiguous().flatten(), p[pre + "fc.2.bias"],
            p[pre + "norm2.weight"], p[pre + "norm2.bias"],
        ]
    fcb = torch.zeros(4, dtype=torch.float32, device=dev)
    fcb[:3] = p["network.fc.bias"]
    parts += [p["network.fc.weight"].contiguous().flatten(), fcb]
    return torch.cat([x.to(torch.float32).flatten() for x in parts]).contiguous()


# ------------------------------------------------------------------------------------------------
# Tensor-core blob for csrc/sky_shade_tc.cu: the per-tile weight STREAM (147 stages x 16 KB of fp16
# operand tiles in the exact order the MMA issuer consumes them) followed by the fp32 epilogue vectors.
# ------------------------------------------------------------------------------------------------
TC_STAGE_BYTES = 16384
TC_STAGES_PER_TILE = 147
TC_BIAS_FLOATS = 15 * 256 + 256 + 4


def _stage_images(W: Tensor, kps: int):
    """W [N][K] (K multiple of kps) -> list of fp16 [kps/8][N][8] images (no-swizzle K-major canonical layout)."""
    N, K = W.shape
    Wh = W.to(torch.float16)
    out = []
    for k0 in range(0, K, kps):
        out.append(Wh[:, k0 : k0 + kps].reshape(N, kps // 8, 8).permute(1, 0, 2).contiguous().flatten())
    return out


def _pad_k(W: Tensor, K: int) -> Tensor:
    out = torch.zeros((W.shape[0], K), dtype=W.dtype, device=W.device)
    out[:, : W.shape[1]] = W
    return out


def fold_film(p: Dict[str, Tensor]):
    """freq' = 15 f + 30 and the trunk bias folded into the FiLM GEMM (see sky_shade_tc.cu header).
    Returns Wf [1280,256], bf [1280], Wp [1280,256], bp [1280] in fp32."""
    W6 = p[f"ddf.mapping_network.network.{2 * DDF_LAYERS}.weight"].to(torch.float32)
    b6 = p[f"ddf.mapping_network.network.{2 * DDF_LAYERS}.bias"].to(torch.float32)
    half = DDF_LAYERS * DDF_HID
    bt = torch.cat([p[f"ddf.net.{l}.layer.bias"].to(torch.float32) for l in range(DDF_LAYERS)])  # [1280]
    Wf = 15.0 * W6[:half]
    bf = 15.0 * b6[:half] + 30.0
    Wp = W6[half:] + bt[:, None] * Wf
    bp = b6[half:] + bf * bt
    return Wf, bf, Wp, bp


def pack_ddf_tc(p: Dict[str, Tensor]) -> Tensor:
    """uint8 blob [147*16384 + 4100*4] for nsk_sky_shade_tc_fwd."""
    dev = p["ddf.final_layer.weight"].device
    Wf, bf, Wp, bp = fold_film(p)
    stages = []
    # mapping network: M1 (K 35 -> 64), M2..M5
    stages += _stage_images(_pad_k(p["ddf.mapping_network.network.0.weight"].to(torch.float32), 64), 32)
    for i in range(1, DDF_LAYERS):
        stages += _stage_images(p[f"ddf.mapping_network.network.{2 * i}.weight"].to(torch.float32), 32)

    def fp(l, c):
        rows = slice(l * DDF_HID + c * 64, l * DDF_HID + c * 64 + 64)
        return _stage_images(torch.cat([Wf[rows], Wp[rows]], 0), 64)

    def z(l):
        W = p[f"ddf.net.{l}.layer.weight"].to(torch.float32)
        return _stage_images(_pad_k(W, 32) if l == 0 else W, 32)

    stages += fp(0, 0) + z(0) + fp(0, 1)
    for l in range(DDF_LAYERS):
        stages += fp(l, 2) + fp(l, 3)
        if l + 1 < DDF_LAYERS:
            stages += fp(l + 1, 0) + z(l + 1) + fp(l + 1, 1)
    assert len(stages) == TC_STAGES_PER_TILE, len(stages)
    assert all(s.numel() * 2 == TC_STAGE_BYTES for s in stages)
    stream = torch.cat(stages).contiguous().view(torch.uint8)
    vec = torch.zeros(TC_BIAS_FLOATS, dtype=torch.float32, device=dev)
    for i in range(DDF_LAYERS):
        vec[i * 256 : (i + 1) * 256] = p[f"ddf.mapping_network.network.{2 * i}.bias"]
    vec[5 * 256 : 10 * 256] = bf
    vec[10 * 256 : 15 * 256] = bp
    vec[15 * 256 : 16 * 256] = p["ddf.final_layer.weight"].flatten()
    vec[16 * 256] = p["ddf.final_layer.bias"].flatten()[0]
    return torch.cat([stream, vec.view(torch.uint8)]).contiguous()
