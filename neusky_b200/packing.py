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
