"""Host-side packing of module parameters into the flat blobs the kernels read.

Layouts mirror ``simt_layout()`` (csrc/sky_shade_simt.cu), ``reni_layout()`` (csrc/reni_decode.cu)
and the tensor-core operand images of csrc/sky_shade_tc.cu.  Packing is done once per weight
update on the host/GPU with torch ops; the kernels never see nn.Parameters directly.
"""
from __future__ import annotations

import functools
from typing import Dict

import torch

Tensor = torch.Tensor
DDF_HID, DDF_LAYERS = 256, 5


def _host_packed(fn):
    """Run a packer on the HOST and upload the finished blob once.  The packers are a few hundred tiny slicing / cast / concat
    steps (171 operand stages for the DDF stream alone); as device ops that is ~1000 torch kernel launches per weight set, which
    buried the product's own kernels in every launch trace.  On the host they cost a few milliseconds and the device sees one
    memcpy.  `device`: where the blob should live (default: the device of the parameters passed in).  Hash tables are not
    touched by any packer and stay where they are."""

    @functools.wraps(fn)
    def wrapped(p: Dict[str, Tensor], *args, device=None, **kw):
        tensors = [v for v in p.values() if isinstance(v, Tensor)]
        dev = torch.device(device) if device is not None else (tensors[0].device if tensors else torch.device("cpu"))
        ph = {k: (v.detach().to("cpu") if isinstance(v, Tensor) and "hash_table" not in k else v) for k, v in p.items()}
        out = fn(ph, *args, **kw)
        if isinstance(out, dict):
            return {k: v.to(dev) for k, v in out.items()}
        return out.to(dev)

    return wrapped


@_host_packed
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


@_host_packed
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


@_host_packed
def pack_reni_gemm(p: Dict[str, Tensor], num_layers: int = 6) -> Dict[str, Tensor]:
    """Decoder weights in the [out, in] layout nsk_gemm_tf32_nt takes (torch's own), for ops.reni_rows_tc: the residual
    projection zero-padded from 510 to 512 input columns, per layer norm1 / fc.0 / fc.2 / norm2, and the 128 -> 3 head."""
    f = lambda k: p[k].to(torch.float32).contiguous()
    Wr = f("network.residual_projection.weight")
    Wp = Wr.new_zeros((Wr.shape[0], 512))
    Wp[:, :Wr.shape[1]] = Wr
    out = {"res_w": Wp.contiguous(), "res_b": f("network.residual_projection.bias"), "fc_w": f("network.fc.weight"), "fc_b": f("network.fc.bias")}
    for i in range(num_layers):
        pre = f"network.layers.{i}."
        out.update({f"n1w{i}": f(pre + "norm1.weight"), f"n1b{i}": f(pre + "norm1.bias"), f"f0w{i}": f(pre + "fc.0.weight"), f"f0b{i}": f(pre + "fc.0.bias"),
                    f"f2w{i}": f(pre + "fc.2.weight"), f"f2b{i}": f(pre + "fc.2.bias"), f"n2w{i}": f(pre + "norm2.weight"), f"n2b{i}": f(pre + "norm2.bias")})
    return out


@_host_packed
def pack_reni_bwd(p: Dict[str, Tensor], num_layers: int = 6) -> Tensor:
    """fp32 blob for nsk_reni_decode_bwd (see reni_bwd_layout()): the decoder's linear weights in torch's own [out][in]
    layout, which is the coalesced one for the transposed products of the backward pass."""
    parts = [p["network.residual_projection.weight"]]
    for i in range(num_layers):
        pre = f"network.layers.{i}."
        parts += [p[pre + "mha.value.weight"], p[pre + "mha.fc_out.weight"], p[pre + "fc.0.weight"], p[pre + "fc.2.weight"]]
    return torch.cat([x.to(torch.float32).contiguous().flatten() for x in parts]).contiguous()


@_host_packed
def pack_ddf_tc2(p: Dict[str, Tensor]) -> Tensor:
    """uint8 blob for nsk_sky_shade_tc2_fwd (CTA-pair kernel, csrc/sky_shade_tc2.cu): the same operand matrices as
    pack_ddf_tc, but every [N][K] tile is split by rows between the two CTAs of a pair (rank r streams rows
    [r N/2, (r+1) N/2) -- for a FiLM chunk that is the freq' rows for rank 0 and the phase' rows for rank 1) and cut
    into 16 KB stages of twice the K extent.  Layout: [rank-0 stream | rank-1 stream | fp32 tail]."""
    dev = p["ddf.final_layer.weight"].device
    f32 = lambda k: p[k].to(torch.float32)
    Wf, bf, Wp, bp = fold_film(p)
    ops_ = []   # (matrix [N][K], nfull, kps, ktail) in MMA issue order
    ops_.append((_with_bias_cols(f32("ddf.mapping_network.network.0.weight"), f32("ddf.mapping_network.network.0.bias"), 48), 0, 64, 48))
    for i in range(1, DDF_LAYERS):
        ops_.append((_with_bias_cols(f32(f"ddf.mapping_network.network.{2 * i}.weight"), f32(f"ddf.mapping_network.network.{2 * i}.bias"), 272), 4, 64, 16))

    def fp(l, c):
        rows = slice(l * DDF_HID + c * 64, l * DDF_HID + c * 64 + 64)
        return (torch.cat([_with_bias_cols(Wf[rows], bf[rows], 272), _with_bias_cols(Wp[rows], bp[rows], 272)], 0), 2, 128, 16)

    def z(l):
        W = f32(f"ddf.net.{l}.layer.weight")
        if l == 0:
            # [W_hi | W_hi | W_lo] against the kernel's input tile [x_hi | x_lo | x_hi]: x W^T = x_hi W_hi + x_lo W_hi + x_hi W_lo (+ 2^-22)
            W0 = torch.zeros((DDF_HID, 16), dtype=torch.float32, device=dev)
            W0[:, :15] = W
            hi = W0.to(torch.float16).to(torch.float32)
            return (torch.cat([hi, hi, W0 - hi], 1), 0, 64, 48)
        return (W, 4, 64, 0)

    ops_ += [fp(0, 0), z(0), fp(0, 1)]
    for l in range(DDF_LAYERS):
        ops_ += [fp(l, 2), fp(l, 3)]
        if l + 1 < DDF_LAYERS:
            ops_ += [fp(l + 1, 0), z(l + 1), fp(l + 1, 1)]
    streams = []
    for r in range(2):
        st = []
        for W, nfull, kps, ktail in ops_:
            h = W.shape[0] // 2
            st += _stages(W[r * h:(r + 1) * h], nfull, kps, ktail)
        streams.append(torch.cat(st).contiguous().view(torch.uint8))
    assert streams[0].numel() == TC2_STREAM_BYTES // 2 == streams[1].numel(), (streams[0].numel(), TC2_STREAM_BYTES)
    vec = torch.zeros(TC_TAIL_FLOATS, dtype=torch.float32, device=dev)
    vec[:256] = f32("ddf.final_layer.weight").flatten()
    vec[256] = f32("ddf.final_layer.bias").flatten()[0]
    return torch.cat([streams[0], streams[1], vec.view(torch.uint8)]).contiguous()


def fold_weight_norm(p: Dict[str, Tensor], name: str) -> Tensor:
    """nn.utils.weight_norm(dim=0) fold W = g * v / ||v||_row (SDFFieldConfig.weight_norm=True [NS-mem A.4];
    neusky/fields/sdf_albedo_field.py:159-160); plain ``.weight`` is accepted too."""
    if name + ".weight" in p:
        return p[name + ".weight"].to(torch.float32)
    v, g = p[name + ".weight_v"].to(torch.float32), p[name + ".weight_g"].to(torch.float32)
    return v * (g / v.norm(dim=1, keepdim=True))


@_host_packed
def pack_sdf_simt(p: Dict[str, Tensor]) -> Tensor:
    """fp32 blob for nsk_sdf_field_simt_fwd (see sdf_layout() in csrc/sdf_field_simt.cu): forward weights
    transposed to [K][N], reverse-pass weights in torch's [out][in] layout.  NeuSky shape only
    (neusky_config.py:66-77: 2 hidden geo layers, 2 hidden colour layers, width 256, geo feature 256)."""
    W0, W1, W2 = (fold_weight_norm(p, f"glin{l}") for l in range(3))
    C0, C1, C2 = (fold_weight_norm(p, f"clin{l}") for l in range(3))
    if tuple(W0.shape) != (256, 71) or tuple(W1.shape) != (256, 256) or tuple(W2.shape) != (257, 256) or tuple(C0.shape) != (256, 295) or tuple(C2.shape) != (3, 256):
        raise ValueError("pack_sdf_simt: expected the NeuSky SDFAlbedoField shape (71->256->256->257, 295->256->256->3)")
    f32 = lambda k: p[k].to(torch.float32).flatten()
    dev = W0.device
    pad4 = lambda v: torch.cat([v.flatten(), torch.zeros(4 - v.numel(), dtype=torch.float32, device=dev)])
    b2 = f32("glin2.bias")
    parts = [
        W0.t().contiguous().flatten(), f32("glin0.bias"), W1.t().contiguous().flatten(), f32("glin1.bias"),
        W2[1:].t().contiguous().flatten(), b2[1:], W2[0].contiguous(), pad4(b2[:1]),
        W1.contiguous().flatten(), W0.contiguous().flatten(),
        C0.t().contiguous().flatten(), f32("clin0.bias"), C1.t().contiguous().flatten(), f32("clin1.bias"),
        C2.contiguous().flatten(), pad4(f32("clin2.bias")),
    ]
    return torch.cat(parts).contiguous()


# ------------------------------------------------------------------------------------------------
# Tensor-core blob for csrc/sky_shade_tc.cu: the per-tile weight STREAM (fp16 operand tiles in the exact
# order the MMA issuer consumes them, see the stage table in the kernel header) followed by the fp32
# final-layer vector.  Every bias rides inside the GEMMs as two extra K columns (fp16 hi + lo) that
# multiply the constant-1 columns of the activation tiles.
# ------------------------------------------------------------------------------------------------
TC_STREAM_BYTES = (16384 + 8192) + 4 * (8 * 16384 + 8192) + 20 * (4 * 16384 + 4096) + 8192 + 4 * 8 * 16384
TC_TAIL_FLOATS = 256 + 4
TC2_STREAM_BYTES = TC_STREAM_BYTES + 16384      # CTA-pair kernel: the first trunk layer carries an fp16 hi/lo split of both operands (K = 48)


def _image(W: Tensor) -> Tensor:
    """W [N][kc] fp16 (kc multiple of 8) -> [kc/8][N][8] flattened: no-swizzle K-major canonical tile."""
    N, kc = W.shape
    return W.reshape(N, kc // 8, 8).permute(1, 0, 2).contiguous().flatten()


def _stages(W: Tensor, nfull: int, kps: int, ktail: int):
    Wh = W.to(torch.float16)
    out = [_image(Wh[:, i * kps : (i + 1) * kps]) for i in range(nfull)]
    if ktail:
        out.append(_image(Wh[:, nfull * kps : nfull * kps + ktail]))
    assert nfull * kps + ktail == W.shape[1]
    return out


def _with_bias_cols(W: Tensor, b: Tensor, K: int) -> Tensor:
    """[N][K]: W, then bias split into fp16 hi and lo columns, zero padded."""
    N, k = W.shape
    out = torch.zeros((N, K), dtype=torch.float32, device=W.device)
    out[:, :k] = W
    hi = b.to(torch.float16).to(torch.float32)
    out[:, k] = hi
    out[:, k + 1] = b.to(torch.float32) - hi
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


@_host_packed
def pack_ddf_tc(p: Dict[str, Tensor]) -> Tensor:
    """uint8 blob [TC_STREAM_BYTES + 260*4] for nsk_sky_shade_tc_fwd."""
    dev = p["ddf.final_layer.weight"].device
    f32 = lambda k: p[k].to(torch.float32)
    Wf, bf, Wp, bp = fold_film(p)
    st = []
    st += _stages(_with_bias_cols(f32("ddf.mapping_network.network.0.weight"), f32("ddf.mapping_network.network.0.bias"), 48), 1, 32, 16)
    for i in range(1, DDF_LAYERS):
        st += _stages(_with_bias_cols(f32(f"ddf.mapping_network.network.{2 * i}.weight"), f32(f"ddf.mapping_network.network.{2 * i}.bias"), 272), 8, 32, 16)

    def fp(l, c):
        rows = slice(l * DDF_HID + c * 64, l * DDF_HID + c * 64 + 64)
        W = torch.cat([_with_bias_cols(Wf[rows], bf[rows], 272), _with_bias_cols(Wp[rows], bp[rows], 272)], 0)
        return _stages(W, 4, 64, 16)

    def z(l):
        W = f32(f"ddf.net.{l}.layer.weight")
        if l == 0:
            W0 = torch.zeros((DDF_HID, 16), dtype=torch.float32, device=dev)
            W0[:, :15] = W
            return _stages(W0, 0, 32, 16)
        return _stages(W, 8, 32, 0)

    st += fp(0, 0) + z(0) + fp(0, 1)
    for l in range(DDF_LAYERS):
        st += fp(l, 2) + fp(l, 3)
        if l + 1 < DDF_LAYERS:
            st += fp(l + 1, 0) + z(l + 1) + fp(l + 1, 1)
    stream = torch.cat(st).contiguous().view(torch.uint8)
    assert stream.numel() == TC_STREAM_BYTES, (stream.numel(), TC_STREAM_BYTES)
    vec = torch.zeros(TC_TAIL_FLOATS, dtype=torch.float32, device=dev)
    vec[:256] = f32("ddf.final_layer.weight").flatten()
    vec[256] = f32("ddf.final_layer.bias").flatten()[0]
    return torch.cat([stream, vec.view(torch.uint8)]).contiguous()


# ------------------------------------------------------------------------------------------------
# Tensor-core blob for csrc/sdf_field_tc.cu: one region of fp16 operand stages per GEMM (the order and the stage
# shapes of the segment table in the kernel), then the fp32 sdf row of the last geo layer.
# ------------------------------------------------------------------------------------------------
SDF_TC_STREAM_BYTES = 256 * 80 * 2 + 3 * 256 * 272 * 2 + 256 * 256 * 2 + 80 * 256 * 2 + 256 * (48 + 272) * 2 + 16 * 272 * 2


@_host_packed
def pack_sdf_tc(p: Dict[str, Tensor]) -> Tensor:
    """uint8 blob [SDF_TC_STREAM_BYTES + 260*4] for nsk_sdf_field_tc_fwd (weight_norm folded)."""
    W0, W1, W2 = (fold_weight_norm(p, f"glin{l}") for l in range(3))
    C0, C1, C2 = (fold_weight_norm(p, f"clin{l}") for l in range(3))
    if tuple(W0.shape) != (256, 71) or tuple(W1.shape) != (256, 256) or tuple(W2.shape) != (257, 256) or tuple(C0.shape) != (256, 295) or tuple(C2.shape) != (3, 256):
        raise ValueError("pack_sdf_tc: expected the NeuSky SDFAlbedoField shape (71->256->256->257, 295->256->256->3)")
    dev = W0.device
    f32 = lambda k: p[k].to(torch.float32).flatten()
    z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
    b2 = f32("glin2.bias")
    st = []
    # G0 / G0': K = [x_hi 3 | PE 36 | feat 32 | bias hi, lo | x_lo 3 | 0 x4] = 80
    G0 = _with_bias_cols(W0, f32("glin0.bias"), 80)
    G0[:, 73:76] = W0[:, 0:3]
    st += _stages(G0, 2, 32, 16)
    st += _stages(_with_bias_cols(W1, f32("glin1.bias"), 272), 8, 32, 16)          # G1
    st += _stages(_with_bias_cols(W2[1:], b2[1:], 272), 8, 32, 16)                 # G2 (geo feature rows)
    st += _stages(W1.t().contiguous(), 8, 32, 0)                                   # B1: B[n=j][k=i] = W1[i][j]
    B0 = z(80, 256)                                                                # B0: rows [x 3 | PE 36 | 0 | feat 32 | 0 x8]
    B0[0:39] = W0[:, 0:39].t()
    B0[40:72] = W0[:, 39:71].t()
    st += _stages(B0, 4, 64, 0)
    Ca = z(256, 48)                                                                # C0: IN columns 0..47 (x, PE; feat columns get zeros)
    Ca[:, 0:39] = C0[:, 0:39]
    st += _stages(Ca, 1, 32, 16)
    st += _stages(_with_bias_cols(C0[:, 39:], f32("clin0.bias"), 272), 8, 32, 16)
    st += _stages(_with_bias_cols(C1, f32("clin1.bias"), 272), 8, 32, 16)          # C1
    Cc = z(16, 272)                                                                # C2: 3 real rows
    Cc[0:3] = _with_bias_cols(C2, f32("clin2.bias"), 272)
    st += _stages(Cc, 1, 272, 0)
    stream = torch.cat(st).contiguous().view(torch.uint8)
    assert stream.numel() == SDF_TC_STREAM_BYTES, (stream.numel(), SDF_TC_STREAM_BYTES)
    vec = z(260)
    vec[:256] = W2[0]
    vec[256] = b2[0]
    return torch.cat([stream, vec.view(torch.uint8)]).contiguous()


# ------------------------------------------------------------------------------------------------
# Blob for csrc/reni_fused_tc.cu: 32 fp16 weight stages of [128 N][64 K] (first layer 510 -> 128 zero-padded to K = 512: 8 stages; then per
# decoder layer fc.0 and fc.2, 2 stages each) in the order the MMA issuer consumes them, followed by the fp32 constants
# (biases, LayerNorm weights, output head).  The attention constants a_i(k) come from nsk_reni_prep (per latent code).
# ------------------------------------------------------------------------------------------------
RENI_FUSED_WEIGHT_BYTES = 32 * 16384
RENI_FUSED_CONST_FLOATS = 128 + 6 * 6 * 128 + 3 * 128 + 4


@_host_packed
def pack_reni_fused(p: Dict[str, Tensor], num_layers: int = 6) -> Tensor:
    """uint8 blob [RENI_FUSED_WEIGHT_BYTES + RENI_FUSED_CONST_FLOATS * 4] for nsk_reni_rows_fused_fwd."""
    if num_layers != 6:
        raise ValueError("pack_reni_fused: the fused kernel implements the 6-layer RENI++ decoder NeuSky ships")
    f = lambda k: p[k].to(torch.float32)
    Wr = f("network.residual_projection.weight")
    if Wr.shape[0] != 128 or Wr.shape[1] > 512:
        raise ValueError("pack_reni_fused: expected a [128, 5 (L + 2) <= 512] residual projection")
    if Wr.shape[1] != 510:
        raise ValueError("pack_reni_fused: the fused kernel is specialised for latent_dim = 100 (a [128, 510] residual projection)")
    nj = 102                                     # 100 latent inputs + d_z + |d_xy|: [sin a, sin 4a] x nj | [cos a, cos 4a] x nj | x x nj  (reni_illumination_field.py:345-348)
    orig = lambda j: [2 * j, 2 * j + 1, 2 * nj + 2 * j, 2 * nj + 2 * j + 1, 4 * nj + j]       # sin a, sin 4a, cos a, cos 4a, x
    extras = orig(100) + orig(101) + [-1, -1]
    perm = []
    for q in range(4):                           # the kernel's per-quarter layout (csrc/reni_fused_tc.cu, positional-encoding warps)
        cols = []
        for jj in range(25):
            cols += orig(25 * q + jj)[:4]
        cols += [orig(25 * q + jj)[4] for jj in range(25)]
        cols += extras[3 * q:3 * q + 3]
        perm += cols
    assert len(perm) == 512 and sorted(c for c in perm if c >= 0) == list(range(510))
    Wp = torch.zeros((128, 512), dtype=torch.float32)
    idx = torch.tensor([c for c in perm if c >= 0])
    Wp[:, torch.tensor([i for i, c in enumerate(perm) if c >= 0])] = Wr[:, idx]
    st = _stages(Wp, 8, 64, 0)
    consts = [f("network.residual_projection.bias")]
    for i in range(num_layers):
        pre = f"network.layers.{i}."
        st += _stages(f(pre + "fc.0.weight"), 2, 64, 0) + _stages(f(pre + "fc.2.weight"), 2, 64, 0)
        consts += [f(pre + "fc.0.bias"), f(pre + "fc.2.bias"), f(pre + "norm1.weight"), f(pre + "norm1.bias"), f(pre + "norm2.weight"), f(pre + "norm2.bias")]
    fcb = torch.zeros(4, dtype=torch.float32)
    fcb[:3] = f("network.fc.bias")
    consts += [f("network.fc.weight").reshape(-1), fcb]
    stream = torch.cat(st).contiguous().view(torch.uint8)
    cvec = torch.cat([c.reshape(-1) for c in consts]).contiguous()
    assert stream.numel() == RENI_FUSED_WEIGHT_BYTES and cvec.numel() == RENI_FUSED_CONST_FLOATS, (stream.numel(), cvec.numel())
    return torch.cat([stream, cvec.view(torch.uint8)]).contiguous()
