"""Thin, validated Python wrappers over the C ABI (include/neusky_b200.h).

Every op takes CUDA fp32 contiguous tensors, allocates its outputs with torch's caching
allocator, launches on torch's current stream and raises ``ValueError`` / ``RuntimeError`` on
misuse -- there is no CPU path and no silent fallback (north_star).
"""
from __future__ import annotations

import ctypes
import functools
import types
from typing import Dict, Optional, Tuple

import torch

from . import _lib

Tensor = torch.Tensor
c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float


_EMPTY_SENTINEL: Dict[torch.device, Tensor] = {}


def _ptr(t: Optional[Tensor]):
    """Device address of `t` (NULL for None).  An EMPTY tensor has no storage (data_ptr() == 0), which the C ABI would take
    for a missing argument; it gets the address of a 16-byte sentinel instead -- never dereferenced, the ops return before
    launching when a count is zero."""
    if t is None:
        return c_void_p(0)
    if t.numel() == 0 and t.is_cuda:
        s = _EMPTY_SENTINEL.get(t.device)
        if s is None:
            s = _EMPTY_SENTINEL[t.device] = torch.zeros(4, device=t.device, dtype=torch.float32)
        return c_void_p(s.data_ptr())
    return c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(t: Tensor):
    """torch's CURRENT stream on the tensor's device as a cudaStream_t.  The raw getter skips the Stream object round trip
    (a few microseconds per launch: the training step makes ~1400 launches)."""
    if _raw_stream is not None:
        idx = t.device.index
        return c_void_p(_raw_stream(torch.cuda.current_device() if idx is None else idx))
    return c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _chk(name: str, t: Tensor, dtype=torch.float32, shape: Optional[Tuple] = None) -> Tensor:
    if not isinstance(t, torch.Tensor):
        raise ValueError(f"{name}: expected a tensor")
    if not t.is_cuda:
        raise ValueError(f"{name}: must be a CUDA tensor (neusky_b200 has no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if shape is not None:
        ts = t.shape
        if len(ts) != len(shape):
            raise ValueError(f"{name}: expected shape {shape}, got {tuple(ts)}")
        for s, d in zip(shape, ts):
            if s is not None and s != d:
                raise ValueError(f"{name}: expected shape {shape}, got {tuple(ts)}")
    return t if t.is_contiguous() else t.contiguous()


def _chk_out(name: str, t: Tensor, dtype=torch.float32, shape: Optional[Tuple] = None) -> Tensor:
    """Validation for OUTPUT / accumulator arguments: like `_chk`, but a non-contiguous tensor is an error -- `_chk` would hand the
    kernel a contiguous temporary and the caller's buffer would silently never be written."""
    if isinstance(t, torch.Tensor) and t.is_cuda and not t.is_contiguous():
        raise ValueError(f"{name}: output / accumulator tensors must be contiguous (got strides {t.stride()})")
    return _chk(name, t, dtype, shape)


def _device_guard(fn):
    """Run an op with the CUDA device of its tensor arguments current.  The kernels, the SM-count / shared-memory opt-in caches
    (csrc/nsk_common.cuh: device_once) and the stream lookup all act on the process's CURRENT device; a module built with
    device='cuda:1' while device 0 is current would otherwise launch on GPU 0 against GPU 1 memory.  Mixed-device tensor
    arguments are an error."""

    @functools.wraps(fn)
    def wrapped(*args, **kw):
        dev = None
        for a in args + tuple(kw.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise ValueError(f"{fn.__name__}: tensor arguments live on different devices ({dev} and {a.device})")
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kw)
        with torch.cuda.device(dev):
            return fn(*args, **kw)

    return wrapped


# ------------------------------------------------------------------------------------------- K1
def hash_encode(x: Tensor, table: Tensor, scalings: Tensor, log2_T: int) -> Tensor:
    """x [...,3] -> [..., 2L]; nerfstudio torch hash-grid semantics."""
    lead = x.shape[:-1]
    x2 = _chk("x", x.reshape(-1, 3), shape=(None, 3))
    L = scalings.numel()
    table = _chk("table", table, shape=(L << log2_T, 2))
    scalings = _chk("scalings", scalings)
    out = torch.empty((x2.shape[0], 2 * L), device=x.device, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.nsk_hash_encode_fwd(_ptr(x2), c_int64(x2.shape[0]), _ptr(table), _ptr(scalings), c_int(L), c_int(log2_T), _ptr(out), _stream(x)), "nsk_hash_encode_fwd")
    return out.reshape(*lead, 2 * L)


def hash_encode_bwd(x: Tensor, scalings: Tensor, log2_T: int, grad_out: Tensor, grad_table: Optional[Tensor] = None) -> Tensor:
    x2 = _chk("x", x.reshape(-1, 3), shape=(None, 3))
    L = scalings.numel()
    g = _chk("grad_out", grad_out.reshape(-1, 2 * L), shape=(x2.shape[0], 2 * L))
    if grad_table is None:
        grad_table = torch.zeros((L << log2_T, 2), device=x.device, dtype=torch.float32)
    grad_table = _chk_out("grad_table", grad_table, shape=(L << log2_T, 2))
    lib = _lib.load()
    _lib.check(lib.nsk_hash_encode_bwd(_ptr(x2), c_int64(x2.shape[0]), _ptr(_chk("scalings", scalings)), c_int(L), c_int(log2_T), _ptr(g), _ptr(grad_table), _stream(x)), "nsk_hash_encode_bwd")
    return grad_table


def hash_encode_grad_x(x: Tensor, table: Tensor, scalings: Tensor, log2_T: int, grad_out: Tensor) -> Tensor:
    """d L / d x [n,3] of the encode for a cotangent grad_out [n,2L]."""
    x2 = _chk("x", x.reshape(-1, 3), shape=(None, 3))
    L = scalings.numel()
    table = _chk("table", table, shape=(L << log2_T, 2))
    g = _chk("grad_out", grad_out.reshape(-1, 2 * L), shape=(x2.shape[0], 2 * L))
    gx = torch.empty((x2.shape[0], 3), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_hash_encode_grad_x(_ptr(x2), c_int64(x2.shape[0]), _ptr(table), _ptr(_chk("scalings", scalings)), c_int(L), c_int(log2_T), _ptr(g), _ptr(gx), _stream(x)), "nsk_hash_encode_grad_x")
    return gx.reshape(x.shape)


def hash_encode_grad_x_bwd(x: Tensor, table: Tensor, scalings: Tensor, log2_T: int, grad_out: Tensor, cot_x: Tensor, want_g: bool = True, want_table: bool = True):
    """Backward of hash_encode_grad_x: (d_grad_out [n,2L] | None, d_table [L*T,2] | None)."""
    x2 = _chk("x", x.reshape(-1, 3), shape=(None, 3))
    n, L = x2.shape[0], scalings.numel()
    table = _chk("table", table, shape=(L << log2_T, 2))
    g = _chk("grad_out", grad_out.reshape(-1, 2 * L), shape=(n, 2 * L))
    c = _chk("cot_x", cot_x.reshape(-1, 3), shape=(n, 3))
    d_g = torch.empty((n, 2 * L), device=x.device, dtype=torch.float32) if want_g else None
    d_t = torch.zeros((L << log2_T, 2), device=x.device, dtype=torch.float32) if want_table else None
    if want_g or want_table:
        _lib.check(_lib.load().nsk_hash_encode_grad_x_bwd(_ptr(x2), c_int64(n), _ptr(table), _ptr(_chk("scalings", scalings)), c_int(L), c_int(log2_T), _ptr(g), _ptr(c), _ptr(d_g), _ptr(d_t), _stream(x)), "nsk_hash_encode_grad_x_bwd")
    return d_g, d_t


def hash_encode_tcnn(x: Tensor, table: Tensor, level_meta: Tensor, log2_T: int, smoothstep: bool = True) -> Tensor:
    """x [...,3] in [0,1] -> [..., 2L] with tiny-cuda-nn's grid semantics (imported reference checkpoints, tcnn_import.py)."""
    lead = x.shape[:-1]
    x2 = _chk("x", x.reshape(-1, 3), shape=(None, 3))
    L = level_meta.shape[0]
    table = _chk("table", table, shape=(L << log2_T, 2))
    level_meta = _chk("level_meta", level_meta, dtype=torch.int32, shape=(L, 4))
    out = torch.empty((x2.shape[0], 2 * L), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_hash_encode_tcnn_fwd(_ptr(x2), c_int64(x2.shape[0]), _ptr(table), _ptr(level_meta), c_int(L), c_int(log2_T), c_int(int(smoothstep)),
                                                    _ptr(out), _stream(x)), "nsk_hash_encode_tcnn_fwd")
    return out.reshape(*lead, 2 * L)


def hash_indices(x: Tensor, scalings: Tensor, log2_T: int) -> Tuple[Tensor, Tensor]:
    x2 = _chk("x", x.reshape(-1, 3), shape=(None, 3))
    L = scalings.numel()
    idx = torch.empty((x2.shape[0], L, 8), device=x.device, dtype=torch.int64)
    off = torch.empty((x2.shape[0], L, 3), device=x.device, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.nsk_hash_indices(_ptr(x2), c_int64(x2.shape[0]), _ptr(_chk("scalings", scalings)), c_int(L), c_int(log2_T), _ptr(idx), _ptr(off), _stream(x)), "nsk_hash_indices")
    return idx, off


# ------------------------------------------------------------------------------------------- K2
def sdf_field(x: Tensor, blob: Tensor, hash_table: Tensor, scalings: Tensor, log2_T: int, want_grad: bool = True, want_albedo: bool = True,
              want_geo: bool = False, impl: str = "simt", grid_meta: Optional[Tensor] = None, smoothstep: bool = True) -> Dict[str, Tensor]:
    """x [...,3] -> {"sdf" [...,1], "gradient" [...,3], "albedo" [...,3], "geo" [...,256]} (SDFAlbedoField.get_outputs
    without alpha, neusky/fields/sdf_albedo_field.py:211-269; the gradient is analytic, not autograd).
    ``grid_meta`` (int32 [L,4], tcnn_import.tcnn_level_meta): evaluate the hash table with tiny-cuda-nn's grid semantics (imported
    reference checkpoint) instead of the nerfstudio torch grid; ``smoothstep`` is tcnn's interpolation flag."""
    lead = x.shape[:-1]
    x2 = _chk("x", x.reshape(-1, 3), shape=(None, 3))
    n = x2.shape[0]
    L = scalings.numel()
    hash_table = _chk("hash_table", hash_table, shape=(L << log2_T, 2))
    scalings = _chk("scalings", scalings)
    lib = _lib.load()
    f = dict(device=x.device, dtype=torch.float32)
    sdf = torch.empty((n,), **f)
    if grid_meta is not None:
        grid_meta = _chk("grid_meta", grid_meta, dtype=torch.int32, shape=(L, 4))
    gm = (_ptr(grid_meta), c_int(int(smoothstep)))
    if impl == "tc":
        if want_geo:
            raise ValueError("sdf_field: the tensor-core path keeps the geometry feature on chip (want_geo needs impl='simt')")
        blob = _chk("blob", blob, dtype=torch.uint8, shape=(lib.nsk_sdf_tc_weights_bytes(),))
        grad, alb = torch.empty((n, 3), **f), torch.empty((n, 3), **f)
        _lib.check(lib.nsk_sdf_field_tc_fwd_ex(_ptr(x2), c_int64(n), _ptr(blob), _ptr(hash_table), _ptr(scalings), c_int(L), c_int(log2_T), *gm, _ptr(sdf), _ptr(grad), _ptr(alb), _stream(x)), "nsk_sdf_field_tc_fwd_ex")
        return {"sdf": sdf.reshape(*lead, 1), "gradient": grad.reshape(*lead, 3), "albedo": alb.reshape(*lead, 3)}
    if impl != "simt":
        raise ValueError(f"impl must be 'tc' or 'simt', got {impl!r}")
    blob = _chk("blob", blob, shape=(lib.nsk_sdf_simt_weights_floats(),))
    grad = torch.empty((n, 3), **f) if want_grad else None
    alb = torch.empty((n, 3), **f) if want_albedo else None
    geo = torch.empty((n, 256), **f) if want_geo else None
    _lib.check(lib.nsk_sdf_field_simt_fwd_ex(_ptr(x2), c_int64(n), _ptr(blob), _ptr(hash_table), _ptr(scalings), c_int(L), c_int(log2_T), *gm, _ptr(sdf), _ptr(grad), _ptr(alb), _ptr(geo), _stream(x)), "nsk_sdf_field_simt_fwd_ex")
    out = {"sdf": sdf.reshape(*lead, 1)}
    if grad is not None:
        out["gradient"] = grad.reshape(*lead, 3)
    if alb is not None:
        out["albedo"] = alb.reshape(*lead, 3)
    if geo is not None:
        out["geo"] = geo.reshape(*lead, 256)
    return out


# ------------------------------------------------------------------------------------------- K3
_MINMAX_INIT: Dict[torch.device, Tensor] = {}


def _minmax_init(dev) -> Tensor:
    """A fresh {+inf, -inf} accumulator for the composite kernel's running min / max: cloned on the device from a per-device constant
    (torch.tensor([...], device=cuda) would be a pageable host->device copy per call: a host sync, and illegal in a graph capture)."""
    dev = torch.device(dev)
    c = _MINMAX_INIT.get(dev)
    if c is None:
        c = _MINMAX_INIT[dev] = torch.tensor([float("inf"), float("-inf")], device=dev, dtype=torch.float32)
    return c.clone()


def _inv_s_arg(inv_s, dev):
    """(by-value float, device pointer) for the composite calls: a tensor stays on the device (`*_dv` entry points)."""
    if isinstance(inv_s, torch.Tensor):
        t = _chk("inv_s", inv_s.detach().reshape(-1), shape=(1,))
        if t.device != dev:
            raise ValueError(f"inv_s: tensor on {t.device}, samples on {dev}")
        return None, t
    return float(inv_s), None


def neus_composite(sdf, grad, albedo, ray_dirs, starts, ends, deltas, dnorm, inv_s, cos_anneal_ratio: float = 1.0, training: bool = False,
                   steps_minmax: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """sdf/starts/ends/deltas [R,S(,1)], grad/albedo [R,S,3], ray_dirs [R,3], dnorm [R(,1)]; inv_s a float or a 1-element CUDA tensor
    (read on the device: no host synchronisation).
    The expected depth is clipped to [min, max] of the sample mid-points like nerfstudio's DepthRenderer: over THIS
    batch by default (the reference renders 256-ray chunks, so its clip range is per chunk), or to the caller's
    ``steps_minmax`` [2] when given (tile- and rank-invariant eval renders)."""
    R, S = sdf.shape[0], sdf.shape[1]
    dev = sdf.device
    sdf = _chk("sdf", sdf.reshape(R, S))
    grad = _chk("grad", grad, shape=(R, S, 3))
    albedo = _chk("albedo", albedo, shape=(R, S, 3))
    ray_dirs = _chk("ray_dirs", ray_dirs, shape=(R, 3))
    starts, ends, deltas = (_chk(n, t.reshape(R, S)) for n, t in (("starts", starts), ("ends", ends), ("deltas", deltas)))
    dnorm = _chk("dnorm", dnorm.reshape(R))
    f = dict(device=dev, dtype=torch.float32)
    out = {
        "weights": torch.empty((R, S), **f), "wa": torch.empty((R, S, 3), **f), "normals": torch.empty((R, S, 3), **f),
        "accumulation": torch.empty((R,), **f), "p2p_raw": torch.empty((R,), **f), "normal": torch.empty((R, 3), **f),
        "albedo": torch.empty((R, 3), **f), "bg_transmittance": torch.empty((R,), **f),
        "p2p_dist": torch.empty((R,), **f), "depth": torch.empty((R,), **f),
    }
    mm = _minmax_init(dev)
    lib = _lib.load()
    st = _stream(sdf)
    inv_f, inv_t = _inv_s_arg(inv_s, dev)
    tail = (c_float(cos_anneal_ratio), c_int(int(training)), _ptr(out["weights"]), _ptr(out["wa"]), _ptr(out["normals"]), _ptr(out["accumulation"]), _ptr(out["p2p_raw"]), _ptr(out["normal"]), _ptr(out["albedo"]), _ptr(out["bg_transmittance"]), _ptr(mm), st)
    head = (_ptr(sdf), _ptr(grad), _ptr(albedo), _ptr(ray_dirs), _ptr(starts), _ptr(ends), _ptr(deltas), c_int64(R), c_int(S))
    if inv_t is None:
        _lib.check(lib.nsk_neus_composite_fwd(*head, c_float(inv_f), *tail), "nsk_neus_composite_fwd")
    else:
        _lib.check(lib.nsk_neus_composite_fwd_dv(*head, _ptr(inv_t), *tail), "nsk_neus_composite_fwd_dv")
    if steps_minmax is not None:
        mm = _chk("steps_minmax", steps_minmax, shape=(2,))
    _lib.check(lib.nsk_neus_finalize_depth(_ptr(out["p2p_raw"]), _ptr(dnorm), _ptr(mm), c_int64(R), _ptr(out["p2p_dist"]), _ptr(out["depth"]), st), "nsk_neus_finalize_depth")
    out["steps_minmax"] = mm
    return out


def neus_composite_bwd(sdf, grad, albedo, ray_dirs, starts, ends, deltas, inv_s, cos_anneal_ratio: float, g: Dict[str, Optional[Tensor]]):
    """Cotangents g[{"weights","wa","normals","accumulation","p2p_raw","normal","albedo","bg_transmittance"}] (missing / None = 0)
    -> (d_sdf [R,S], d_grad [R,S,3], d_albedo [R,S,3], d_inv_s [1])."""
    R, S = sdf.shape[0], sdf.shape[1]
    sdf = _chk("sdf", sdf.reshape(R, S))
    grad, albedo = _chk("grad", grad, shape=(R, S, 3)), _chk("albedo", albedo, shape=(R, S, 3))
    ray_dirs = _chk("ray_dirs", ray_dirs, shape=(R, 3))
    starts, ends, deltas = (_chk(n, t.reshape(R, S)) for n, t in (("starts", starts), ("ends", ends), ("deltas", deltas)))
    shapes = {"weights": (R, S), "wa": (R, S, 3), "normals": (R, S, 3), "accumulation": (R,), "p2p_raw": (R,), "normal": (R, 3), "albedo": (R, 3), "bg_transmittance": (R,)}
    gg = {k: (None if g.get(k) is None else _chk("g_" + k, g[k].reshape(shp), shape=shp)) for k, shp in shapes.items()}
    f = dict(device=sdf.device, dtype=torch.float32)
    d_sdf, d_grad, d_alb, d_inv = torch.empty((R, S), **f), torch.empty((R, S, 3), **f), torch.empty((R, S, 3), **f), torch.zeros((1,), **f)
    inv_f, inv_t = _inv_s_arg(inv_s, sdf.device)
    head = (_ptr(sdf), _ptr(grad), _ptr(albedo), _ptr(ray_dirs), _ptr(starts), _ptr(ends), _ptr(deltas), c_int64(R), c_int(S))
    tail = (c_float(cos_anneal_ratio), _ptr(gg["weights"]), _ptr(gg["wa"]), _ptr(gg["normals"]), _ptr(gg["accumulation"]), _ptr(gg["p2p_raw"]), _ptr(gg["normal"]), _ptr(gg["albedo"]), _ptr(gg["bg_transmittance"]),
            _ptr(d_sdf), _ptr(d_grad), _ptr(d_alb), _ptr(d_inv), _stream(sdf))
    if inv_t is None:
        _lib.check(_lib.load().nsk_neus_composite_bwd(*head, c_float(inv_f), *tail), "nsk_neus_composite_bwd")
    else:
        _lib.check(_lib.load().nsk_neus_composite_bwd_dv(*head, _ptr(inv_t), *tail), "nsk_neus_composite_bwd_dv")
    return d_sdf, d_grad, d_alb, d_inv


def surface_points(origins: Tensor, ray_dirs: Tensor, p2p: Tensor, radius: float) -> Tensor:
    R = origins.shape[0]
    origins, ray_dirs = _chk("origins", origins, shape=(R, 3)), _chk("ray_dirs", ray_dirs, shape=(R, 3))
    p2p = _chk("p2p", p2p.reshape(R))
    out = torch.empty((R, 3), device=origins.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_surface_points(_ptr(origins), _ptr(ray_dirs), _ptr(p2p), c_int64(R), c_float(radius), _ptr(out), _stream(origins)), "nsk_surface_points")
    return out


# ------------------------------------------------------------------------------------------- RENI++
def reni_radiance_table(dirs: Tensor, latents: Tensor, scale: Optional[Tensor], packed: Tensor, rotation: Optional[Tensor] = None, hidden: int = 128, num_layers: int = 6, log_domain: bool = True) -> Tensor:
    """dirs [D,3], latents [K,L,3], scale [K] -> HDR radiance [K,D,3]."""
    D = dirs.shape[0]
    K, L = latents.shape[0], latents.shape[1]
    dirs = _chk("dirs", dirs, shape=(D, 3))
    latents = _chk("latents", latents, shape=(K, L, 3))
    if scale is not None:
        scale = _chk("scale", scale, shape=(K,))
    if rotation is not None:
        if rotation.dim() == 3:
            raise NotImplementedError("Batched rotation not implemented yet")  # reni_illumination_field.py:520-521
        rotation = _chk("rotation", rotation, shape=(3, 3))
    lib = _lib.load()
    need = lib.nsk_reni_weights_floats(c_int(L), c_int(hidden), c_int(num_layers))
    packed = _chk("packed", packed, shape=(need,))
    ws = torch.empty((K * num_layers * hidden + K * L * 2,), device=dirs.device, dtype=torch.float32)
    out = torch.empty((K, D, 3), device=dirs.device, dtype=torch.float32)
    _lib.check(lib.nsk_reni_decode_fwd(_ptr(dirs), c_int64(D), _ptr(latents), _ptr(scale), c_int64(K), _ptr(rotation), _ptr(packed), c_int(L), c_int(hidden), c_int(num_layers), c_int(int(log_domain)), _ptr(ws), _ptr(out), _stream(dirs)), "nsk_reni_decode_fwd")
    return out


def reni_rows_tc(dirs: Tensor, latents: Tensor, scale: Optional[Tensor], packed: Tensor, gemm_w: Dict[str, Tensor], rotation: Optional[Tensor] = None,
                 row_cam: Optional[Tensor] = None, hidden: int = 128, num_layers: int = 6, log_domain: bool = True, chunk: int = 1 << 20) -> Tensor:
    """dirs [N,3] (+ row_cam [N] int32 when K > 1), latents [K,L,3], scale [K] -> HDR radiance [N,3]: the decoder's 13 dense layers on
    the 3xTF32 tensor-core GEMM (fp32-accurate), prep / input rows / LayerNorm on the kernels of csrc/reni_rows_tc.cu.  Same
    result as reni_radiance_rows up to fp32 summation order; meant for large N (a frame's background rays).  `packed` =
    packing.pack_reni(...), `gemm_w` = packing.pack_reni_gemm(...).  Rows are processed in chunks to bound the [N,512] input buffer."""
    N = dirs.shape[0]
    K, L = latents.shape[0], latents.shape[1]
    dirs = _chk("dirs", dirs, shape=(N, 3))
    latents = _chk("latents", latents, shape=(K, L, 3))
    if row_cam is not None:
        row_cam = _chk("row_cam", row_cam, dtype=torch.int32, shape=(N,))
    elif K != 1:
        raise ValueError("reni_rows_tc: several latent codes need row_cam")
    if scale is not None:
        scale = _chk("scale", scale, shape=(K,))
    if rotation is not None:
        if rotation.dim() == 3:
            raise NotImplementedError("Batched rotation not implemented yet")  # reni_illumination_field.py:520-521
        rotation = _chk("rotation", rotation, shape=(3, 3))
    lib = _lib.load()
    need = lib.nsk_reni_weights_floats(c_int(L), c_int(hidden), c_int(num_layers))
    packed = _chk("packed", packed, shape=(need,))
    dev = dirs.device
    ws = torch.empty((K * num_layers * hidden + K * L * 2,), device=dev, dtype=torch.float32)
    _lib.check(lib.nsk_reni_prep(_ptr(latents), _ptr(rotation), c_int64(K), _ptr(packed), c_int(L), c_int(hidden), c_int(num_layers), _ptr(ws), _stream(dirs)), "nsk_reni_prep")
    attn = ws[:K * num_layers * hidden].view(K, num_layers, hidden)
    zxy = ws[K * num_layers * hidden:]
    out = torch.empty((N, 3), device=dev, dtype=torch.float32)
    st = _stream(dirs)
    for a in range(0, N, chunk):
        b = min(N, a + chunk)
        n = b - a
        d_c = dirs[a:b]
        rc = None if row_cam is None else row_cam[a:b]
        pe = torch.empty((n, 512), device=dev, dtype=torch.float32)
        _lib.check(lib.nsk_reni_pe_rows(_ptr(d_c), _ptr(rc), c_int64(n), _ptr(zxy), c_int(L), _ptr(pe), st), "nsk_reni_pe_rows")
        x = gemm_nt(pe, gemm_w["res_w"], bias=gemm_w["res_b"], split=3)                       # [n,128]
        del pe
        stride = c_int(num_layers * hidden)
        _lib.check(lib.nsk_reni_ln_rows(_ptr(x), c_int64(n), _ptr(attn[:, 0]), stride, _ptr(rc), _ptr(gemm_w["n1w0"]), _ptr(gemm_w["n1b0"]), _ptr(None), _ptr(None), _ptr(None), st),
                   "nsk_reni_ln_rows")                                                         # x = LN1_0(attn_0 + x)
        for i in range(num_layers):
            h = gemm_nt(x, gemm_w[f"f0w{i}"], bias=gemm_w[f"f0b{i}"], act="relu", split=3)
            gemm_nt(h, gemm_w[f"f2w{i}"], bias=gemm_w[f"f2b{i}"], out=x, accumulate=True, split=3)   # x += fc2(relu(fc1 x))
            if i + 1 < num_layers:    # x = LN1_{i+1}(attn_{i+1} + LN2_i(x)) in one pass
                _lib.check(lib.nsk_reni_ln_rows(_ptr(x), c_int64(n), _ptr(None), stride, _ptr(rc), _ptr(gemm_w[f"n2w{i}"]), _ptr(gemm_w[f"n2b{i}"]), _ptr(attn[:, i + 1]),
                                                _ptr(gemm_w[f"n1w{i + 1}"]), _ptr(gemm_w[f"n1b{i + 1}"]), st), "nsk_reni_ln_rows")
            else:
                _lib.check(lib.nsk_reni_ln_rows(_ptr(x), c_int64(n), _ptr(None), stride, _ptr(rc), _ptr(gemm_w[f"n2w{i}"]), _ptr(gemm_w[f"n2b{i}"]), _ptr(None), _ptr(None), _ptr(None), st),
                           "nsk_reni_ln_rows")
        o = gemm_nt(x, gemm_w["fc_w"], bias=gemm_w["fc_b"], split=3)                            # [n,3]
        if scale is not None:
            sc = scale if rc is None else scale[rc.long()][:, None]
            o = (o + sc) if log_domain else (o * torch.exp(sc))                                 # + log(exp(scale)) in the log domain (:561-565)
        out[a:b] = torch.exp(o) if log_domain else o
    return out


def reni_rows_fused(dirs: Tensor, latents: Tensor, scale: Optional[Tensor], packed: Tensor, fused_blob: Tensor, rotation: Optional[Tensor] = None,
                    row_cam: Optional[Tensor] = None, log_domain=True) -> Tensor:
    """dirs [N,3] (+ row_cam [N] int32 when K > 1), latents [K,L,3], scale [K] -> HDR radiance [N,3] through the FUSED tcgen05 kernel
    (csrc/reni_fused_tc.cu: fp16 operands, fp32 accumulate / LayerNorm; a row's activations never leave the SM).  `packed` =
    packing.pack_reni(...) (per-code prologue), `fused_blob` = packing.pack_reni_fused(...).  Meant for frame-sized N."""
    N = dirs.shape[0]
    K, L = latents.shape[0], latents.shape[1]
    dirs = _chk("dirs", dirs, shape=(N, 3))
    latents = _chk("latents", latents, shape=(K, L, 3))
    if row_cam is not None:
        row_cam = _chk("row_cam", row_cam, dtype=torch.int32, shape=(N,))
    elif K != 1:
        raise ValueError("reni_rows_fused: several latent codes need row_cam")
    if scale is not None:
        scale = _chk("scale", scale, shape=(K,))
    if rotation is not None:
        if rotation.dim() == 3:
            raise NotImplementedError("Batched rotation not implemented yet")  # reni_illumination_field.py:520-521
        rotation = _chk("rotation", rotation, shape=(3, 3))
    lib = _lib.load()
    hidden, num_layers = 128, 6
    packed = _chk("packed", packed, shape=(lib.nsk_reni_weights_floats(c_int(L), c_int(hidden), c_int(num_layers)),))
    fused_blob = _chk("fused_blob", fused_blob, dtype=torch.uint8, shape=(lib.nsk_reni_fused_weights_bytes(),))
    dev = dirs.device
    ws = torch.empty((K * num_layers * hidden + K * L * 2,), device=dev, dtype=torch.float32)
    st = _stream(dirs)
    _lib.check(lib.nsk_reni_prep(_ptr(latents), _ptr(rotation), c_int64(K), _ptr(packed), c_int(L), c_int(hidden), c_int(num_layers), _ptr(ws), st), "nsk_reni_prep")
    attn, zxy = ws[:K * num_layers * hidden], ws[K * num_layers * hidden:]
    out = torch.empty((N, 3), device=dev, dtype=torch.float32)
    _lib.check(lib.nsk_reni_rows_fused_fwd(_ptr(dirs), _ptr(row_cam), c_int64(N), _ptr(zxy), _ptr(attn), _ptr(scale), _ptr(fused_blob), c_int(L), c_int(int(log_domain)),
                                           _ptr(out), st), "nsk_reni_rows_fused_fwd")
    return out


def reni_radiance_rows(dirs: Tensor, row_cam: Tensor, latents: Tensor, scale: Optional[Tensor], packed: Tensor, rotation: Optional[Tensor] = None, hidden: int = 128, num_layers: int = 6, log_domain: bool = True) -> Tensor:
    """dirs [N,3], row_cam [N] int32 (latent code of each row), latents [K,L,3], scale [K] -> HDR radiance [N,3]
    (the per-ray background colours of a mixed-camera batch, neusky_model.py:535-549)."""
    N = dirs.shape[0]
    K, L = latents.shape[0], latents.shape[1]
    dirs = _chk("dirs", dirs, shape=(N, 3))
    row_cam = _chk("row_cam", row_cam, dtype=torch.int32, shape=(N,))
    latents = _chk("latents", latents, shape=(K, L, 3))
    if scale is not None:
        scale = _chk("scale", scale, shape=(K,))
    if rotation is not None:
        rotation = _chk("rotation", rotation, shape=(3, 3))
    lib = _lib.load()
    packed = _chk("packed", packed, shape=(lib.nsk_reni_weights_floats(c_int(L), c_int(hidden), c_int(num_layers)),))
    ws = torch.empty((K * num_layers * hidden + K * L * 2,), device=dirs.device, dtype=torch.float32)
    out = torch.empty((N, 3), device=dirs.device, dtype=torch.float32)
    _lib.check(lib.nsk_reni_decode_rows_fwd(_ptr(dirs), _ptr(row_cam), c_int64(N), _ptr(latents), _ptr(scale), c_int64(K), _ptr(rotation), _ptr(packed), c_int(L), c_int(hidden), c_int(num_layers), c_int(int(log_domain)), _ptr(ws), _ptr(out), _stream(dirs)), "nsk_reni_decode_rows_fwd")
    return out


def reni_decode_bwd(dirs: Tensor, row_cam: Optional[Tensor], latents: Tensor, scale: Optional[Tensor], packed: Tensor, packed_bwd: Tensor, out: Tensor, g_out: Tensor,
                    d_latents: Tensor, d_scale: Optional[Tensor], rotation: Optional[Tensor] = None, hidden: int = 128, num_layers: int = 6, log_domain: bool = True) -> None:
    """Accumulates d loss / d latents [K,L,3] and d loss / d scale [K] of reni_radiance_table (row_cam None; out, g_out [K,D,3])
    or reni_radiance_rows (out, g_out [N,3]); the decoder weights are frozen (RENIField.hold_decoder_fixed)."""
    D = dirs.shape[0]
    K, L = latents.shape[0], latents.shape[1]
    dirs = _chk("dirs", dirs, shape=(D, 3))
    latents = _chk("latents", latents, shape=(K, L, 3))
    oshape = (D, 3) if row_cam is not None else (K, D, 3)
    out, g_out = _chk("out", out, shape=oshape), _chk("g_out", g_out, shape=oshape)
    d_latents = _chk("d_latents", d_latents, shape=(K, L, 3))
    if row_cam is not None:
        row_cam = _chk("row_cam", row_cam, dtype=torch.int32, shape=(D,))
    if (scale is None) != (d_scale is None):
        raise ValueError("reni_decode_bwd: scale and d_scale must be given together")
    if scale is not None:
        scale, d_scale = _chk("scale", scale, shape=(K,)), _chk("d_scale", d_scale, shape=(K,))
    if rotation is not None:
        rotation = _chk("rotation", rotation, shape=(3, 3))
    lib = _lib.load()
    packed = _chk("packed", packed, shape=(lib.nsk_reni_weights_floats(c_int(L), c_int(hidden), c_int(num_layers)),))
    packed_bwd = _chk("packed_bwd", packed_bwd, shape=(lib.nsk_reni_bwd_weights_floats(c_int(L), c_int(hidden), c_int(num_layers)),))
    ws = torch.empty((lib.nsk_reni_bwd_workspace_floats(c_int64(K), c_int(L), c_int(hidden), c_int(num_layers)),), device=dirs.device, dtype=torch.float32)
    _lib.check(lib.nsk_reni_decode_bwd(_ptr(dirs), _ptr(row_cam), c_int64(D), _ptr(latents), _ptr(scale), c_int64(K), _ptr(rotation), _ptr(packed), _ptr(packed_bwd),
                                       c_int(L), c_int(hidden), c_int(num_layers), c_int(int(log_domain)), _ptr(out), _ptr(g_out), _ptr(ws), _ptr(d_latents), _ptr(d_scale),
                                       _stream(dirs)), "nsk_reni_decode_bwd")


# ------------------------------------------------------------------------------------------- Lambert / K4
def lambert_prep(normals, wa, dirs, ddf_mask, radiance, cam=None, unoccluded_vis: float = 1.0):
    R, S = normals.shape[0], normals.shape[1]
    D = dirs.shape[0]
    normals, wa = _chk("normals", normals, shape=(R, S, 3)), _chk("wa", wa, shape=(R, S, 3))
    dirs = _chk("dirs", dirs, shape=(D, 3))
    ddf_mask = _chk("ddf_mask", ddf_mask, dtype=torch.uint8, shape=(D,))
    radiance = _chk("radiance", radiance, shape=(None, D, 3))
    if cam is not None:
        cam = _chk("cam", cam, dtype=torch.int32, shape=(R,))
    inv_count = torch.empty((R, S), device=normals.device, dtype=torch.float32)
    rgb_lin = torch.empty((R, 3), device=normals.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_lambert_prep(_ptr(normals), _ptr(wa), c_int64(R), c_int(S), _ptr(dirs), _ptr(ddf_mask), c_int(D), _ptr(radiance), _ptr(cam), c_float(unoccluded_vis), _ptr(inv_count), _ptr(rgb_lin), _stream(normals)), "nsk_lambert_prep")
    return inv_count, rgb_lin


def lambert_relight(normals, wa, inv_count, dirs, sel_index, radiance, vis_sel, cam=None, unoccluded_vis: float = 1.0) -> Tensor:
    """Lambertian sum with cached per-ray visibility vis_sel [R,Dp] -> linear rgb [R,3] (config 5: relighting)."""
    R, S = normals.shape[0], normals.shape[1]
    D = dirs.shape[0]
    Dp = vis_sel.shape[1]
    normals, wa = _chk("normals", normals, shape=(R, S, 3)), _chk("wa", wa, shape=(R, S, 3))
    inv_count = _chk("inv_count", inv_count, shape=(R, S))
    dirs = _chk("dirs", dirs, shape=(D, 3))
    sel_index = _chk("sel_index", sel_index, dtype=torch.int32, shape=(D,))
    radiance = _chk("radiance", radiance, shape=(None, D, 3))
    vis_sel = _chk("vis_sel", vis_sel, shape=(R, Dp))
    if cam is not None:
        cam = _chk("cam", cam, dtype=torch.int32, shape=(R,))
    rgb_lin = torch.empty((R, 3), device=normals.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_lambert_relight(_ptr(normals), _ptr(wa), _ptr(inv_count), c_int64(R), c_int(S), _ptr(dirs), _ptr(sel_index), c_int(D), c_int(Dp), _ptr(radiance), _ptr(cam), _ptr(vis_sel), c_float(unoccluded_vis), _ptr(rgb_lin), _stream(normals)), "nsk_lambert_relight")
    return rgb_lin


def lambert_collapse(normals, wa, inv_count, dirs, sel_index, vis_sel, unoccluded_vis: float = 1.0) -> Tensor:
    """Per-(ray, direction) shading coefficients H [R,D,3] with visibility folded in: the relighting cache (config 5)."""
    R, S = normals.shape[0], normals.shape[1]
    D, Dp = dirs.shape[0], vis_sel.shape[1]
    normals, wa = _chk("normals", normals, shape=(R, S, 3)), _chk("wa", wa, shape=(R, S, 3))
    inv_count = _chk("inv_count", inv_count, shape=(R, S))
    dirs = _chk("dirs", dirs, shape=(D, 3))
    sel_index = _chk("sel_index", sel_index, dtype=torch.int32, shape=(D,))
    vis_sel = _chk("vis_sel", vis_sel, shape=(R, Dp))
    H = torch.empty((R, D, 3), device=normals.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_lambert_collapse(_ptr(normals), _ptr(wa), _ptr(inv_count), c_int64(R), c_int(S), _ptr(dirs), _ptr(sel_index), c_int(D), c_int(Dp),
                                                _ptr(vis_sel), c_float(unoccluded_vis), _ptr(H), _stream(normals)), "nsk_lambert_collapse")
    return H


def lambert_collapse_sel(normals, wa, inv_count, dirs_sel) -> Tensor:
    """G [R,Dp,3]: per (ray, DDF direction) Lambert coefficients summed over the ray's samples (K4 with S = 0 reads them)."""
    R, S = normals.shape[0], normals.shape[1]
    Dp = dirs_sel.shape[0]
    normals, wa = _chk("normals", normals, shape=(R, S, 3)), _chk("wa", wa, shape=(R, S, 3))
    inv_count = _chk("inv_count", inv_count, shape=(R, S))
    dirs_sel = _chk("dirs_sel", dirs_sel, shape=(Dp, 3))
    G = torch.empty((R, Dp, 3), device=normals.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_lambert_collapse_sel(_ptr(normals), _ptr(wa), _ptr(inv_count), c_int64(R), c_int(S), _ptr(dirs_sel), c_int(Dp), _ptr(G), _stream(normals)),
               "nsk_lambert_collapse_sel")
    return G


def relight_collapsed(H: Tensor, radiance: Tensor, cam: Optional[Tensor] = None) -> Tensor:
    """H [R,D,3], radiance [K,D,3] -> linear rgb [R,3]: one streaming pass per new illumination."""
    R, D = H.shape[0], H.shape[1]
    H = _chk("H", H, shape=(R, D, 3))
    radiance = _chk("radiance", radiance, shape=(None, D, 3))
    if cam is not None:
        cam = _chk("cam", cam, dtype=torch.int32, shape=(R,))
    elif radiance.shape[0] != 1:
        raise ValueError("radiance has several tables but no per-ray camera index was given")
    rgb_lin = torch.empty((R, 3), device=H.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_relight_collapsed(_ptr(H), c_int64(R), c_int(D), _ptr(radiance), _ptr(cam), _ptr(rgb_lin), _stream(H)), "nsk_relight_collapsed")
    return rgb_lin


def relight_collapsed_multi(H: Tensor, radiance: Tensor) -> Tensor:
    """H [R,D,3], radiance [NL,D,3] (one table per illumination) -> linear rgb [NL,R,3], reading H once per 4 illuminations."""
    R, D = H.shape[0], H.shape[1]
    H = _chk("H", H, shape=(R, D, 3))
    radiance = _chk("radiance", radiance, shape=(None, D, 3))
    NL = radiance.shape[0]
    out = torch.empty((NL, R, 3), device=H.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_relight_collapsed_multi(_ptr(H), c_int64(R), c_int(D), _ptr(radiance), c_int(NL), _ptr(out), _stream(H)), "nsk_relight_collapsed_multi")
    return out


def relight_pack_h16(H: Tensor, rows: Tensor) -> Tuple[Tensor, Tensor]:
    """H [R,D,3] fp32 coefficients + rows [Rs] int32 (the rays that own a cache row) -> (H16 [Rs, 3*DP] fp16 channel-planar rows
    normalised by their maximum, hscale [Rs]): the compact relighting cache (csrc/relight_compact.cu)."""
    R, D = H.shape[0], H.shape[1]
    H = _chk("H", H, shape=(R, D, 3))
    rows = _chk("rows", rows, dtype=torch.int32, shape=(None,))
    Rs, DP = rows.shape[0], (D + 15) // 16 * 16
    H16 = torch.empty((Rs, 3 * DP), device=H.device, dtype=torch.float16)
    hscale = torch.empty((Rs,), device=H.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_relight_pack_h16(_ptr(H), _ptr(rows), c_int64(Rs), c_int(D), _ptr(H16), _ptr(hscale), _stream(H)), "nsk_relight_pack_h16")
    return H16, hscale


def relight_h16_multi(H16: Tensor, hscale: Tensor, rows: Tensor, R: int, D: int, radiance: Tensor) -> Tensor:
    """Compact cache + radiance [NL,D,3] -> linear rgb [NL,R,3] (zero for rays without a cache row), up to 32 illuminations per pass."""
    Rs, DP = rows.shape[0], (D + 15) // 16 * 16
    H16 = _chk("H16", H16, dtype=torch.float16, shape=(Rs, 3 * DP))
    hscale = _chk("hscale", hscale, shape=(Rs,))
    rows = _chk("rows", rows, dtype=torch.int32, shape=(Rs,))
    radiance = _chk("radiance", radiance, shape=(None, D, 3))
    NL = radiance.shape[0]
    out = torch.empty((NL, R, 3), device=radiance.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_relight_h16_multi(_ptr(H16), _ptr(hscale), _ptr(rows), c_int64(Rs), c_int64(R), c_int(D), _ptr(radiance), c_int(NL), _ptr(out),
                                                 _stream(radiance)), "nsk_relight_h16_multi")
    return out


def lambert_relight_bwd(normals, wa, inv_count, dirs, sel_index, radiance, vis_sel, g_rgb_lin, cam=None, unoccluded_vis: float = 1.0, want_vis: bool = True, want_radiance: bool = True):
    """-> (d_wa [R,S,3], d_normals [R,S,3], d_vis_sel [R,Dp] | None, d_radiance [K,D,3] | None)."""
    R, S = normals.shape[0], normals.shape[1]
    D, Dp = dirs.shape[0], vis_sel.shape[1]
    normals, wa = _chk("normals", normals, shape=(R, S, 3)), _chk("wa", wa, shape=(R, S, 3))
    inv_count = _chk("inv_count", inv_count, shape=(R, S))
    dirs = _chk("dirs", dirs, shape=(D, 3))
    sel_index = _chk("sel_index", sel_index, dtype=torch.int32, shape=(D,))
    radiance = _chk("radiance", radiance, shape=(None, D, 3))
    vis_sel = _chk("vis_sel", vis_sel, shape=(R, Dp))
    g = _chk("g_rgb_lin", g_rgb_lin, shape=(R, 3))
    if cam is not None:
        cam = _chk("cam", cam, dtype=torch.int32, shape=(R,))
    f = dict(device=normals.device, dtype=torch.float32)
    d_wa, d_n = torch.empty((R, S, 3), **f), torch.empty((R, S, 3), **f)
    d_vis = torch.empty((R, Dp), **f) if want_vis else None
    d_rad = torch.zeros(tuple(radiance.shape), **f) if want_radiance else None
    _lib.check(_lib.load().nsk_lambert_relight_bwd(_ptr(normals), _ptr(wa), _ptr(inv_count), c_int64(R), c_int(S), _ptr(dirs), _ptr(sel_index), c_int(D), c_int(Dp), _ptr(radiance), _ptr(cam), _ptr(vis_sel),
                                                   c_float(unoccluded_vis), _ptr(g), _ptr(d_wa), _ptr(d_n), _ptr(d_vis), _ptr(d_rad), _stream(normals)), "nsk_lambert_relight_bwd")
    return d_wa, d_n, d_vis, d_rad


def shade_finalize_bwd(rgb_lin, bg, acc, g_rgb):
    R = rgb_lin.shape[0]
    rgb_lin, bg, g_rgb = _chk("rgb_lin", rgb_lin, shape=(R, 3)), _chk("bg", bg, shape=(R, 3)), _chk("g_rgb", g_rgb, shape=(R, 3))
    acc = _chk("acc", acc.reshape(R))
    f = dict(device=rgb_lin.device, dtype=torch.float32)
    d_lin, d_bg, d_acc = torch.empty((R, 3), **f), torch.empty((R, 3), **f), torch.empty((R,), **f)
    _lib.check(_lib.load().nsk_shade_finalize_bwd(_ptr(rgb_lin), _ptr(bg), _ptr(acc), _ptr(g_rgb), c_int64(R), _ptr(d_lin), _ptr(d_bg), _ptr(d_acc), _stream(rgb_lin)), "nsk_shade_finalize_bwd")
    return d_lin, d_bg, d_acc


def shade_finalize(rgb_lin, bg, acc, training: bool = False) -> Tensor:
    R = rgb_lin.shape[0]
    rgb_lin, bg = _chk("rgb_lin", rgb_lin, shape=(R, 3)), _chk("bg", bg, shape=(R, 3))
    acc = _chk("acc", acc.reshape(R))
    out = torch.empty((R, 3), device=rgb_lin.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_shade_finalize(_ptr(rgb_lin), _ptr(bg), _ptr(acc), c_int64(R), c_int(int(training)), _ptr(out), _stream(rgb_lin)), "nsk_shade_finalize")
    return out


COLLAPSE_MIN_SAMPLES = 8     # K4 (tc2): from this many samples per ray on, pre-collapse the Lambert coefficients


def sky_shade(points, normals, wa, inv_count, dirs_sel, radiance_sel, ddf_blob, hash_table, scalings, log2_T: int, radius: float, threshold: float, sigmoid_scale: float, rgb_lin: Tensor, cam=None, want_vis: bool = False, want_ddf: bool = False, impl: str = "tc",
              grid_meta: Optional[Tensor] = None, smoothstep: bool = True):
    """K4.  Accumulates into ``rgb_lin`` [R,3]; returns (vis [R,Dp] | None, ddf [R*Dp] | None, term | None).
    ``grid_meta`` / ``smoothstep``: the DDF position grid is an imported tiny-cuda-nn grid (see ``sdf_field``; impl 'tc2' and 'simt')."""
    R, S = normals.shape[0], normals.shape[1]
    Dp = dirs_sel.shape[0]
    points = _chk("points", points, shape=(R, 3))
    normals, wa = _chk("normals", normals, shape=(R, S, 3)), _chk("wa", wa, shape=(R, S, 3))
    inv_count = _chk("inv_count", inv_count, shape=(R, S))
    dirs_sel = _chk("dirs_sel", dirs_sel, shape=(Dp, 3))
    radiance_sel = _chk("radiance_sel", radiance_sel, shape=(None, Dp, 3))
    if cam is not None:
        cam = _chk("cam", cam, dtype=torch.int32, shape=(R,))
    L = scalings.numel()
    hash_table = _chk("hash_table", hash_table, shape=(L << log2_T, 2))
    scalings = _chk("scalings", scalings)
    if not (rgb_lin.is_cuda and rgb_lin.dtype == torch.float32 and rgb_lin.is_contiguous() and tuple(rgb_lin.shape) == (R, 3)):
        raise ValueError("rgb_lin: expected a contiguous CUDA fp32 [R,3] accumulator")
    dev = points.device
    vis = torch.empty((R, Dp), device=dev, dtype=torch.float32) if want_vis else None
    ddf = torch.empty((R * Dp,), device=dev, dtype=torch.float32) if want_ddf else None
    term = torch.empty((R * Dp,), device=dev, dtype=torch.float32) if want_ddf else None
    lib = _lib.load()
    if impl == "tc2" and S >= COLLAPSE_MIN_SAMPLES:
        # full renders: sum the ray's samples once per (ray, direction) up front instead of per pair inside K4's tail
        wa = lambert_collapse_sel(normals, wa, inv_count, dirs_sel)
        normals = inv_count = None
        S = 0
    if impl == "simt":
        blob = _chk("ddf_blob", ddf_blob, shape=(lib.nsk_ddf_simt_weights_floats(),))
        fn, name = lib.nsk_sky_shade_simt_fwd, "nsk_sky_shade_simt_fwd"
    elif impl == "tc":
        blob = _chk("ddf_blob", ddf_blob, dtype=torch.uint8, shape=(lib.nsk_ddf_tc_weights_bytes(),))
        fn, name = lib.nsk_sky_shade_tc_fwd, "nsk_sky_shade_tc_fwd"
    elif impl == "tc2":
        blob = _chk("ddf_blob", ddf_blob, dtype=torch.uint8, shape=(lib.nsk_ddf_tc2_weights_bytes(),))
        fn, name = lib.nsk_sky_shade_tc2_fwd, "nsk_sky_shade_tc2_fwd"
    else:
        raise ValueError(f"impl must be 'tc2', 'tc' or 'simt', got {impl!r}")
    if grid_meta is not None:
        if impl == "tc":
            raise ValueError("sky_shade: an imported tiny-cuda-nn grid needs impl 'tc2' (default) or 'simt'")
        grid_meta = _chk("grid_meta", grid_meta, dtype=torch.int32, shape=(L, 4))
        fn, name = (lib.nsk_sky_shade_tc2_fwd_ex, "nsk_sky_shade_tc2_fwd_ex") if impl == "tc2" else (lib.nsk_sky_shade_simt_fwd_ex, "nsk_sky_shade_simt_fwd_ex")
        _lib.check(fn(_ptr(points), c_int64(R), _ptr(normals), _ptr(wa), _ptr(inv_count), c_int(S), _ptr(dirs_sel), c_int(Dp), _ptr(radiance_sel), _ptr(cam), _ptr(blob), _ptr(hash_table), _ptr(scalings), c_int(L), c_int(log2_T), _ptr(grid_meta), c_int(int(smoothstep)), c_float(radius), c_float(threshold), c_float(sigmoid_scale), _ptr(rgb_lin), _ptr(vis), _ptr(ddf), _ptr(term), _stream(points)), name)
        return vis, ddf, term
    _lib.check(fn(_ptr(points), c_int64(R), _ptr(normals), _ptr(wa), _ptr(inv_count), c_int(S), _ptr(dirs_sel), c_int(Dp), _ptr(radiance_sel), _ptr(cam), _ptr(blob), _ptr(hash_table), _ptr(scalings), c_int(L), c_int(log2_T), c_float(radius), c_float(threshold), c_float(sigmoid_scale), _ptr(rgb_lin), _ptr(vis), _ptr(ddf), _ptr(term), _stream(points)), name)
    return vis, ddf, term


# ------------------------------------------------------------------------------------------- training path
ACT = {"none": 0, "relu": 1, "leaky": 2, "softplus100": 3, "sigmoid": 4}


def _mat(name: str, t: Tensor, rows: Optional[int] = None, cols: Optional[int] = None) -> Tuple[Tensor, int]:
    """A row-major 2-D CUDA fp32 view with unit column stride (row stride = leading dimension); slices of wider
    buffers are passed without a copy."""
    if not isinstance(t, torch.Tensor) or t.dim() != 2:
        raise ValueError(f"{name}: expected a 2-D tensor")
    if not t.is_cuda or t.dtype != torch.float32:
        raise ValueError(f"{name}: expected a CUDA fp32 tensor (neusky_b200 has no CPU path), got {t.device} {t.dtype}")
    if t.shape[1] > 1 and t.stride(1) != 1:
        raise ValueError(f"{name}: columns must be contiguous (stride {t.stride()})")
    if (rows is not None and t.shape[0] != rows) or (cols is not None and t.shape[1] != cols):
        raise ValueError(f"{name}: expected shape ({rows}, {cols}), got {tuple(t.shape)}")
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))
    return t, int(ld)


def gemm_nt(A: Tensor, B: Tensor, bias: Optional[Tensor] = None, act: str = "none", out: Optional[Tensor] = None, aux: Optional[Tensor] = None,
            dact: str = "none", accumulate: bool = False, split: int = 1) -> Tensor:
    """out[M,N] = dact'(aux) * act(A[M,K] @ B[N,K]^T + bias) (+ out).  tcgen05 kind::tf32; split=3 is the fp32-accurate 3xTF32 mode."""
    A, lda = _mat("A", A)
    B, ldb = _mat("B", B, cols=A.shape[1])
    M, K = A.shape
    N = B.shape[0]
    if out is None:
        if accumulate:
            raise ValueError("gemm_nt: accumulate needs `out`")
        out = torch.empty((M, N), device=A.device, dtype=torch.float32)
    out, ldc = _mat("out", out, rows=M, cols=N)
    if bias is not None:
        bias = _chk("bias", bias, shape=(N,))
    auxp, ldaux = None, 0
    if dact != "none":
        auxp, ldaux = _mat("aux", aux, rows=M, cols=N)
    _lib.check(_lib.load().nsk_gemm_tf32_nt(_ptr(A), c_int(lda), _ptr(B), c_int(ldb), _ptr(out), c_int(ldc), c_int64(M), c_int(N), c_int(K), _ptr(bias), c_int(ACT[act]),
                                            _ptr(auxp), c_int(ldaux), c_int(ACT[dact]), c_int(int(accumulate)), c_int(split), _stream(A)), "nsk_gemm_tf32_nt")
    return out


def gemm_tn(A: Tensor, B: Tensor, out: Tensor, split: int = 1) -> Tensor:
    """out[P,Q] += A[M,P]^T @ B[M,Q]  (weight gradient; `out` holds zeros or a running sum)."""
    A, lda = _mat("A", A)
    B, ldb = _mat("B", B, rows=A.shape[0])
    out, ldc = _mat("out", out, rows=A.shape[1], cols=B.shape[1])
    _lib.check(_lib.load().nsk_gemm_tf32_tn(_ptr(A), c_int(lda), _ptr(B), c_int(ldb), _ptr(out), c_int(ldc), c_int64(A.shape[0]), c_int(A.shape[1]), c_int(B.shape[1]),
                                            c_int(split), _stream(A)), "nsk_gemm_tf32_tn")
    return out


def colsum(X: Tensor, out: Tensor) -> Tensor:
    """out[c] += sum_r X[r, c]."""
    X, ld = _mat("X", X)
    out = _chk_out("out", out, shape=(X.shape[1],))
    _lib.check(_lib.load().nsk_colsum(_ptr(X), c_int(ld), c_int64(X.shape[0]), c_int(X.shape[1]), _ptr(out), _stream(X)), "nsk_colsum")
    return out


def ddf_pairs(points: Tensor, dirs_sel: Tensor, hash_table: Tensor, scalings: Tensor, log2_T: int, radius: float):
    """(points [R,3], dirs [D',3]) -> cond [N,40], xin [N,16], q [N,3], term_dist [N] with N = R*D' (pair i = r*D' + j)."""
    R, D = points.shape[0], dirs_sel.shape[0]
    points, dirs_sel = _chk("points", points, shape=(R, 3)), _chk("dirs_sel", dirs_sel, shape=(D, 3))
    L = scalings.numel()
    hash_table, scalings = _chk("hash_table", hash_table, shape=(L << log2_T, 2)), _chk("scalings", scalings)
    N, dev = R * D, points.device
    cond = torch.empty((N, 40), device=dev, dtype=torch.float32)
    xin = torch.empty((N, 16), device=dev, dtype=torch.float32)
    q = torch.empty((N, 3), device=dev, dtype=torch.float32)
    term = torch.empty((N,), device=dev, dtype=torch.float32)
    _lib.check(_lib.load().nsk_ddf_pairs_fwd(_ptr(points), c_int64(R), _ptr(dirs_sel), c_int(D), _ptr(hash_table), _ptr(scalings), c_int(L), c_int(log2_T), c_float(radius),
                                             _ptr(cond), _ptr(xin), _ptr(q), _ptr(term), _stream(points)), "nsk_ddf_pairs_fwd")
    return cond, xin, q, term


def ddf_rows(origins: Tensor, directions: Tensor, hash_table: Tensor, scalings: Tensor, log2_T: int):
    """(origins [N,3] on the DDF sphere, world directions [N,3]) -> cond [N,40], xin [N,16] (row-wise DDF inputs)."""
    N = origins.shape[0]
    origins, directions = _chk("origins", origins, shape=(N, 3)), _chk("directions", directions, shape=(N, 3))
    L = scalings.numel()
    hash_table, scalings = _chk("hash_table", hash_table, shape=(L << log2_T, 2)), _chk("scalings", scalings)
    cond = torch.empty((N, 40), device=origins.device, dtype=torch.float32)
    xin = torch.empty((N, 16), device=origins.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_ddf_rows_fwd(_ptr(origins), _ptr(directions), c_int64(N), _ptr(hash_table), _ptr(scalings), c_int(L), c_int(log2_T),
                                            _ptr(cond), _ptr(xin), _stream(origins)), "nsk_ddf_rows_fwd")
    return cond, xin


def film_sin(z: Tensor, film: Tensor, layer: int) -> Tensor:
    N = z.shape[0]
    z, film = _chk("z", z, shape=(N, 256)), _chk("film", film, shape=(N, None))
    a = torch.empty_like(z)
    _lib.check(_lib.load().nsk_film_sin_fwd(_ptr(z), _ptr(film), c_int(film.shape[1]), c_int(layer), c_int64(N), _ptr(a), _stream(z)), "nsk_film_sin_fwd")
    return a


def film_sin_bwd(da: Tensor, z: Tensor, film: Tensor, layer: int, dfilm: Tensor, sum_dz: Optional[Tensor] = None, sum_dfilm: Optional[Tensor] = None) -> Tensor:
    """Returns d z [N,256]; writes the layer's frequency / phase column blocks of ``dfilm`` [N, ldf].  With ``sum_dz`` [256] and
    ``sum_dfilm`` [ldf] (both or neither; ACCUMULATED INTO, caller zero-fills) the column sums of d z and of those ``dfilm`` blocks
    come out of the same pass (the bias gradients of the trunk layer and of the last mapping layer)."""
    N = z.shape[0]
    da, z, film = _chk("da", da, shape=(N, 256)), _chk("z", z, shape=(N, 256)), _chk("film", film, shape=(N, None))
    if not (dfilm.is_cuda and dfilm.dtype == torch.float32 and dfilm.is_contiguous() and dfilm.shape == film.shape):
        raise ValueError("dfilm: expected a contiguous CUDA fp32 tensor shaped like film")
    dz = torch.empty_like(z)
    if (sum_dz is None) != (sum_dfilm is None):
        raise ValueError("film_sin_bwd: pass both sum_dz and sum_dfilm or neither")
    if sum_dz is None:
        _lib.check(_lib.load().nsk_film_sin_bwd(_ptr(da), _ptr(z), _ptr(film), c_int(film.shape[1]), c_int(layer), c_int64(N), _ptr(dz), _ptr(dfilm), _stream(z)), "nsk_film_sin_bwd")
    else:
        sum_dz, sum_dfilm = _chk_out("sum_dz", sum_dz, shape=(256,)), _chk_out("sum_dfilm", sum_dfilm, shape=(film.shape[1],))
        _lib.check(_lib.load().nsk_film_sin_bwd_sums(_ptr(da), _ptr(z), _ptr(film), c_int(film.shape[1]), c_int(layer), c_int64(N), _ptr(dz), _ptr(dfilm),
                                                     _ptr(sum_dz), _ptr(sum_dfilm), _stream(z)), "nsk_film_sin_bwd_sums")
    return dz


def ddf_head(a5: Tensor, w_final: Tensor, b_final: Tensor, term_dist: Tensor, radius: float, threshold: Tensor, sigmoid_scale: float):
    """-> (that [N] expected termination distance, vis [N])."""
    N = a5.shape[0]
    a5, w_final, b_final = _chk("a5", a5, shape=(N, 256)), _chk("w_final", w_final.reshape(-1), shape=(256,)), _chk("b_final", b_final.reshape(-1), shape=(1,))
    term_dist, threshold = _chk("term_dist", term_dist, shape=(N,)), _chk("threshold", threshold.reshape(-1), shape=(1,))
    that, vis = torch.empty_like(term_dist), torch.empty_like(term_dist)
    _lib.check(_lib.load().nsk_ddf_head_fwd(_ptr(a5), _ptr(w_final), _ptr(b_final), _ptr(term_dist), c_int64(N), c_float(radius), _ptr(threshold), c_float(sigmoid_scale),
                                            _ptr(that), _ptr(vis), _stream(a5)), "nsk_ddf_head_fwd")
    return that, vis


def ddf_head_bwd(a5: Tensor, w_final: Tensor, that: Tensor, term_dist: Tensor, d_vis: Optional[Tensor], d_that_extra: Optional[Tensor], radius: float, threshold: Tensor,
                 sigmoid_scale: float, d_w_final: Tensor, d_b_final: Tensor, d_threshold: Optional[Tensor]) -> Tensor:
    """-> d a5 [N,256]; accumulates d w_final [256], d b_final [1], d threshold [1]."""
    N = a5.shape[0]
    a5, w_final = _chk("a5", a5, shape=(N, 256)), _chk("w_final", w_final.reshape(-1), shape=(256,))
    that, term_dist, threshold = _chk("that", that, shape=(N,)), _chk("term_dist", term_dist, shape=(N,)), _chk("threshold", threshold.reshape(-1), shape=(1,))
    if d_vis is not None:
        d_vis = _chk("d_vis", d_vis.reshape(-1), shape=(N,))
    if d_that_extra is not None:
        d_that_extra = _chk("d_that_extra", d_that_extra.reshape(-1), shape=(N,))
    for nm, t, n in (("d_w_final", d_w_final, 256), ("d_b_final", d_b_final, 1)) + ((("d_threshold", d_threshold, 1),) if d_threshold is not None else ()):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n):
            raise ValueError(f"{nm}: expected a contiguous CUDA fp32 accumulator with {n} elements")
    da5 = torch.empty_like(a5)
    _lib.check(_lib.load().nsk_ddf_head_bwd(_ptr(a5), _ptr(w_final), _ptr(that), _ptr(term_dist), _ptr(d_vis), _ptr(d_that_extra), c_int64(N), c_float(radius), _ptr(threshold),
                                            c_float(sigmoid_scale), _ptr(da5), _ptr(d_w_final), _ptr(d_b_final), _ptr(d_threshold), _stream(a5)), "nsk_ddf_head_bwd")
    return da5


def colsum_w(X: Tensor, v: Tensor, out: Tensor) -> Tensor:
    """out[c] += sum_r v[r] X[r, c]."""
    X, ld = _mat("X", X)
    v = _chk("v", v.reshape(-1), shape=(X.shape[0],))
    out = _chk_out("out", out, shape=(X.shape[1],))
    _lib.check(_lib.load().nsk_colsum_w(_ptr(X), c_int(ld), _ptr(v), c_int64(X.shape[0]), c_int(X.shape[1]), _ptr(out), _stream(X)), "nsk_colsum_w")
    return out


def sdf_inputs(x: Tensor, hash_table: Tensor, scalings: Tensor, log2_T: int, tail: Optional[Tensor] = None):
    """x [n,3] -> (H0 [n,72], pos [n,3], J [n,9]); also fills tail[:, 0:40] (a column slice of the colour-net input) if given."""
    n = x.shape[0]
    x = _chk("x", x, shape=(n, 3))
    L = scalings.numel()
    hash_table, scalings = _chk("hash_table", hash_table, shape=(L << log2_T, 2)), _chk("scalings", scalings)
    dev = x.device
    H0 = torch.empty((n, 72), device=dev, dtype=torch.float32)
    pos = torch.empty((n, 3), device=dev, dtype=torch.float32)
    J = torch.empty((n, 9), device=dev, dtype=torch.float32)
    ld_tail = 0
    if tail is not None:
        tail, ld_tail = _mat("tail", tail, rows=n, cols=40)
    _lib.check(_lib.load().nsk_sdf_inputs_fwd(_ptr(x), c_int64(n), _ptr(hash_table), _ptr(scalings), c_int(L), c_int(log2_T), _ptr(H0), _ptr(tail), c_int(ld_tail),
                                              _ptr(pos), _ptr(J), _stream(x)), "nsk_sdf_inputs_fwd")
    return H0, pos, J


def sdf_grad_assemble(x: Tensor, G: Tensor, gpos: Tensor, J: Tensor) -> Tensor:
    n = x.shape[0]
    x, G, gpos, J = _chk("x", x, shape=(n, 3)), _chk("G", G, shape=(n, 72)), _chk("gpos", gpos, shape=(n, 3)), _chk("J", J, shape=(n, 9))
    grad = torch.empty((n, 3), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_sdf_grad_assemble(_ptr(x), _ptr(G), _ptr(gpos), _ptr(J), c_int64(n), _ptr(grad), _stream(x)), "nsk_sdf_grad_assemble")
    return grad


def sdf_grad_assemble_bwd(x: Tensor, c: Tensor, J: Tensor):
    """-> (dG [n,72] with the hash columns 39..70 left for the caller, cpos [n,3])."""
    n = x.shape[0]
    x, c, J = _chk("x", x, shape=(n, 3)), _chk("c", c, shape=(n, 3)), _chk("J", J, shape=(n, 9))
    dG = torch.empty((n, 72), device=x.device, dtype=torch.float32)
    cpos = torch.empty((n, 3), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_sdf_grad_assemble_bwd(_ptr(x), _ptr(c), _ptr(J), c_int64(n), _ptr(dG), _ptr(cpos), _stream(x)), "nsk_sdf_grad_assemble_bwd")
    return dG, cpos


EW = {"sp_chain": 0, "mul_dsp": 1, "sp_bwd2": 2, "sp_bwd2_w": 3, "outer_add": 4}


def ew256(op: str, n: int, a=None, b=None, c=None, d=None, w=None, s=None) -> Tensor:
    """Pointwise family over [n,256] tensors (see include/neusky_b200.h, nsk_ew256)."""
    ref = next(t for t in (a, b, c, d) if t is not None) if any(t is not None for t in (a, b, c, d)) else w
    chk = lambda nm, t: None if t is None else _chk(nm, t, shape=(n, 256))
    a, b, c, d = chk("a", a), chk("b", b), chk("c", c), chk("d", d)
    if w is not None:
        w = _chk("w", w.reshape(-1), shape=(256,))
    if s is not None:
        s = _chk("s", s.reshape(-1), shape=(n,))
    out = torch.empty((n, 256), device=ref.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_ew256(c_int(EW[op]), c_int64(n), _ptr(a), _ptr(b), _ptr(c), _ptr(d), _ptr(w), _ptr(s), _ptr(out), _stream(out)), "nsk_ew256")
    return out


def rowdot256(X: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    n = X.shape[0]
    X, w = _chk("X", X, shape=(n, 256)), _chk("w", w.reshape(-1), shape=(256,))
    if b is not None:
        b = _chk("b", b.reshape(-1), shape=(1,))
    out = torch.empty((n,), device=X.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_rowdot256(_ptr(X), _ptr(w), _ptr(b), c_int64(n), _ptr(out), _stream(X)), "nsk_rowdot256")
    return out


# ------------------------------------------------------------------------------------------- light-sum shaders (csrc/shaders.cu)
def _shade_args(mode: int, albedo, normals, specular, shininess, view_dirs, dirs, radiance, cam, vis):
    N = albedo.shape[0]
    albedo, normals = _chk("albedo", albedo, shape=(N, 3)), _chk("normals", normals, shape=(N, 3))
    per_row = dirs.dim() == 3
    M = dirs.shape[-2]
    dirs = _chk("light_directions", dirs, shape=(N, M, 3) if per_row else (M, 3))
    radiance = _chk("radiance", radiance, shape=(None, M, 3))
    if cam is not None:
        cam = _chk("cam", cam, dtype=torch.int32, shape=(N,))
    elif radiance.shape[0] != 1:
        raise ValueError("radiance has several light tables but no row -> table index (cam) was given")
    if mode >= 1:
        shininess, view_dirs = _chk("shininess", shininess, shape=(N,)), _chk("view_directions", view_dirs, shape=(N, 3))
    if mode == 1:
        specular = _chk("specular", specular, shape=(N, 3))
    rows_per_vis = 1
    if vis is not None:
        if vis.dim() != 2 or vis.shape[1] != M or vis.shape[0] == 0 or N % vis.shape[0] != 0:
            raise ValueError(f"visibility: expected [N / S, {M}] with N = {N} a multiple of its rows, got {tuple(vis.shape)}")
        vis = _chk("visibility", vis)
        rows_per_vis = N // vis.shape[0]
    return N, M, per_row, albedo, normals, specular, shininess, view_dirs, dirs, radiance, cam, vis, rows_per_vis


def shade_lights(mode: int, albedo: Tensor, normals: Tensor, dirs: Tensor, radiance: Tensor, cam: Optional[Tensor] = None, specular: Optional[Tensor] = None,
                 shininess: Optional[Tensor] = None, view_dirs: Optional[Tensor] = None, vis: Optional[Tensor] = None, normalize_dirs: bool = False,
                 weights: Optional[Tensor] = None, rgb_lin: Optional[Tensor] = None):
    """Per-row light sums (see include/neusky_b200.h, nsk_shade_lights_fwd).  Returns (out_a, out_b | None)."""
    N, M, per_row, albedo, normals, specular, shininess, view_dirs, dirs, radiance, cam, vis, rpv = _shade_args(
        mode, albedo, normals, specular, shininess, view_dirs, dirs, radiance, cam, vis)
    out_a = torch.empty((N, 3), device=albedo.device, dtype=torch.float32)
    out_b = torch.empty((N, 3), device=albedo.device, dtype=torch.float32) if mode == 0 else None
    S = 0
    if weights is not None and rgb_lin is not None:
        weights = _chk("weights", weights.reshape(-1), shape=(N,))
        rgb_lin = _chk("rgb_lin", rgb_lin, shape=(None, 3))
        if rgb_lin.shape[0] == 0 or N % rgb_lin.shape[0] != 0:
            raise ValueError("rgb_lin rows must divide the number of samples")
        S = N // rgb_lin.shape[0]
    _lib.check(_lib.load().nsk_shade_lights_fwd(c_int(mode), _ptr(albedo), _ptr(normals), _ptr(specular), _ptr(shininess), _ptr(view_dirs), _ptr(dirs), c_int(int(per_row)),
                                                c_int(int(normalize_dirs)), _ptr(radiance), _ptr(cam), _ptr(vis), c_int(rpv), c_int64(N), c_int(M), _ptr(out_a), _ptr(out_b),
                                                _ptr(weights if S else None), _ptr(rgb_lin if S else None), c_int(S), _stream(albedo)), "nsk_shade_lights_fwd")
    return out_a, out_b


def shade_lights_bwd(mode: int, albedo: Tensor, normals: Tensor, dirs: Tensor, radiance: Tensor, g_a: Optional[Tensor], g_b: Optional[Tensor] = None,
                     cam: Optional[Tensor] = None, specular: Optional[Tensor] = None, shininess: Optional[Tensor] = None, view_dirs: Optional[Tensor] = None,
                     vis: Optional[Tensor] = None, normalize_dirs: bool = False, want_radiance: bool = True, want_vis: bool = True):
    """Backward of shade_lights: (d_albedo, d_normals, d_specular | None, d_shininess | None, d_radiance | None, d_vis | None)."""
    N, M, per_row, albedo, normals, specular, shininess, view_dirs, dirs, radiance, cam, vis, rpv = _shade_args(
        mode, albedo, normals, specular, shininess, view_dirs, dirs, radiance, cam, vis)
    dev = albedo.device
    g_a = None if g_a is None else _chk("g_a", g_a, shape=(N, 3))
    g_b = None if g_b is None else _chk("g_b", g_b, shape=(N, 3))
    d_alb, d_nrm = torch.empty((N, 3), device=dev), torch.empty((N, 3), device=dev)
    d_spec = torch.empty((N, 3), device=dev) if mode == 1 else None
    d_shin = torch.empty((N,), device=dev) if mode >= 1 else None
    d_rad = torch.zeros_like(radiance) if want_radiance else None
    d_vis = torch.zeros_like(vis) if (want_vis and vis is not None and mode == 2) else None
    _lib.check(_lib.load().nsk_shade_lights_bwd(c_int(mode), _ptr(albedo), _ptr(normals), _ptr(specular), _ptr(shininess), _ptr(view_dirs), _ptr(dirs), c_int(int(per_row)),
                                                c_int(int(normalize_dirs)), _ptr(radiance), _ptr(cam), _ptr(vis), c_int(rpv), c_int64(N), c_int(M), _ptr(g_a), _ptr(g_b),
                                                _ptr(d_alb), _ptr(d_nrm), _ptr(d_spec), _ptr(d_shin), _ptr(d_rad), _ptr(d_vis), _stream(albedo)), "nsk_shade_lights_bwd")
    return d_alb, d_nrm, d_spec, d_shin, d_rad, d_vis


# every public op runs with the device of its tensor arguments current (see _device_guard)
for _name, _fn in list(globals().items()):
    if isinstance(_fn, types.FunctionType) and _fn.__module__ == __name__ and not _name.startswith("_"):
        globals()[_name] = _device_guard(_fn)
del _name, _fn
