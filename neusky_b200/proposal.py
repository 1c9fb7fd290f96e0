"""Proposal-network sampler on the GPU: the sample placement directly in front of the
render-and-shade path.  Host mirror of nerfstudio's ``UniformSampler`` / ``PDFSampler`` / ``ProposalNetworkSampler`` /
``HashMLPDensityField`` and ``interlevel_loss`` as NeuSky drives them (neusky/models/neusky_model.py:561, 575-576,
987-988; defaults in SURVEY A.6), with the C-ABI kernels of ``csrc/proposal_sampler.cu`` behind them.

Placement is kept in the SPACING domain (bins in [0,1], ``[R, S+1]`` edges per ray); euclidean edges are
``bin*far + (1-bin)*near`` (UniformSampler: identity spacing function).  No torch arithmetic on the hot path: torch
allocates, the kernels compute.
"""
from __future__ import annotations

import ctypes
from types import SimpleNamespace
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .init import hash_scalings
from .ops import _chk, _chk_out, _ptr, _stream

Tensor = torch.Tensor
c_int, c_int64, c_float = ctypes.c_int, ctypes.c_int64, ctypes.c_float

HIDDEN = 16

_LINSPACE_CACHE: Dict[tuple, Tensor] = {}


def _linspace(end: float, steps: int, device) -> Tensor:
    """torch.linspace(0, end, steps) with the CPU kernel's rounding (what the oracle and the reference's CPU path produce), cached per device."""
    key = (float(end), int(steps), str(device))
    t = _LINSPACE_CACHE.get(key)
    if t is None:
        t = _LINSPACE_CACHE[key] = torch.linspace(0.0, end, steps, dtype=torch.float32).to(device)
    return t


# ------------------------------------------------------------------------------------------------ raw ops
def pack_proposal_mlp(p: Dict[str, Tensor]) -> Tensor:
    """``mlp.0.weight`` [16,2L], ``mlp.0.bias`` [16], ``mlp.1.weight`` [1,16], ``mlp.1.bias`` [1] -> blob
    W0 [2L][16] (input-major) | b0 | W1 | b1."""
    return torch.cat([p["mlp.0.weight"].t().reshape(-1), p["mlp.0.bias"].reshape(-1), p["mlp.1.weight"].reshape(-1), p["mlp.1.bias"].reshape(-1)]).float().contiguous()


def unpack_proposal_mlp_grad(d_mlp: Tensor, L: int) -> Dict[str, Tensor]:
    n0 = 2 * L * HIDDEN
    return {"mlp.0.weight": d_mlp[:n0].reshape(2 * L, HIDDEN).t().contiguous(), "mlp.0.bias": d_mlp[n0:n0 + HIDDEN].clone(),
            "mlp.1.weight": d_mlp[n0 + HIDDEN:n0 + 2 * HIDDEN].reshape(1, HIDDEN).clone(), "mlp.1.bias": d_mlp[n0 + 2 * HIDDEN:].clone()}


def uniform_bins(R: int, S: int, device, jitter: Optional[Tensor] = None) -> Tensor:
    """SpacedSampler spacing bins [R,S+1]; jitter [R] in [0,1) = the per-ray t_rand of single_jitter training."""
    base = _linspace(1.0, S + 1, device)
    out = torch.empty((R, S + 1), device=device, dtype=torch.float32)
    if jitter is not None:
        jitter = _chk("jitter", jitter.reshape(-1), shape=(R,))
    _lib.check(_lib.load().nsk_uniform_bins(_ptr(base), _ptr(jitter), c_int64(R), c_int(S), _ptr(out), _stream(out)), "nsk_uniform_bins")
    return out


def proposal_density(origins: Tensor, dirs: Optional[Tensor], near: Optional[Tensor], far: Optional[Tensor], bins: Optional[Tensor],
                     table: Tensor, scalings: Tensor, log2_T: int, mlp: Tensor) -> Tensor:
    """Ray mode: origins/dirs [R,3], near/far [R], bins [R,S+1] -> density [R,S] at the bin mid-points.
    Positions mode (dirs None): origins [n,3] world positions -> density [n]."""
    L = scalings.numel()
    table = _chk("table", table, shape=(L << log2_T, 2))
    scalings = _chk("scalings", scalings)
    mlp = _chk("mlp", mlp, shape=(2 * L * HIDDEN + 2 * HIDDEN + 1,))
    if dirs is None:
        o = _chk("positions", origins.reshape(-1, 3), shape=(None, 3))
        R, S = o.shape[0], 1
        out = torch.empty((R,), device=o.device, dtype=torch.float32)
        near = far = bins = None
    else:
        R = origins.shape[0]
        o, dirs = _chk("origins", origins, shape=(R, 3)), _chk("dirs", dirs, shape=(R, 3))
        near, far = _chk("near", near.reshape(-1), shape=(R,)), _chk("far", far.reshape(-1), shape=(R,))
        bins = _chk("bins", bins, shape=(R, None))
        S = bins.shape[1] - 1
        out = torch.empty((R, S), device=o.device, dtype=torch.float32)
    _lib.check(_lib.load().nsk_proposal_density_fwd(_ptr(o), _ptr(dirs), _ptr(near), _ptr(far), _ptr(bins), c_int64(R), c_int(S), _ptr(table), _ptr(scalings),
                                                    c_int(L), c_int(log2_T), _ptr(mlp), c_int(HIDDEN), _ptr(out), _stream(o)), "nsk_proposal_density_fwd")
    return out


def proposal_density_bwd(origins: Tensor, dirs: Tensor, near: Tensor, far: Tensor, bins: Tensor, table: Tensor, scalings: Tensor, log2_T: int,
                         mlp: Tensor, g_density: Tensor, d_table: Optional[Tensor] = None, d_mlp: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    L = scalings.numel()
    R, S = g_density.shape
    d_table = torch.zeros_like(table) if d_table is None else _chk_out("d_table", d_table, shape=(L << log2_T, 2))
    d_mlp = torch.zeros_like(mlp) if d_mlp is None else _chk_out("d_mlp", d_mlp, shape=(2 * L * HIDDEN + 2 * HIDDEN + 1,))
    _lib.check(_lib.load().nsk_proposal_density_bwd(_ptr(_chk("origins", origins, shape=(R, 3))), _ptr(_chk("dirs", dirs, shape=(R, 3))), _ptr(_chk("near", near.reshape(-1), shape=(R,))),
                                                    _ptr(_chk("far", far.reshape(-1), shape=(R,))), _ptr(_chk("bins", bins, shape=(R, S + 1))), c_int64(R), c_int(S),
                                                    _ptr(_chk("table", table, shape=(L << log2_T, 2))), _ptr(_chk("scalings", scalings)), c_int(L), c_int(log2_T), _ptr(_chk("mlp", mlp)),
                                                    c_int(HIDDEN), _ptr(_chk("g_density", g_density)), _ptr(d_table), _ptr(d_mlp), _stream(g_density)), "nsk_proposal_density_bwd")
    return d_table, d_mlp


def pdf_resample(bins: Tensor, near: Tensor, far: Tensor, N: int, density: Optional[Tensor] = None, weights: Optional[Tensor] = None, anneal: float = 1.0,
                 jitter: Optional[Tensor] = None, histogram_padding: float = 0.01, eps: float = 1e-5, want_weights: bool = True, want_euclid: bool = True):
    """PDFSampler over existing spacing bins [R,S+1]: give ``density`` [R,S] (RaySamples.get_weights fused) or ``weights`` [R,S].
    -> (new_bins [R,N+1], new_euclid [R,N+1] | None, weights [R,S] | None)."""
    R, S = bins.shape[0], bins.shape[1] - 1
    bins = _chk("bins", bins, shape=(R, S + 1))
    near, far = _chk("near", near.reshape(-1), shape=(R,)), _chk("far", far.reshape(-1), shape=(R,))
    if (density is None) == (weights is None):
        raise ValueError("pdf_resample: give exactly one of density / weights")
    if density is not None:
        density = _chk("density", density, shape=(R, S))
    else:
        weights = _chk("weights", weights, shape=(R, S))
    nb = N + 1
    u_base = _linspace(1.0 - (1.0 / nb), nb, bins.device)
    if jitter is not None:
        jitter = _chk("jitter", jitter.reshape(-1), shape=(R,))
    w_out = torch.empty((R, S), device=bins.device, dtype=torch.float32) if (want_weights and density is not None) else None
    new_bins = torch.empty((R, nb), device=bins.device, dtype=torch.float32)
    new_e = torch.empty((R, nb), device=bins.device, dtype=torch.float32) if want_euclid else None
    _lib.check(_lib.load().nsk_pdf_resample(_ptr(bins), _ptr(density), _ptr(weights), _ptr(near), _ptr(far), c_int64(R), c_int(S), c_int(N), c_float(anneal),
                                            c_float(histogram_padding), c_float(eps), _ptr(u_base), c_float(1.0 / (2 * nb)), _ptr(jitter), _ptr(w_out), _ptr(new_bins),
                                            _ptr(new_e), _stream(bins)), "nsk_pdf_resample")
    return new_bins, new_e, (w_out if density is not None else weights)


def density_weights_bwd(bins: Tensor, density: Tensor, near: Tensor, far: Tensor, g_weights: Tensor) -> Tensor:
    R, S = density.shape
    out = torch.empty_like(density)
    _lib.check(_lib.load().nsk_density_weights_bwd(_ptr(_chk("bins", bins, shape=(R, S + 1))), _ptr(_chk("density", density)), _ptr(_chk("near", near.reshape(-1), shape=(R,))),
                                                   _ptr(_chk("far", far.reshape(-1), shape=(R,))), c_int64(R), c_int(S), _ptr(_chk("g_weights", g_weights, shape=(R, S))), _ptr(out),
                                                   _stream(density)), "nsk_density_weights_bwd")
    return out


def interlevel_loss_level(c: Tensor, w: Tensor, cp: Tensor, wp: Tensor, want_grad: bool = False) -> Tuple[Tensor, Optional[Tensor]]:
    """One term of nerfstudio's interlevel_loss: mean over [R,Sf] of lossfun_outer(c, w, cp, wp).  Returns (loss scalar tensor,
    d loss / d wp [R,Sp] | None)."""
    R, Sf = w.shape
    Sp = wp.shape[1]
    c, w = _chk("c", c, shape=(R, Sf + 1)), _chk("w", w, shape=(R, Sf))
    cp, wp = _chk("cp", cp, shape=(R, Sp + 1)), _chk("wp", wp, shape=(R, Sp))
    loss_ray = torch.empty((R,), device=w.device, dtype=torch.float32)
    g = torch.empty((R, Sp), device=w.device, dtype=torch.float32) if want_grad else None
    _lib.check(_lib.load().nsk_interlevel_loss(_ptr(c), _ptr(w), c_int(Sf), _ptr(cp), _ptr(wp), c_int(Sp), c_int64(R), _ptr(loss_ray), _ptr(g), _stream(w)), "nsk_interlevel_loss")
    scale = 1.0 / float(R * Sf)
    return loss_ray.sum() * scale, (g * scale if want_grad else None)


# ------------------------------------------------------------------------------------------------ reference-named surface
class _Table(torch.nn.Module):
    def __init__(self, table: Tensor):
        super().__init__()
        self.hash_table = torch.nn.Parameter(table)


class HashMLPDensityField(torch.nn.Module):
    """nerfstudio ``HashMLPDensityField`` as NeuSFactoModel builds its proposal networks [SURVEY A.6]: hash grid (5 levels, T = 2^17,
    base_res 16, max_res 64 / 256) + Linear(10,16) + ReLU + Linear(16,1) + trunc_exp, on the L-inf-contracted position mapped to
    [0,1]^3, density zeroed outside (0,1)^3.  ``params``: ``encoding.hash_table`` / ``mlp.{0,1}.{weight,bias}`` (the field's
    ``mlp_base`` split into its two halves); they are registered under exactly those names (``state_dict`` round-trips)."""

    def __init__(self, params: Dict[str, Tensor], max_res: int, num_levels: int = 5, base_res: int = 16, log2_hashmap_size: int = 17, device="cuda"):
        super().__init__()
        self.device = torch.device(device)
        self.num_levels, self.log2_T, self.max_res = num_levels, log2_hashmap_size, max_res
        self.register_buffer("scalings", hash_scalings(num_levels, base_res, max_res).to(self.device), persistent=False)
        f = lambda k: params[k].detach().to(self.device, torch.float32).contiguous()
        self.encoding = _Table(f("encoding.hash_table"))
        lins = []
        for i in range(2):
            W, b = f(f"mlp.{i}.weight"), f(f"mlp.{i}.bias")
            lin = torch.nn.Linear(W.shape[1], W.shape[0], device=self.device)
            with torch.no_grad():
                lin.weight.copy_(W)
                lin.bias.copy_(b)
            lins.append(lin)
        self.mlp = torch.nn.ModuleList(lins)
        self._mlp_key = None
        self.refresh()

    @property
    def params(self) -> Dict[str, torch.nn.Parameter]:
        return dict(self.named_parameters())

    def _apply(self, fn, *args, **kwargs):
        """Module.to / .cuda: the packed blob and the cached device follow the parameters."""
        r = super()._apply(fn, *args, **kwargs)
        self.device = self.encoding.hash_table.device
        self._mlp_key = None
        self.refresh()
        return r

    def refresh(self) -> None:
        """Re-pack the MLP blob when a parameter changed (optimizer step, load_state_dict)."""
        ps = self.params
        key = tuple((v.data_ptr(), v._version) for v in ps.values())
        if key != self._mlp_key:
            self.mlp_blob = pack_proposal_mlp({k: v.detach() for k, v in ps.items()})
            self._mlp_key = key
        self.table = ps["encoding.hash_table"].detach()

    def backward_on_rays(self, origins, dirs, near, far, bins, g_density: Tensor) -> None:
        """Manual backward of ``density_on_rays``: accumulate d loss / d params into ``.grad`` (stand-alone use; the training step goes
        through autograd, ``ProposalNetworkSampler.interlevel_loss``)."""
        self.refresh()
        _accumulate_field_grads(self, origins, dirs, near, far, bins, g_density)

    def density_fn(self, positions: Tensor) -> Tensor:
        """positions [...,3] -> density [...,1] (what the reference passes as ``density_fns[i]``)."""
        lead = positions.shape[:-1]
        return proposal_density(positions.reshape(-1, 3).contiguous(), None, None, None, None, self.table, self.scalings, self.log2_T, self.mlp_blob).reshape(*lead, 1)

    def density_on_rays(self, origins: Tensor, dirs: Tensor, near: Tensor, far: Tensor, bins: Tensor) -> Tensor:
        return proposal_density(origins, dirs, near, far, bins, self.table, self.scalings, self.log2_T, self.mlp_blob)


def _accumulate_field_grads(field: "HashMLPDensityField", origins, dirs, near, far, bins, g_density: Tensor) -> None:
    d_table, d_mlp = proposal_density_bwd(origins, dirs, near, far, bins, field.table, field.scalings, field.log2_T, field.mlp_blob, g_density)
    grads = unpack_proposal_mlp_grad(d_mlp, field.num_levels)
    grads["encoding.hash_table"] = d_table
    for k, p in field.params.items():
        g = grads[k].reshape(p.shape)
        if p.grad is None:
            p.grad = g
        else:
            p.grad.add_(g)          # in place: .grad may be a view into a GradBucketReducer communication bucket


def _ray_samples(origins, dirs, euclid, spacing, near, far):
    """Duck-typed nerfstudio RaySamples [SURVEY A.1] over [R,S+1] edges (views, no copies)."""
    if euclid is None:      # level 0 of the proposal loop: only its spacing bins are consumed downstream (interlevel loss)
        fr = SimpleNamespace(origins=origins[:, None, :], directions=dirs[:, None, :], starts=None, ends=None)
        deltas = None
    else:
        fr = SimpleNamespace(origins=origins[:, None, :], directions=dirs[:, None, :], starts=euclid[:, :-1, None], ends=euclid[:, 1:, None])
        deltas = (euclid[:, 1:] - euclid[:, :-1])[..., None]
    return SimpleNamespace(frustums=fr, deltas=deltas, spacing_starts=spacing[:, :-1, None], spacing_ends=spacing[:, 1:, None],
                           spacing_bins=spacing, euclidean_bins=euclid, nears=near, fars=far, camera_indices=None)


class ProposalNetworkSampler:
    """nerfstudio ``ProposalNetworkSampler`` with NeuS-facto's configuration [SURVEY A.6]: ``UniformSampler(single_jitter)`` ->
    density_0 -> weights -> ``PDFSampler`` -> density_1 -> weights -> ``PDFSampler``; ``weights ** anneal`` before each PDF.
    ``generate_ray_samples`` returns ``(ray_samples, weights_list, ray_samples_list)`` like the reference's call at
    neusky_model.py:561; ``weights_list`` entries are ``[R,S,1]``."""

    def __init__(self, num_nerf_samples_per_ray: int = 48, num_proposal_samples_per_ray: Sequence[int] = (256, 96), num_proposal_network_iterations: int = 2,
                 single_jitter: bool = True, histogram_padding: float = 0.01):
        if len(num_proposal_samples_per_ray) != num_proposal_network_iterations:
            raise ValueError("num_proposal_samples_per_ray must have one entry per proposal iteration")
        if not single_jitter:
            raise NotImplementedError("only single_jitter=True (the NeuS-facto default, use_single_jitter) is implemented")
        self.num_nerf_samples_per_ray = num_nerf_samples_per_ray
        self.num_proposal_samples_per_ray = tuple(num_proposal_samples_per_ray)
        self.num_proposal_network_iterations = num_proposal_network_iterations
        self.histogram_padding = histogram_padding
        self.training = False
        self._anneal = 1.0
        self._state: List[dict] = []

    def set_anneal(self, anneal: float) -> None:
        self._anneal = float(anneal)

    def generate_ray_samples(self, origins: Tensor, directions: Tensor, nears: Tensor, fars: Tensor, density_fields: Sequence[HashMLPDensityField],
                             jitters: Optional[Sequence[Tensor]] = None):
        R = origins.shape[0]
        n = self.num_proposal_network_iterations
        if len(density_fields) != n:
            raise ValueError(f"expected {n} density fields, got {len(density_fields)}")
        if self.training and jitters is None:
            jitters = [torch.rand(R, device=origins.device) for _ in range(n + 1)]
        near, far = nears.reshape(-1).contiguous(), fars.reshape(-1).contiguous()
        weights_list, samples_list, self._state = [], [], []
        bins = uniform_bins(R, self.num_proposal_samples_per_ray[0], origins.device, None if jitters is None else jitters[0])
        euclid = None
        for lvl in range(n):
            dens = density_fields[lvl].density_on_rays(origins, directions, near, far, bins)
            N = self.num_proposal_samples_per_ray[lvl + 1] if lvl + 1 < n else self.num_nerf_samples_per_ray
            new_bins, new_e, w = pdf_resample(bins, near, far, N, density=dens, anneal=self._anneal, jitter=None if jitters is None else jitters[lvl + 1],
                                              histogram_padding=self.histogram_padding)
            samples_list.append(_ray_samples(origins, directions, euclid, bins, near, far))
            weights_list.append(w[..., None])
            self._state.append({"bins": bins, "density": dens, "weights": w, "field": density_fields[lvl]})
            bins, euclid = new_bins, new_e
        return _ray_samples(origins, directions, euclid, bins, near, far), weights_list, samples_list

    __call__ = generate_ray_samples

    def interlevel_loss(self, fine_weights: Tensor, fine_samples, origins, directions, nears, fars) -> Tensor:
        """nerfstudio ``interlevel_loss(weights_list, ray_samples_list)`` (neusky_model.py:987-988) with the fine NeuS weights appended as
        the reference does (:575-576).  The (unscaled) loss is returned through autograd: its backward runs weights -> density ->
        MLP + hash table (nsk_density_weights_bwd, nsk_proposal_density_bwd) scaled by the incoming cotangent, so loss scaling,
        gradient accumulation and ``torch.no_grad()`` behave as for any torch loss; the fine histogram is detached like the reference's."""
        near, far = nears.reshape(-1).contiguous(), fars.reshape(-1).contiguous()
        c = fine_samples.spacing_bins
        w = fine_weights.reshape(c.shape[0], -1).detach()
        params = [p for st in self._state for p in st["field"].params.values()]
        return _InterlevelLoss.apply(self._state, c, w, origins, directions, near, far, *params)


class _InterlevelLoss(torch.autograd.Function):
    """Sum over proposal levels of lossfun_outer(fine, proposal); inputs of the graph = the proposal fields' parameters."""

    @staticmethod
    def forward(ctx, state, c, w, origins, directions, near, far, *params):
        total = torch.zeros((), device=c.device)
        need = any(ctx.needs_input_grad[7:])
        g_list = []
        for st in state:
            loss, g_wp = interlevel_loss_level(c, w, st["bins"], st["weights"], want_grad=need)       # weights_list holds the un-annealed weights
            total = total + loss
            g_list.append(g_wp)
        ctx.levels = [(st["bins"], st["density"], st["field"], g) for st, g in zip(state, g_list)]
        ctx.rays = (origins, directions, near, far)
        return total

    @staticmethod
    def backward(ctx, g):
        origins, directions, near, far = ctx.rays
        grads = []
        for bins, dens, field, g_wp in ctx.levels:
            g_d = density_weights_bwd(bins, dens, near, far, (g_wp * g).contiguous())
            d_table, d_mlp = proposal_density_bwd(origins, directions, near, far, bins, field.table, field.scalings, field.log2_T, field.mlp_blob, g_d)
            gm = unpack_proposal_mlp_grad(d_mlp, field.num_levels)
            gm["encoding.hash_table"] = d_table
            grads += [gm[k].reshape(p.shape) for k, p in field.params.items()]
        return (None,) * 7 + tuple(grads)
