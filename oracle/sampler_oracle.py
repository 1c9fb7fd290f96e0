"""CPU oracle for the proposal-network sampler that sits directly in front of the render-and-shade path
(SURVEY.md 8f row f1): nerfstudio's ProposalNetworkSampler as NeuS-facto configures it and NeuSky calls it at
neusky/models/neusky_model.py:561 (``self.proposal_sampler(ray_bundle, density_fns=self.density_fns)``).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as neusky_oracle.py).

**Parity unpinned**: every function here restates nerfstudio behaviour from memory ([NS-mem], SURVEY Appendix A.6 --
nerfstudio is an un-vendored, un-pinned dependency and the reference has no tests or golden vectors for it).  The
reference-side facts that ARE in the tree: the sampler is called once per forward with the model's ``density_fns``
(neusky_model.py:561), its outputs feed ``self.field(ray_samples, ...)`` (:563) and the interlevel loss consumes
``weights_list`` / ``ray_samples_list`` (:575-576, 987-988); the proposal-net hyper-parameters are nerfstudio's
NeuSFactoModelConfig defaults (A.6): two HashMLPDensityFields (hidden 16, log2 T = 17, 5 levels, max_res 64 / 256),
256 -> 96 proposal samples, 48 NeuS samples, UniformSampler as the initial sampler, single jitter.

Accumulation order (what "bit-exact sample placement" is defined against): torch's CPU ``cumsum`` accumulates fp32
inputs in fp64 and rounds every prefix to fp32; this file does exactly that with an explicit loop-free numpy statement
(``np.cumsum(x.astype(float64)).astype(float32)``), and uses the same fp64-sequential rule for the two ``sum``s
(torch's vectorised CPU ``sum`` order is ISA-dependent, so it cannot serve as a definition).  Every other op is a single
correctly-rounded fp32 operation, written out one op at a time so that no FMA contraction is implied.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import neusky_oracle as O

Tensor = torch.Tensor
F32 = np.float32


# --------------------------------------------------------------------------------------
# HashMLPDensityField [NS-mem]: nerfstudio/fields/density_fields.py as built by NeuSFactoModel.populate_modules with
# spatial_distortion = the model's L-inf SceneContraction
# --------------------------------------------------------------------------------------

def proposal_scalings(max_res: int, num_levels: int = 5, base_res: int = 16) -> Tensor:
    return O.hash_scalings(num_levels, base_res, max_res)


def proposal_density(positions: Tensor, p: Dict[str, Tensor], scalings: Tensor, log2_T: int = 17) -> Tensor:
    """positions [...,3] (world) -> density [...] .
    get_density: x = contract_Linf(pos); x = (x+2)/4; selector = all(0 < x < 1); x *= selector;
    h = relu(W0 hash(x) + b0); raw = W1 h + b1; density = trunc_exp(raw) * selector  (average_init_density = 1).
    ``p``: ``encoding.hash_table`` [L*T,2], ``mlp.0.weight`` [16,2L], ``mlp.0.bias``, ``mlp.1.weight`` [1,16], ``mlp.1.bias``."""
    shp = positions.shape[:-1]
    x = O.scene_contraction_linf(positions.reshape(-1, 3).to(torch.float32))
    x = (x + 2.0) / 4.0
    sel = ((x > 0.0) & (x < 1.0)).all(dim=-1)
    x = x * sel[:, None]
    feat = O.hash_encode(x, p["encoding.hash_table"], scalings, log2_T)
    h = torch.relu(feat @ p["mlp.0.weight"].T + p["mlp.0.bias"])
    raw = (h @ p["mlp.1.weight"].T + p["mlp.1.bias"])[:, 0]
    return (torch.exp(raw) * sel).reshape(shp)


# --------------------------------------------------------------------------------------
# Samplers [NS-mem A.6]: nerfstudio/model_components/ray_samplers.py
# --------------------------------------------------------------------------------------

def _seq_cumsum(x: np.ndarray) -> np.ndarray:
    """torch CPU cumsum semantics: fp64 running sum, every prefix rounded to fp32."""
    return np.cumsum(x.astype(np.float64), axis=-1).astype(F32)


def _seq_sum(x: np.ndarray) -> np.ndarray:
    return _seq_cumsum(x)[..., -1:]


def spacing_to_euclidean(bins: np.ndarray, near: np.ndarray, far: np.ndarray) -> np.ndarray:
    """UniformSampler: spacing_fn = identity, so x -> x*far + (1-x)*near (three rounded fp32 ops + one add)."""
    bins = bins.astype(F32)
    return (bins * far.astype(F32) + (F32(1.0) - bins) * near.astype(F32)).astype(F32)


def uniform_bins(R: int, S: int, jitter: Optional[np.ndarray] = None) -> np.ndarray:
    """SpacedSampler.generate_ray_samples: bins = linspace(0,1,S+1); training with single_jitter: one t_rand per ray,
    bins = lower + (upper-lower)*t_rand with lower/upper = the bin-centre brackets.  -> spacing bins [R,S+1]."""
    bins = torch.linspace(0.0, 1.0, S + 1, dtype=torch.float32).numpy()[None, :]
    if jitter is not None:
        t = jitter.astype(F32).reshape(R, 1)
        centers = ((bins[:, 1:] + bins[:, :-1]) / F32(2.0)).astype(F32)
        upper = np.concatenate([centers, bins[:, -1:]], -1)
        lower = np.concatenate([bins[:, :1], centers], -1)
        bins = (lower + (upper - lower) * t).astype(F32)
    return np.broadcast_to(bins, (R, S + 1)).astype(F32).copy()


def density_weights(density: np.ndarray, deltas: np.ndarray) -> np.ndarray:
    """RaySamples.get_weights [NS-mem]: dd = delta*density; alpha = 1-exp(-dd); T = exp(-cumsum([0, dd[:-1]]));
    w = nan_to_num(alpha*T).  ``exp`` is float32 ``expf`` (numpy's and CUDA's differ by <= 1 ulp, so weights are compared
    with a tolerance; everything from the weights on is bit-exact)."""
    dd = (deltas.astype(F32) * density.astype(F32)).astype(F32)
    alpha = (F32(1.0) - np.exp(-dd, dtype=F32)).astype(F32)
    cs = _seq_cumsum(dd[..., :-1])
    cs = np.concatenate([np.zeros_like(dd[..., :1]), cs], -1)
    T = np.exp(-cs, dtype=F32)
    return np.nan_to_num((alpha * T).astype(F32))


def pdf_resample(spacing_bins: np.ndarray, weights: np.ndarray, N: int, jitter: Optional[np.ndarray] = None,
                 histogram_padding: float = 0.01, eps: float = 1e-5) -> np.ndarray:
    """PDFSampler.generate_ray_samples(include_original=False): existing spacing bins [R,S+1], (annealed) weights [R,S]
    -> new spacing bins [R,N+1]."""
    R, S = weights.shape
    nb = N + 1
    w = (weights.astype(F32) + F32(histogram_padding)).astype(F32)
    wsum = _seq_sum(w)
    pad = np.maximum(F32(eps) - wsum, F32(0.0)).astype(F32)
    w = (w + (pad / F32(S)).astype(F32)).astype(F32)
    wsum = (wsum + pad).astype(F32)
    pdf = (w / wsum).astype(F32)
    cdf = np.minimum(F32(1.0), _seq_cumsum(pdf))
    cdf = np.concatenate([np.zeros((R, 1), F32), cdf], -1)                       # [R,S+1]
    u = torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb, dtype=torch.float32).numpy()[None, :]
    if jitter is not None:
        u = (u + (jitter.astype(F32).reshape(R, 1) / F32(nb)).astype(F32)).astype(F32)
    else:
        u = (u + F32(1.0 / (2 * nb))).astype(F32)
    u = np.broadcast_to(u, (R, nb))
    inds = np.stack([np.searchsorted(cdf[r], u[r], side="right") for r in range(R)])
    below = np.clip(inds - 1, 0, S)
    above = np.clip(inds, 0, S)
    c0, c1 = np.take_along_axis(cdf, below, -1), np.take_along_axis(cdf, above, -1)
    b0, b1 = np.take_along_axis(spacing_bins.astype(F32), below, -1), np.take_along_axis(spacing_bins.astype(F32), above, -1)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = ((u - c0).astype(F32) / (c1 - c0).astype(F32)).astype(F32)
    t = np.clip(np.nan_to_num(t, nan=0.0), F32(0.0), F32(1.0))      # nan_to_num: nan->0, +inf->fp32 max (clipped to 1)
    return (b0 + (t * (b1 - b0).astype(F32)).astype(F32)).astype(F32)


def proposal_sample(origins: Tensor, directions: Tensor, near: Tensor, far: Tensor, nets: List[Dict[str, Tensor]],
                    num_proposal: Tuple[int, ...] = (256, 96), num_final: int = 48, max_res: Tuple[int, ...] = (64, 256),
                    log2_T: int = 17, anneal: float = 1.0, jitters: Optional[List[np.ndarray]] = None):
    """ProposalNetworkSampler.generate_ray_samples: uniform(256) -> density_0 -> weights -> pdf(96) -> density_1 ->
    weights -> pdf(48).  Returns (final euclidean bins [R,num_final+1], weights_list, spacing_bins_list, euclid_list)."""
    R = origins.shape[0]
    nr, fr = near.reshape(R, 1).numpy().astype(F32), far.reshape(R, 1).numpy().astype(F32)
    weights_list, spacing_list, euclid_list = [], [], []
    bins = None
    w = None
    n_iter = len(nets)
    for lvl in range(n_iter + 1):
        n = num_proposal[lvl] if lvl < n_iter else num_final
        jit = None if jitters is None else jitters[lvl]
        if lvl == 0:
            bins = uniform_bins(R, n, jit)
        else:
            aw = w if anneal == 1.0 else np.power(w, F32(anneal), dtype=F32)
            bins = pdf_resample(bins, aw, n, jit)
        e = spacing_to_euclidean(bins, nr, fr)
        spacing_list.append(bins)
        euclid_list.append(e)
        if lvl < n_iter:
            et = torch.from_numpy(e)
            mids = (et[:, :-1] + et[:, 1:]) / 2                                            # Frustums.get_positions
            pos = origins[:, None, :] + directions[:, None, :] * mids[..., None]
            dens = proposal_density(pos, nets[lvl], proposal_scalings(max_res[lvl]), log2_T).numpy()
            w = density_weights(dens, e[:, 1:] - e[:, :-1])
            weights_list.append(w)
    return euclid_list[-1], weights_list, spacing_list, euclid_list


def init_proposal_net(seed: int, num_levels: int = 5, log2_T: int = 17, hidden: int = 16, table_scale: float = 1e-3,
                      density_bias: float = 0.0) -> Dict[str, Tensor]:
    """Random-init HashMLPDensityField state: hash table U(-1,1)*table_scale (nerfstudio uses 1e-3), torch Linear default
    init for the MLP.  Tests pass a larger table_scale so that the densities (and hence the sample placement) are not
    trivially uniform."""
    g = torch.Generator().manual_seed(seed)
    T = 1 << log2_T
    p = {"encoding.hash_table": (torch.rand(num_levels * T, 2, generator=g) * 2 - 1) * table_scale}
    for i, (fin, fout) in enumerate(((2 * num_levels, hidden), (hidden, 1))):
        b = 1.0 / np.sqrt(fin)
        p[f"mlp.{i}.weight"] = (torch.rand(fout, fin, generator=g) * 2 - 1) * b
        p[f"mlp.{i}.bias"] = (torch.rand(fout, generator=g) * 2 - 1) * b
    p["mlp.1.bias"] = p["mlp.1.bias"] + density_bias
    return p


# --------------------------------------------------------------------------------------
# Interlevel (proposal) loss [NS-mem]: nerfstudio/model_components/losses.py interlevel_loss / lossfun_outer / outer,
# called at neusky/models/neusky_model.py:987-988 with the NeuS weights and samples appended (:575-576).  torch ops, so that
# fp64 autograd gives the reference gradient for the proposal weights.
# --------------------------------------------------------------------------------------

def outer(t0_starts: Tensor, t0_ends: Tensor, t1_starts: Tensor, t1_ends: Tensor, y1: Tensor) -> Tensor:
    cy1 = torch.cat([torch.zeros_like(y1[..., :1]), torch.cumsum(y1, dim=-1)], dim=-1)
    idx_lo = torch.searchsorted(t1_starts.contiguous(), t0_starts.contiguous(), side="right") - 1
    idx_lo = torch.clamp(idx_lo, min=0, max=y1.shape[-1] - 1)
    idx_hi = torch.searchsorted(t1_ends.contiguous(), t0_ends.contiguous(), side="right")
    idx_hi = torch.clamp(idx_hi, min=0, max=y1.shape[-1] - 1)
    cy1_lo = torch.take_along_dim(cy1[..., :-1], idx_lo, dim=-1)
    cy1_hi = torch.take_along_dim(cy1[..., 1:], idx_hi, dim=-1)
    return cy1_hi - cy1_lo


def lossfun_outer(t: Tensor, w: Tensor, t_env: Tensor, w_env: Tensor) -> Tensor:
    eps = 1.1920928955078125e-07
    w_outer = outer(t[..., :-1], t[..., 1:], t_env[..., :-1], t_env[..., 1:], w_env)
    return torch.clip(w - w_outer, min=0) ** 2 / (w + eps)


def interlevel_loss(weights_list: List[Tensor], sdist_list: List[Tensor]) -> Tensor:
    """weights_list[i] [R,S_i], sdist_list[i] [R,S_i+1] (spacing bins); the last entry is the fine (NeuS) level, detached."""
    c, w = sdist_list[-1].detach(), weights_list[-1].detach()
    loss = 0.0
    for cp, wp in zip(sdist_list[:-1], weights_list[:-1]):
        loss = loss + torch.mean(lossfun_outer(c, w, cp, wp))
    return loss


def density_weights_torch(density: Tensor, deltas: Tensor) -> Tensor:
    """RaySamples.get_weights in torch ops (differentiable twin of density_weights)."""
    dd = deltas * density
    alphas = 1 - torch.exp(-dd)
    T = torch.cumsum(dd[..., :-1], dim=-1)
    T = torch.exp(-torch.cat([torch.zeros_like(dd[..., :1]), T], dim=-1))
    return torch.nan_to_num(alphas * T)


def pdf_resample_torch(spacing_bins: Tensor, weights: Tensor, N: int, histogram_padding: float = 0.01, eps: float = 1e-5) -> Tensor:
    """PDFSampler eval placement written with the torch calls nerfstudio makes (torch.sum / cumsum / searchsorted): the
    cross-check that the explicit-order numpy statement above restates the same algorithm (agreement to a few ulp)."""
    nb = N + 1
    w = weights + histogram_padding
    ws = torch.sum(w, dim=-1, keepdim=True)
    padding = torch.relu(eps - ws)
    w = w + padding / w.shape[-1]
    ws = ws + padding
    pdf = w / ws
    cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, dim=-1))
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
    u = torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb) + 1.0 / (2 * nb)
    u = u.expand(*cdf.shape[:-1], nb).contiguous()
    inds = torch.searchsorted(cdf, u, side="right")
    below = torch.clamp(inds - 1, 0, spacing_bins.shape[-1] - 1)
    above = torch.clamp(inds, 0, spacing_bins.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    b0, b1 = torch.gather(spacing_bins, -1, below), torch.gather(spacing_bins, -1, above)
    t = torch.clip(torch.nan_to_num((u - c0) / (c1 - c0), 0), 0, 1)
    return b0 + t * (b1 - b0)
