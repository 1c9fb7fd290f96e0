"""Ray-tile partitioning for multi-GPU eval renders.

Every ray is independent given replicated weights, so an image is cut into contiguous tiles of `tile` rays,
tile i goes to rank i mod world, and there is NO data-path collective while rendering; the only exchange is the
final gather of the per-ray outputs ([H*W, C] floats).  The reference has nothing like this: it renders an image
serially in 256-ray chunks on one GPU (neusky/models/neusky_model.py:1413-1437); its only multi-GPU code is the
DDP wrap for training (neusky/pipelines/neusky_pipeline.py:198-200).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Tuple

import torch

Tensor = torch.Tensor


def tiles_of_rank(n_rays: int, tile: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """[(start, end)) ray ranges rendered by `rank`: tile i -> rank i mod world (round-robin keeps sky-heavy and
    surface-heavy image regions spread over all ranks)."""
    if tile <= 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("tiles_of_rank: bad tile / rank / world")
    n_tiles = (n_rays + tile - 1) // tile
    return [(i * tile, min((i + 1) * tile, n_rays)) for i in range(rank, n_tiles, world)]


def local_ray_indices(n_rays: int, tile: int, rank: int, world: int, device=None) -> Tensor:
    r = [torch.arange(a, b, device=device) for a, b in tiles_of_rank(n_rays, tile, rank, world)]
    return torch.cat(r) if r else torch.zeros(0, dtype=torch.long, device=device)


def max_local_rays(n_rays: int, tile: int, world: int) -> int:
    return max(sum(b - a for a, b in tiles_of_rank(n_rays, tile, r, world)) for r in range(world))


def gather_rays(local: Tensor, n_rays: int, tile: int, group=None) -> Tensor:
    """local [n_local, C] (rows in the order of local_ray_indices) -> [n_rays, C] on every rank.
    One all_gather of equally padded buffers (NCCL over NVLink on GPUs, gloo on CPU); world_size 1 is a no-op."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if local.shape[0] != n_rays:
            raise ValueError("gather_rays: single process must hold every ray")
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    C = local.shape[1]
    cap = max_local_rays(n_rays, tile, world)
    buf = torch.zeros((cap, C), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty((world * cap, C), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    full = torch.empty((n_rays, C), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = local_ray_indices(n_rays, tile, r, world, device=local.device)
        full[idx] = out[r * cap : r * cap + idx.shape[0]]
    return full


def render_sharded(render_fn: Callable[[Tensor], Dict[str, Tensor]], n_rays: int, tile: int, keys: Tuple[str, ...], device=None, group=None) -> Dict[str, Tensor]:
    """Run `render_fn(ray_indices) -> {key: [n, c]}` on this rank's tiles and gather every key to all ranks."""
    import torch.distributed as dist

    on = dist.is_available() and dist.is_initialized()
    world, rank = (dist.get_world_size(group), dist.get_rank(group)) if on else (1, 0)
    idx = local_ray_indices(n_rays, tile, rank, world, device=device)
    out = render_fn(idx)
    n = idx.shape[0]
    widths = [int(torch.Size(out[k].shape[1:]).numel()) for k in keys]
    packed = torch.cat([out[k].reshape(n, w).to(torch.float32) for k, w in zip(keys, widths)], 1)
    full = gather_rays(packed, n_rays, tile, group)
    res, o = {}, 0
    for k, w in zip(keys, widths):
        res[k] = full[:, o : o + w]
        o += w
    return res
