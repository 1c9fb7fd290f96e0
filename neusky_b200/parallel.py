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


def balanced_tile(n_rays: int, tile: int, world: int) -> int:
    """Largest tile size <= `tile` that cuts `n_rays` into a multiple of `world` (almost) equal tiles, so that the round-robin
    partition gives every rank the same number of rays (a 1280x720 frame is 56.25 tiles of 16384: with that size one rank carries
    a quarter tile more than the others and everybody waits for it at the gather).  world == 1 keeps `tile`."""
    if world <= 1 or n_rays <= 0:
        return tile
    k = (n_rays + world * tile - 1) // (world * tile)          # tiles per rank
    return (n_rays + world * k - 1) // (world * k)


def local_ray_indices(n_rays: int, tile: int, rank: int, world: int, device=None) -> Tensor:
    r = [torch.arange(a, b, device=device) for a, b in tiles_of_rank(n_rays, tile, rank, world)]
    return torch.cat(r) if r else torch.zeros(0, dtype=torch.long, device=device)


def max_local_rays(n_rays: int, tile: int, world: int) -> int:
    return max(sum(b - a for a, b in tiles_of_rank(n_rays, tile, r, world)) for r in range(world))


_GATHER_MAP: Dict[tuple, Tensor] = {}
GATHER_EVENTS = None      # bench hook: when a list, gather_rays appends a (start, end) CUDA event pair around the collective + reorder


def _gather_map(n_rays: int, tile: int, world: int, device) -> Tensor:
    """src [n_rays]: row of the padded all-gather buffer ([world * cap, C]) that holds ray i.  Built once per (image, tile,
    world, device): tile t sits on rank t mod world at local tile slot t div world."""
    key = (n_rays, tile, world, str(device))
    m = _GATHER_MAP.get(key)
    if m is None:
        cap = max_local_rays(n_rays, tile, world)
        i = torch.arange(n_rays, device=device)
        t = i // tile
        m = _GATHER_MAP[key] = ((t % world) * cap + (t // world) * tile + (i - t * tile)).contiguous()
        if len(_GATHER_MAP) > 64:
            _GATHER_MAP.pop(next(iter(_GATHER_MAP)))
    return m


def gather_rays(local: Tensor, n_rays: int, tile: int, group=None) -> Tensor:
    """local [n_local, C] (rows in the order of local_ray_indices) -> [n_rays, C] on every rank.
    One all_gather of equally padded buffers (NCCL over NVLink on GPUs, gloo on CPU) and ONE indexed copy that puts the
    rows back into image order; world_size 1 is a no-op."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if local.shape[0] != n_rays:
            raise ValueError("gather_rays: single process must hold every ray")
        return local
    world = dist.get_world_size(group)
    C = local.shape[1]
    cap = max_local_rays(n_rays, tile, world)
    if local.shape[0] == cap:
        buf = local.contiguous()
    else:
        buf = torch.zeros((cap, C), dtype=local.dtype, device=local.device)
        buf[: local.shape[0]] = local
    out = torch.empty((world * cap, C), dtype=local.dtype, device=local.device)
    ev = None
    if GATHER_EVENTS is not None and local.is_cuda:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    dist.all_gather_into_tensor(out, buf, group=group)
    full = out.index_select(0, _gather_map(n_rays, tile, world, local.device))
    if ev is not None:
        ev[1].record()
        GATHER_EVENTS.append(ev)
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


# =====================================================================================================================
# Training: data-parallel gradient all-reduce (the reference wraps the model in torch DDP,
# neusky/pipelines/neusky_pipeline.py:198-200)
# =====================================================================================================================
class GradBucketReducer:
    """Bucketed gradient all-reduce (sum -> mean) for identical model replicas, one process per GPU.

    Every parameter's .grad is a view into one of a few flat fp32 buckets, so backward kernels accumulate straight into
    communication buffers and nothing is packed or unpacked.  The two hash tables (64 MiB each at T = 2^19) get a bucket
    of their own; the ~5 MB of MLP weights, latents and scalars share one.  A bucket's all-reduce is issued from a
    post-accumulate-grad hook as soon as its last gradient has landed, on the process group's own stream (NCCL over
    NVLink / NVSwitch on GPUs, gloo on CPU), so the DDF buckets are reduced while the SDF backward is still running.
    Buckets are launched in a FIXED order on every rank (bucket i only after buckets 0..i-1, as torch DDP does): a parameter
    that is unused on one rank only must not reorder that rank's collectives.  One `backward()` per `zero_grad()`: a second
    backward into already-launched buckets raises instead of silently dropping its gradients.
    `finish()` launches what is left in order, waits for the outstanding work and applies the 1/world factor.

        red = GradBucketReducer(params)           # once
        red.zero_grad(); loss.backward(); red.finish()          # per step: grads are now the mean over ranks
    """

    def __init__(self, params, big_bytes: int = 8 << 20, group=None):
        import torch.distributed as dist

        self.group = group
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.on else 1
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucketReducer: no trainable parameters")
        dev = self.params[0].device
        small, assign = [], []
        # big parameters in REVERSE registration order (like DDP: gradients of the modules used last in forward arrive first in
        # backward -- here the DDF hash table, whose 64 MiB all-reduce then overlaps the SDF backward), the small ones last
        for p in reversed(self.params):
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("GradBucketReducer: parameters must be fp32 on one device")
            if p.numel() * 4 >= big_bytes:
                assign.append([p])
            else:
                small.append(p)
        small.reverse()
        if small:
            assign.append(small)
        self.buckets: List[Tensor] = []
        self._pending: List[int] = []
        self._bucket_of: Dict[int, int] = {}
        self._sizes: List[int] = []
        for bi, ps in enumerate(assign):
            n = sum((p.numel() + 3) // 4 * 4 for p in ps)                      # 16-byte aligned views
            flat = torch.zeros(n, dtype=torch.float32, device=dev)
            o = 0
            for p in ps:
                p.grad = flat[o : o + p.numel()].view_as(p)
                self._bucket_of[id(p)] = bi
                o += (p.numel() + 3) // 4 * 4
            self.buckets.append(flat)
            self._sizes.append(len(ps))
        self._left = list(self._sizes)
        self._work: List = []
        self.launched_order: List[int] = []
        self._next = 0                      # first bucket not launched yet
        self._finished = False
        self.capturing = False              # set by graphed.GraphedTrainIteration while it captures forward + backward
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._hook)

    @property
    def bytes_per_step(self) -> int:
        return sum(b.numel() * 4 for b in self.buckets)

    def zero_grad(self) -> None:
        """Zero the buckets in place (the .grad views stay attached) and re-arm the ready counters."""
        for b in self.buckets:
            b.zero_()
        self.rearm()

    def rearm(self) -> None:
        """Re-arm the ready counters WITHOUT touching the buckets: what a replay of a captured iteration needs (the zero-fill and
        the backward are inside the CUDA graph, `neusky_b200/graphed.py`; the hooks only ran while it was captured), followed by
        `finish()`, which then launches every bucket in order."""
        self._left = list(self._sizes)
        self._work = []
        self.launched_order = []
        self._next = 0
        self._finished = False

    def _launch(self, bi: int) -> None:
        import torch.distributed as dist

        self.launched_order.append(bi)
        if self.on:
            self._work.append(dist.all_reduce(self.buckets[bi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _hook(self, p) -> None:
        bi = self._bucket_of[id(p)]
        if p.grad is None or p.grad.untyped_storage().data_ptr() != self.buckets[bi].untyped_storage().data_ptr():
            raise RuntimeError("GradBucketReducer: a parameter's .grad was replaced; use reducer.zero_grad(), not optimizer.zero_grad(set_to_none=True)")
        if self.capturing:          # stream capture of the iteration: no collective inside the graph, `finish()` after each replay reduces
            return
        if self._left[bi] <= 0 or self._finished:
            raise RuntimeError("GradBucketReducer: a gradient arrived for a bucket that is already being reduced -- call reducer.zero_grad() "
                               "before every backward() (one backward per step; accumulate micro-batches into the loss instead)")
        self._left[bi] -= 1
        # in-order launch: bucket i goes out only when buckets 0..i-1 are out, so every rank issues the same collective sequence
        while self._next < len(self.buckets) and self._left[self._next] == 0:
            self._launch(self._next)
            self._next += 1

    def finish(self) -> None:
        """Reduce buckets whose hooks never completed (parameters unused this step keep zero gradients, as DDP's
        find_unused_parameters would), wait, and turn sums into means."""
        while self._next < len(self.buckets):
            self._left[self._next] = 0
            self._launch(self._next)
            self._next += 1
        self._finished = True
        for w in self._work:
            w.wait()
        self._work = []
        if self.world > 1:
            for b in self.buckets:
                b.mul_(1.0 / self.world)
