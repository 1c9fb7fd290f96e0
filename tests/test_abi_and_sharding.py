"""CPU-only checks: the C-ABI library loads and exports every symbol the public header declares (no compute
calls without a GPU), the ops refuse CPU tensors, and the world_size-2 ray-tile sharding logic over gloo."""
import os
import socket
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    from neusky_b200 import _lib

    g.build()
    lib = _lib.load()
    syms = _lib.declared_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/neusky_b200.h but not exported"
    assert lib.nsk_version() == 1
    assert lib.nsk_ddf_tc_weights_bytes() > 1_000_000


def test_ops_refuse_cpu_tensors():
    from neusky_b200 import ops

    with pytest.raises(ValueError):
        ops.hash_encode(torch.zeros(4, 3), torch.zeros(16 << 4, 2), torch.ones(16), 4)
    with pytest.raises(ValueError):
        ops.surface_points(torch.zeros(2, 3), torch.zeros(2, 3), torch.zeros(2), 1.0)


def test_tile_partition_covers_every_ray_once():
    from neusky_b200 import parallel

    for n, tile, world in ((0, 4, 2), (1, 4, 2), (10, 4, 2), (1000, 64, 8), (921600, 4096, 8), (7, 100, 4)):
        seen = torch.zeros(n, dtype=torch.int32)
        for r in range(world):
            idx = parallel.local_ray_indices(n, tile, r, world)
            seen[idx] += 1
        assert bool((seen == 1).all())
        assert parallel.max_local_rays(n, tile, world) * world >= n


def _worker(rank, world, port, n, tile, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from neusky_b200 import parallel

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    def fake_render(idx):   # stands in for RayRenderer.render on this rank's rays
        f = idx.to(torch.float32)
        return {"rgb": torch.stack([f, f * 2, f * 3], 1), "depth": (f + 0.5)[:, None]}

    out = parallel.render_sharded(fake_render, n, tile, ("rgb", "depth"))
    ref = fake_render(torch.arange(n))
    ok = torch.equal(out["rgb"], ref["rgb"]) and torch.equal(out["depth"], ref["depth"])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,tile", [(1000, 64), (130, 64), (64, 64)])
def test_sharded_render_gathers_identically_world2_gloo(n, tile):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, tile, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_balanced_tile_equalises_ranks():
    """parallel.balanced_tile: the tile size the eval split uses at N > 1 gives every rank the same number of rays (+-1 tile rounding)
    and never exceeds the requested tile; world 1 keeps the requested tile."""
    from neusky_b200.parallel import balanced_tile, tiles_of_rank

    assert balanced_tile(921600, 16384, 1) == 16384
    for n, tile in ((921600, 16384), (230400, 16384), (1000, 64), (17, 4)):
        for world in (2, 3, 4, 8):
            t = balanced_tile(n, tile, world)
            assert 0 < t <= tile
            per_rank = [sum(b - a for a, b in tiles_of_rank(n, t, r, world)) for r in range(world)]
            assert sum(per_rank) == n
            assert max(per_rank) - min(per_rank) <= max(1, world * ((n + t - 1) // t) // world), (n, tile, world, per_rank)
            assert max(per_rank) <= -(-n // world) + t            # nobody carries more than its share plus one tile's rounding
