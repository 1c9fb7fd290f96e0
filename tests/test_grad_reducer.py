"""CPU checks of the data-parallel gradient path (BASELINE config 4): GradBucketReducer over gloo with
world_size 2 must give every rank the gradient of the mean loss over both ranks' ray shards, keep .grad views attached
to the communication buckets across steps, and reduce the big (hash-table-sized) bucket separately from the small one."""
import os
import socket
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(seed):
    g = torch.Generator().manual_seed(seed)
    table = torch.nn.Parameter(torch.randn(4096, 2, generator=g) * 0.1)       # "hash table": its own bucket (big_bytes below)
    W = torch.nn.Parameter(torch.randn(8, 5, generator=g))
    b = torch.nn.Parameter(torch.randn(8, generator=g))
    unused = torch.nn.Parameter(torch.randn(3, generator=g))                   # never touched by the loss
    return table, W, b, unused


def _loss(params, idx, x):
    table, W, b, _ = params
    feat = table[idx].reshape(x.shape[0], -1)                                  # gather -> sparse gradient rows
    return ((torch.cat([feat, x], 1) @ W.t() + b) ** 2).mean()


def _data(seed, n):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 4096, (n,), generator=g), torch.randn(n, 3, generator=g)


def _worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from neusky_b200 import parallel

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    params = _model(0)
    red = parallel.GradBucketReducer(params, big_bytes=16 << 10)
    ok = len(red.buckets) == 2
    for step in range(2):                                                      # two steps: views must survive zero_grad()
        red.zero_grad()
        idx, x = _data(100 + 10 * step + rank, 64)
        _loss(params, idx, x).backward()
        red.finish()
        # reference: the mean over ranks of each rank's loss gradient, computed locally on a fresh replica
        ref = _model(0)
        tot = 0
        for r in range(world):
            i2, x2 = _data(100 + 10 * step + r, 64)
            tot = tot + _loss(ref, i2, x2) / world
        tot.backward()
        for p, pr in zip(params[:3], ref[:3]):
            ok = ok and torch.allclose(p.grad, pr.grad, rtol=1e-5, atol=1e-7)
        ok = ok and bool((params[3].grad == 0).all())
        ok = ok and sorted(red.launched_order) == [0, 1]
        ok = ok and params[0].grad.untyped_storage().data_ptr() == red.buckets[0].untyped_storage().data_ptr()
    # graph-replay protocol (neusky_b200/graphed.py): while an iteration is CAPTURED the hooks must not launch a collective; after a
    # replay (here: the same backward, run with `capturing` set) `rearm()` + `finish()` reduce every bucket in order
    red.zero_grad()
    red.capturing = True
    idx, x = _data(500 + rank, 64)
    _loss(params, idx, x).backward()
    red.capturing = False
    ok = ok and red.launched_order == []
    red.rearm()
    red.finish()
    ref = _model(0)
    tot = 0
    for r in range(world):
        i2, x2 = _data(500 + r, 64)
        tot = tot + _loss(ref, i2, x2) / world
    tot.backward()
    for p, pr in zip(params[:3], ref[:3]):
        ok = ok and torch.allclose(p.grad, pr.grad, rtol=1e-5, atol=1e-7)
    ok = ok and red.launched_order == [0, 1]
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_grad_bucket_reducer_world2_gloo():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_grad_bucket_reducer_single_process_is_identity():
    from neusky_b200 import parallel

    params = _model(1)
    red = parallel.GradBucketReducer(params, big_bytes=16 << 10)
    red.zero_grad()
    idx, x = _data(5, 32)
    _loss(params, idx, x).backward()
    red.finish()
    ref = _model(1)
    _loss(ref, idx, x).backward()
    for p, pr in zip(params[:3], ref[:3]):
        assert torch.allclose(p.grad, pr.grad, rtol=1e-6, atol=1e-8)
    with pytest.raises(RuntimeError):
        params[1].grad = None
        _loss(params, idx, x).backward()
