"""NCCL transport / bandwidth probe for the two collectives of the path (eval all-gather of per-ray outputs, training gradient all-reduce).
torchrun --nproc-per-node N tools/nccl_probe.py ; run with NCCL_DEBUG=INFO to see the transport (P2P/NVLS vs SHM)."""
import os, time, torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)


def timeit(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


n_rays, C = 921600, 12
cap = (n_rays + world - 1) // world
buf = torch.randn(cap, C, device=dev)
out = torch.empty(world * cap, C, device=dev)
ms = timeit(lambda: dist.all_gather_into_tensor(out, buf))
idx = torch.randperm(world * cap, device=dev)[:n_rays]
ms_sel = timeit(lambda: out.index_select(0, idx))
g = torch.randn(16 * 2**20, device=dev)          # 64 MiB bucket
ms_ar = timeit(lambda: dist.all_reduce(g))
if rank == 0:
    print(f"world {world}: all_gather_into_tensor {out.numel()*4/1e6:.1f} MB total: {ms:.3f} ms ({out.numel()*4/ms/1e6:.1f} GB/s out); index_select {ms_sel:.3f} ms; "
          f"all_reduce 64 MiB: {ms_ar:.3f} ms (bus {2*(world-1)/world*g.numel()*4/ms_ar/1e6:.1f} GB/s)", flush=True)
dist.destroy_process_group()
