"""How much of a training step is GPU-idle?  torch.profiler over a few steps of bench.py's train arm: sum of kernel durations vs wall."""
import os, sys, time, copy, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torch.profiler import profile, ProfilerActivity

sys.argv = ["bench.py", "--workload", "train", "--steps", "4", "--warmup", "3", "--no-cpu-baseline"]
# reuse bench's arm but intercept the timed loop with the profiler: run once normally for the reference time, then profiled
ap_main = bench.main
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    bench.main()
torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total for e in evs) / 1e3
ka = prof.key_averages()
print(f"profiled run: total CUDA kernel time {tot:.1f} ms over {len(evs)} device events (7 steps incl. warm-up + setup)")
rows = sorted(ka, key=lambda a: -(getattr(a, 'device_time_total', 0) or getattr(a, 'cuda_time_total', 0)))[:12]
for a in rows:
    print(f"  {(getattr(a, 'device_time_total', 0) or a.cuda_time_total)/1e3:9.2f} ms  x{a.count:5d}  {a.key[:90]}")
