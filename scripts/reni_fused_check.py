"""Bring-up check of the fused RENI++ row kernel: values vs the fp32 SIMT decode (and the reference golden), then timing."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neusky_b200 import init as nb_init, ops, packing

dev = torch.device("cuda:0")
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reni.npz"))
p = nb_init.init_reni_params(int(g["seed"]))
blob, fused = packing.pack_reni(p, device=dev), packing.pack_reni_fused(p, device=dev)
dirs, Z, sc, rot = (torch.from_numpy(g[k]).to(dev) for k in ("dirs", "latents", "scale", "rotation"))
K, D = Z.shape[0], dirs.shape[0]
rows = dirs[None].expand(K, D, 3).reshape(-1, 3).contiguous()
cam = torch.arange(K, device=dev, dtype=torch.int32)[:, None].expand(K, D).reshape(-1).contiguous()
for key, R in (("radiance", None), ("radiance_rot", rot)):
    out = ops.reni_rows_fused(rows, Z, sc, blob, fused, rotation=R, row_cam=cam).reshape(K, D, 3).cpu()
    ref = torch.from_numpy(g[key])
    rel = ((out - ref).abs() / ref.abs()).max()
    print(json.dumps({"check": f"golden {key}", "max_rel": float(rel), "finite": bool(torch.isfinite(out).all())}), flush=True)
gen = torch.Generator().manual_seed(0)
for N in (1, 255, 256, 257, 1000, 40000):
    d = torch.nn.functional.normalize(torch.randn(N, 3, generator=gen), dim=-1).to(dev)
    Z1 = torch.randn(1, 100, 3, generator=gen).to(dev)
    s1 = torch.zeros(1, device=dev)
    a = ops.reni_rows_fused(d, Z1, s1, blob, fused)
    b = ops.reni_radiance_table(d, Z1, s1, blob)[0]
    torch.cuda.synchronize()
    rel = ((a - b).abs() / b.abs()).max()
    la = ops.reni_rows_fused(d, Z1, s1, blob, fused, log_domain=2)
    print(json.dumps({"check": f"vs simt N={N}", "max_rel": float(rel), "log_abs": float((la - torch.log(b)).abs().max()), "finite": bool(torch.isfinite(a).all())}), flush=True)
Nf = 1280 * 720
d = torch.nn.functional.normalize(torch.randn(Nf, 3, generator=gen), dim=-1).to(dev)
Z1, s1 = torch.randn(1, 100, 3, generator=gen).to(dev), torch.zeros(1, device=dev)
gw = packing.pack_reni_gemm(p, device=dev)
for name, fn in (("fused", lambda: ops.reni_rows_fused(d, Z1, s1, blob, fused)), ("3xtf32 chain", lambda: ops.reni_rows_tc(d, Z1, s1, blob, gw))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"timing": name, "N": Nf, "ms": ms, "TFLOPs": Nf * 524544 / (ms * 1e-3) / 1e12}), flush=True)
a, b = ops.reni_rows_fused(d, Z1, s1, blob, fused), ops.reni_rows_tc(d, Z1, s1, blob, gw)
rel = ((a - b).abs() / b.abs())
print(json.dumps({"check": "frame vs 3xtf32", "max_rel": float(rel.max()), "mean_rel": float(rel.mean()), "p999": float(rel.flatten().kthvalue(int(0.999 * rel.numel())).values)}), flush=True)
