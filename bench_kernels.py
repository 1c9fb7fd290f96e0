"""Per-kernel roofline microbenchmarks of the hot path (K1 hash encode, K2 SDF field, K3 compositing, RENI++ decode, proposal
sampler): CUDA events on the launching stream, 3 warm-up launches, a 256 MiB scratch write between timed launches (L2 flush).
`kernel_rooflines(dev, peaks, scale)` returns one dict per kernel: algorithmic bytes (HBM-bound) or FLOP (tensor-bound) per unit
as SURVEY.md 8(d) states them, x units, / the mean launch time, against the measured peak.  bench.py attaches the list to the
driver-run JSON line as `kernels`; scripts/kernel_bench.py prints it at full size; the matching ncu dram__bytes captures are
profiles/r02_ncu_*.  Nothing here touches oracle/."""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def kernel_rooflines(dev, peaks, scale: float = 1.0, emit=None):
    from neusky_b200 import init as nb_init, ops, packing
    from neusky_b200.samplers import EquirectangularSampler

    out = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, iters=5, warm=3):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(iters):
            flush.fill_(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sum(ts) / len(ts)

    def report(name, bound, ms, work, unit_work, units, extra=None):
        ach = work / (ms * 1e-3) / (1e9 if bound == "hbm" else 1e12)
        peak = peaks["hbm_gbs"] if bound == "hbm" else peaks["tf_burst"]
        d = {"kernel": name, "bound": bound, "ms": ms, "units": units, "work_per_unit": unit_work, "achieved": ach,
             "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "peak": peak, "frac": ach / peak, "peak_source": peaks["source"]}
        if extra:
            d.update(extra)
        out.append(d)
        if emit is not None:
            emit(d)

    sc = nb_init.hash_scalings().to(dev)
    g = torch.Generator().manual_seed(0)

    # ---- K1 hash encode forward: 1164 B/point ------------------------------------------------------------
    n = int(8_000_000 * scale)
    table = nb_init.init_hash_table(1).to(dev)
    x = torch.rand(n, 3, generator=g).to(dev)
    ms = timeit(lambda: ops.hash_encode(x, table, sc, 19))
    report("hash_encode_fwd (K1, L=16 F=2 T=2^19)", "hbm", ms, n * 1164.0, 1164, n, {"note": "uniform random points in [0,1]^3: worst case for gather locality"})
    # ray-ordered points (samples along rays: consecutive points are neighbours in space), the access pattern of the render path
    R, S = int(62_500 * scale), 128
    o = torch.tensor([0.5, -0.4, 0.6], device=dev)
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(dev)
    t = torch.linspace(0.0, 0.45, S, device=dev)
    xr = (o + d[:, None, :] * t[None, :, None]).reshape(-1, 3).contiguous()
    ms = timeit(lambda: ops.hash_encode(xr, table, sc, 19))
    report("hash_encode_fwd (K1), ray-ordered samples", "hbm", ms, xr.shape[0] * 1164.0, 1164, xr.shape[0])
    # backward (scatter): 12 + 128 read + 1024 B read-modify-write
    gout = torch.randn(n, 32, generator=g).to(dev)
    gt = torch.zeros_like(table)
    ms = timeit(lambda: ops.hash_encode_bwd(x, sc, 19, gout, gt), iters=3)
    report("hash_encode_bwd (K1 scatter)", "hbm", ms, n * (12 + 128 + 2 * 1024.0), 12 + 128 + 2048, n, {"note": "atomics: read-modify-write counted as 2 x 1024 B"})
    del gout, gt, x, xr

    # ---- K3 composite: 56*S + 80 B/ray -------------------------------------------------------------------
    for R, S in ((int(1_000_000 * scale), 128), (int(2_000_000 * scale), 48)):
        tt = torch.sort(torch.rand(R, S + 1, generator=g) * 2 + 0.05, dim=1).values.to(dev)
        starts, ends = tt[:, :-1].contiguous(), tt[:, 1:].contiguous()
        deltas = ends - starts
        rd = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(dev)
        sdf = ((1.0 - (starts + ends) / 2) * 0.3).contiguous()
        grad = (-rd[:, None, :] * torch.ones(R, S, 1, device=dev)).contiguous()
        alb = torch.rand(R, S, 3, device=dev)
        dn = torch.ones(R, device=dev)
        ms = timeit(lambda: ops.neus_composite(sdf, grad, alb, rd, starts, ends, deltas, dn, 20.0))
        # this launch also writes normals [R,S,3] and wa [R,S,3] for K4: actual traffic is (52 + 4 + 24)*S + 80
        report(f"neus_composite (K3) S={S}", "hbm", ms, R * (56.0 * S + 80), 56 * S + 80, R, {"actual_bytes_per_ray": (52 + 4 + 24) * S + 80,
               "achieved_actual_GBps": R * ((52 + 4 + 24.0) * S + 80) / (ms * 1e-3) / 1e9})
        del tt, starts, ends, deltas, rd, sdf, grad, alb, dn

    # ---- K2 exact fp32 path: 881,664 FLOP/sample (algorithmic) -----------------------------------
    p = nb_init.init_sdf_params(0)
    blob = packing.pack_sdf_simt(p, device=dev)
    tab = p["encoding.hash_table"].to(dev)
    n = int(1_000_000 * scale)
    x = ((torch.rand(n, 3, generator=g) * 2 - 1) * 0.6).to(dev)
    ms = timeit(lambda: ops.sdf_field(x, blob, tab, sc, 19), iters=3)
    report("sdf_field_simt (K2, exact fp32 CUDA cores)", "tensor", ms, n * 881664.0, 881664, n, {"note": "fp32 FMA path; fraction is against the dense fp16/bf16 tensor peak"})

    blob_tc = packing.pack_sdf_tc(p, device=dev)
    n = int(8_000_000 * scale)
    x = ((torch.rand(n, 3, generator=g) * 2 - 1) * 0.6).to(dev)
    ms = timeit(lambda: ops.sdf_field(x, blob_tc, tab, sc, 19, impl="tc"), iters=3)
    report("sdf_field_tc (K2, tcgen05 fp16xfp16->fp32), random points", "tensor", ms, n * 881664.0, 881664, n)
    R, S = int(62_500 * scale), 128
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(dev)
    t = torch.linspace(0.05, 0.9, S, device=dev)
    xr = (torch.tensor([0.0, -0.3, 0.1], device=dev) + d[:, None, :] * t[None, :, None] * 0.7).reshape(-1, 3).contiguous()
    ms = timeit(lambda: ops.sdf_field(xr, blob_tc, tab, sc, 19, impl="tc"), iters=3)
    report("sdf_field_tc (K2), ray-ordered samples", "tensor", ms, xr.shape[0] * 881664.0, 881664, xr.shape[0])
    del x, xr

    # ---- RENI++ decode: 524,544 FLOP per (camera, direction) ------------------------------------------------
    rp = nb_init.init_reni_params(1)
    rblob = packing.pack_reni(rp, device=dev)
    dirs = EquirectangularSampler(64)().frustums.directions.to(dev)
    for K in (1, 64):
        Z = torch.randn(K, 100, 3, generator=g).to(dev)
        s0 = torch.zeros(K, device=dev)
        ms = timeit(lambda: ops.reni_radiance_table(dirs, Z, s0, rblob))
        report(f"reni_decode K={K} D=2048", "tensor", ms, K * 2048 * 524544.0 + K * 657408.0, 524544, K * 2048, {"note": "fp32 SIMT (direction tables, training)"})
    # frame-sized row batch (the per-ray background of a 1280x720 render): 13 dense layers on the 3xTF32 tcgen05 GEMM chain; the
    # fraction is against the dense fp16/bf16 peak although every contraction runs three tf32 passes (1/6 of that peak at best)
    rgw = packing.pack_reni_gemm(rp, device=dev)
    Nf = 1280 * 720
    rows = torch.nn.functional.normalize(torch.randn(Nf, 3, generator=g), dim=-1).to(dev)
    Z1, s1 = torch.randn(1, 100, 3, generator=g).to(dev), torch.zeros(1, device=dev)
    ms = timeit(lambda: ops.reni_rows_tc(rows, Z1, s1, rblob, rgw), iters=3)
    report("reni_rows_tc N=921600 (3xTF32 GEMM chain)", "tensor", ms, Nf * 524544.0, 524544, Nf, {"note": "algorithmic FLOP; 3xTF32 issues 3 tf32 MMAs per product"})
    ms = timeit(lambda: ops.reni_radiance_table(rows, Z1, s1, rblob), iters=3)
    report("reni_decode rows N=921600 (fp32 SIMT, same work)", "tensor", ms, Nf * 524544.0, 524544, Nf)
    # the product path for frame-sized row batches since round 2: the whole decoder as ONE tcgen05 kernel (fp16 operands, fp32 LayerNorm)
    rfused = packing.pack_reni_fused(rp, device=dev)
    ms = timeit(lambda: ops.reni_rows_fused(rows, Z1, s1, rblob, rfused), iters=5)
    report("reni_rows_fused N=921600 (one tcgen05 kernel, fp16 operands)", "tensor", ms, Nf * 524544.0, 524544, Nf,
           {"note": "LayerNorm epilogues (two per decoder layer) bound it, not the tensor pipe: profiles/r02_reni_fused_phase_cycles.log"})
    del rows

    # ---- relighting pass over the compact cache (8f row f3): 6 DP + 4 bytes per cached ray and pass of eight latent codes --------------
    Rc, Dr = int(921_600 * scale), 642
    DPr = (Dr + 15) // 16 * 16
    H16 = (torch.rand(Rc, 3 * DPr, generator=g) * 0.9).to(torch.float16).to(dev)
    hscale = torch.rand(Rc, generator=g).to(dev)
    rws = torch.arange(Rc, dtype=torch.int32, device=dev)
    rad8 = torch.rand(32, Dr, 3, generator=g).to(dev)
    ms = timeit(lambda: ops.relight_h16_multi(H16, hscale, rws, Rc, Dr, rad8), iters=5)
    byt = 6 * DPr + 8 + 32 * 12
    report("relight_h16 (compact cache pass, 32 latent codes per read)", "hbm", ms, Rc * float(byt), byt, Rc,
           {"note": "warp-level mma m16n8k16 on 16-row tiles; bytes = fp16 cache rows in + 32 x 12 B results out per ray"})
    del H16, hscale, rws

    # ---- proposal-network sampler (8f row f1): density field 372 B/sample algorithmic (12 + 5*8*8 gather + 40 features);
    # ---- PDF resampling: reads bins (S+1)*4 + density S*4, writes weights S*4 + new bins/euclid 2*(N+1)*4 per ray --------------------
    from neusky_b200 import proposal as P
    R = 921_600 // 4
    c2 = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    o = (torch.tensor([0.0, -0.9, 0.25]) + torch.zeros(R, 3)).to(dev).contiguous()
    d = torch.nn.functional.normalize(-o.cpu() + 0.4 * c2, dim=-1).to(dev).contiguous()
    from neusky_b200.render import sphere_collider
    near, far = (t.reshape(-1).contiguous() for t in sphere_collider(o, d))
    for S, N, mr in ((256, 96, 64), (96, 48, 256)):
        f = P.HashMLPDensityField(nb_init.init_proposal_params(mr, table_scale=1.0, density_bias=1.0), mr, device=dev)
        bins = P.uniform_bins(R, S, dev, torch.rand(R, device=dev))
        ms = timeit(lambda: f.density_on_rays(o, d, near, far, bins))
        # NOT an HBM kernel: the 5 MB table is L2 / L1 resident, so its 372 algorithmic bytes per sample are cache traffic (1.2-1.7x the HBM
        # peak) and no HBM fraction is claimed; the DRAM side is ~8 B per sample (ncu: profiles/r02_ncu_hbm_kernels_summary.txt)
        report(f"proposal_density_fwd (P1) S={S} max_res={mr}", "hbm", ms, R * S * 372.0, 372, R * S,
               {"bound": "l2 (5 MB table is cache resident)", "frac": None, "peak": None, "achieved_is": "algorithmic gather + stream bytes per second (cache traffic)",
                "note": "table 5 MB: gathers are L2/L1 hits; actual HBM traffic is ~8 B/sample", "achieved_actual_GBps": R * S * 8.0 / (ms * 1e-3) / 1e9,
                "frac_hbm_actual": R * S * 8.0 / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "Gsamples_per_s": R * S / (ms * 1e-3) / 1e9})
        dens = f.density_on_rays(o, d, near, far, bins)
        ms = timeit(lambda: P.pdf_resample(bins, near, far, N, density=dens))
        byt = (S + 1) * 4 + S * 4 + S * 4 + 2 * (N + 1) * 4 + 8
        report(f"pdf_resample (P2) S={S} -> N={N}", "hbm", ms, R * float(byt), byt, R, {"Mrays_per_s": R / (ms * 1e-3) / 1e6})

    return out


if __name__ == "__main__":
    from bench import _peaks

    kernel_rooflines(torch.device("cuda:0"), _peaks(), 1.0, emit=lambda d: print(json.dumps(d), flush=True))
