#!/usr/bin/env python
"""bench.py -- shaded rays/s of the NeuSky shading hot path (BASELINE.json config 2).

Workload (N = 1, `config.workload`): 1,000,000 synthetic surface points x 2048 RENI++ directions
(32x64 equirectangular grid, `EquirectangularSampler(width=64)`), DDF sky visibility on the 1024
upper-hemisphere directions (`only_upperhemisphere_visibility=True`, neusky_config.py:157), fused
forward.  One step = RENI++ radiance decode -> Lambert pre-pass -> fused DDF visibility + cosine-weighted
sum (K4, tcgen05) -> sRGB, for every point.  N > 1: every rank shades its own 1M points with replicated
weights (weak scaling, no data-path collective -- SURVEY.md 8e).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` = points of all ranks / device time (max over ranks), inputs
resident in HBM; `e2e` = the same through `SkyShader.shade_points_host` with pinned HOST buffers,
H2D + D2H inside the timed region; `roofline` = the K4 kernel's algorithmic FLOP/s (2,385,408 FLOP per
(point, direction) pair) against the measured dense tensor peak; `cpu_baseline` = the CPU
oracle (a port of the reference algorithm: `oracle/`) on a bounded sample on this box's host cores.
`--impl reference` times that CPU implementation alone (rank 0 only).

Nothing here reads /root/reference.  The oracle is used only for the cpu_baseline / reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR = 2_385_408          # DDF network, SURVEY.md 8(d): 1,192,704 MAC per (point, direction) pair
METRIC = "shaded rays/s (NeuS+RENI+++DDF visibility)"
UNIT = "rays/s"
WIDTH = 64                         # equirect 32 x 64 = 2048 directions
SEED_W, SEED_P = 0, 1


# ------------------------------------------------------------------------------------------ helpers
def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "tf_burst": float(d.get("bf16_tflops", 1590.0)),
                "tf_sustained": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "source": "measured"}
    # /opt/skills/guides/B200_PROFILING.md fallback (earlier measurement on this pool)
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


def _ncu_traffic(pairs: int):
    """dram bytes per launch of the K4 kernel from the committed ncu capture (profiles/k4_ncu_summary.json); only
    reported when the capture was taken at this launch size."""
    path = os.path.join(ROOT, "profiles", "k4_ncu_summary.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                d = json.load(f)
            return d.get("dram_bytes_per_launch") if int(d.get("pairs_per_launch", -1)) == pairs else None
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi sampling of SM clocks / throttle reasons during the timed region."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w": round(statistics.median(pw), 1),
                "reasons": sorted(reasons), "samples": len(sm)}


def _inputs(P: int, seed: int):
    """points uniform in the ball |p| < 0.95, unit normals uniform on S^2, albedo U(0,1)."""
    g = torch.Generator().manual_seed(seed)
    pts = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=-1) * torch.rand(P, 1, generator=g) ** (1.0 / 3.0) * 0.95
    g2 = torch.Generator().manual_seed(seed + 1)
    nrm = torch.nn.functional.normalize(torch.randn(P, 3, generator=g2), dim=-1)
    alb = torch.rand(P, 3, generator=torch.Generator().manual_seed(seed + 2))
    return pts.contiguous(), nrm.contiguous(), alb.contiguous()


def _latents():
    return torch.randn(1, 100, 3, generator=torch.Generator().manual_seed(3)), torch.zeros(1)


def _equirect_directions(width: int):
    """EquirectangularSampler(width) directions, z-up (ns_reni illumination_samplers.py:373-432)."""
    from neusky_b200.samplers import EquirectangularSampler

    return EquirectangularSampler(width)().frustums.directions


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_rate(target_s: float = 15.0, max_points: int = 8192, chunk: int = 64):
    """Points/s of the CPU oracle (reference algorithm, torch fp32 on all host threads) on a bounded
    sample of the config-2 workload.  Returns (rate, cores, sample description, seconds)."""
    from oracle import neusky_oracle as O
    from neusky_b200 import init as nb_init

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ddf = nb_init.init_ddf_params(SEED_W)
    reni = nb_init.init_reni_params(SEED_W + 1)
    dirs = O.equirect_directions(WIDTH)
    Z, sc = _latents()
    sca = O.hash_scalings()

    def run(P, seed):
        pts, nrm, alb = _inputs(P, seed)
        t0 = time.perf_counter()
        with torch.no_grad():
            rad = O.reni_radiance_table(dirs, Z, sc, reni)[0]
            for i in range(0, P, chunk):
                lin, _ = O.shade_points(pts[i:i + chunk], nrm[i:i + chunk], alb[i:i + chunk], dirs, rad, ddf, sca, 19, 1.0, 0.1, 25.0)
                O.linear_to_srgb(lin)
        return time.perf_counter() - t0

    run(chunk, 100)                       # warm-up (thread pool, allocator)
    t_cal = run(chunk, 101)
    P = int(max(chunk, min(max_points, (target_s / max(t_cal, 1e-6)) * chunk)))
    P = (P // chunk) * chunk
    t = run(P, SEED_P)
    return P / t, cores, f"{P} of the 1,000,000 points x 2048 directions (D'=1024 through the DDF), fp32 torch CPU, one pass", t


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rates, secs, desc, cores = [], [], "", 1
    per_step = max(5.0, min(30.0, 120.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        r, cores, desc, t = cpu_oracle_rate(target_s=per_step)
        if i >= args.warmup:
            rates.append(r); secs.append(t)
    v = sum(rates) / len(rates)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": _config(args, 1),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _config(args, world):
    return {"workload": f"BASELINE.json configs[1]: shading microbench, {args.points:,} synthetic surface points/GPU x 2048 RENI++ directions "
                        f"(32x64 equirect; D'=1024 upper-hemisphere directions through the DDF visibility field), fused forward",
            "points_per_gpu": args.points, "directions": 2048, "ddf_directions": 1024, "latent_codes": 1,
            "parallelism": f"ray-sharded x{world}, weights replicated, no collective",
            "l2": "256 MiB scratch write between timed steps (L2 flush); the 1.7 MB fp16 weight stream is L2-resident by design",
            "k4_numerics": "fp16 operands, fp32 accumulate (tcgen05.mma.cta_group::2 kind::f16, CTA pairs), fp32 epilogues"}


# ------------------------------------------------------------------------------------------ process / device context
class Ctx:
    """One process per GPU: rank / device / process group, initialised once per bench.py invocation."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, *xs):
        if self.dist is None:
            return xs if len(xs) > 1 else xs[0]
        t = torch.tensor(list(xs), device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        v = [float(x) for x in t]
        return v if len(v) > 1 else v[0]

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ eval-render arm (configs[2])
def eval_arm(args, ctx):
    """Full-image eval render of a 1280x720 synthetic camera, S uniform samples per ray, D = 642 icosphere directions
    (308 through the DDF), ray tiles round-robin over the ranks, per-ray outputs gathered at the end (strong scaling:
    the image is fixed).  Supplementary line: the driver's headline is the default workload."""
    rank, local, world, dev, dist = ctx.rank, ctx.local, ctx.world, ctx.dev, ctx.dist
    import math
    from neusky_b200 import _lib, init as nb_init, samplers
    from neusky_b200.render import RayRenderer, pinhole_rays, render_image

    sdf_p = nb_init.init_sdf_params(SEED_W + 2, bias=0.45)
    sdf_p["deviation_network.variance"] = torch.tensor(0.3)
    prop = None
    if args.sampler == "proposal":      # the shipped NeuS-facto placement: 256 -> 96 -> S samples through two HashMLPDensityFields
        prop = [nb_init.init_proposal_params(SEED_W + 3, table_scale=1.0, density_bias=1.0), nb_init.init_proposal_params(SEED_W + 4, table_scale=1.0, density_bias=2.0)]
    r = RayRenderer(sdf_p, nb_init.init_ddf_params(SEED_W), nb_init.init_reni_params(SEED_W + 1), device=dev, proposal_params=prop)
    r.set_directions(samplers.IcosahedronSampler(512)().frustums.directions)
    H, W = args.height, args.width
    fx = (W / 2) / math.tan(math.radians(30.0))
    eye = torch.tensor([0.0, -0.9, 0.25]); f = torch.nn.functional.normalize(-eye, dim=0)
    rt = torch.nn.functional.normalize(torch.linalg.cross(f, torch.tensor([0.0, 0.0, 1.0])), dim=0)
    c2w = torch.cat([torch.stack([rt, torch.linalg.cross(rt, f), -f], 1), eye[:, None]], 1)
    o, d, dn = pinhole_rays(H, W, fx, fx, W / 2, H / 2, c2w, dev)
    Z, sc = (t.to(dev) for t in _latents())
    Dp = int(r.shader.mask.sum())
    from neusky_b200 import parallel as _par

    tile = _par.balanced_tile(H * W, args.tile, world)      # equal ray counts per rank (N = 1 keeps --tile)

    def step():
        return render_image(r, o, d, dn, args.samples, Z[0], sc[0], tile=tile)

    for _ in range(args.warmup):
        step()
    ctx.barrier()
    l0 = _lib.launches
    ts, tg = [], []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _par.GATHER_EVENTS = []
        e0.record(); out = step(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 1e3)
        tg.append(sum(a.elapsed_time(b) for a, b in _par.GATHER_EVENTS) / 1e3)
    _par.GATHER_EVENTS = None
    t, t_gather = ctx.max_over_ranks(sum(ts), sum(tg))
    line = None
    if rank == 0:
        n = H * W
        acc = out["accumulation"]
        line = ({"metric": METRIC, "value": n * args.steps / t, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16xf16->f32", "data": "synthetic",
                          "config": {"workload": f"BASELINE.json configs[2]: full-image eval render {W}x{H}, {args.samples} {'proposal-network (256->96->' + str(args.samples) + ')' if args.sampler == 'proposal' else 'uniform'} samples/ray, 642 icosphere directions (D'={Dp}), "
                                                 f"ray tiles of {tile} round-robin over {world} GPU(s), outputs gathered", "parallelism": f"ray tiles x{world}, weights replicated",
                                     "l2": "inputs (118 M samples/frame) exceed L2", "surface_coverage": float((acc > 0.5).float().mean())},
                          "gpu_launches": _lib.launches - l0, "pairs_per_frame": n * Dp, "samples_per_frame": n * args.samples,
                          "gather_ms_per_step": 1e3 * t_gather / args.steps, "gather_bytes_per_step": n * 12 * 4 if world > 1 else 0,
                          "collective": "all_gather_into_tensor of the per-ray outputs (12 floats/ray) + one indexed copy" if world > 1 else "none (single rank)"})
    return line



# ------------------------------------------------------------------------------------------ relighting sweep (configs[4])
def relight_arm(args, ctx):
    """BASELINE.json configs[4]: fixed geometry, 64 RENI++ latent codes re-shaded per view.  The frame is rendered once with
    `want_cache=True` (per-sample shading inputs + per-ray visibility of the D' DDF directions, kept in HBM), then every
    latent code is one RENI++ decode + one Lambertian pass over the cache per tile -- no SDF field, no compositing, no DDF.
    The timed region is the 64-code sweep; the cache build is reported separately.  Ray tiles are partitioned over the ranks
    (strong scaling), each rank relights its own tiles; no collective in the timed region."""
    rank, local, world, dev, dist = ctx.rank, ctx.local, ctx.world, ctx.dev, ctx.dist
    import math
    from neusky_b200 import _lib, init as nb_init, samplers
    from neusky_b200.render import RayRenderer, global_steps_minmax, pinhole_rays

    sdf_p = nb_init.init_sdf_params(SEED_W + 2, bias=0.45)
    sdf_p["deviation_network.variance"] = torch.tensor(0.3)
    r = RayRenderer(sdf_p, nb_init.init_ddf_params(SEED_W), nb_init.init_reni_params(SEED_W + 1), device=dev)
    r.set_directions(samplers.IcosahedronSampler(512)().frustums.directions)
    H, W, S, NL = args.height, args.width, args.samples, args.latents
    fx = (W / 2) / math.tan(math.radians(30.0))
    eye = torch.tensor([0.0, -0.9, 0.25]); f = torch.nn.functional.normalize(-eye, dim=0)
    rt = torch.nn.functional.normalize(torch.linalg.cross(f, torch.tensor([0.0, 0.0, 1.0])), dim=0)
    c2w = torch.cat([torch.stack([rt, torch.linalg.cross(rt, f), -f], 1), eye[:, None]], 1)
    o, d, dn = pinhole_rays(H, W, fx, fx, W / 2, H / 2, c2w, dev)
    Z0, sc = (t.to(dev) for t in _latents())
    codes = torch.randn(NL, 100, 3, generator=torch.Generator().manual_seed(3)).to(dev)     # 64 latent codes N(0,1), seed 3
    n = H * W
    tiles = [(a, min(n, a + args.tile)) for a in range(0, n, args.tile)][rank::world]
    my_dirs = torch.cat([d[a:b] for a, b in tiles]).contiguous() if tiles else d[:0]
    offs = [0]
    for a, b in tiles:
        offs.append(offs[-1] + (b - a))
    mm = global_steps_minmax(o, d, S)
    compact = not args.per_sample_cache and not args.fp32_cache
    if tiles:      # untimed warm-up of the render path (allocator, lazy kernel attributes) so that cache_build_s is a steady-state figure
        a, b = tiles[0]
        r.render(o[a:b].contiguous(), d[a:b].contiguous(), dn[a:b].contiguous(), S, Z0[0], sc[0], steps_minmax=mm, want_cache=True,
                 collapse_cache=not args.per_sample_cache, compact_cache=compact)
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    caches = []
    for a, b in tiles:
        out = r.render(o[a:b].contiguous(), d[a:b].contiguous(), dn[a:b].contiguous(), S, Z0[0], sc[0], steps_minmax=mm, want_cache=True,
                       collapse_cache=not args.per_sample_cache, compact_cache=compact)
        caches.append(out["relight_cache"])
    del out
    if compact and caches:
        frame_cache = r.merge_caches(caches)      # one cache for this rank's rays: a sweep is a handful of launches, not one set per tile
        caches = [frame_cache]
    e1.record()
    torch.cuda.synchronize()
    build_s = e0.elapsed_time(e1) / 1e3
    cache_bytes = sum(sum(v.numel() * v.element_size() for v in c.values() if torch.is_tensor(v)) for c in caches)

    sc_all = torch.zeros(NL, device=dev)

    def sweep():
        last = None
        if compact:
            return r.relight_sweep(caches[0], codes, sc_all) if caches else None
        if args.per_sample_cache:
            for k in range(NL):
                rad, bg = r.illumination_for(codes[k], sc[0], my_dirs)          # two RENI++ decodes per latent, not per tile
                for i, c in enumerate(caches):
                    last = r.relight(c, codes[k], sc[0], radiance=rad, background=bg[offs[i]:offs[i + 1]])
            return last
        for k0 in range(0, NL, 4):                                              # four latent codes per pass over the collapsed cache
            ill = [r.illumination_for(codes[k], sc[0], my_dirs) for k in range(k0, min(NL, k0 + 4))]
            rad = torch.cat([a for a, _ in ill], 0)
            bg = torch.stack([b for _, b in ill], 0)
            for i, c in enumerate(caches):
                last = r.relight_many(c, rad, bg[:, offs[i]:offs[i + 1]])
        return last

    for _ in range(max(1, args.warmup // 3)):
        sweep()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = _lib.launches
    ts = []
    for _ in range(args.steps):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); sweep(); s1.record()
        torch.cuda.synchronize()
        ts.append(s0.elapsed_time(s1) / 1e3)
    t = sum(ts)
    stages = None
    if compact and caches:      # one more sweep with per-stage CUDA events: where a latent code's time goes
        r.sweep_events = {}
        sweep(); torch.cuda.synchronize()
        stages = {k: sum(a.elapsed_time(b) for a, b in v) / NL for k, v in r.sweep_events.items()}
        r.sweep_events = None
    if dist is not None:
        tt = torch.tensor([t, build_s], device=dev, dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX); t, build_s = (float(x) for x in tt)
    line = None
    if rank == 0:
        peaks = _peaks()
        # bytes one relight pass must read per ray: the collapsed coefficients D x 3 x 4 (or, per-sample cache: normals 12 + wa 12 +
        # inv_count 4 per sample and D' x 4 visibilities) + accumulation + direction; writes 12
        Dp = int(r.shader.mask.sum())
        per_ray = (S * 28 + Dp * 4) if args.per_sample_cache else 642 * 12
        algo = n * (per_ray + 16 + 12) * NL * args.steps
        if compact:      # compact cache: 6 DP + 4 bytes per hit ray and pass of up to 32 codes; 12 B direction + 12 B result per ray and code
            hit = float(caches[0]["rows"].shape[0]) if caches else 0.0
            algo = (hit * world * (6 * 656 + 4) * ((NL + 31) // 32) + n * 24.0 * NL) * args.steps
        roof = {"kernel": "lambert_relight_kernel" if args.per_sample_cache else ("relight_h16_kernel" if compact else "relight_collapsed_kernel"), "bound": "hbm",
                "achieved": algo / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": algo / t / 1e9 / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"],
                "note": "whole-sweep rate over the cache bytes: includes the RENI++ decodes of the direction set and of the per-ray background"}
        if compact and stages and caches:
            # the dominant kernel of a sweep is the per-ray background decode (the fused RENI++ row kernel, tensor-core bound); the pass over the
            # cache is reported beside it against the HBM roofline, each from its own CUDA-event time
            bg_rows = float(caches[0]["bg_rows"].shape[0]) * world
            tf = bg_rows * 524544.0 / (stages["background"] * 1e-3) / 1e12 if stages.get("background") else 0.0
            pass_bytes = hit * world * (6 * 656 + 4) / 32.0 + n * 12.0          # per latent code: 1/32 of a read of the cache + 12 B of result per ray
            pass_gbs = pass_bytes / (stages["pass"] * 1e-3) / 1e9 if stages.get("pass") else 0.0
            roof = {"kernel": "reni_rows_fused_kernel (per-ray background radiance, %.0f %% of a latent code's time)" % (100.0 * stages["background"] / max(1e-9, sum(stages.values()))),
                    "bound": "tensor", "achieved": tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["tf_sustained"], "traffic": None,
                    "peak_source": f"{peaks['source']} dense bf16/fp16 cuBLAS, sustained", "flop_per_row": 524544, "rows_per_latent": bg_rows,
                    "note": "bound by its LayerNorm epilogues, not the tensor pipe (DESIGN.md 3)"}
            line_pass = {"kernel": "relight_h16_kernel (cache pass, 32 codes per read)", "bound": "hbm", "achieved": pass_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": pass_gbs / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"]}
        else:
            line_pass = None
        line = ({"metric": "relit rays/s (fixed geometry, new RENI++ latent code per pass)", "value": n * NL * args.steps / t, "unit": UNIT, "n_gpus": world,
                          "steps": args.steps, "warmup": max(1, args.warmup // 3), "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"BASELINE.json configs[4]: relighting sweep, {W}x{H} frame, {S} samples/ray, {NL} latent codes per step, "
                                                 f"642 icosphere directions (D'={Dp} cached visibilities per ray), tiles of {args.tile} rays over {world} GPU(s)",
                                     "parallelism": f"ray tiles x{world}, weights replicated, no collective", "l2": f"relight cache {cache_bytes / 2**30:.2f} GiB per rank exceeds L2",
                                     "cache_build_s": build_s, "ms_per_latent_frame": 1e3 * t / args.steps / NL,
                                     "ms_per_latent_by_stage": stages, "background_rows": int(caches[0]["bg_rows"].shape[0]) if (compact and caches) else None,
                                     "cache_rows": int(caches[0]["rows"].shape[0]) if (compact and caches) else None},
                          "cache_format": "compact: fp16 channel-planar rows for hit rays + per-row scale" if compact else ("per-sample fp32" if args.per_sample_cache else "collapsed fp32 [R,D,3]"),
                          "roofline": roof, "roofline_cache_pass": line_pass,
                          "gpu_launches": _lib.launches - l0})
    return line


# ------------------------------------------------------------------------------------------ training-step arm (configs[3])
def _train_batch(R: int, K: int, seed: int):
    """R rays sampled from K synthetic cameras on a ring of radius 0.9 around the origin (z-up,
    height 0.25), each looking at a random point of the ball |x| < 0.5; random target colours and masks."""
    g = torch.Generator().manual_seed(seed)
    cam = torch.randint(0, K, (R,), generator=g)
    ang = cam.to(torch.float32) * (2 * 3.141592653589793 / K)
    o = torch.stack([0.9 * torch.cos(ang), 0.9 * torch.sin(ang), torch.full((R,), 0.25)], 1)
    tgt = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1) * torch.rand(R, 1, generator=g) ** (1 / 3) * 0.5
    d = torch.nn.functional.normalize(tgt - o, dim=-1)
    return {"origins": o.contiguous(), "directions": d.contiguous(), "dnorm": torch.ones(R, 1), "cam": cam.to(torch.int32),
            "image": torch.rand(R, 3, generator=g), "fg": (torch.rand(R, generator=g) > 0.3).float(), "ground": (torch.rand(R, generator=g) > 0.7).float(),
            "sky": (torch.rand(R, generator=g) > 0.8).float()}


def train_arm(args, ctx):
    """One `ns-train neusky` iteration per step at R = 1024 rays per GPU (README default): forward (uniform S = 48 samples,
    SDF/albedo field with analytic normals, NeuS compositing, RENI++ radiance, DDF visibility on the R x D' pairs of a randomly
    rotated 642-direction icosphere, sdf_at_termination, Lambertian shading), the reference's losses, backward into every
    parameter group, bucketed gradient all-reduce (NCCL) and a fused Adam step.  Weak scaling: every rank has its own R rays."""
    rank, local, world, dev, dist = ctx.rank, ctx.local, ctx.world, ctx.dev, ctx.dist
    import numpy as np
    from scipy.spatial.transform import Rotation
    from neusky_b200 import _lib, init as nb_init, samplers
    from neusky_b200.parallel import GradBucketReducer
    from neusky_b200.train import NeuSkyTrainStep

    _lib.load()
    R, K, S = args.rays, 32, args.train_samples
    sdf_p = nb_init.init_sdf_params(SEED_W + 2, bias=0.45)
    sdf_p["deviation_network.variance"] = torch.tensor(0.3)
    prop = None
    if args.train_sampler == "proposal":   # the shipped NeuS-facto placement: 256 -> 96 -> S samples through two HashMLPDensityFields + interlevel loss
        prop = [nb_init.init_proposal_params(SEED_W + 3, table_scale=1.0, density_bias=1.0), nb_init.init_proposal_params(SEED_W + 4, table_scale=1.0, density_bias=2.0)]
    step_mod = NeuSkyTrainStep(sdf_p, nb_init.init_ddf_params(SEED_W), nb_init.init_reni_params(SEED_W + 1), num_cameras=K, device=dev,
                               num_samples=S, split_geo=3, split=args.split, ddf_split_bwd=(args.ddf_split_bwd or None), threshold_init=0.4, proposal_params=prop)
    with torch.no_grad():
        step_mod.latents.copy_(torch.randn(K, 100, 3, generator=torch.Generator().manual_seed(3)).to(dev))
    params = [p for p in step_mod.parameters() if p.requires_grad]
    red = GradBucketReducer(params)
    opt = torch.optim.Adam(params, lr=1e-4, fused=True)
    base_dirs = samplers.IcosahedronSampler(512)().frustums.directions.to(torch.float64).numpy()
    rots = Rotation.random(args.steps + args.warmup + 8, random_state=np.random.RandomState(11 + rank)).as_matrix()
    batch_h = {k: v.pin_memory() for k, v in _train_batch(R, K, 100 + rank).items()}
    gres = 10
    lin = torch.linspace(-1.0, 1.0, gres)
    gpos0 = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(-1, 3)
    gg = torch.Generator().manual_seed(200 + rank)
    it = {"i": 0}
    loss_h = torch.zeros(1).pin_memory()
    Dp_seen = []
    fit = None
    if not args.no_ddf_fit:
        # the reference's iteration also fits the DDF to the scene (fit_visibility_field=True, neusky_pipeline.py:272-289): 8 x 128 vMF
        # rays rendered through the SDF field + DDF on those rays, on the multi-view rows and on 256 sky rays, stop_sdf_gradients=False
        from neusky_b200.ddf_fit import DDFFit
        fit = DDFFit(step_mod)
        gs = torch.Generator().manual_seed(300 + rank)
        sky_o = (torch.tensor([0.0, -0.6, 0.1]).expand(256, 3) + 0.1 * torch.randn(256, 3, generator=gs)).to(dev)
        sky_d = torch.nn.functional.normalize(torch.randn(256, 3, generator=gs) + torch.tensor([0.0, 0.0, 1.0]), dim=-1).to(dev)

    # zero-fill + forward + DDF fitting pass + backward as ONE CUDA graph after two eager iterations (neusky_b200/graphed.py); the host's
    # random draws, the h2d of the batch, the gradient all-reduce and the optimizer step stay outside it.  --no-train-graph: all eager.
    from neusky_b200.graphed import GraphedTrainIteration
    iteration = GraphedTrainIteration(step_mod, red, opt, fit=fit, graph=not args.no_train_graph, overlap_fit=not args.no_fit_overlap)

    def one_step():
        i = it["i"]; it["i"] += 1
        dirs = torch.from_numpy((base_dirs @ rots[i % len(rots)]).astype(np.float32))          # IcosahedronSampler random rotation (host, :339-341)
        gp = gpos0 + (torch.rand(gpos0.shape, generator=gg) - 0.5) * (2.0 / gres)
        gd = torch.nn.functional.normalize(torch.randn(gpos0.shape, generator=gg), dim=-1)
        loss = iteration(batch_h, dirs, gp, gd, sky_o if fit is not None else None, sky_d if fit is not None else None)   # h2d of the ray batch inside
        Dp_seen.append(int(step_mod.dirs_sel.shape[0]))
        loss_h.copy_(loss.reshape(1), non_blocking=True)                                        # d2h of the step's result

    for _ in range(max(args.warmup, 0 if args.no_train_graph else 3)):      # two eager iterations + the capture come before the timed region
        one_step()
    ctx.barrier()
    replays0, eager0 = iteration.replays, iteration.eager_steps
    sampler = ClockSampler(local) if (rank == 0 and getattr(args, "sample_clocks", True)) else None
    if sampler:
        sampler.start()
    Dp_seen.clear()
    l0 = _lib.launches
    wall0 = time.perf_counter()
    evs = []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); one_step(); e1.record()
        evs.append((e0, e1))
    host_launch = time.perf_counter() - wall0        # the loop never synchronises: time until the last launch of the last step was queued
    torch.cuda.synchronize()
    t = sum(a.elapsed_time(b) for a, b in evs) / 1e3
    ctx.barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if sampler else None
    launches = _lib.launches - l0
    t, wall = ctx.max_over_ranks(t, wall)
    if os.environ.get("NSK_TRAIN_PROFILE") and rank == 0:
        # diagnostics: per-kernel device time of two more iterations (graph replays) through CUPTI, NOT part of any reported number
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            one_step(); one_step()
            torch.cuda.synchronize()
        with open(os.environ["NSK_TRAIN_PROFILE"], "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
    # exposed all-reduce time: the same steps with the collective switched off (every rank keeps its local gradients; identical compute)
    exposed_ms = None
    if world > 1:
        red.on = False
        n_off = max(2, args.steps // 2)
        one_step()
        ctx.barrier()
        ev2 = []
        for _ in range(n_off):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); one_step(); e1.record()
            ev2.append((e0, e1))
        torch.cuda.synchronize()
        t_off = ctx.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev2) / 1e3)
        red.on = True
        exposed_ms = 1e3 * (t / args.steps - t_off / n_off)
    line = None
    if rank == 0:
        peaks = _peaks()
        Dp = sum(Dp_seen) / max(1, len(Dp_seen))
        # algorithmic FLOP of one iteration, forward figures of SURVEY 8(d) x 3 (forward + two backward contractions):
        flop = 3.0 * (R * Dp * (FLOP_PER_PAIR + 2 * 149_504) + R * S * 881_664 + gres**3 * 2 * 2 * 149_504)
        if fit is not None:      # fitting pass: 1024 rays x S samples through the geometry network with normals, 2304 DDF rows, 1024 sdf_at_termination rows
            flop += 3.0 * (1024 * S * 4 * 149_504 + 2304 * FLOP_PER_PAIR + 1024 * 2 * 149_504)
        h2d = sum(v.numel() * v.element_size() for v in batch_h.values()) + 2 * gpos0.numel() * 4 + 642 * 3 * 4
        line = {"metric": "training rays/s (forward + losses + backward + gradient all-reduce + Adam)", "value": world * R * args.steps / t, "unit": UNIT,
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "wall_ms_per_step": 1e3 * wall / args.steps,
                "host_launch_ms_per_step": 1e3 * host_launch / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": f"tf32 operands (3xTF32 on the SDF geometry network, split={args.split} elsewhere"
                         + (f", DDF backward contractions split={args.ddf_split_bwd}" if args.ddf_split_bwd else "") + "), fp32 accumulate / activations / gradients", "data": "synthetic",
                "config": {"workload": f"BASELINE.json configs[3]: training step, {R} rays/GPU from {K} cameras, S={S} {'proposal-network (256->96->' + str(S) + ', interlevel loss)' if prop is not None else 'uniform'} samples/ray, 642-direction icosphere with a random "
                                       f"rotation per step (mean D'={Dp:.0f} through the DDF), sdf_at_termination branch, hashgrid density loss on {gres**3} grid points, "
                                       + ("DDF fitting pass (8 x 128 vMF rays rendered through the SDF field, DDF on 1024 + 1024 multi-view + 256 sky rows, gradients into both fields), " if fit is not None else "no DDF fitting pass, ")
                                       + 
                                       f"hash tables 2 x 2^19 x 16 x 2 fp32", "rays_per_gpu": R, "samples_per_ray": S,
                           "parallelism": f"data-parallel x{world}, bucketed gradient all-reduce ({red.bytes_per_step / 2**20:.0f} MiB/step, {len(red.buckets)} buckets)"
                                          + (" over NCCL" if world > 1 else " (single rank: no collective)"),
                           "l2": "per-step activations (~15 GB) exceed L2"},
                "e2e": {"value": world * R * args.steps / wall, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "api": "neusky_b200.graphed.GraphedTrainIteration (train.NeuSkyTrainStep + ddf_fit.DDFFit + parallel.GradBucketReducer + fused Adam); e2e value is wall-clock over the timed steps (host random draws, h2d of the ray batch, d2h of the loss included); `value` is CUDA-event time of the same steps"},
                "cuda_graph": {"enabled": not args.no_train_graph, "replayed_steps": iteration.replays - replays0, "eager_steps": iteration.eager_steps - eager0,
                               "abi_kernels_in_graph": iteration.kernels_in_graph, "captures": iteration.captures},
                "gpu_launches": launches, "algorithmic_tflops": flop * args.steps / t / 1e12, "tf32_peak_tflops_sustained": peaks["tf_sustained"] / 2,
                "all_reduce_exposed_ms": exposed_ms, "all_reduce_bytes_per_step": red.bytes_per_step if world > 1 else 0,
                "clocks": clocks, "loss": float(loss_h.item())}
    return line


# ------------------------------------------------------------------------------------------ fp32-parity figures for K4
def fp32_parity_roofline(dev):
    """The same (point, direction) pairs through the paths that meet north_star's 1e-3 on per-pair visibility WHATEVER the weights
    (the default fp16-operand kernel meets it on the reference's initialisation, 4.4e-4 measured, and is at 3e-3 under the x8
    stress gain of the goldens): the exact fp32 CUDA-core kernel, and the layer-wise 3xTF32 tcgen05 chain (fp32-accurate tensor-core
    path, unfused: activations round-trip HBM).  Algorithmic FLOP per pair as for the headline; fractions against the same
    sustained dense fp16/bf16 peak (a 3xTF32 contraction issues three half-rate MMAs per product: 1/6 of that peak at best)."""
    from neusky_b200 import init as nb_init, train as T
    from neusky_b200.render import SkyShader

    peaks = _peaks()
    ddf = nb_init.init_ddf_params(SEED_W)
    sh = SkyShader(ddf, None, device=dev)
    sh.set_directions(_equirect_directions(WIDTH))
    Dp = int(sh.mask.sum())
    out = {"peak": peaks["tf_sustained"], "unit": "TFLOP/s", "flop_per_pair": FLOP_PER_PAIR}

    def run(fn, pairs, iters=3):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        ach = pairs * FLOP_PER_PAIR / (ms * 1e-3) / 1e12
        return {"achieved": ach, "frac": ach / peaks["tf_sustained"], "ms_per_launch": ms, "pairs_per_launch": pairs}

    P = 16384
    pts, nrm, alb = (t.to(dev) for t in _inputs(P, SEED_P))
    rad = torch.ones(1, sh.dirs.shape[0], 3, device=dev)
    out["simt_fp32"] = {"kernel": "sky_shade_simt_kernel (fp32 FMA, fused)", "visibility_err_vs_reference": "<= 5e-4 (tests/test_gpu_parity.py)",
                        **run(lambda: sh.shade(pts, nrm.reshape(P, 1, 3), alb.reshape(P, 1, 3), rad, impl="simt"), P * Dp)}
    P2 = 512      # 524k rows: the [rows, 2560] FiLM tensor of the chain is 5.4 GB in fp32
    p_dev = {k: v.to(dev) for k, v in ddf.items()}
    cfg = T.DDFConfig(scalings=sh.scalings, log2_T=19, radius=1.0, sigmoid_scale=25.0, split=3)
    thr = torch.tensor(0.1, device=dev)
    mlp = T.ddf_param_list(p_dev)

    def chain():
        with torch.no_grad():
            T.ddf_visibility(cfg, pts[:P2].contiguous(), sh.dirs_sel, thr, p_dev["position_encoding.hash_table"], p_dev["ddf.final_layer.weight"], p_dev["ddf.final_layer.bias"], mlp)

    out["tc_3xtf32_chain"] = {"kernel": "gemm_tf32_kernel<3> x 11 + film_sin / ddf_head (tcgen05 kind::tf32, 3xTF32, layer-wise)",
                              "visibility_err_vs_reference": "<= 2e-4 (tests/test_gpu_train.py, split=3)", **run(chain, P2 * Dp)}
    return out


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config-3 / config-4 runs (extra.eval, extra.train) after the headline")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel roofline microbenchmarks (kernels[], roofline_fp32_parity)")
    ap.add_argument("--extra-steps", type=int, default=3, help="timed frames of the config-3 extra (the config-4 extra times twice as many steps)")
    ap.add_argument("--workload", default="shade", choices=["shade", "eval", "train", "relight"],
                    help="shade = BASELINE.json configs[1] (the headline line); eval = configs[2] full-image render, train = configs[3] training step (supplementary lines)")
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--samples", type=int, default=128)
    ap.add_argument("--tile", type=int, default=16384)
    ap.add_argument("--sampler", default="uniform", choices=["uniform", "proposal"], help="eval: sample placement (proposal = shipped NeuS-facto default, use --samples 48)")
    ap.add_argument("--latents", type=int, default=64, help="relight: latent codes per sweep")
    ap.add_argument("--per-sample-cache", action="store_true", help="relight: keep the per-sample cache instead of the collapsed [R,D,3] coefficients")
    ap.add_argument("--fp32-cache", action="store_true", help="relight: the round-1 collapsed fp32 [R,D,3] cache instead of the compact fp16 one")
    ap.add_argument("--rays", type=int, default=1024, help="train: rays per GPU")
    ap.add_argument("--train-samples", type=int, default=48)
    ap.add_argument("--split", type=int, default=3, choices=[1, 3], help="train: 1 = tf32 GEMMs, 3 = 3xTF32 (fp32-accurate)")
    ap.add_argument("--ddf-split-bwd", type=int, default=0, choices=[0, 1, 3], help="train: precision of the DDF's backward contractions only (0 = same as --split)")
    ap.add_argument("--train-sampler", default="proposal", choices=["proposal", "uniform"], help="train: sample placement (proposal = shipped NeuS-facto default)")
    ap.add_argument("--no-ddf-fit", action="store_true", help="train: leave the DDF fitting pass (fit_visibility_field=True in the reference) out of the step")
    ap.add_argument("--no-fit-overlap", action="store_true", help="train: run the DDF fitting pass after the main pass on the same stream instead of as a parallel branch")
    ap.add_argument("--no-train-graph", action="store_true", help="train: run every iteration eagerly (~2900 launches from the host) instead of replaying the captured CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return
    ctx = Ctx()
    if args.workload != "shade":
        line = {"eval": eval_arm, "relight": relight_arm, "train": train_arm}[args.workload](args, ctx)
        if line is not None:
            print(json.dumps(line), flush=True)
        ctx.close()
        return
    rank, local, world, dev, dist = ctx.rank, ctx.local, ctx.world, ctx.dev, ctx.dist

    from neusky_b200 import _lib, init as nb_init
    from neusky_b200.render import SkyShader

    _lib.load()
    shader = SkyShader(nb_init.init_ddf_params(SEED_W), nb_init.init_reni_params(SEED_W + 1), device=dev)
    shader.set_directions(_equirect_directions(WIDTH))
    Dp = int(shader.mask.sum())
    P = args.points
    pts_h, nrm_h, alb_h = (t.pin_memory() for t in _inputs(P, SEED_P + 10 * rank))
    pts, nrm, alb = (t.to(dev) for t in (pts_h, nrm_h, alb_h))
    Z, sc = (t.to(dev) for t in _latents())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    barrier, max_over_ranks = ctx.barrier, ctx.max_over_ranks

    def timed(fn, steps):
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / 1e3

    # ---- device-resident throughput --------------------------------------------------------------
    step = lambda: shader.shade_points(pts, nrm, alb, Z, sc)
    for _ in range(args.warmup):
        flush.fill_(1); step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    shader.k4_events = []
    l0 = _lib.launches
    wall0 = time.perf_counter()
    t_dev = timed(step, args.steps)
    barrier()
    wall = time.perf_counter() - wall0
    launches = _lib.launches - l0
    k4_s = sum(a.elapsed_time(b) for a, b in shader.k4_events) / 1e3 / max(1, len(shader.k4_events))
    shader.k4_events = None
    clocks = sampler.stop() if sampler else None
    t_dev = max_over_ranks(t_dev)

    # ---- end to end from pinned host buffers -----------------------------------------------------
    out_h = torch.empty(P, 3, dtype=torch.float32).pin_memory()
    e2e_step = lambda: shader.shade_points_host(pts_h, nrm_h, alb_h, Z, sc, out_h=out_h)
    e2e_steps = args.steps                 # the same K steps as the device-resident figure
    e2e_step()
    barrier()
    t_e2e = max_over_ranks(timed(e2e_step, e2e_steps))
    barrier()

    if rank == 0:
        peaks = _peaks()
        pairs = P * Dp
        achieved = pairs * FLOP_PER_PAIR / k4_s / 1e12
        peak = peaks["tf_sustained"]       # K4 runs for seconds inside the step: the sustained (power-capped) figure applies
        line = {
            "metric": METRIC, "value": world * P * args.steps / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16xf16->f32",
            "data": "synthetic", "config": _config(args, world),
            "e2e": {"value": world * P * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": 3 * P * 12, "d2h_bytes_per_step": P * 12,
                    "api": "neusky_b200.render.SkyShader.shade_points_host", "steps": e2e_steps},
            "gpu_launches": launches,
            "roofline": {"kernel": "sky_shade_tc2_kernel", "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": _ncu_traffic(pairs), "traffic_source": "committed ncu capture of this command, not measured in the run: profiles/k4_ncu_summary.json (r02_k4_tc2_dram_full_final.csv)",
                         "peak_source": f"{peaks['source']} dense bf16/fp16 cuBLAS, sustained ({peaks['tf_burst']:.0f} burst)",
                         "frac_of_burst": achieved / peaks["tf_burst"], "pairs_per_launch": pairs, "flop_per_pair": FLOP_PER_PAIR,
                         "ms_per_launch": 1e3 * k4_s, "k4_share_of_step": k4_s * args.steps / t_dev if world == 1 else None},
            "clocks": clocks, "wall_s_timed_region": wall,
        }
        if not args.no_cpu_baseline and world == 1:
            r, cores, desc, t = cpu_oracle_rate()
            line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc, "seconds": t}
    else:
        line = None
    del pts, nrm, alb, pts_h, nrm_h, alb_h, out_h, shader
    torch.cuda.empty_cache()

    # ---- the splits that carry a collective, under the same clock: config 3 (eval frame, ray tiles over the ranks +
    # ---- all-gather of the outputs, strong scaling) and config 4 (training step, data-parallel gradient all-reduce, weak scaling) ---------
    if not args.no_extras:
        import copy

        ea = copy.copy(args); ea.steps, ea.warmup = args.extra_steps, 2
        ev = eval_arm(ea, ctx)
        torch.cuda.empty_cache()
        ta = copy.copy(args); ta.steps, ta.warmup, ta.sample_clocks = max(4, 2 * args.extra_steps), 3, False
        tr = train_arm(ta, ctx)
        torch.cuda.empty_cache()
        # the same iteration with the DDF's backward contractions in single-pass tf32 (forward and every loss value stay 3xTF32):
        # stated separately, tolerance in tests/test_gpu_train.py
        tb = copy.copy(ta); tb.ddf_split_bwd = 1
        trb = train_arm(tb, ctx)
        torch.cuda.empty_cache()
        if line is not None:
            keep = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "gpu_launches")
            line["extra"] = {
                "eval": {**{k: ev[k] for k in keep}, "workload": ev["config"]["workload"], "gather_ms_per_step": ev["gather_ms_per_step"],
                         "gather_bytes_per_step": ev["gather_bytes_per_step"], "collective": ev["collective"]},
                "train": {**{k: tr[k] for k in keep}, "workload": tr["config"]["workload"], "parallelism": tr["config"]["parallelism"],
                          "all_reduce_exposed_ms": tr["all_reduce_exposed_ms"], "all_reduce_bytes_per_step": tr["all_reduce_bytes_per_step"],
                          "algorithmic_tflops": tr["algorithmic_tflops"], "wall_ms_per_step": tr["wall_ms_per_step"],
                          "tf32_ddf_backward": {"value": trb["value"], "ms_per_step": trb["ms_per_step"], "dtype": trb["dtype"],
                                                "all_reduce_exposed_ms": trb["all_reduce_exposed_ms"]}},
            }
    # ---- per-kernel rooflines of the rest of the path (K1, K2, K3, RENI++, proposal sampler): rank 0, the other ranks wait at the barrier ----
    if not args.no_kernels:
        if rank == 0:
            from bench_kernels import kernel_rooflines

            line["kernels"] = kernel_rooflines(dev, _peaks(), scale=0.25)
            line["roofline_fp32_parity"] = fp32_parity_roofline(dev)
        barrier()
    if line is not None:
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
