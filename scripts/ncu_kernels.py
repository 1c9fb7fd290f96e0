"""Launch every bandwidth-bound kernel of the path once at benchmark size (after one warm-up launch) so that
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,...` can attribute DRAM traffic per launch.  Sizes match bench_kernels.py.
Prints the algorithmic bytes per launch of each kernel as JSON lines for the summary script."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neusky_b200 import init as nb_init, ops, packing
from neusky_b200 import proposal as P
from neusky_b200.render import sphere_collider
from neusky_b200.samplers import EquirectangularSampler

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
sc = nb_init.hash_scalings().to(dev)
rec = []


def note(kernel, units, bytes_per_unit, what):
    rec.append({"kernel": kernel, "units": units, "algorithmic_bytes_per_launch": units * bytes_per_unit, "what": what})


def twice(fn):
    fn(); torch.cuda.synchronize(); fn(); torch.cuda.synchronize()


n = 8_000_000
table = nb_init.init_hash_table(1).to(dev)
x = torch.rand(n, 3, generator=g).to(dev)
twice(lambda: ops.hash_encode(x, table, sc, 19)); note("hash_encode_fwd_kernel", n, 1164, "K1 forward, uniform random points")
R, S = 62_500, 128
o = torch.tensor([0.5, -0.4, 0.6], device=dev)
d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(dev)
t = torch.linspace(0.0, 0.45, S, device=dev)
xr = (o + d[:, None, :] * t[None, :, None]).reshape(-1, 3).contiguous()
twice(lambda: ops.hash_encode(xr, table, sc, 19)); note("hash_encode_fwd_kernel", xr.shape[0], 1164, "K1 forward, ray-ordered samples")
gout = torch.randn(n, 32, generator=g).to(dev)
gt = torch.zeros_like(table)
twice(lambda: ops.hash_encode_bwd(x, sc, 19, gout, gt)); note("hash_encode_bwd_kernel", n, 12 + 128 + 2048, "K1 backward scatter (RMW counted as 2 x 1024 B)")
del gout, gt, x, xr
R, S = 1_000_000, 128
tt = torch.sort(torch.rand(R, S + 1, generator=g) * 2 + 0.05, dim=1).values.to(dev)
starts, ends = tt[:, :-1].contiguous(), tt[:, 1:].contiguous()
rd = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(dev)
sdf = ((1.0 - (starts + ends) / 2) * 0.3).contiguous()
grad = (-rd[:, None, :] * torch.ones(R, S, 1, device=dev)).contiguous()
alb = torch.rand(R, S, 3, device=dev)
twice(lambda: ops.neus_composite(sdf, grad, alb, rd, starts, ends, ends - starts, torch.ones(R, device=dev), 20.0)); note("neus_composite_fwd_kernel", R, 56 * S + 80, "K3 S=128 (kernel also writes normals / wa for K4: 80 S + 80 B actually moved)")
del tt, starts, ends, rd, sdf, grad, alb
rp = nb_init.init_reni_params(1)
rblob, rgw = packing.pack_reni(rp, device=dev), packing.pack_reni_gemm(rp, device=dev)
Nf = 1280 * 720
rows = torch.nn.functional.normalize(torch.randn(Nf, 3, generator=g), dim=-1).to(dev)
Z1, s1 = torch.randn(1, 100, 3, generator=g).to(dev), torch.zeros(1, device=dev)
twice(lambda: ops.reni_rows_tc(rows, Z1, s1, rblob, rgw)); note("reni rows chain (gemm_tf32_kernel + reni_* kernels)", Nf, 12 + 12, "RENI++ rows, frame-sized; algorithmic bytes = directions in + radiance out")
dirs = EquirectangularSampler(64)().frustums.directions.to(dev)
Z64 = torch.randn(64, 100, 3, generator=g).to(dev)
twice(lambda: ops.reni_radiance_table(dirs, Z64, torch.zeros(64, device=dev), rblob)); note("reni_rows_kernel", 64 * 2048, 12, "RENI++ table K=64 D=2048 (fp32 SIMT)")
del rows
R = 921_600 // 4
c2 = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
o = (torch.tensor([0.0, -0.9, 0.25]) + torch.zeros(R, 3)).to(dev).contiguous()
d = torch.nn.functional.normalize(-o.cpu() + 0.4 * c2, dim=-1).to(dev).contiguous()
near, far = (t.reshape(-1).contiguous() for t in sphere_collider(o, d))
f = P.HashMLPDensityField(nb_init.init_proposal_params(64, table_scale=1.0, density_bias=1.0), 64, device=dev)
bins = P.uniform_bins(R, 256, dev, torch.rand(R, device=dev))
twice(lambda: f.density_on_rays(o, d, near, far, bins)); note("proposal_density_fwd_kernel", R * 256, 372, "P1 S=256 (gathers hit L2: 5 MB table)")
dens = f.density_on_rays(o, d, near, far, bins)
twice(lambda: P.pdf_resample(bins, near, far, 96, density=dens)); note("pdf_resample_kernel", R, (257 + 256 + 256 + 2 * 97) * 4 + 8, "P2 256 -> 96")
# relight pass over the collapsed cache
Rr, D = 131072, 642
H = torch.rand(Rr, D, 3, device=dev)
rad = torch.rand(4, D, 3, device=dev)
twice(lambda: ops.relight_collapsed_multi(H, rad)); note("relight_collapsed_multi_kernel", Rr, D * 12 + 4 * 12, "config-5 pass: 4 latent codes per read of H [R,642,3] fp32")
# round 2: compact cache pass (8 codes per read) and the fused RENI++ row kernel
Rc, DPc = 921_600 // 2, 656
H16 = (torch.rand(Rc, 3 * DPc, device=dev) * 0.9).to(torch.float16)
hs = torch.rand(Rc, device=dev)
rws = torch.arange(Rc, dtype=torch.int32, device=dev)
rad8 = torch.rand(32, D, 3, device=dev)
twice(lambda: ops.relight_h16_multi(H16, hs, rws, Rc, D, rad8)); note("relight_h16_kernel", Rc, 6 * DPc + 8 + 32 * 12, "config-5 pass over the compact cache: 32 latent codes per read of fp16 rows [3][656]")
del H16, hs, rws
rfused = packing.pack_reni_fused(rp, device=dev)
rows = torch.nn.functional.normalize(torch.randn(Nf, 3, generator=g), dim=-1).to(dev)
twice(lambda: ops.reni_rows_fused(rows, Z1, s1, rblob, rfused)); note("reni_rows_fused_kernel", Nf, 24, "RENI++ rows, fused tcgen05 kernel (tensor / epilogue bound; DRAM = directions in + radiance out + 0.5 MB weights)")
for r_ in rec:
    print(json.dumps(r_), flush=True)
