"""Quick timing of the tensor-core shading kernel (not the official bench)."""
import sys, time
import torch
sys.path.insert(0, ".")
from neusky_b200 import init as nb_init
from neusky_b200.render import SkyShader

R = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
impl = sys.argv[3] if len(sys.argv) > 3 else "tc"
dev = torch.device("cuda:0")
p = nb_init.init_ddf_params(0)
g = torch.Generator().manual_seed(1)
pts = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1) * torch.rand(R, 1, generator=g) ** (1 / 3) * 0.95
normals = torch.nn.functional.normalize(torch.randn(R, 1, 3, generator=g), dim=-1)
wa = torch.rand(R, 1, 3, generator=g)
from oracle.neusky_oracle import equirect_directions
dirs = equirect_directions(64) if D == 2048 else torch.nn.functional.normalize(torch.randn(D, 3, generator=g), dim=-1)
rad = torch.exp(torch.randn(1, dirs.shape[0], 3, generator=g))
sh = SkyShader(p, None, device=dev, impl=impl)
sh.set_directions(dirs)
a = (pts.to(dev), normals.to(dev), wa.to(dev), rad.to(dev))
Dp = int(sh.mask.sum())
for _ in range(2):
    sh.shade(*a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3
for _ in range(n):
    sh.shade(*a)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
pairs = R * Dp
print(f"impl={impl} R={R} D={dirs.shape[0]} Dp={Dp}: {ms:.2f} ms/step, {pairs/(ms*1e-3)/1e6:.1f} M pairs/s, {pairs*2385408/(ms*1e-3)/1e12:.0f} TFLOP/s algorithmic, {R/ms*1e3:.0f} points/s")
