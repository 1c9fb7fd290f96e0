import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neusky_b200 import init as nb_init, ops, packing
dev = torch.device("cuda:0")
p = nb_init.init_reni_params(1)
blob, fused = packing.pack_reni(p, device=dev), packing.pack_reni_fused(p, device=dev)
gen = torch.Generator().manual_seed(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 256 * 6
d = torch.nn.functional.normalize(torch.randn(N, 3, generator=gen), dim=-1).to(dev)
Z1, s1 = torch.randn(1, 100, 3, generator=gen).to(dev), torch.zeros(1, device=dev)
for _ in range(3):
    ops.reni_rows_fused(d, Z1, s1, blob, fused)
torch.cuda.synchronize()
