"""Per-role cycle accounting of the fused RENI++ row kernel (PROF instantiation): where each warp role of a CTA spends its time."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neusky_b200 import _lib, init as nb_init, ops, packing
dev = torch.device("cuda:0")
lib = _lib.load()
p = nb_init.init_reni_params(1)
blob, fused = packing.pack_reni(p, device=dev), packing.pack_reni_fused(p, device=dev)
gen = torch.Generator().manual_seed(0)
N = 148 * 256 * 8
d = torch.nn.functional.normalize(torch.randn(N, 3, generator=gen), dim=-1).to(dev)
Z1, s1 = torch.randn(1, 100, 3, generator=gen).to(dev), torch.zeros(1, device=dev)
ops.reni_rows_fused(d, Z1, s1, blob, fused); torch.cuda.synchronize()
buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
lib.nsk_debug_set_reni_fused_prof(ctypes.c_void_p(buf.data_ptr()))
ops.reni_rows_fused(d, Z1, s1, blob, fused); torch.cuda.synchronize()
lib.nsk_debug_set_reni_fused_prof(ctypes.c_void_p(0))
m = buf.view(148, 16).double().mean(0) / 8.0     # cycles per tile pair
names = ["issuer total", "issuer wait PE (first layer)", "issuer wait relu epilogue (before F2)", "issuer wait LayerNorm epilogue (before F0)", "issuer wait weights", "issuer wait X free",
         "PE total", "PE wait buffer free", "epilogue tile0 total", "epilogue tile0 wait accumulator", "epilogue tile1 total", "epilogue tile1 wait accumulator"]
for n, v in zip(names, m.tolist()):
    print(f"{n:45s} {v:10.0f} cycles / pair")
