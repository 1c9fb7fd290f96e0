import sys, ctypes, torch
sys.path.insert(0, ".")
from neusky_b200 import init as nb_init, _lib, packing
from neusky_b200.render import SkyShader
from oracle import neusky_oracle as O
dev = torch.device("cuda:0")
p = nb_init.init_ddf_params(21, final_gain=8.0)
g = torch.Generator().manual_seed(3)
R, D = 4, 40
pts = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1) * 0.5
normals = torch.nn.functional.normalize(torch.randn(R, 1, 3, generator=g), dim=-1)
wa = torch.rand(R, 1, 3, generator=g)
dirs = torch.nn.functional.normalize(torch.randn(D, 3, generator=g).abs(), dim=-1)
rad = torch.ones(1, D, 3)
sh = SkyShader(p, None, device=dev); sh.set_directions(dirs)
a = (pts.to(dev), normals.to(dev), wa.to(dev), rad.to(dev))
dump = torch.full((10, 128, 256), float("nan"), device=dev)
lib = _lib.load()
lib.nsk_debug_set_tc_dump(ctypes.c_void_p(dump.data_ptr()))
out = sh.shade(*a, want_vis=True, want_ddf=True, impl="tc")
torch.cuda.synchronize()
lib.nsk_debug_set_tc_dump(ctypes.c_void_p(0))
ref = sh.shade(*a, want_vis=True, want_ddf=True, impl="simt")
# host reference activations for the first 128 pairs
Dp = int(sh.mask.sum()); n = min(128, R * Dp)
pos = pts[:, None, :].expand(R, Dp, 3).reshape(-1, 3)[:n]
dd = sh.dirs_sel.cpu()[None].expand(R, Dp, 3).reshape(-1, 3)[:n]
q = O.ray_sphere_intersection(pos, dd, 1.0)
dl = O.ddf_local_directions(q, -dd)
cond = torch.cat([q, O.hash_encode(q, p["position_encoding.hash_table"], O.hash_scalings(), 19)], -1)
x = torch.cat([dl, O.nerf_encode(dl, 2, 0.0, 2.0, False)], -1)
h = cond; acts = []
for i in range(5):
    h = torch.nn.functional.leaky_relu(h @ p[f"ddf.mapping_network.network.{2*i}.weight"].T + p[f"ddf.mapping_network.network.{2*i}.bias"], 0.2); acts.append(h)
fp = h @ p["ddf.mapping_network.network.10.weight"].T + p["ddf.mapping_network.network.10.bias"]
freq, phase = fp[:, :1280] * 15 + 30, fp[:, 1280:]
for l in range(5):
    x = torch.sin(freq[:, l*256:(l+1)*256] * (x @ p[f"ddf.net.{l}.layer.weight"].T + p[f"ddf.net.{l}.layer.bias"]) + phase[:, l*256:(l+1)*256]); acts.append(x)
d = dump.cpu()
for k in range(10):
    e = (d[k, :n] - acts[k]).abs()
    print(f"stage {k} ({'map' if k<5 else 'trunk'} {k%5}): max err {e.max().item():.4g}, nan rows {int(torch.isnan(d[k,:n]).any(1).sum())}, ref absmax {acts[k].abs().max().item():.3g}")
    if not (e.max() < 0.05):
        bad = torch.nonzero(~(e < 0.05)); print("   first bad (row,col):", bad[:6].tolist(), " bad cols per row0:", int((~(e[0] < 0.05)).sum()))
        break
print("max ddf err", (ref["expected_termination_dist"] - out["expected_termination_dist"]).abs().max().item())
