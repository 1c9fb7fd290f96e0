"""Per-warp-role cycle accounting of the K4 tensor-core kernel (diagnostic; uses the nsk_debug_set_tc_prof hook)."""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from neusky_b200 import _lib, init as nb_init
from neusky_b200.render import SkyShader
from bench import _equirect_directions, _inputs

R = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
impl = sys.argv[2] if len(sys.argv) > 2 else "tc"
dev = torch.device("cuda:0")
sh = SkyShader(nb_init.init_ddf_params(0), None, device=dev, impl=impl)
setter = lib_setter = None
sh.set_directions(_equirect_directions(64))
pts, nrm, alb = (t.to(dev) for t in _inputs(R, 1))
rad = torch.rand(1, 2048, 3, device=dev)
lib = _lib.load()
setter = lib.nsk_debug_set_tc_prof if impl == "tc" else lib.nsk_debug_set_tc2_prof
prof = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
for it in range(2):
    setter(ctypes.c_void_p(prof.data_ptr() if it == 1 else 0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sh.shade(pts, nrm[:, None], alb[:, None], rad); e1.record()
    torch.cuda.synchronize()
    print(f"{impl} run {it}: {e0.elapsed_time(e1):.2f} ms, {R*1024/e0.elapsed_time(e1)/1e3:.1f} M pairs/s")
setter(ctypes.c_void_p(0))
p = prof.view(148, 16).double().cpu()
tiles = (R * 1024 + 127) // 128 / 148
names = ["producer total", "producer wait ring-empty", "mma total", "mma wait dependency(all)", "mma wait weights", "mma wait dependency(mapping ops)",
         "epi total", "epi wait mapping acc", "epi wait Z", "epi wait FP", "epi tail", "prologue total", "prologue wait in-empty",
         "mma issue: mapping ops (71 MMAs N=256)", "mma issue: FiLM chunks (340 MMAs N=128)", "mma issue: trunk Z (67 MMAs N=256)"]
lead = p[0::2] if impl == "tc2" else p      # tc2: only the leader CTA of a pair issues MMAs (the peer's issuer slots stay 0)
for i, n in enumerate(names):
    src = lead if n.startswith("mma") else p
    print(f"{n:44s} {src[:, i].mean():14.0f} cycles/CTA   {src[:, i].mean()/tiles:10.0f} cycles/tile")
