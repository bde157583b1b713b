"""Time field fwd(stash) and bwd separately (CUDA events): python scripts/perf_bwd.py [mode] [rays] [N]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_sandbox_b200 as nsb
from nerf_sandbox_b200 import _lib
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
N = int(sys.argv[3]) if len(sys.argv) > 3 else 192
dev = "cuda"; L = _lib.lib()
net = nsb.NeRF(63, 27, mode=mode).to(dev)
o = torch.randn(B, 3, device=dev); d = torch.nn.functional.normalize(torch.randn(B, 3, device=dev), dim=-1)
z = torch.sort(torch.rand(B, N, device=dev) * 4 + 2, -1).values.contiguous(); rn = torch.ones(B, device=dev)
Q = B * N
wsb = L.nsb_field_workspace_bytes(Q, net.mode, 1); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
raw = torch.empty(Q, 4, device=dev); d_raw = torch.randn(Q, 4, device=dev) * 1e-3; g = torch.zeros(_lib.N_PARAMS, device=dev)
pk = net.packed(); st = _lib.stream()
fwd = lambda: _lib.check(L.nsb_field_fwd_rays(_lib.ptr(o), _lib.ptr(d), _lib.ptr(z), _lib.ptr(rn), _lib.ptr(d), _lib.ptr(pk), _lib.ptr(raw), _lib.ptr(ws), wsb, B, N, net.mode, 1, st))
bwd = lambda: _lib.check(L.nsb_field_bwd(_lib.ptr(d_raw), _lib.ptr(pk), _lib.ptr(g), _lib.ptr(ws), wsb, Q, net.mode, st))
def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
tf, tb = timeit(fwd), timeit(bwd)
print(json.dumps({"mode": mode, "points": Q, "fwd_ms": tf, "bwd_ms": tb, "fwd_TF": Q * 1186816 / tf / 1e9, "bwd_TF": Q * 2302208 / tb / 1e9,
                  "ws_GB": wsb / 1e9}))
