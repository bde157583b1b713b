"""Time the three roles of the split GEMM at the fine-pass size: python scripts/perf_split_gemm.py [points]"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_sandbox_b200 import _lib
L = _lib.lib(); fn = L.nsb_debug_split_gemm; fn.restype = C.c_int
p, i64, i32 = C.c_void_p, C.c_int64, C.c_int
fn.argtypes = [p, i64, p, i64, p, i64, i64, i64, i64, i32, p, i32, p, i64, p, i64, i32, p, p]
dev = torch.device("cuda", 0)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 196608
X = torch.randn(P, 256, device=dev); W = torch.randn(256, 256, device=dev) * 0.05; b = torch.randn(256, device=dev)
Y = torch.empty(P, 256, device=dev); dY = torch.randn(P, 256, device=dev); dX = torch.empty(P, 256, device=dev); gW = torch.zeros(256, 256, device=dev)
st = _lib.stream()
calls = {
    "fwd": lambda: fn(_lib.ptr(X), 256, _lib.ptr(W), 256, _lib.ptr(Y), 256, P, 256, 256, 0, _lib.ptr(b), 1, None, 0, None, 0, 0, None, st),
    "dgrad": lambda: fn(_lib.ptr(dY), 256, _lib.ptr(W), 256, _lib.ptr(dX), 256, P, 256, 256, 1, None, 0, _lib.ptr(Y), 256, None, 0, 0, None, st),
    "wgrad": lambda: fn(_lib.ptr(dY), 256, _lib.ptr(X), 256, _lib.ptr(gW), 256, 256, 256, P, 2, None, 0, None, 0, None, 0, 256, None, st),
}
flop = 2.0 * P * 256 * 256
for name, f in calls.items():
    for _ in range(3): _lib.check(f(), name)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:6s} {P} x 256 x 256: {ms*1e3:8.1f} us = {flop/ms/1e9:6.1f} TFLOP/s fp32-equivalent; fp32 bytes moved (A + C, once) {P*256*8/ms/1e6:.0f} GB/s")
