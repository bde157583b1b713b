"""Accuracy anatomy of the bf16-split tensor-core GEMM (csrc/field_split.cu): error and its SIGN against fp64 as a function of
the contraction length, for same-sign and mixed-sign operands.  python scripts/dbg_split_gemm.py"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_sandbox_b200 import _lib
L = _lib.lib(); fn = L.nsb_debug_split_gemm; fn.restype = C.c_int
p, i64, i32 = C.c_void_p, C.c_int64, C.c_int
fn.argtypes = [p, i64, p, i64, p, i64, i64, i64, i64, i32, p, i32, p, i64, p, i64, i32, p, p]
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device=dev).manual_seed(0)
M, N = 1024, 256
b = torch.zeros(N, device=dev)
for pos in (True, False):
    for K in (64, 256, 1024, 4096):
        X = torch.rand(M, K, device=dev, generator=g) if pos else torch.randn(M, K, device=dev, generator=g)
        W = torch.rand(N, K, device=dev, generator=g) if pos else torch.randn(N, K, device=dev, generator=g)
        Y = torch.empty(M, N, device=dev)
        _lib.check(fn(_lib.ptr(X), K, _lib.ptr(W), K, _lib.ptr(Y), N, M, N, K, 0, _lib.ptr(b), 0, None, 0, None, 0, 0, None, _lib.stream()), "gemm")
        ref = X.double() @ W.double().T
        y32 = X @ W.T
        scale = (X.double().abs() @ W.double().abs().T)           # sum of |terms|: the natural error scale
        e = (Y.double() - ref) / scale; e32 = (y32.double() - ref) / scale
        print(f"{'same-sign' if pos else 'mixed    '} K={K:5d}: split mean {float(e.mean()):+.2e} rms {float(e.pow(2).mean().sqrt()):.2e} max {float(e.abs().max()):.2e}"
              f" | cuBLAS fp32 mean {float(e32.mean()):+.2e} rms {float(e32.pow(2).mean().sqrt()):.2e} max {float(e32.abs().max()):.2e}  (units of sum|terms|)")
