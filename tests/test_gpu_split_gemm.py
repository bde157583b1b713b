"""The bf16-split tensor-core GEMM of the fp32 mode (csrc/field_split.cu) against float64 matmuls: the three roles with
their epilogues, ragged row counts, the padded / partial column tiles of the real layers, tiny gradients."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

FWD, DGRAD, WGRAD = 0, 1, 2


def _fn():
    from nerf_sandbox_b200 import _lib
    L = _lib.lib()
    fn = L.nsb_debug_split_gemm
    fn.restype = C.c_int
    p, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    fn.argtypes = [p, i64, p, i64, p, i64, i64, i64, i64, i32, p, i32, p, i64, p, i64, i32, p, p]
    return fn, _lib


def _rel(a, b):
    return float((a.double() - b).norm() / b.norm().clamp_min(1e-300))


@pytest.mark.parametrize("M,N,K", [(1000, 256, 320), (128, 128, 288), (257, 256, 64), (4096, 256, 256)])
def test_forward_role(M, N, K):
    fn, _lib = _fn()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(M + K)
    X = torch.randn(M, K, device=dev, generator=g)
    W = torch.randn(N, K, device=dev, generator=g) * 0.1
    b = torch.randn(N, device=dev, generator=g)
    for relu in (0, 1):
        Y = torch.full((M, N), float("nan"), device=dev)
        _lib.check(fn(_lib.ptr(X), K, _lib.ptr(W), K, _lib.ptr(Y), N, M, N, K, FWD, _lib.ptr(b), relu, None, 0, None, 0, 0, None, _lib.stream()),
                   "split_gemm fwd")
        ref = X.double() @ W.double().T + b.double()
        if relu:
            ref = ref.clamp_min(0)
        ref32 = torch.addmm(b, X, W.T)
        if relu:
            ref32 = ref32.clamp_min(0)
        e, e32 = _rel(Y, ref), _rel(ref32, ref)
        assert torch.isfinite(Y).all()
        assert e < 2e-6, (M, N, K, relu, e, e32)      # fp32-grade: the cuBLAS fp32 product of the same operands sits at ~1e-7..1e-6


def test_dgrad_role_mask_addend_and_tiny_gradients():
    fn, _lib = _fn()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(7)
    M, n_out, kpad = 777, 256, 320
    for scale in (1.0, 1e-9):          # back-propagated gradients are tiny: bf16 terms keep fp32's exponent range
        dY = torch.randn(M, n_out, device=dev, generator=g) * scale
        W = torch.randn(n_out, kpad, device=dev, generator=g) * 0.1
        mask = torch.randn(M, 300, device=dev, generator=g)           # ldm != N
        add = torch.randn(M, 256, device=dev, generator=g) * scale
        dX = torch.full((M, 256), float("nan"), device=dev)
        _lib.check(fn(_lib.ptr(dY), n_out, _lib.ptr(W), kpad, _lib.ptr(dX), 256, M, 256, n_out, DGRAD, None, 0, _lib.ptr(mask), 300,
                      _lib.ptr(add), 256, 0, None, _lib.stream()), "split_gemm dgrad")
        ref = (dY.double() @ W.double()[:, :256] + add.double()) * (mask[:, :256] > 0)
        assert _rel(dX, ref) < 2e-6, (scale, _rel(dX, ref))
        # no mask, no addend, n_out = 128 (color_fc)
        dC = torch.randn(M, 128, device=dev, generator=g) * scale
        Wc = torch.randn(128, 288, device=dev, generator=g) * 0.1
        dF = torch.full((M, 256), float("nan"), device=dev)
        _lib.check(fn(_lib.ptr(dC), 128, _lib.ptr(Wc), 288, _lib.ptr(dF), 256, M, 256, 128, DGRAD, None, 0, None, 0, None, 0, 0, None, _lib.stream()),
                   "split_gemm dgrad")
        ref = dC.double() @ Wc.double()[:, :256]
        assert _rel(dF, ref) < 2e-6, (scale, _rel(dF, ref))


@pytest.mark.parametrize("P,n_out,K,Kpad", [(5000, 256, 319, 320), (12345, 128, 283, 288), (300, 256, 63, 64), (40000, 256, 256, 256)])
def test_wgrad_role(P, n_out, K, Kpad):
    fn, _lib = _fn()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(P)
    dY = torch.randn(P, n_out, device=dev, generator=g) * 1e-4
    X = torch.randn(P, Kpad, device=dev, generator=g)
    X[:, K:] = 0
    gW = torch.zeros(n_out, K, device=dev)
    gb = torch.zeros(n_out, device=dev)                        # bias gradient = column sums of dY, accumulated by the same launch
    ref = dY.double().T @ X.double()[:, :K]
    for times in (1, 2):                                       # accumulates INTO the buffer (split-K atomics)
        _lib.check(fn(_lib.ptr(dY), n_out, _lib.ptr(X), Kpad, _lib.ptr(gW), K, n_out, Kpad, P, WGRAD, None, 0, None, 0, None, 0, K,
                      _lib.ptr(gb), _lib.stream()), "split_gemm wgrad")
        assert _rel(gb, times * dY.double().sum(0)) < 1e-5
        assert _rel(gW, times * ref) < 1e-5, (times, _rel(gW, times * ref))    # 40,000 random-sign terms per element: fp32 accumulation noise


def test_error_anatomy_matches_an_fp32_gemm():
    """What DESIGN section 2 claims about the split product, as assertions: in units of sum|terms| per output element, mixed-sign
    operands land at the error of cuBLAS's fp32 GEMM, and the truncating tensor-core accumulate shows up only as a small negative
    bias on same-sign sums (x0 * y0 alone in the main accumulator: K / 16 truncations, not 6 K / 16)."""
    fn, _lib = _fn()
    dev = torch.device("cuda", 0)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        g = torch.Generator(device=dev).manual_seed(0)
        M, N, K = 1024, 256, 256
        b = torch.zeros(N, device=dev)
        for same_sign in (False, True):
            X = torch.rand(M, K, device=dev, generator=g) if same_sign else torch.randn(M, K, device=dev, generator=g)
            W = torch.rand(N, K, device=dev, generator=g) if same_sign else torch.randn(N, K, device=dev, generator=g)
            Y = torch.empty(M, N, device=dev)
            _lib.check(fn(_lib.ptr(X), K, _lib.ptr(W), K, _lib.ptr(Y), N, M, N, K, FWD, _lib.ptr(b), 0, None, 0, None, 0, 0, None, _lib.stream()), "gemm")
            ref = X.double() @ W.double().T
            scale = X.double().abs() @ W.double().abs().T
            e = (Y.double() - ref) / scale
            e32 = ((X @ W.T).double() - ref) / scale
            rms, rms32 = float(e.pow(2).mean().sqrt()), float(e32.pow(2).mean().sqrt())
            if same_sign:
                assert -6e-7 < float(e.mean()) < 0 and rms < 3 * rms32, (float(e.mean()), rms, rms32)     # measured -3.1e-7 / 3.2e-7 vs 2.2e-7
            else:
                assert abs(float(e.mean())) < 1e-9 and rms < 2 * rms32, (float(e.mean()), rms, rms32)     # measured 2.3e-8 vs 2.8e-8
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
