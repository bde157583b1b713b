"""One pass of the K2 (samplers) / K3 (compositor) kernels at a bandwidth-bound ray count, for `ncu --set full`:
python scripts/k2k3_case.py [rays] [nc] [nf] [reps]   (prints CUDA-event GB/s per kernel against the SURVEY 8d bytes)"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_sandbox_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nf = int(sys.argv[3]) if len(sys.argv) > 3 else 128
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
nt = nc + nf
dev = "cuda"; L = _lib.lib(); st = _lib.stream()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HBM = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else 6650.0
rn = torch.rand(B, device=dev) * 0.12 + 1.0
zc = torch.empty(B, nc, device=dev); z_all = torch.empty(B, nt, device=dev)
raw_c = torch.randn(B * nc, 4, device=dev); raw = torch.randn(B * nt, 4, device=dev)
comp = torch.empty(B, 3, device=dev); acc = torch.empty(B, device=dev); dep = torch.empty(B, device=dev)
wts = torch.empty(B, nc, device=dev); g = torch.randn(B, 3, device=dev); d_raw = torch.empty(B * nt, 4, device=dev)
flags = 7      # white | infinite last bin | training
calls = {
    "stratified": (lambda: L.nsb_stratified_z(_lib.ptr(zc), None, B, nc, 2.0, 6.0, 1, 1, 0, st), 4 * nc),
    "composite_fwd_coarse": (lambda: L.nsb_composite_raw_fwd(_lib.ptr(raw_c), None, 1.0, _lib.ptr(zc), _lib.ptr(rn), _lib.ptr(comp), _lib.ptr(wts), None, None, B, nc, flags, 1, 0, st), 20 * nc + 24 + 4 * nc),
    "resample_merge": (lambda: L.nsb_resample_merge(_lib.ptr(zc), _lib.ptr(wts), None, _lib.ptr(z_all), None, B, nc, nf, 0, 1, 0, st), 12 * nc + 4 * nf),
    "resample_merge_det": (lambda: L.nsb_resample_merge(_lib.ptr(zc), _lib.ptr(wts), None, _lib.ptr(z_all), None, B, nc, nf, 1, 1, 0, st), 12 * nc + 4 * nf),
    "composite_fwd_fine": (lambda: L.nsb_composite_raw_fwd(_lib.ptr(raw), None, 1.0, _lib.ptr(z_all), _lib.ptr(rn), _lib.ptr(comp), None, _lib.ptr(acc), _lib.ptr(dep), B, nt, flags, 1, 0, st), 20 * nt + 24),
    "composite_fwd_fine_eval": (lambda: L.nsb_composite_raw_fwd(_lib.ptr(raw), None, 0.0, _lib.ptr(z_all), _lib.ptr(rn), _lib.ptr(comp), None, _lib.ptr(acc), _lib.ptr(dep), B, nt, 1, 0, 0, st), 20 * nt + 24),
    "composite_bwd_fine": (lambda: L.nsb_composite_raw_bwd(_lib.ptr(raw), None, 1.0, _lib.ptr(z_all), _lib.ptr(rn), _lib.ptr(g), _lib.ptr(d_raw), B, nt, flags, 1, 0, st), 36 * nt + 16),
}
out = {"rays": B, "Nc": nc, "Nf": nf, "hbm_peak_gbs": HBM}
for name, (f, bytes_per_ray) in calls.items():
    _lib.check(f()); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.check(f())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = B * bytes_per_ray / (ms * 1e-3) / 1e9
    out[name] = {"ms": round(ms, 4), "algorithmic_bytes": B * bytes_per_ray, "GBps": round(gbs, 1), "frac_hbm": round(gbs / HBM, 3)}
print(json.dumps(out))
