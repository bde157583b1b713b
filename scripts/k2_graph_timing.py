"""Kernel-only times of the small K2 / K3 launches: 20 launches captured in one CUDA graph, replayed, so that the host's
launch rate (ctypes + Python, ~15-25 us per call) does not cap a 20 us kernel.  python scripts/k2_graph_timing.py [rays] [Nc] [Nf]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_sandbox_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nf = int(sys.argv[3]) if len(sys.argv) > 3 else 128
dev = torch.device("cuda", 0); L = _lib.lib()
HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6452.5) if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6452.5
nt = nc + nf
zc = torch.sort(torch.rand(B, nc, device=dev) * 4 + 2, -1).values.contiguous(); wc = torch.rand(B, nc, device=dev)
za = torch.empty(B, nt, device=dev); raw_c = torch.randn(B * nc, 4, device=dev); rn = torch.ones(B, device=dev)
comp = torch.empty(B, 3, device=dev); w = torch.empty(B, nc, device=dev)
s = torch.cuda.Stream()
calls = {
    "stratified": (lambda st: L.nsb_stratified_z(_lib.ptr(zc), None, B, nc, 2.0, 6.0, 1, 1, 0, st), 4 * nc),
    "resample_merge": (lambda st: L.nsb_resample_merge(_lib.ptr(zc), _lib.ptr(wc), None, _lib.ptr(za), None, B, nc, nf, 0, 1, 0, st), 12 * nc + 4 * nf),
    "composite_fwd_coarse": (lambda st: L.nsb_composite_raw_fwd(_lib.ptr(raw_c), None, 1.0, _lib.ptr(zc), _lib.ptr(rn), _lib.ptr(comp), _lib.ptr(w), None, None, B, nc, 7, 1, 0, st), 24 * nc + 24),
}
out = {"rays": B, "Nc": nc, "Nf": nf, "hbm_peak_gbs": HBM, "method": "20 launches per CUDA-graph replay, 5 replays"}
with torch.cuda.stream(s):
    for name, (f, bpr) in calls.items():
        st = s.cuda_stream
        _lib.check(f(st)); s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(20):
                _lib.check(f(torch.cuda.current_stream().cuda_stream))
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s); s.synchronize()
        ms = e0.elapsed_time(e1) / 100
        gbs = B * bpr / (ms * 1e-3) / 1e9
        out[name] = {"ms": round(ms, 4), "GBps": round(gbs, 1), "frac_hbm": round(gbs / HBM, 3)}
print(json.dumps(out))
