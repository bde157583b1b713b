"""BASELINE configs[4]: sampler / compositor HBM GB/s (vs measured peak) and MLP tensor TFLOP/s over an
Nc x Nf sweep at a bandwidth-bound ray count.  python scripts/sweep_k2k3.py [rays] > profiles/...json"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_sandbox_b200 as nsb
from nerf_sandbox_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
quick = len(sys.argv) > 2
dev = "cuda"; L = _lib.lib(); st = _lib.stream()
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
HBM, TF = peaks["hbm_gbs"], peaks["bf16_tflops"]


def timeit(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


net = nsb.NeRF(63, 27, mode="bf16").to(dev)
rows = []
grid = [(64, 128)] if quick else [(nc, nf) for nc in (32, 64, 128, 256) for nf in (64, 128, 256, 512)]
rn = torch.rand(B, device=dev) * 0.12 + 1.0
for nc, nf in grid:
    nt = nc + nf
    zc = torch.empty(B, nc, device=dev); w_c = torch.rand(B, nc, device=dev) ** 4; z_all = torch.empty(B, nt, device=dev)
    t_strat = timeit(lambda: _lib.check(L.nsb_stratified_z(_lib.ptr(zc), None, B, nc, 2.0, 6.0, 1, 1, 0, st)))
    t_res = timeit(lambda: _lib.check(L.nsb_resample_merge(_lib.ptr(zc), _lib.ptr(w_c), None, _lib.ptr(z_all), None, B, nc, nf, 0, 1, 0, st)))
    raw = torch.randn(B * nt, 4, device=dev); comp = torch.empty(B, 3, device=dev); acc = torch.empty(B, device=dev); dep = torch.empty(B, device=dev)
    wts = torch.empty(B, nc, device=dev); g = torch.randn(B, 3, device=dev); d_raw = torch.empty(B * nt, 4, device=dev)
    t_cc = timeit(lambda: _lib.check(L.nsb_composite_raw_fwd(_lib.ptr(raw), None, 1.0, _lib.ptr(zc), _lib.ptr(rn), _lib.ptr(comp), _lib.ptr(wts), _lib.ptr(acc), _lib.ptr(dep), B, nc, 7, 1, 0, st)))
    t_cf = timeit(lambda: _lib.check(L.nsb_composite_raw_fwd(_lib.ptr(raw), None, 1.0, _lib.ptr(z_all), _lib.ptr(rn), _lib.ptr(comp), None, _lib.ptr(acc), _lib.ptr(dep), B, nt, 7, 1, 0, st)))
    t_cb = timeit(lambda: _lib.check(L.nsb_composite_raw_bwd(_lib.ptr(raw), None, 1.0, _lib.ptr(z_all), _lib.ptr(rn), _lib.ptr(g), _lib.ptr(d_raw), B, nt, 7, 1, 0, st)))
    # algorithmic bytes per ray, SURVEY 8d
    by = {"stratified": 4 * nc, "resample_merge": 12 * nc + 4 * nf, "composite_fwd_coarse": 20 * nc + 24 + 4 * nc,
          "composite_fwd_fine": 20 * nt + 24, "composite_bwd_fine": 36 * nt + 16}
    ts = {"stratified": t_strat, "resample_merge": t_res, "composite_fwd_coarse": t_cc, "composite_fwd_fine": t_cf, "composite_bwd_fine": t_cb}
    row = {"rays": B, "Nc": nc, "Nf": nf}
    for k in by:
        gbs = B * by[k] / ts[k] / 1e9
        row[k] = {"ms": ts[k] * 1e3, "GBps": gbs, "frac_hbm": gbs / HBM}
    # MLP forward on the fine pass (points = min(B,65536) * nt to bound memory)
    Bm = min(B, 65536)
    o = torch.randn(Bm, 3, device=dev); d = torch.nn.functional.normalize(torch.randn(Bm, 3, device=dev), dim=-1)
    zz = torch.sort(torch.rand(Bm, nt, device=dev) * 4 + 2, -1).values.contiguous(); rawm = torch.empty(Bm * nt, 4, device=dev)
    ws = torch.empty(L.nsb_field_workspace_bytes(Bm * nt, 1, 0), dtype=torch.uint8, device=dev)
    t_m = timeit(lambda: _lib.check(L.nsb_field_fwd_rays(_lib.ptr(o), _lib.ptr(d), _lib.ptr(zz), _lib.ptr(rn[:Bm].contiguous()), _lib.ptr(d), _lib.ptr(net.packed()), _lib.ptr(rawm), _lib.ptr(ws), ws.numel(), Bm, nt, 1, 0, st)), reps=3)
    tf = Bm * nt * 1186816 / t_m / 1e12
    row["mlp_fwd_fine"] = {"ms": t_m * 1e3, "TFLOPs": tf, "frac_bf16_peak": tf / TF}
    rows.append(row)
    print(json.dumps(row), flush=True)
    del raw, d_raw, rawm, ws
    torch.cuda.empty_cache()
