"""Where an 800x800 eval frame spends its time: python scripts/frame_breakdown.py  (run under ncu --metrics gpu__time_duration.sum
for the per-kernel list, or plain for the wall time)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_sandbox_b200 as nsb
dev = torch.device("cuda", 0)
tr = nsb.VanillaTrainer(dev, mode="bf16", seed=0, sigma_bias=0.3)
H = W = 800
n = H * W
g = torch.Generator(device=dev); g.manual_seed(0)
o = torch.nn.functional.normalize(torch.randn(n, 3, device=dev, generator=g), dim=-1) * 4.0311
d = torch.nn.functional.normalize(-o + 0.35 * torch.randn(n, 3, device=dev, generator=g), dim=-1)
rn = torch.ones(n, device=dev)
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
def frame():
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        nsb.render_rays(o[s:e], d[s:e], rn[s:e], d[s:e], tr.nerf_c, tr.nerf_f, near=2.0, far=6.0, nc=64, nf=128, white_bkgd=True)
frame(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); frame(); e1.record(); torch.cuda.synchronize()
print(f"chunk {chunk}: {e0.elapsed_time(e1):.2f} ms / frame")
