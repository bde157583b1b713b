"""Soak run: thousands of graph-replayed training steps on a fixed synthetic scene + repeated frames; checks that nothing
hangs, no NaN appears and the loss goes down.  python scripts/soak.py [steps]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_sandbox_b200 as nsb
from oracle import nerf_oracle as O
dev = torch.device("cuda", 0)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
tr = nsb.VanillaTrainer(dev, mode="bf16", seed=0, sigma_bias=0.4, lr_scheduler="cosine", lr_scheduler_params={"T_max": steps, "eta_min": 5e-5})
pool = []
for s in range(16):
    r = O.synthetic_rays(np.random.default_rng(s), 1024)
    d = r["rays_d_world_unit"]
    r["rgb"] = (0.5 + 0.5 * np.sin(3.0 * d + np.array([0.0, 2.0, 4.0], np.float32))).astype(np.float32)     # a smooth target
    pool.append({k: T(v) for k, v in r.items()})
t0 = time.time(); hist = []
for i in range(steps):
    sc = tr.step_graph(pool[i % 16])
    if i % 250 == 0 or i == steps - 1:
        hist.append(float(sc[0]))
        assert np.isfinite(hist[-1]), (i, hist)
torch.cuda.synchronize()
print(f"{steps} steps in {time.time() - t0:.2f} s; loss {hist[0]:.4f} -> {hist[-1]:.4f}; history {[round(h, 4) for h in hist]}")
assert hist[-1] < 0.5 * hist[0]
b = pool[0]
for _ in range(20):
    out = nsb.render_rays(b["rays_o_marching"], b["rays_d_marching_unit"], b["rays_d_marching_norm"].reshape(-1), b["rays_d_world_unit"],
                          tr.nerf_c, tr.nerf_f, near=2.0, far=6.0, nc=64, nf=128, white_bkgd=True)
assert torch.isfinite(out[0]).all()
print("eval mse vs target", float(((out[0] - b["rgb"]) ** 2).mean()))
