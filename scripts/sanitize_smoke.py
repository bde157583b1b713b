"""Small end-to-end run for compute-sanitizer: python scripts/sanitize_smoke.py
(compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_smoke.py)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_sandbox_b200 as nsb
from oracle import nerf_oracle as O
dev = torch.device("cuda", 0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for mode in (sys.argv[1:] or ["bf16", "fp32"]):
    tr = nsb.VanillaTrainer(dev, mode=mode, seed=0, sigma_bias=0.4)
    rays = {k: T(v) for k, v in O.synthetic_rays(np.random.default_rng(1), 200).items()}      # ragged: 200*64 / 200*192 points
    for _ in range(2):
        tr.step(rays)
    for _ in range(2):
        tr.step_graph(rays)
    out = nsb.render_rays(rays["rays_o_marching"], rays["rays_d_marching_unit"], rays["rays_d_marching_norm"].reshape(-1),
                          rays["rays_d_world_unit"], tr.nerf_c, tr.nerf_f, near=2.0, far=6.0, nc=64, nf=128, white_bkgd=True)
    torch.cuda.synchronize()
    print(mode, "loss", float(tr.scalars[0]), "rgb mean", float(out[0].mean()))
