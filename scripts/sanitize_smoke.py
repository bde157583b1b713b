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
    # ragged sample counts: the compositor's partial-run kernels (N not a multiple of 32), the general resampler, grad clip
    from nerf_sandbox_b200 import _lib
    L = _lib.lib(); st = _lib.stream()
    for N in (37, 100, 192, 300):
        raw = torch.randn(50 * N, 4, device=dev); z = torch.sort(torch.rand(50, N, device=dev) * 4 + 2, -1).values.contiguous()
        rn = torch.ones(50, device=dev); comp = torch.empty(50, 3, device=dev); w = torch.empty(50, N, device=dev)
        _lib.check(L.nsb_composite_raw_fwd(_lib.ptr(raw), None, 1.0, _lib.ptr(z), _lib.ptr(rn), _lib.ptr(comp), _lib.ptr(w), None, None, 50, N, 7, 1, 0, st))
        g = torch.randn(50, 3, device=dev); d_raw = torch.empty_like(raw)
        _lib.check(L.nsb_composite_raw_bwd(_lib.ptr(raw), None, 1.0, _lib.ptr(z), _lib.ptr(rn), _lib.ptr(g), _lib.ptr(d_raw), 50, N, 7, 1, 0, st))
    for nc, nf in ((64, 128), (40, 70), (32, 64), (256, 512)):
        zc = torch.sort(torch.rand(50, nc, device=dev) * 4 + 2, -1).values.contiguous(); wc = torch.rand(50, nc, device=dev)
        za = torch.empty(50, nc + nf, device=dev)
        for det in (0, 1):
            _lib.check(L.nsb_resample_merge(_lib.ptr(zc), _lib.ptr(wc), None, _lib.ptr(za), None, 50, nc, nf, det, 1, 0, st))
    gg = torch.randn(2 * _lib.N_PARAMS, device=dev); sc = torch.zeros(1, device=dev)
    _lib.check(L.nsb_grad_clip(_lib.ptr(gg), gg.numel(), 1.0, 1.0, _lib.ptr(sc), st))
    torch.cuda.synchronize()
    print(mode, "loss", float(tr.scalars[0]), "rgb mean", float(out[0].mean()))
