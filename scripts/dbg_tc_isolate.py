import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_sandbox_b200 as nsb
from oracle import nerf_oracle as O
DEV = torch.device("cuda", 0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
big = O.synthetic_rays(np.random.default_rng(8), 1024)
args = (T(big["rays_o_marching"]), T(big["rays_d_marching_unit"]), T(big["rays_d_marching_norm"]).reshape(-1), T(big["rays_d_world_unit"]))
tr = nsb.VanillaTrainer(DEV, mode="bf16", seed=7, sigma_bias=1.0)
torch.cuda.synchronize(); print("pack ok", flush=True)
r = nsb.render_rays(*args, tr.nerf_c, tr.nerf_f, near=2.0, far=6.0, nc=64, nf=128, white_bkgd=True)
torch.cuda.synchronize(); print("forward (no stash) ok", float(r[0].mean()), flush=True)
from nerf_sandbox_b200 import _lib
L = _lib.lib()
Q = 1024 * 64
z = torch.sort(torch.rand((1024, 64), device=DEV) * 4 + 2, -1).values.contiguous()
wsb = L.nsb_field_workspace_bytes(Q, 1, 1); ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
raw = torch.empty((Q, 4), device=DEV)
_lib.check(L.nsb_field_fwd_rays(_lib.ptr(args[0]), _lib.ptr(args[1]), _lib.ptr(z), _lib.ptr(args[2]), _lib.ptr(args[3]), _lib.ptr(tr.nerf_f.packed()), _lib.ptr(raw), _lib.ptr(ws), wsb, 1024, 64, 1, 1, _lib.stream()))
torch.cuda.synchronize(); print("forward (stash) ok", flush=True)
os.environ["NSB_DBG_STAGE"] = "dgrad"
g = torch.zeros(_lib.N_PARAMS, device=DEV); d_raw = torch.randn((Q, 4), device=DEV) * 1e-3
_lib.check(L.nsb_field_bwd(_lib.ptr(d_raw), _lib.ptr(tr.nerf_f.packed()), _lib.ptr(g), _lib.ptr(ws), wsb, Q, 1, _lib.stream()))
torch.cuda.synchronize(); print("backward ok", float(g.norm()), flush=True)
