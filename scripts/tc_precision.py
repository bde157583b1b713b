"""Numerical distance of the tensor-core mode (fp16 forward operands, bf16 gradients, fp32 accumulate) from the fp32 mode /
the fp32 oracle, on the shapes the tests use.  Prints one JSON object (kept under profiles/ as the basis of the test bars)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import nerf_sandbox_b200 as nsb
from oracle import nerf_oracle as O
DEV = torch.device("cuda", 0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
N = lambda t: t.detach().float().cpu().numpy()
rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))
out = {}
# 1. one 1024-ray train step vs the fp32 oracle: loss, composites, gradient direction and norm per net
rng = np.random.default_rng(101)
batch = O.synthetic_rays(rng, 1024)
draws = dict(U=rng.uniform(0, 1, (1024, 64)).astype(np.float32), u_fine=rng.uniform(0, 1, (1024, 128)).astype(np.float32),
             noise_c=rng.standard_normal(1024 * 64).astype(np.float32), noise_f=rng.standard_normal(1024 * 192).astype(np.float32))
res = {}
for mode in ("fp32", "bf16"):
    tr = nsb.VanillaTrainer(DEV, mode=mode, seed=5, sigma_bias=0.4)
    o = tr._train_step({k: T(v) for k, v in batch.items()}, {k: T(v) for k, v in draws.items()})
    o["loss"].backward()
    res[mode] = dict(loss=float(o["loss"].detach()), comp_f=N(o["comp_f"]), gc=N(torch.cat([q.grad.reshape(-1) for q in tr.nerf_c.parameters()])),
                     gf=N(torch.cat([q.grad.reshape(-1) for q in tr.nerf_f.parameters()])),
                     norms_f=np.array([float(q.grad.norm()) for q in tr.nerf_f.parameters()]))
a, b = res["bf16"], res["fp32"]
cos = lambda x, y: float(x.astype(np.float64) @ y.astype(np.float64) / (np.linalg.norm(x.astype(np.float64)) * np.linalg.norm(y.astype(np.float64))))
out["step_1024"] = dict(loss_rel=abs(a["loss"] - b["loss"]) / b["loss"], comp_f_maxabs=float(np.abs(a["comp_f"] - b["comp_f"]).max()),
                        comp_f_psnr=float(-10 * np.log10(np.mean((a["comp_f"] - b["comp_f"]) ** 2))),
                        grad_cos_c=cos(a["gc"], b["gc"]), grad_cos_f=cos(a["gf"], b["gf"]), grad_rel_c=rel(a["gc"], b["gc"]), grad_rel_f=rel(a["gf"], b["gf"]),
                        grad_norm_ratio_f=float(np.linalg.norm(a["gf"]) / np.linalg.norm(b["gf"])),
                        per_param_norm_dev_max=float(np.abs(a["norms_f"] / b["norms_f"] - 1).max()))
# 2. eval render: 4096 rays, both modes, same weights
big = O.synthetic_rays(np.random.default_rng(8), 4096)
args = (T(big["rays_o_marching"]), T(big["rays_d_marching_unit"]), T(big["rays_d_marching_norm"]).reshape(-1), T(big["rays_d_world_unit"]))
imgs = {}
for mode in ("fp32", "bf16"):
    tr = nsb.VanillaTrainer(DEV, mode=mode, seed=7, sigma_bias=1.0)
    imgs[mode] = [N(x) for x in nsb.render_rays(*args, tr.nerf_c, tr.nerf_f, near=2.0, far=6.0, nc=64, nf=128, white_bkgd=True)]
out["eval_4096"] = dict(rgb_psnr=float(-10 * np.log10(np.mean((imgs["bf16"][0] - imgs["fp32"][0]) ** 2))), rgb_maxabs=float(np.abs(imgs["bf16"][0] - imgs["fp32"][0]).max()),
                        acc_maxabs=float(np.abs(imgs["bf16"][1] - imgs["fp32"][1]).max()),
                        depth_maxabs=float(np.abs(imgs["bf16"][2] * imgs["bf16"][1] - imgs["fp32"][2] * imgs["fp32"][1]).max()))
# 3. 20 steps on 512 fixed rays from the same init
fixed = {k: T(v) for k, v in O.synthetic_rays(np.random.default_rng(21), 512).items()}
fin = {}
for mode in ("fp32", "bf16"):
    t2 = nsb.VanillaTrainer(DEV, mode=mode, seed=3, sigma_bias=0.4)
    for _ in range(20):
        sc = t2.step(fixed)
    fin[mode] = float(sc[0])
out["loss_after_20_steps"] = dict(fp32=fin["fp32"], tc=fin["bf16"], rel=abs(fin["bf16"] - fin["fp32"]) / fin["fp32"])
# 4. trained scene (tests/test_gpu_fullsize.py)
import test_gpu_fullsize as fs
out["trained_scene"] = fs.train_and_eval(nsb, int(os.environ.get("NSB_TRAIN_STEPS", "600")))
print(json.dumps(out))
