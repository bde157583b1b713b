import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_sandbox_b200 as nsb
from oracle import nerf_oracle as O
DEV = "cuda"
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV, torch.float32)
rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12))
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
rng = np.random.default_rng(11)
p = O.init_params(np.random.default_rng(7), sigma_bias=0.3)
ep = O.positional_encode(rng.uniform(-4, 4, (Q, 3)).astype(np.float32), 10)
ed = O.positional_encode(O._normalize(rng.standard_normal((Q, 3)).astype(np.float32)), 4)
d_out = rng.standard_normal((Q, 4)).astype(np.float32)
raw, caches = O.mlp_forward(p, ep, ed, keep=True)
refg = O.mlp_backward(p, caches, d_out)
for mode in ("fp32", "bf16"):
    net = nsb.NeRF(63, 27, mode=mode).to(DEV)
    net.load_state_dict({k: T(v) for k, v in p.items()})
    out = net(T(ep), T(ed)); out.backward(T(d_out)); torch.cuda.synchronize()
    print(mode, "raw rel", rel(out.detach().cpu().numpy(), raw))
    for (name, _), q in zip(O.PARAM_SHAPES, net.parameters()):
        g = q.grad.cpu().numpy(); r = refg[name]
        extra = ""
        if name == "mlp.4.weight":
            extra = f" [h part {rel(g[:, :256], r[:, :256]):.4f} | gx part {rel(g[:, 256:], r[:, 256:]):.4f}]"
        if name == "color_fc.weight":
            extra = f" [feat part {rel(g[:, :256], r[:, :256]):.4f} | gd part {rel(g[:, 256:], r[:, 256:]):.4f}]"
        print(f"  {name:18s} rel {rel(g, r):.4f}  |g| {np.linalg.norm(g):.4e} |ref| {np.linalg.norm(r):.4e}{extra}")
