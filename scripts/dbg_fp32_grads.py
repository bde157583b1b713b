"""Per-parameter gradient comparison of the fp32 mode's two GEMM back ends on the `mlp` golden:
   NSB_FP32_GEMM=ffma python scripts/dbg_fp32_grads.py save /tmp/a.pt ; python scripts/dbg_fp32_grads.py cmp /tmp/a.pt"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_sandbox_b200 as nsb
from oracle import nerf_oracle as O
dev = torch.device("cuda", 0)
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "mlp.npz"))
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
p = O.init_params(np.random.default_rng(int(g["seed"])), sigma_bias=float(g["sigma_bias"]))
net = nsb.NeRF(63, 27, 8, 256, skip_pos=4, sigma_activation="relu", mode="fp32").to(dev)
net.load_state_dict({k: T(v) for k, v in p.items()})
out = net(T(g["enc_pos"]), T(g["enc_dir"]))
out.backward(T(g["d_out"]))
grads = {n: q.grad.detach().cpu() for n, q in net.named_parameters()}
grads["__out"] = out.detach().cpu()
if sys.argv[1] == "save":
    torch.save(grads, sys.argv[2])
else:
    ref = torch.load(sys.argv[2])
    for n in grads:
        a, b = grads[n].double(), ref[n].double()
        d = (a - b).abs()
        i = int(d.argmax())
        print(f"{n:22s} shape {tuple(a.shape)}  max|diff| {float(d.max()):.3e} at flat {i} (this {float(a.reshape(-1)[i]):+.6e} other {float(b.reshape(-1)[i]):+.6e})  max|ref| {float(b.abs().max()):.3e}  rel-L2 {float((a-b).norm()/b.norm().clamp_min(1e-30)):.2e}")
