"""Train-step throughput at other batch sizes (e.g. BASELINE configs[3]: 8192 rays/GPU): python scripts/perf_step.py [rays] [mode]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_sandbox_b200 as nsb
from oracle import nerf_oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
dev = torch.device("cuda", 0)
tr = nsb.VanillaTrainer(dev, mode=mode, seed=0, sigma_bias=0.3)
batches = [{k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in O.synthetic_rays(np.random.default_rng(s), B).items()} for s in range(2)]
for i in range(5):
    tr.step_graph(batches[i & 1])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20
e0.record()
for i in range(K):
    tr.step_graph(batches[i & 1])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"{mode} {B} rays/step: {ms:.3f} ms/step = {B / ms * 1e3:,.0f} rays/s; MLP {B * 256 * 3489024 / (ms * 1e-3) / 1e12:.0f} TFLOP/s algorithmic; loss {float(tr.scalars[0]):.4f}")
