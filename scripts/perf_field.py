"""Time the field kernels alone (CUDA events): usage  python scripts/perf_field.py [mode] [rays] [N]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_sandbox_b200 as nsb
from nerf_sandbox_b200 import _lib

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
N = int(sys.argv[3]) if len(sys.argv) > 3 else 192
stash = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = "cuda"
L = _lib.lib()
net = nsb.NeRF(63, 27, mode=mode).to(dev)
rng = np.random.default_rng(0)
o = torch.randn(B, 3, device=dev); d = torch.nn.functional.normalize(torch.randn(B, 3, device=dev), dim=-1)
z = torch.sort(torch.rand(B, N, device=dev) * 4 + 2, -1).values.contiguous()
rn = torch.ones(B, device=dev)
Q = B * N
wsb = L.nsb_field_workspace_bytes(Q, net.mode, stash)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
raw = torch.empty(Q, 4, device=dev)
pk = net.packed()


def run():
    _lib.check(L.nsb_field_fwd_rays(_lib.ptr(o), _lib.ptr(d), _lib.ptr(z), _lib.ptr(rn), _lib.ptr(d), _lib.ptr(pk), _lib.ptr(raw),
                                    _lib.ptr(ws), wsb, B, N, net.mode, stash, _lib.stream()))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(json.dumps({"mode": mode, "rays": B, "N": N, "points": Q, "stash": stash, "ms": ms, "Mpts_per_s": Q / ms / 1e3,
                  "fwd_TFLOPs": Q * 1186816 / (ms * 1e-3) / 1e12}))
