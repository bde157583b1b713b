"""2-GPU NCCL test (skipped on a 1-GPU box): ray-sharded training keeps the replicas identical and the pixel-tiled
frame equals the single-GPU frame."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["NSB_ROOT"])
import nerf_sandbox_b200 as nsb
from nerf_sandbox_b200 import dist as D
from oracle import nerf_oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
tr = nsb.VanillaTrainer(dev, mode=os.environ.get("NSB_MODE", "bf16"), seed=0, sigma_bias=0.4)
for step in range(3):
    rays = O.synthetic_rays(np.random.default_rng(100 * step + rank), 256)      # different rays per rank
    tr.step({k: T(v) for k, v in rays.items()})
flat = torch.cat([tr.nerf_c.flat_params(), tr.nerf_f.flat_params()])
ref = flat.clone(); dist.broadcast(ref, 0)
assert torch.equal(flat, ref), "replicas diverged"
H, W = 20, 31
rays = O.synthetic_rays(np.random.default_rng(7), H * W)
args = (T(rays["rays_o_marching"]), T(rays["rays_d_marching_unit"]), T(rays["rays_d_marching_norm"]).reshape(-1))
out = D.render_image_sharded(*args, H, W, 2.0, 6.0, tr.nerf_c, tr.nerf_f, 64, 128, True, eval_chunk=128,
                             viewdirs_world_unit=T(rays["rays_d_world_unit"]))
rgb, acc, depth = nsb.render_rays(*args, T(rays["rays_d_world_unit"]), tr.nerf_c, tr.nerf_f, near=2.0, far=6.0, nc=64, nf=128, white_bkgd=True)
assert torch.equal(out["rgb"].reshape(-1, 3), rgb) and torch.equal(out["depth"].reshape(-1), depth)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_training_and_tiled_eval(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, NSB_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
