"""2-GPU test (skipped on a 1-GPU box): ray-sharded training keeps the replicas identical with both gradient exchanges (NCCL
all-reduce; all-reduce fused into the Adam kernel over peer memory) and the pixel-tiled frame equals the single-GPU frame."""
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
flats = {}
for ar in ("nccl", "p2p"):          # NCCL sum-allreduce + Adam   vs   nsb_adam_allreduce_step (peer loads, one kernel)
    tr = nsb.VanillaTrainer(dev, mode=os.environ.get("NSB_MODE", "bf16"), seed=0, sigma_bias=0.4, allreduce=ar)
    assert (tr.peer is not None) == (ar == "p2p")
    for step in range(4):
        rays = O.synthetic_rays(np.random.default_rng(100 * step + rank), 256)      # different rays per rank
        tr.step({k: T(v) for k, v in rays.items()})
    if ar == "p2p":                 # the same exchange inside the CUDA-graph step (one graph per gradient-buffer parity)
        for step in range(4, 9):
            rays = O.synthetic_rays(np.random.default_rng(100 * step + rank), 256)
            tr.step_graph({k: T(v) for k, v in rays.items()})
        assert tr.adam_t == 9 and int(tr._step_dev.item()) == 9
    flat = torch.cat([tr.nerf_c.flat_params(), tr.nerf_f.flat_params()])
    ref = flat.clone(); dist.broadcast(ref, 0)
    assert torch.equal(flat, ref), f"replicas diverged ({ar})"
    flats[ar] = flat
# (the two trainers are not compared with each other: the wgrad flush uses float atomics, and Adam turns last-bit gradient
# differences into +-lr parameter differences in the first steps)
# nsb_adam_allreduce_step against NCCL all-reduce + nsb_adam_step on the SAME gradients: two summands commute, so at world
# size 2 the results must agree bit for bit; otherwise to rounding of the different summation order
import ctypes as C
from nerf_sandbox_b200 import _lib
L = _lib.lib(); n = 4096
pg = D.PeerGrads(2 * n, dev)
g = torch.Generator(device=dev); g.manual_seed(5)
st = [[torch.randn(n, device=dev, generator=g) * s for s in (1.0, 0.0, 0.0)] for _ in range(2)]     # p, m, v of 2 nets ...
st += [[t.clone() for t in trip] for trip in st]                                                    # ... for the two paths
arr = lambda ts: (C.c_void_p * len(ts))(*[_lib.ptr(t) for t in ts])
for epoch in range(1, 4):
    gr = torch.Generator(device=dev); gr.manual_seed(1000 * epoch + rank)
    grads = torch.randn(2 * n, device=dev, generator=gr)
    pg.buffer(epoch).copy_(grads)
    mc = pg.multicast(epoch) if (os.environ.get("NSB_NVLS") == "1" or world >= 4) else None      # NVLS variant: multimem.ld_reduce
    red_mc, red, lsync = pg.two_phase() if mc else (None, None, None)                            # ... two-phase from 4 ranks
    _lib.check(L.nsb_adam_allreduce_step(arr([st[0][0], st[1][0]]), arr([st[0][1], st[1][1]]), arr([st[0][2], st[1][2]]), 2,
                                         pg.pointers(epoch), mc, red_mc, _lib.ptr(red), _lib.ptr(lsync), pg.flag_array, rank, world, epoch, n,
                                         5e-4, 0.9, 0.999, 1e-8, epoch, 1.0 / world, None, _lib.stream()), "fused")
    red = grads.clone(); dist.all_reduce(red)
    for k in range(2):
        _lib.check(L.nsb_adam_step(_lib.ptr(st[2 + k][0]), _lib.ptr(red[k * n:(k + 1) * n]), _lib.ptr(st[2 + k][1]), _lib.ptr(st[2 + k][2]),
                                   n, 5e-4, 0.9, 0.999, 1e-8, epoch, 1.0 / world, None, _lib.stream()), "adam")
    for k in range(2):
        for a, b in zip(st[k], st[2 + k]):
            if world == 2:
                assert torch.equal(a, b), (epoch, k, float((a - b).abs().max()))
            else:
                torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
# a non-finite loss on ONE rank makes EVERY rank skip the update (the reference's skip, trainer.py:713-716, made collective)
before = [t.clone() for t in st[0]]
for epoch, bad_rank in ((4, 1), (5, None)):
    loss = torch.tensor([float("nan") if rank == bad_rank else 0.5], device=dev)
    pg.buffer(epoch).copy_(torch.ones(2 * n, device=dev))
    mc = pg.multicast(epoch) if world >= 4 else None
    red_mc, red, lsync = pg.two_phase() if mc else (None, None, None)
    _lib.check(L.nsb_adam_allreduce_step(arr([st[0][0], st[1][0]]), arr([st[0][1], st[1][1]]), arr([st[0][2], st[1][2]]), 2,
                                         pg.pointers(epoch), mc, red_mc, _lib.ptr(red), _lib.ptr(lsync), pg.flag_array, rank, world, epoch, n,
                                         5e-4, 0.9, 0.999, 1e-8, epoch, 1.0 / world, _lib.ptr(loss), _lib.stream()), "fused+guard")
    torch.cuda.synchronize()
    same = all(torch.equal(a, b) for a, b in zip(st[0], before))
    assert same == (bad_rank is not None), (epoch, rank, same)
# resume: load_state_dict on a trainer whose flag blocks are ahead of the restored step count, then keep training in lockstep
sd = tr.state_dict()
for step in range(9, 12):
    tr.step_graph({k: T(v) for k, v in O.synthetic_rays(np.random.default_rng(100 * step + rank), 256).items()})
tr.load_state_dict(sd)                      # back to t = 9 while the flags say 12
assert tr.adam_t == 9
for step in range(9, 12):
    f = tr.step if step % 2 else tr.step_graph
    f({k: T(v) for k, v in O.synthetic_rays(np.random.default_rng(100 * step + rank), 256).items()})
tr.check_peers()
flat = torch.cat([tr.nerf_c.flat_params(), tr.nerf_f.flat_params()]); ref = flat.clone(); dist.broadcast(ref, 0)
assert torch.equal(flat, ref), "replicas diverged after resume"
H, W = 20, 31
rays = O.synthetic_rays(np.random.default_rng(7), H * W)
args = (T(rays["rays_o_marching"]), T(rays["rays_d_marching_unit"]), T(rays["rays_d_marching_norm"]).reshape(-1))
out = D.render_image_sharded(*args, H, W, 2.0, 6.0, tr.nerf_c, tr.nerf_f, 64, 128, True, eval_chunk=128,
                             viewdirs_world_unit=T(rays["rays_d_world_unit"]))
rgb, acc, depth = nsb.render_rays(*args, T(rays["rays_d_world_unit"]), tr.nerf_c, tr.nerf_f, near=2.0, far=6.0, nc=64, nf=128, white_bkgd=True)
assert torch.equal(out["rgb"].reshape(-1, 3), rgb) and torch.equal(out["depth"].reshape(-1), depth)
dist.barrier()
if os.environ.get("NSB_TEST_TIMEOUT") == "1":
    # a peer that never arrives: the waiting ranks skip the update after NSB_PEER_TIMEOUT_S, stay alive (no trap) and raise on
    # the host from check_peers(); the late rank finds everyone's flags and completes normally
    import time
    before = torch.cat([tr.nerf_c.flat_params(), tr.nerf_f.flat_params()]).clone()
    rays = {k: T(v) for k, v in O.synthetic_rays(np.random.default_rng(999), 256).items()}
    if rank == world - 1:
        time.sleep(float(os.environ["NSB_PEER_TIMEOUT_S"]) + 4.0)
    tr.step(rays)
    torch.cuda.synchronize()
    after = torch.cat([tr.nerf_c.flat_params(), tr.nerf_f.flat_params()])
    if rank != world - 1:
        assert torch.equal(before, after), "a timed-out exchange must skip the update"
        try:
            tr.check_peers(); raise SystemExit("check_peers() did not raise")
        except RuntimeError as e:
            assert "timed out" in str(e), e
        x = torch.ones(4, device=dev) * 2; assert float(x.sum()) == 8.0          # the context is still usable
        time.sleep(12.0)             # keep this rank's symmetric buffers alive until the late rank has finished reading them
    print("rank", rank, "timeout path ok", flush=True)
    os._exit(0)                                  # the ranks are out of step now: no collective teardown
dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_training_and_tiled_eval(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, NSB_ROOT=ROOT)
    if env.get("NSB_TEST_TIMEOUT") == "1":
        env.setdefault("NSB_PEER_TIMEOUT_S", "15")
    nproc = os.environ.get("NSB_TEST_NPROC", "2")            # e.g. 8 on a full node (NVLS path: NSB_NVLS=1 or >= 4 ranks)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", nproc, "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
