"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: ray sharding + one gradient all-reduce per step
reproduces the single-process gradient; pixel-tile sharding + gather reproduces the full frame.
Compute is done by the oracle here (no GPU in this container); the CUDA path uses the same host functions."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from nerf_sandbox_b200 import dist as D
    from oracle import nerf_oracle as O
    B, nc, nf = 8, 16, 16
    rng = np.random.default_rng(0)                      # identical on both ranks
    pc, pf = O.init_params(rng, 0.4), O.init_params(rng, 0.4)
    rays = O.synthetic_rays(rng, B)
    draws = dict(U=rng.uniform(0, 1, (B, nc)).astype(np.float32), u_fine=rng.uniform(0, 1, (B, nf)).astype(np.float32),
                 noise_c=rng.standard_normal(B * nc).astype(np.float32), noise_f=rng.standard_normal(B * (nc + nf)).astype(np.float32))
    full = O.train_step(pc, pf, rays, near=2.0, far=6.0, nc=nc, nf=nf, **draws)
    # ---- training: each rank takes its ray shard, grads summed by ONE all-reduce, scaled 1/world
    s, e = D.shard_range(B, rank, world)
    sh_rays = {k: v[s:e] for k, v in rays.items()}
    sh_draws = dict(U=draws["U"][s:e], u_fine=draws["u_fine"][s:e], noise_c=draws["noise_c"][s * nc:e * nc],
                    noise_f=draws["noise_f"][s * (nc + nf):e * (nc + nf)])
    loc = O.train_step(pc, pf, sh_rays, near=2.0, far=6.0, nc=nc, nf=nf, **sh_draws)
    flat = torch.from_numpy(np.concatenate([O.flatten_params(loc["grads_c"]), O.flatten_params(loc["grads_f"])]))
    w = D.allreduce_grads(flat)
    got = flat.numpy() / w
    ref = np.concatenate([O.flatten_params(full["grads_c"]), O.flatten_params(full["grads_f"])])
    ok_grad = np.abs(got - ref).max() <= 2e-5 * np.abs(ref).max() + 1e-8
    # ---- eval: pixel-tile shards gathered back (ragged last shard, alignment 4)
    n = 37
    frame = np.arange(n * 5, dtype=np.float32).reshape(n, 5)
    s2, e2 = D.shard_range(n, rank, world, align=4)
    out = D.gather_shards(torch.from_numpy(frame[s2:e2].copy()), n, align=4).numpy()
    ok_gather = np.array_equal(out, frame)
    cover = [D.shard_range(n, r, world, align=4) for r in range(world)]
    ok_cover = cover[0][0] == 0 and cover[-1][1] == n and all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
    q.put((rank, w, bool(ok_grad), bool(ok_gather), bool(ok_cover)))
    dist.destroy_process_group()


def test_ray_sharded_grads_and_tile_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, w, ok_grad, ok_gather, ok_cover in res:
        assert w == 2 and ok_grad and ok_gather and ok_cover, (rank, w, ok_grad, ok_gather, ok_cover)


def test_shard_range_properties():
    from nerf_sandbox_b200.dist import shard_range
    for n, world, align in [(640000, 8, 128), (640000, 3, 128), (1024, 8, 1), (5, 8, 1), (0, 2, 1), (190512, 4, 128)]:
        spans = [shard_range(n, r, world, align) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all((e - s) % align == 0 for s, e in spans[:-1] if e < n)
