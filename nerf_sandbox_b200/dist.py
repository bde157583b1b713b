"""Multi-GPU host logic (one process per GPU, torch.distributed over NCCL/NVLink; gloo in the CPU tests).

The path shards by rays: every reduction of the hot path runs along the per-ray sample axis,
so training needs exactly one exchange per step -- a sum-allreduce of the flat fp32 gradient buffer of both
nets (2 x 595,844 floats = 4.77 MB, latency-bound on NVSwitch) -- and evaluation needs one gather of the
rendered pixel tiles.  Nothing inside a ray ever crosses a device."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n: int, rank: int, world: int, align: int = 1):
    """Contiguous [start, end) of `n` items owned by `rank`; every shard but the last is a multiple of `align`
    (eval uses align = the 128-point tile so no tile straddles two GPUs)."""
    per = -(-n // world)
    per = -(-per // align) * align
    start = min(n, rank * per)
    return start, min(n, start + per)


def allreduce_grads(flat_grads: torch.Tensor, group=None) -> int:
    """Sum-allreduce the flat gradient buffer in place; returns the world size (the Adam kernel applies 1/world)."""
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return world


def gather_shards(local: torch.Tensor, n_total: int, group=None, align: int = 1) -> torch.Tensor:
    """All-gather row shards produced with shard_range(n_total, rank, world, align) back into [n_total, ...]."""
    rank, world = world_info(group)
    if world == 1:
        return local
    per = shard_range(n_total, 0, world, align)[1]
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat(out, 0)[:n_total]


@torch.no_grad()
def render_image_sharded(rays_o, rays_d_unit, ray_norms, H, W, near, far, nerf_c, nerf_f, nc_eval, nf_eval, white_bkgd,
                         eval_chunk=65536, *, viewdirs_world_unit=None, infinite_last_bin=False, group=None) -> dict:
    """render_image_chunked (utils/render_utils.py:285-424) with the H*W rays split into one contiguous pixel block per
    rank; every rank returns the full frame."""
    from .render import render_rays
    rank, world = world_info(group)
    n = H * W
    s, e = shard_range(n, rank, world, align=128)
    o, d, rn = rays_o.reshape(n, 3)[s:e], rays_d_unit.reshape(n, 3)[s:e], ray_norms.reshape(n)[s:e]
    vd = None if viewdirs_world_unit is None else viewdirs_world_unit.reshape(n, 3)[s:e]
    m = e - s
    out = torch.empty((m, 5), device=rays_o.device, dtype=torch.float32)       # rgb | acc | depth
    for a in range(0, m, eval_chunk):
        b = min(m, a + eval_chunk)
        rgb, acc, depth = render_rays(o[a:b].contiguous(), d[a:b].contiguous(), rn[a:b].contiguous(),
                                      None if vd is None else vd[a:b].contiguous(), nerf_c, nerf_f, near=near, far=far,
                                      nc=nc_eval, nf=nf_eval, white_bkgd=white_bkgd, infinite_last_bin=infinite_last_bin)
        out[a:b, :3] = rgb; out[a:b, 3] = acc; out[a:b, 4] = depth
    full = gather_shards(out, n, group, align=128)
    return {"rgb": full[:, :3].reshape(H, W, 3), "acc": full[:, 3].reshape(H, W, 1), "depth": full[:, 4].reshape(H, W, 1)}
