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


class PeerGrads:
    """Double-buffered flat gradient buffers in symmetric memory (every rank can load every other rank's buffer over
    NVLink) plus a flag block, for `nsb_adam_allreduce_step`: the all-reduce and Adam run as one kernel with no NCCL
    call on the step path.  `buffer(epoch)` is the gradient buffer of that step (epoch parity), `pointers(epoch, ofs)`
    the ctypes arrays of all ranks' addresses of it."""

    def __init__(self, n_floats: int, device, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        self.rank, self.world = world_info(group)
        group = group if group is not None else dist.group.WORLD
        self.n = int(n_floats)
        self.buf = symm_mem.empty(2 * self.n, dtype=torch.float32, device=device)
        self.flags = symm_mem.empty(64, dtype=torch.int32, device=device)
        self.buf.zero_(); self.flags.zero_()
        self._h_buf = symm_mem.rendezvous(self.buf, group.group_name)
        self._h_flags = symm_mem.rendezvous(self.flags, group.group_name)
        self.buf_ptrs = [int(p) for p in self._h_buf.buffer_ptrs]
        # NVLS: multicast address of the gradient buffers (0 when the fabric has no multicast support).  NSB_NVLS=0 disables
        # it, =1 requires it; by default it is used from 4 ranks up, where it beats reading world-1 peer buffers
        import os
        want = os.environ.get("NSB_NVLS", "auto")
        mc = int(getattr(self._h_buf, "multicast_ptr", 0) or 0) if want != "0" else 0
        if want == "1" and not mc:
            raise RuntimeError("NSB_NVLS=1 but the symmetric-memory handle reports no multicast support")
        self.mc_ptr = mc if (want == "1" or self.world >= 4) else 0
        # two-phase exchange (reduce-scatter + multicast store, nsb_adam_allreduce_step): one more symmetric buffer that receives
        # the REDUCED gradient on every rank, + two device ints of local bookkeeping.  Used whenever the NVLS sum is
        # (NSB_TWO_PHASE=0 falls back to every rank pulling the whole buffer through multimem.ld_reduce).
        self.red = self.red_mc = self.local_sync = None
        if self.mc_ptr and os.environ.get("NSB_TWO_PHASE", "1") != "0":
            self.red = symm_mem.empty(self.n, dtype=torch.float32, device=device)
            self.red.zero_()
            self._h_red = symm_mem.rendezvous(self.red, group.group_name)
            rmc = int(getattr(self._h_red, "multicast_ptr", 0) or 0)
            if rmc:
                self.red_mc = rmc
                self.local_sync = torch.zeros(4, dtype=torch.int32, device=device)
            else:
                self.red = None
        self.flag_ptrs = [int(p) for p in self._h_flags.buffer_ptrs]
        self._C = C
        self.flag_array = (C.c_void_p * self.world)(*self.flag_ptrs)
        torch.cuda.synchronize(device)
        dist.barrier(group)                       # every rank's buffers are zeroed before anyone's first epoch

    def buffer(self, epoch: int) -> torch.Tensor:
        h = epoch & 1
        return self.buf[h * self.n:(h + 1) * self.n]

    def multicast(self, epoch: int):
        """Multicast address of the gradient buffer of `epoch` (None without NVLS)."""
        return (self.mc_ptr + (epoch & 1) * self.n * 4) if self.mc_ptr else None

    def two_phase(self):
        """(multicast address of the reduced-gradient buffers, this rank's copy, local sync ints) or (None, None, None)."""
        if self.red is None:
            return None, None, None
        return self.red_mc, self.red, self.local_sync

    def pointers(self, epoch: int, float_offset: int = 0):
        byte_ofs = ((epoch & 1) * self.n + float_offset) * 4
        return (self._C.c_void_p * self.world)(*[p + byte_ofs for p in self.buf_ptrs])


def gather_shards(local: torch.Tensor, n_total: int, group=None, align: int = 1) -> torch.Tensor:
    """All-gather row shards produced with shard_range(n_total, rank, world, align) back into [n_total, ...]: one
    all_gather_into_tensor into a single buffer of world equal slots (the last shard is padded), no per-rank copies."""
    rank, world = world_info(group)
    if world == 1:
        return local
    per = shard_range(n_total, 0, world, align)[1]
    if local.shape[0] != per:                                   # only the last rank's shard can be short
        pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        local = pad
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:n_total]


@torch.no_grad()
def render_image_sharded(rays_o, rays_d_unit, ray_norms, H, W, near, far, nerf_c, nerf_f, nc_eval, nf_eval, white_bkgd,
                         eval_chunk=65536, *, viewdirs_world_unit=None, infinite_last_bin=False, sigma_activation="relu",
                         group=None, timings=None) -> dict:
    """render_image_chunked (utils/render_utils.py:285-424) with the H*W rays split into one contiguous pixel block per
    rank; every rank returns the full frame.  ``timings`` (optional dict) receives CUDA events around this rank's
    compute and around the all-gather (keys "compute", "gather": (start, end) event pairs)."""
    from .render import render_rays
    rank, world = world_info(group)
    n = H * W
    s, e = shard_range(n, rank, world, align=128)
    o, d, rn = rays_o.reshape(n, 3)[s:e], rays_d_unit.reshape(n, 3)[s:e], ray_norms.reshape(n)[s:e]
    vd = None if viewdirs_world_unit is None else viewdirs_world_unit.reshape(n, 3)[s:e]
    m = e - s
    out = torch.empty((m, 5), device=rays_o.device, dtype=torch.float32)       # rgb | acc | depth
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timings is not None else None
    if ev:
        ev[0].record()
    for a in range(0, m, eval_chunk):
        b = min(m, a + eval_chunk)
        rgb, acc, depth = render_rays(o[a:b].contiguous(), d[a:b].contiguous(), rn[a:b].contiguous(),
                                      None if vd is None else vd[a:b].contiguous(), nerf_c, nerf_f, near=near, far=far,
                                      nc=nc_eval, nf=nf_eval, white_bkgd=white_bkgd, infinite_last_bin=infinite_last_bin,
                                      sigma_activation=sigma_activation)
        out[a:b, :3] = rgb; out[a:b, 3] = acc; out[a:b, 4] = depth
    if ev:
        ev[1].record(); ev[2].record()
    full = gather_shards(out, n, group, align=128)
    if ev:
        ev[3].record()
        timings["compute"], timings["gather"] = (ev[0], ev[1]), (ev[2], ev[3])
    return {"rgb": full[:, :3].reshape(H, W, 3), "acc": full[:, 3].reshape(H, W, 1), "depth": full[:, 4].reshape(H, W, 1)}
