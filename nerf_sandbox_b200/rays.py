"""get_camera_rays with the reference's signature (utils/ray_utils.py:10-136) on the nsb_camera_rays kernel, and
render_pose (utils/render_utils.py:426-526) composed from it and render_image_chunked."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_CONV = {"opengl": 0, "blender": 0, "nerf": 0, "opencv": 1, "colmap": 1, "pytorch3d": 2, "p3d": 2}


@torch.no_grad()
def get_camera_rays(image_h: int, image_w: int, intrinsic_matrix, transform_camera_to_world, *, device=None,
                    dtype: torch.dtype = torch.float32, convention: str = "opengl", pixel_center: bool = False,
                    as_ndc: bool = False, near_plane: float = 1.0, pixels_xy=None):
    """Returns (rays_o_world, rays_d_world_unit, rays_d_world_norm, rays_o_marching, rays_d_marching_unit,
    rays_d_marching_norm) like the reference; shapes (N,3)/(N,1)."""
    K = np.ascontiguousarray(torch.as_tensor(intrinsic_matrix).detach().cpu().numpy(), dtype=np.float32)
    c2w = np.ascontiguousarray(torch.as_tensor(transform_camera_to_world).detach().cpu().numpy(), dtype=np.float32)
    if K.shape[-2:] != (3, 3):
        raise ValueError(f"K must be (3,3), got {K.shape}")
    if c2w.shape[-2:] not in {(3, 4), (4, 4)}:
        raise ValueError(f"c2w must be (3,4) or (4,4), got {c2w.shape}")
    conv = (convention or "opengl").lower()
    if conv not in _CONV:
        raise ValueError(f"Unknown convention '{convention}'")
    dev = torch.device(device) if device is not None else torch.device("cuda")
    if dev.type != "cuda":
        raise RuntimeError("nerf_sandbox_b200 generates rays on CUDA devices only (no CPU fallback)")
    px = None
    if pixels_xy is not None:
        px = torch.as_tensor(pixels_xy).to(device=dev, dtype=torch.float32).reshape(-1, 2).contiguous()
    n = int(px.shape[0]) if px is not None else int(image_h) * int(image_w)
    outs = [torch.empty((n, k), device=dev, dtype=torch.float32) for k in (3, 3, 1, 3, 3, 1)]
    c2w34 = np.ascontiguousarray(c2w[:3, :4])
    _lib.check(_lib.lib().nsb_camera_rays(
        int(image_h), int(image_w), K.ctypes.data_as(C.c_void_p), c2w34.ctypes.data_as(C.c_void_p), 4, _CONV[conv],
        int(bool(pixel_center)), int(bool(as_ndc)), float(near_plane), _lib.ptr(px), n, *[_lib.ptr(o) for o in outs],
        torch.cuda.current_stream(dev).cuda_stream), "nsb_camera_rays")
    return tuple(o.to(dtype) for o in outs)


@torch.no_grad()
def render_pose(c2w, H, W, K, near: float, far: float, pos_enc, dir_enc, nerf_c, nerf_f, device, white_bkgd: bool = True,
                nc_eval: int = 64, nf_eval: int = 128, eval_chunk: int = 8192, perturb: bool = False,
                sigma_activation: str = "relu", *, use_ndc: bool = False, convention: str = "opengl",
                near_plane: float | None = None, samp_near: float | None = None, samp_far: float | None = None,
                infinite_last_bin: bool = False):
    """Render one pose (render_utils.py:426-526): WORLD rays feed the directional encoding, MARCHING rays (NDC if
    requested) are sampled and composited."""
    from .render import render_image_chunked
    world = get_camera_rays(H, W, K, c2w, device=device, convention=convention, pixel_center=True, as_ndc=False,
                            near_plane=near)
    if use_ndc:
        ndc = get_camera_rays(H, W, K, c2w, device=device, convention=convention, pixel_center=True, as_ndc=True,
                              near_plane=float(near if near_plane is None else near_plane))
        o, d, nrm = ndc[3], ndc[4], ndc[5]
        s_near, s_far = (0.0 if samp_near is None else float(samp_near)), (1.0 if samp_far is None else float(samp_far))
    else:
        o, d, nrm = world[0], world[1], world[2]
        s_near, s_far = (near if samp_near is None else float(samp_near)), (far if samp_far is None else float(samp_far))
    return render_image_chunked(o, d, nrm, H, W, s_near, s_far, pos_enc, dir_enc, nerf_c, nerf_f, nc_eval, nf_eval, white_bkgd,
                                device, eval_chunk=eval_chunk, perturb=perturb, sigma_activation=sigma_activation,
                                viewdirs_world_unit=world[1], infinite_last_bin=infinite_last_bin)
