"""Eval output path of the reference's ValidationRenderer (utils/validation_renderer.py:485-533) on the device:
uint8 RGB / opacity / normalised-depth images as save_rgb_png / save_gray_png would write them (utils/render_utils.py:28-47)
and the frame PSNR of _compute_psnr (:171-196), from the float buffers of render_image_chunked -- one kernel, and the host
receives 5 bytes per pixel instead of 20."""
from __future__ import annotations

import torch

from . import _lib


@torch.no_grad()
def frame_outputs(res: dict, *, near: float, far: float, use_ndc: bool = False, gt_rgb=None, mask=None) -> dict:
    """res: {"rgb": (H,W,3), "acc": (H,W,1), "depth": (H,W,1)} float CUDA tensors (render_image_chunked / render_pose).
    Returns {"rgb": uint8 (H,W,3), "opacity": uint8 (H,W), "depth": uint8 (H,W)} and, with gt_rgb (H,W,3) [and mask
    (H,W,1) or (H,W), 1 = valid], "psnr" / "mse" as 0-dim device tensors."""
    rgb = _lib.f32c(res["rgb"])
    H, W = rgb.shape[0], rgb.shape[1]
    n = H * W
    acc = _lib.f32c(res["acc"]).reshape(n) if res.get("acc") is not None else None
    depth = _lib.f32c(res["depth"]).reshape(n) if res.get("depth") is not None else None
    dev = rgb.device
    rgb8 = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
    acc8 = torch.empty((H, W), dtype=torch.uint8, device=dev) if acc is not None else None
    depth8 = torch.empty((H, W), dtype=torch.uint8, device=dev) if depth is not None else None
    gt = m = scratch = out = None
    if gt_rgb is not None:
        gt = _lib.f32c(torch.as_tensor(gt_rgb).to(dev))
        if tuple(gt.shape) != (H, W, 3):
            raise ValueError(f"gt_rgb must be ({H},{W},3), got {tuple(gt.shape)}")
        if mask is not None:
            m = _lib.f32c(torch.as_tensor(mask).to(dev))
            m = (m[..., :1] if m.dim() == 3 else m).reshape(n).contiguous()
        scratch = torch.empty(2, dtype=torch.float64, device=dev)
        out = torch.empty(2, dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().nsb_frame_output(_lib.ptr(rgb.reshape(n, 3)), _lib.ptr(acc), _lib.ptr(depth), n, float(near), float(far),
                                           int(bool(use_ndc)), _lib.ptr(rgb8), _lib.ptr(acc8), _lib.ptr(depth8), _lib.ptr(gt), _lib.ptr(m),
                                           _lib.ptr(scratch), _lib.ptr(out), _lib.stream()), "nsb_frame_output")
    ret = {"rgb": rgb8, "opacity": acc8, "depth": depth8}
    if out is not None:
        ret["psnr"], ret["mse"] = out[0], out[1]
    return ret


def compute_psnr(pred_rgb_hw3, gt_rgb_hw3, mask_hw1=None) -> float:
    """ValidationRenderer._compute_psnr (utils/validation_renderer.py:171-196)."""
    H, W = pred_rgb_hw3.shape[0], pred_rgb_hw3.shape[1]
    r = frame_outputs({"rgb": pred_rgb_hw3, "acc": None, "depth": None}, near=0.0, far=1.0, gt_rgb=gt_rgb_hw3, mask=mask_hw1)
    return float(r["psnr"].item())
