"""Drop-in installation behind the reference's by-name seam.

The reference binds the hot-path callables by name at import time (train/trainer.py:46-54,
utils/render_utils.py:24-25, utils/validation_renderer.py:16-24) and its own test patches exactly that seam
(tests/unit/test_trainer.py:172).  ``install()`` rebinds those names to this package's CUDA-backed
implementations so ``scripts/train_nerf.py --vanilla`` runs unmodified on a B200."""
from __future__ import annotations

import importlib
import sys
import types

from . import encoders, mlps, rays, render, sampling

_BINDINGS = {
    "nerf_sandbox.source.train.trainer": {
        "PositionalEncoder": encoders.PositionalEncoder, "get_vanilla_nerf_encoders": encoders.get_vanilla_nerf_encoders,
        "NeRF": mlps.NeRF, "log_nerf_arch": mlps.log_nerf_arch, "volume_render_rays": render.volume_render_rays,
        "nerf_forward_pass": render.nerf_forward_pass, "sample_pdf": sampling.sample_pdf,
        "get_camera_rays": rays.get_camera_rays},
    "nerf_sandbox.source.utils.render_utils": {
        "get_camera_rays": rays.get_camera_rays, "render_pose": rays.render_pose,
        "sample_pdf": sampling.sample_pdf, "volume_render_rays": render.volume_render_rays,
        "nerf_forward_pass": render.nerf_forward_pass, "render_image_chunked": render.render_image_chunked},
    "nerf_sandbox.source.utils.validation_renderer": {"render_image_chunked": render.render_image_chunked,
                                                      "render_pose": rays.render_pose, "get_camera_rays": rays.get_camera_rays},
    "nerf_sandbox.source.data.samplers": {"get_camera_rays": rays.get_camera_rays},
    "nerf_sandbox.source.utils.ray_utils": {"get_camera_rays": rays.get_camera_rays},
    "nerf_sandbox.source.models.encoders": {
        "PositionalEncoder": encoders.PositionalEncoder, "get_vanilla_nerf_encoders": encoders.get_vanilla_nerf_encoders},
    "nerf_sandbox.source.models.mlps": {"NeRF": mlps.NeRF, "log_nerf_arch": mlps.log_nerf_arch},
    "nerf_sandbox.source.utils.sampling_utils": {"sample_pdf": sampling.sample_pdf},
}


def _imageio_shim():
    """render_utils.py:20 imports imageio at module scope; provide a stub when it is not installed
    (PNG/MP4 export is outside the hot path)."""
    if "imageio" in sys.modules:
        return
    try:
        importlib.import_module("imageio")
    except ImportError:
        m = types.ModuleType("imageio"); m.v2 = types.ModuleType("imageio.v2")

        def _missing(*a, **k):
            raise RuntimeError("imageio is not installed: image/video export is unavailable")
        m.imread = m.imwrite = m.mimwrite = m.v2.imread = m.v2.imwrite = _missing
        sys.modules["imageio"], sys.modules["imageio.v2"] = m, m.v2


def install(verbose: bool = False) -> dict:
    """Rebind the reference's names; returns {module: [names rebound]}.  The reference package must be importable."""
    _imageio_shim()
    done = {}
    for modname, names in _BINDINGS.items():
        mod = importlib.import_module(modname)
        for n, obj in names.items():
            setattr(mod, n, obj)
        done[modname] = sorted(names)
        if verbose:
            print(f"[nerf_sandbox_b200] {modname}: {', '.join(sorted(names))}")
    return done
