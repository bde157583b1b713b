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


_ORIGINALS = {}          # (module, name) -> the reference's own object, for uninstall()


def uninstall() -> None:
    """Put the reference's own callables back (e.g. to time its CPU path in the same process after a drop-in run)."""
    for (modname, n), obj in _ORIGINALS.items():
        mod = sys.modules.get(modname)
        if mod is not None:
            setattr(mod, n, obj)
    _ORIGINALS.clear()


def _imageio_shim():
    """render_utils.py:20 and the loaders import imageio at module scope; when it is not installed provide a stand-in:
    PNG read/write through PIL when that is present (enough for the Blender loader and the validation PNGs), video export
    unavailable.  Image IO is outside the hot path."""
    if "imageio" in sys.modules:
        return
    try:
        importlib.import_module("imageio")
    except ImportError:
        m = types.ModuleType("imageio"); m.v2 = types.ModuleType("imageio.v2")

        def _missing(*a, **k):
            raise RuntimeError("imageio is not installed: image/video export is unavailable")
        imread = imwrite = _missing
        try:
            import numpy as np
            from PIL import Image

            def imread(path, *a, **k):
                return np.array(Image.open(path))

            def imwrite(path, arr, *a, **k):
                Image.fromarray(np.asarray(arr)).save(str(path))
        except ImportError:
            pass
        m.imread = m.v2.imread = imread
        m.imwrite = m.v2.imwrite = imwrite
        m.mimwrite = m.v2.mimwrite = m.get_writer = m.v2.get_writer = _missing
        sys.modules["imageio"], sys.modules["imageio.v2"] = m, m.v2


def install(verbose: bool = False, mode: str | None = None) -> dict:
    """Rebind the reference's names; returns {module: [names rebound]}.  The reference package must be importable.

    ``mode`` ("bf16" = tcgen05 tensor cores, "fp32" = FFMA parity kernels; default: NSB_MODE, else "fp32") becomes the
    arithmetic mode of every ``NeRF`` the reference builds afterwards -- its Trainer passes no mode (train/trainer.py:326-341),
    so this is how ``train_nerf.py --vanilla`` reaches the tensor-core path."""
    if mode is not None:
        mlps.set_default_mode(mode)
    _imageio_shim()
    done = {}
    for modname, names in _BINDINGS.items():
        mod = importlib.import_module(modname)
        for n, obj in names.items():
            if (modname, n) not in _ORIGINALS and hasattr(mod, n):
                _ORIGINALS[(modname, n)] = getattr(mod, n)
            setattr(mod, n, obj)
        done[modname] = sorted(names)
        if verbose:
            print(f"[nerf_sandbox_b200] {modname}: {', '.join(sorted(names))}")
    return done
