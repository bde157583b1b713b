"""nerf_sandbox_b200 -- B200-native engine for the vanilla-NeRF ray-march path of evan-wes/nerf-sandbox.

Python host code (this package) mirrors the reference's module interfaces; all arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of ``libnsb.so`` (include/nsb.h).  No CPU fallback."""
from .encoders import PositionalEncoder, get_vanilla_nerf_encoders
from .mlps import NeRF, log_nerf_arch, set_default_mode, get_default_mode
from .install import install
from .sampling import sample_pdf
from .render import volume_render_rays, nerf_forward_pass, render_image_chunked, render_rays
from .rays import get_camera_rays, render_pose
from .samplers import RandomPixelRaySampler
from .trainer import VanillaTrainer
from .validation import frame_outputs, compute_psnr

__all__ = ["PositionalEncoder", "get_vanilla_nerf_encoders", "NeRF", "log_nerf_arch", "sample_pdf", "volume_render_rays",
           "nerf_forward_pass", "render_image_chunked", "render_rays", "get_camera_rays", "render_pose", "RandomPixelRaySampler", "VanillaTrainer",
           "frame_outputs", "compute_psnr", "install", "set_default_mode", "get_default_mode"]
