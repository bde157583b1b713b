"""sample_pdf with the reference's signature (utils/sampling_utils.py:5-64) on the nsb_sample_pdf kernel."""
from __future__ import annotations

import torch

from . import _hooks, _lib


@torch.no_grad()
def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, n_samples: int, *, deterministic: bool = False,
               u: torch.Tensor | None = None, cdf: torch.Tensor | None = None, return_inds: bool = False,
               seed: int | None = None):
    """Hierarchical sampling from a piecewise-constant PDF; returns (B, n_samples).

    Extra keyword-only inputs beyond the reference: ``u`` (explicit uniforms replacing torch.rand at
    sampling_utils.py:48), ``cdf`` (explicit (B,M+1) CDF) and ``return_inds`` (also return
    searchsorted(cdf,u,right=True) as int64) -- the hooks the bit-exact parity test needs."""
    if bins.ndim != 2 or weights.ndim != 2:
        raise ValueError(f"Expected (B,·) tensors: bins={tuple(bins.shape)}, weights={tuple(weights.shape)}")
    B, M = weights.shape
    if bins.shape[-1] not in (M, M + 1) or bins.shape[0] != B:
        raise ValueError(f"Incompatible shapes: bins={tuple(bins.shape)}, weights={tuple(weights.shape)}")
    b, w = _lib.f32c(bins), _lib.f32c(weights)
    out = torch.empty((B, n_samples), device=b.device, dtype=torch.float32)
    inds = torch.empty((B, n_samples), device=b.device, dtype=torch.int64) if return_inds else None
    if u is None and not deterministic and _hooks.uniform is not None:
        u = _hooks.uniform(B, int(n_samples), b.device)
    uu = None if (u is None or deterministic) else _lib.f32c(u)
    cc = None if cdf is None else _lib.f32c(cdf)
    if uu is not None and tuple(uu.shape) != (B, n_samples):
        raise ValueError(f"u must be {(B, n_samples)}, got {tuple(uu.shape)}")
    if cc is not None and tuple(cc.shape) != (B, M + 1):
        raise ValueError(f"cdf must be {(B, M + 1)}, got {tuple(cc.shape)}")
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if (uu is None and not deterministic) else 0
    _lib.check(_lib.lib().nsb_sample_pdf(_lib.ptr(b), b.shape[-1], _lib.ptr(w), M, _lib.ptr(uu), _lib.ptr(cc), _lib.ptr(out),
                                         _lib.ptr(inds), B, int(n_samples), int(bool(deterministic)), seed, 0,
                                         _lib.stream()), "nsb_sample_pdf")
    out = out.to(weights.dtype)
    return (out, inds) if return_inds else out
