"""A/B of the compositor kernels (run layout vs strided) against an fp64 torch restatement on training-like inputs."""
import os, sys, subprocess, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from nerf_sandbox_b200 import _lib
    L = _lib.lib(); st = _lib.stream(); dev = "cuda"
    torch.manual_seed(0)
    B, N = 4096, 192
    raw = torch.randn(B * N, 4, device=dev); raw[:, 3] = raw[:, 3] * 2.0 + 0.4
    z = torch.sort(torch.rand(B, N, device=dev) * 4 + 2, -1).values.contiguous()
    rn = torch.rand(B, device=dev) * 0.12 + 1.0
    noise = torch.randn(B * N, device=dev)
    comp = torch.empty(B, 3, device=dev); w = torch.empty(B, N, device=dev); acc = torch.empty(B, device=dev); dep = torch.empty(B, device=dev)
    _lib.check(L.nsb_composite_raw_fwd(_lib.ptr(raw), _lib.ptr(noise), 1.0, _lib.ptr(z), _lib.ptr(rn), _lib.ptr(comp), _lib.ptr(w), _lib.ptr(acc), _lib.ptr(dep), B, N, 7, 1, 0, st))
    g = torch.randn(B, 3, device=dev); d_raw = torch.empty(B * N, 4, device=dev)
    _lib.check(L.nsb_composite_raw_bwd(_lib.ptr(raw), _lib.ptr(noise), 1.0, _lib.ptr(z), _lib.ptr(rn), _lib.ptr(g), _lib.ptr(d_raw), B, N, 7, 1, 0, st))
    torch.cuda.synchronize()
    # fp64 restatement (render_utils.py:108-167 + :236-246)
    r64 = raw.double().reshape(B, N, 4); rgb = torch.sigmoid(r64[..., :3]).requires_grad_()
    pre = (r64[..., 3] + noise.double().reshape(B, N)).requires_grad_()
    sig = torch.relu(pre)
    zz = z.double(); delta = torch.cat([zz[:, 1:] - zz[:, :-1], torch.full((B, 1), 1e10, device=dev, dtype=torch.float64)], -1) * rn.double()[:, None]
    sdt = (sig * delta).clamp(0, 60); alpha = 1 - torch.exp(-sdt)
    T = torch.cumprod(torch.cat([torch.ones(B, 1, device=dev, dtype=torch.float64), 1 - alpha + 1e-10], -1), -1)[:, :-1]
    ww = T * alpha; a = ww.sum(-1).clamp(0, 1); c = (ww[..., None] * rgb).sum(1) + (1 - a)[:, None]
    c = c.clamp(0, 1)
    (c * g.double()).sum().backward()
    res = {"comp_maxabs": float((comp.double() - c).abs().max()), "w_maxabs": float((w.double() - ww).abs().max()),
           "comp_p99": float(torch.quantile((comp.double() - c).abs().flatten(), 0.99)),
           "d_sigma_rel": float((d_raw[:, 3].double().reshape(B, N) - pre.grad).norm() / pre.grad.norm()),
           "d_rgb_rel": float(((d_raw[:, :3].double().reshape(B, N, 3) / (rgb * (1 - rgb)).clamp_min(1e-12)) - rgb.grad).norm() / rgb.grad.norm())}
    worst = int((comp.double() - c).abs().max(-1).values.argmax())
    res["worst_ray"] = worst; res["worst_acc"] = float(a[worst]); res["worst_w_err"] = float((w[worst].double() - ww[worst]).abs().max())
    res["worst_sample"] = int((w[worst].double() - ww[worst]).abs().argmax())
    print(json.dumps(res))
else:
    for tag, env in (("run", {}), ("strided", {"NSB_K3_STRIDED": "1"})):
        out = subprocess.run([sys.executable, __file__, "child"], env={**os.environ, **env}, capture_output=True, text=True)
        print(tag, out.stdout.strip(), out.stderr.strip()[-300:])
