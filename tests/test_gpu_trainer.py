"""Trainer-shell behaviour around the kernels (train/trainer.py:702-729): gradient clipping, the non-finite-loss skip,
autograd hygiene of the fused step, CUDA-graph invalidation, reference-format checkpoints."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _batch(seed, n=256):
    return {k: T(v) for k, v in O.synthetic_rays(np.random.default_rng(seed), n).items()}


def _flat(tr):
    return torch.cat([tr.nerf_c.flat_params(), tr.nerf_f.flat_params()]).clone()


def _update_distance(p_a, p_b, p_start):
    """Relative L2 distance between two parameter UPDATES from the same start.  Gradients are summed with fp32 atomics
    (run-to-run rounding differs) and Adam turns a 1e-9 difference on a near-zero gradient into a fraction of lr, so two
    runs of the same steps agree on the update as a whole, not element by element."""
    da, db = (p_a - p_start).double(), (p_b - p_start).double()
    return float((da - db).norm() / db.norm())


@pytest.mark.parametrize("graph", [False, True])
def test_grad_clip_matches_torch_clip_grad_norm(graph):
    """grad_clip_norm (trainer.py:719-721) on the fused step == clip_grad_norm_ + torch Adam on the same gradients."""
    import nerf_sandbox_b200 as nsb
    b = _batch(1)
    mk = lambda clip: nsb.VanillaTrainer(DEV, rays_per_batch=256, mode="fp32", seed=3, sigma_bias=0.4, grad_clip_norm=clip)
    probe = mk(0.0)
    out = probe._train_step(b); out["loss"].backward()
    total = float(torch.nn.utils.clip_grad_norm_(probe.parameters(), max_norm=1e9))
    clip = 0.25 * total                                          # a threshold that really clips
    tr = mk(clip)
    start = _flat(tr)
    ref = mk(0.0)                                                # reference semantics: autograd + clip_grad_norm_ + torch Adam
    opt = torch.optim.Adam(ref.parameters(), lr=5e-4)
    for _ in range(3):
        (tr.step_graph if graph else tr.step)(b)
        opt.zero_grad(set_to_none=True)
        ref.global_step = tr.global_step - 1                     # same Philox streams as the fused step just taken
        o = ref._train_step(b); o["loss"].backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm=clip)
        opt.step()
    torch.cuda.synchronize()
    assert float(tr._scal8[4]) > 0                               # the squared norm the clip kernel measured
    assert _update_distance(_flat(tr), _flat(ref), start) <= 2e-2
    # and the clip really changed the gradient Adam saw: scaled by 0.25 against eps = 1e-8 is invisible to Adam's ratio, so
    # check the kernel itself -- the clipped buffer's norm equals the threshold
    g = torch.randn(2 * 595844, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
    norm0 = float(g.double().norm())
    from nerf_sandbox_b200 import _lib
    scratch = torch.zeros(1, device=DEV)
    _lib.check(_lib.lib().nsb_grad_clip(_lib.ptr(g), g.numel(), 3.0, 1.0, _lib.ptr(scratch), _lib.stream()))
    assert abs(float(g.norm()) - 3.0) <= 1e-3 and abs(float(scratch[0]) ** 0.5 - norm0) <= 1e-3 * norm0
    g2 = g.clone()
    _lib.check(_lib.lib().nsb_grad_clip(_lib.ptr(g2), g2.numel(), 10.0, 1.0, _lib.ptr(scratch), _lib.stream()))
    assert torch.equal(g, g2)                                     # below the threshold: untouched


def test_non_finite_loss_skips_the_update():
    """trainer.py:713-716: a non-finite loss leaves parameters and Adam moments untouched.  (The path's own guards --
    nan_to_num + clamp on composites and targets, trainer.py:999-1001 -- make such a loss hard to produce through a step, so the
    optimiser kernels are driven directly with a poisoned / a clean loss word.)"""
    import ctypes as C
    from nerf_sandbox_b200 import _lib
    L = _lib.lib()
    n = 595844
    for fused in (False, True):
        for bad in (float("nan"), float("inf"), None):
            p = torch.randn(2 * n, device=DEV); g = torch.randn(2 * n, device=DEV); m = torch.zeros(2 * n, device=DEV); v = torch.zeros(2 * n, device=DEV)
            p0 = p.clone()
            loss = torch.tensor([0.25 if bad is None else bad], device=DEV)
            if fused:
                arr = lambda ts: (C.c_void_p * len(ts))(*[_lib.ptr(t) for t in ts])
                _lib.check(L.nsb_adam_allreduce_step(arr([p[:n], p[n:]]), arr([m[:n], m[n:]]), arr([v[:n], v[n:]]), 2, arr([g]), None, None, None,
                                                     None, None, 0, 1, 1, n, 5e-4, 0.9, 0.999, 1e-8, 1, 1.0, _lib.ptr(loss), _lib.stream()))
            else:
                _lib.check(L.nsb_adam_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), 2 * n, 5e-4, 0.9, 0.999, 1e-8, 1, 1.0, _lib.ptr(loss),
                                           _lib.stream()))
            torch.cuda.synchronize()
            if bad is None:
                assert float((p - p0).abs().max()) > 1e-4 and float(m.abs().max()) > 0
            else:
                assert torch.equal(p, p0) and float(m.abs().max()) == 0 and float(v.abs().max()) == 0


def test_fused_step_backward_returns_its_own_gradients():
    """Two _train_step calls before backward (gradient accumulation, (l1 + l2).backward()): each backward must return the
    gradients of ITS forward, not whatever the trainer's shared buffers hold by then."""
    import nerf_sandbox_b200 as nsb
    tr = nsb.VanillaTrainer(DEV, rays_per_batch=256, mode="fp32", seed=4, sigma_bias=0.4)
    b1, b2 = _batch(5), _batch(6)
    g = {}
    for tag, b in (("1", b1), ("2", b2)):
        for p in tr.parameters():
            p.grad = None
        tr._train_step(b)["loss"].backward()
        g[tag] = torch.cat([p.grad.reshape(-1) for p in tr.parameters()]).clone()
    for p in tr.parameters():
        p.grad = None
    l1 = tr._train_step(b1)["loss"]; l2 = tr._train_step(b2)["loss"]
    tr.step(b2)                                                   # and the fast path in between, overwriting tr.grads_*
    (l1 + l2).backward()
    both = torch.cat([p.grad.reshape(-1) for p in tr.parameters()])
    np.testing.assert_allclose(both.cpu().numpy(), (g["1"] + g["2"]).cpu().numpy(), rtol=1e-5, atol=1e-9)


def test_graph_is_recaptured_when_workspace_or_params_move():
    import nerf_sandbox_b200 as nsb
    tr = nsb.VanillaTrainer(DEV, rays_per_batch=256, mode="bf16", seed=5, sigma_bias=0.4)
    ref = nsb.VanillaTrainer(DEV, rays_per_batch=256, mode="bf16", seed=5, sigma_bias=0.4)
    start = _flat(tr)
    small, big = _batch(7, 256), _batch(8, 1024)
    tr.step_graph(small); ref.step(small)
    tr.step(big); ref.step(big)                                   # grows the workspace -> the captured graph is stale
    assert tr._graphs is None
    tr.step_graph(small); ref.step(small)
    tr.nerf_c.to(DEV); tr.nerf_c._flat = None; tr.nerf_c.flat_params()          # re-flattened parameters: new addresses
    key_before = tr._graph_key
    tr.step_graph(small); ref.step(small)
    assert tr._graph_key != key_before
    torch.cuda.synchronize()
    assert torch.isfinite(_flat(tr)).all() and _update_distance(_flat(tr), _flat(ref), start) <= 5e-2


def test_checkpoint_round_trip_in_reference_format():
    """state_dict() is what Trainer.save_checkpoint writes (trainer.py:596-621): nets + a torch.optim.Adam state over the 48
    parameters; it is a snapshot (no live aliases), loads into torch Adam and back, and a resumed run continues like the original."""
    import nerf_sandbox_b200 as nsb
    mk = lambda: nsb.VanillaTrainer(DEV, rays_per_batch=256, mode="bf16", seed=6, sigma_bias=0.4, lr_scheduler="cosine",
                                    lr_scheduler_params={"T_max": 50, "eta_min": 1e-5})
    a = mk()
    bs = [_batch(10 + i) for i in range(6)]
    for b in bs[:3]:
        a.step_graph(b)
    sd = a.state_dict()
    snap = sd["opt"]["state"][30]["exp_avg"].clone()
    for b in bs[3:]:
        a.step(b)                                                  # (mixing graph and eager steps after a snapshot)
    assert torch.equal(sd["opt"]["state"][30]["exp_avg"], snap)   # the snapshot did not move with training
    opt = torch.optim.Adam(a.parameters(), lr=5e-4); opt.load_state_dict(sd["opt"])
    r = mk(); r.load_state_dict({**sd, "opt": opt.state_dict()})  # through torch's own format and back
    assert r.adam_t == 3 and r.global_step == 3
    at3 = torch.cat([sd["nerf_c"][k].reshape(-1) for k in sd["nerf_c"]] + [sd["nerf_f"][k].reshape(-1) for k in sd["nerf_f"]])
    assert torch.equal(_flat(r), at3)                              # exactly the saved weights and moments ...
    assert torch.equal(r.nerf_f.unflatten(r.m_f)[6], sd["opt"]["state"][24 + 6]["exp_avg"])
    for b in bs[3:]:
        r.step(b)
    torch.cuda.synchronize()
    assert _update_distance(_flat(r), _flat(a), at3) <= 2e-2       # ... and the same continuation (lr schedule, Adam t, Philox streams)
    legacy = mk(); legacy.load_state_dict({"step": 3, "nerf_c": sd["nerf_c"], "nerf_f": sd["nerf_f"], "opt": None})
    assert legacy.adam_t == 0 and float(legacy.m_c.abs().max()) == 0
