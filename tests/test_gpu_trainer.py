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
    np.testing.assert_allclose(_flat(tr).cpu().numpy(), _flat(ref).cpu().numpy(), rtol=0, atol=3e-6)
    unclipped = mk(0.0)
    for _ in range(3):
        unclipped.step(b)
    assert float((_flat(unclipped) - _flat(tr)).abs().max()) > 1e-5      # (Adam normalises, but m/v histories differ once clipped)


@pytest.mark.parametrize("graph", [False, True])
def test_non_finite_loss_skips_the_update(graph):
    """trainer.py:713-716: a non-finite loss leaves parameters and Adam moments untouched."""
    import nerf_sandbox_b200 as nsb
    tr = nsb.VanillaTrainer(DEV, rays_per_batch=256, mode="bf16", seed=3, sigma_bias=0.4)
    step = tr.step_graph if graph else tr.step
    good, bad = _batch(2), _batch(2)
    bad["rgb"] = torch.full_like(bad["rgb"], float("inf"))       # guard01 maps +inf to 1 ... so poison the rays instead
    bad["rays_o_marching"] = torch.full_like(bad["rays_o_marching"], float("nan"))
    step(good); torch.cuda.synchronize()
    p0, m0 = _flat(tr), tr.m_f.clone()
    sc = step(bad); torch.cuda.synchronize()
    if not np.isfinite(float(sc[0])):                            # NaN rays -> NaN loss: the update must have been skipped
        assert torch.equal(_flat(tr), p0) and torch.equal(tr.m_f, m0)
    else:                                                         # the guards of the path absorbed it: then a normal step happened
        assert torch.isfinite(_flat(tr)).all()
    step(good); torch.cuda.synchronize()
    assert torch.isfinite(_flat(tr)).all() and float((_flat(tr) - p0).abs().max()) > 0


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
    np.testing.assert_allclose(_flat(tr).cpu().numpy(), _flat(ref).cpu().numpy(), rtol=0, atol=2e-5)


def test_checkpoint_round_trip_in_reference_format():
    """state_dict() is what Trainer.save_checkpoint writes (trainer.py:596-621): nets + a torch.optim.Adam state over the 48
    parameters; it is a snapshot (no live aliases), loads into torch Adam and back, and resumes bit-identically."""
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
    for b in bs[3:]:
        r.step(b)
    torch.cuda.synchronize()
    assert torch.equal(_flat(r), _flat(a)) and torch.equal(r.v_c, a.v_c)
    legacy = mk(); legacy.load_state_dict({"step": 3, "nerf_c": sd["nerf_c"], "nerf_f": sd["nerf_f"], "opt": None})
    assert legacy.adam_t == 0 and float(legacy.m_c.abs().max()) == 0
