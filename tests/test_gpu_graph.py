"""step_graph (nsb_train_step captured into a CUDA graph: step count, Philox streams and Adam bias corrections in device
memory) against the eager step() on the same batches."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu


def _batches(n, rays):
    dev = torch.device("cuda", 0)
    out = []
    for s in range(n):
        r = O.synthetic_rays(np.random.default_rng(50 + s), rays)
        out.append({k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in r.items()})
    return out


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_graph_step_matches_eager_step(mode):
    import nerf_sandbox_b200 as nsb
    dev = torch.device("cuda", 0)
    batches = _batches(6, 256)
    eager = nsb.VanillaTrainer(dev, mode=mode, seed=3, sigma_bias=0.4)
    graph = nsb.VanillaTrainer(dev, mode=mode, seed=3, sigma_bias=0.4)
    le, lg = [], []
    for b in batches:
        le.append(eager.step(b).clone())
        lg.append(graph.step_graph(b).clone())
    assert graph.adam_t == eager.adam_t == 6 and int(graph._step_dev.item()) == 6
    le, lg = torch.stack(le).cpu(), torch.stack(lg).cpu()
    # same draws (Philox streams follow the device counter), same bias corrections: the loss curves coincide; in bf16 mode
    # up to the run-to-run noise of the atomically accumulated weight gradients
    tol = 2e-4 if mode == "fp32" else 2e-3       # fp32: the weight gradients are also accumulated with float atomics
    assert torch.allclose(le[:, 0], lg[:, 0], rtol=tol, atol=tol), (le[:, 0], lg[:, 0])
    pe = torch.cat([eager.nerf_c.flat_params(), eager.nerf_f.flat_params()])
    pg = torch.cat([graph.nerf_c.flat_params(), graph.nerf_f.flat_params()])
    if mode == "fp32":
        # Adam turns last-bit differences of near-zero gradients (atomic accumulation order) into +-lr steps for a few
        # parameters in the first iterations, so bound the bulk, not the maximum
        diff = (pe - pg).abs()
        assert float(diff.median()) <= 1e-5 and float((diff > 2e-4).float().mean()) <= 5e-3, (float(diff.median()), float(diff.max()))
    # the packed weights the next forward uses follow the updated parameters
    b = batches[0]
    args = (b["rays_o_marching"], b["rays_d_marching_unit"], b["rays_d_marching_norm"].reshape(-1), b["rays_d_world_unit"])
    r1 = nsb.render_rays(*args, graph.nerf_c, graph.nerf_f, near=2.0, far=6.0, nc=64, nf=128, white_bkgd=True)[0]
    graph.nerf_c.packed(force=True); graph.nerf_f.packed(force=True)
    r2 = nsb.render_rays(*args, graph.nerf_c, graph.nerf_f, near=2.0, far=6.0, nc=64, nf=128, white_bkgd=True)[0]
    assert torch.equal(r1, r2)


def test_graph_step_mixes_with_eager_steps():
    import nerf_sandbox_b200 as nsb
    dev = torch.device("cuda", 0)
    batches = _batches(4, 128)
    a = nsb.VanillaTrainer(dev, mode="fp32", seed=1, sigma_bias=0.4)
    b = nsb.VanillaTrainer(dev, mode="fp32", seed=1, sigma_bias=0.4)
    la = [a.step(x).clone() for x in batches]
    lb = [b.step(batches[0]).clone(), b.step_graph(batches[1]).clone(), b.step(batches[2]).clone(), b.step_graph(batches[3]).clone()]
    assert torch.allclose(torch.stack(la)[:, 0], torch.stack(lb)[:, 0], rtol=2e-4, atol=1e-6)


def test_cosine_schedule_graph_matches_eager_and_torch():
    """lr_scheduler='cosine' (make_scheduler, trainer.py:81-88): the device-side schedule of the graph step, the host-side
    one of the eager step and torch's CosineAnnealingLR agree."""
    import nerf_sandbox_b200 as nsb
    dev = torch.device("cuda", 0)
    kw = dict(mode="fp32", seed=2, sigma_bias=0.4, lr=5e-4, lr_scheduler="cosine", lr_scheduler_params={"T_max": 5, "eta_min": 1e-5})
    a, b = nsb.VanillaTrainer(dev, **kw), nsb.VanillaTrainer(dev, **kw)
    opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=5e-4)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=5, eta_min=1e-5)
    batches = _batches(5, 128)
    la, lb = [], []
    for x in batches:
        assert abs(a.current_lr() - sched.get_last_lr()[0]) <= 1e-9
        la.append(a.step(x).clone()); lb.append(b.step_graph(x).clone())
        opt.step(); sched.step()
    assert torch.allclose(torch.stack(la)[:, 0], torch.stack(lb)[:, 0], rtol=2e-4, atol=1e-6)
    pa = torch.cat([a.nerf_c.flat_params(), a.nerf_f.flat_params()]); pb = torch.cat([b.nerf_c.flat_params(), b.nerf_f.flat_params()])
    assert float((pa - pb).abs().median()) <= 1e-5
    # and the decayed rate really is applied: a constant-lr trainer moves further
    c = nsb.VanillaTrainer(dev, mode="fp32", seed=2, sigma_bias=0.4, lr=5e-4)
    p0 = torch.cat([c.nerf_c.flat_params(), c.nerf_f.flat_params()]).clone()
    for x in batches:
        c.step(x)
    pc = torch.cat([c.nerf_c.flat_params(), c.nerf_f.flat_params()])
    assert float((pc - p0).abs().mean()) > 1.2 * float((pa - p0).abs().mean())
