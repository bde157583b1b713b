"""Parity of the CUDA path (through the C ABI, via the Python host mirror) against the oracle and the
reference-generated golden vectors.  Tolerances: bit-exact for sample_pdf bin indices, stratified z and
the merge; 1e-4 relative (+1e-6 absolute floor) for composited rgb/depth/opacity in fp32 mode."""
import numpy as np
import pytest
import torch

from conftest import golden, pdf_tolerance
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def T(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV, dtype)


def N(t):
    return t.detach().float().cpu().numpy()


def close(a, b, rtol=1e-4, atol=1e-6):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def load_nerf(seed, sigma_bias, mode="fp32"):
    import nerf_sandbox_b200 as nsb
    p = O.init_params(np.random.default_rng(int(seed)), sigma_bias=float(sigma_bias))
    net = nsb.NeRF(63, 27, 8, 256, skip_pos=4, sigma_activation="relu", mode=mode).to(DEV)
    net.load_state_dict({k: T(v) for k, v in p.items()})
    return net, p


@pytest.fixture(scope="module")
def nsb():
    import nerf_sandbox_b200 as m
    from nerf_sandbox_b200 import _lib
    _lib.lib()                                   # loud failure if the CUDA library is missing
    assert torch.cuda.is_available()
    return m


# ---------------------------------------------------------------------------------------------------- K1 encoder
def test_encoder(nsb):
    g = golden("encoder")
    pe, de = nsb.get_vanilla_nerf_encoders()
    ep = N(pe.to(DEV)(T(g["x"]))); ed = N(de.to(DEV)(T(g["d"])))
    close(ep, g["enc_pos"], 0, 3e-4)             # |arg| up to 3072 rad; sinf vs Sleef ulps
    close(ed, g["enc_dir"], 0, 2e-6)
    assert ep.shape == (96, 63) and ed.shape == (96, 27)
    close(N(pe(T(g["x"]).reshape(8, 12, 3))).reshape(96, 63), ep, 0, 0)      # leading dims preserved
    assert pe(torch.zeros((0, 3), device=DEV)).shape == (0, 63)                # empty input


# ---------------------------------------------------------------------------------------------------- K2 samplers
def test_stratified_bit_exact(nsb):
    from nerf_sandbox_b200 import _lib
    g = golden("sampler")
    L = _lib.lib()
    for i in range(int(g["n_cases"])):
        near, far, nc = float(g[f"near{i}"]), float(g[f"far{i}"]), int(g[f"nc{i}"])
        U = T(g[f"U{i}"])
        z = torch.empty_like(U)
        _lib.check(L.nsb_stratified_z(_lib.ptr(z), _lib.ptr(U), U.shape[0], nc, near, far, 1, 0, 0, _lib.stream()))
        assert np.array_equal(N(z), g[f"z{i}"]), f"case {i}"
        _lib.check(L.nsb_stratified_z(_lib.ptr(z), None, U.shape[0], nc, near, far, 0, 0, 0, _lib.stream()))
        assert np.array_equal(N(z)[3], g[f"zlin{i}"]), f"linspace case {i}"
        # Philox jitter: stays inside its stratum, sorted
        _lib.check(L.nsb_stratified_z(_lib.ptr(z), None, U.shape[0], nc, near, far, 1, 123, 7, _lib.stream()))
        zz = N(z)
        assert (np.diff(zz, axis=-1) >= 0).all() and zz.min() >= near and zz.max() <= far
        assert np.unique(zz[:, 1]).size > 1


def test_sample_pdf_indices_bit_exact_and_values(nsb):
    g = golden("sample_pdf")
    M = g["wb"].shape[1]
    edges = O.pdf_edges(g["bins_mid"], M)
    # north star: bit-exact bin indices given the same CDF and uniforms
    out, inds = nsb.sample_pdf(T(g["bins_mid"]), T(g["wb"]), 128, u=T(g["u"]), cdf=T(g["cdf_mid"]), return_inds=True)
    assert np.array_equal(N(inds).astype(np.int64), g["inds_mid_rand128"])
    assert inds.dtype == torch.int64
    close(N(out), g["out_mid_rand128"], 0, 1e-6)
    out, inds = nsb.sample_pdf(T(g["bins_mid"]), T(g["wb"]), 128, deterministic=True, cdf=T(g["cdf_mid"]), return_inds=True)
    assert np.array_equal(N(inds).astype(np.int64), g["inds_mid_det128"])
    close(N(out), g["out_mid_det128"], 0, 1e-6)
    # own CDF (warp scan): indices may flip where u sits on a CDF value; values continuous
    out, inds = nsb.sample_pdf(T(g["bins_mid"]), T(g["wb"]), 128, u=T(g["u"]), return_inds=True)
    same = N(inds).astype(np.int64) == g["inds_mid_rand128"]
    assert same.mean() > 0.999
    tol = pdf_tolerance(edges, g["cdf_mid"], g["inds_mid_rand128"])
    assert (np.abs(N(out) - g["out_mid_rand128"]) <= tol)[same].all()
    # edges input, deterministic (author's recipe, compare_nerf_repos.py:439-463: OK threshold 1e-5)
    close(N(nsb.sample_pdf(T(g["edges"]), T(g["w_edges"]), 64, deterministic=True)), g["out_edges_det64"], 0, 1e-5)
    close(N(nsb.sample_pdf(T(g["m1_bins"]), T(g["m1_w"]), 8, deterministic=True)), g["out_m1_det8"], 0, 1e-6)
    close(N(nsb.sample_pdf(T(g["bins_mid"]), T(g["wb"]), 1, deterministic=True)), g["out_mid_det1"], 0, 1e-5)
    # Philox uniforms: in range, reproducible with a seed, ~uniform in CDF space
    a = nsb.sample_pdf(T(g["bins_mid"]), T(g["wb"]), 128, seed=5); b = nsb.sample_pdf(T(g["bins_mid"]), T(g["wb"]), 128, seed=5)
    assert torch.equal(a, b) and float(a.min()) >= edges.min() - 1e-6 and float(a.max()) <= edges.max() + 1e-6
    with pytest.raises(ValueError):
        nsb.sample_pdf(T(g["bins_mid"])[:, :10], T(g["wb"]), 8)
    assert nsb.sample_pdf(T(g["bins_mid"])[:0], T(g["wb"])[:0], 8, deterministic=True).shape == (0, 8)


def test_resample_merge(nsb):
    from nerf_sandbox_b200 import _lib
    g = golden("sample_pdf")
    L = _lib.lib()
    zc, w_c, u = T(g["zc"]), T(g["w_c"]), T(g["u"])
    B, Nc = zc.shape
    for det in (0, 1):
        z_all = torch.empty((B, Nc + 128), device=DEV); zf = torch.empty((B, 128), device=DEV)
        _lib.check(L.nsb_resample_merge(_lib.ptr(zc), _lib.ptr(w_c), _lib.ptr(u), _lib.ptr(z_all), _lib.ptr(zf), B, Nc, 128,
                                        det, 0, 0, _lib.stream()))
        ref_zf, ref_inds = O.sample_pdf(g["bins_mid"], g["wb"], 128, deterministic=bool(det), u=g["u"], return_inds=True)
        gold = g["out_mid_det128"] if det else g["out_mid_rand128"]
        gi = g["inds_mid_det128"] if det else g["inds_mid_rand128"]
        tol = pdf_tolerance(O.pdf_edges(g["bins_mid"], Nc - 1), g["cdf_mid"], gi)
        ok = np.abs(N(zf) - gold) <= tol
        assert ok.mean() > 0.995                                   # rest: index flips at CDF ties (see oracle test)
        # the merge itself is exact: z_all == sort(cat(zc, the kernel's own zf))
        assert np.array_equal(N(z_all), np.sort(np.concatenate([g["zc"], N(zf)], -1), -1))
    gs = golden("sampler")
    # pure merge check against the reference's sort(cat) output: feed zf through a flat pdf is not possible,
    # so use the kernel's zf path above; here check the golden merge with torch-free numpy sort equality
    assert np.array_equal(np.sort(np.concatenate([gs["merge_zc"], gs["merge_zf"]], -1), -1), gs["merge_out"])


# ---------------------------------------------------------------------------------------------------- K3 compositor
@pytest.mark.parametrize("tag,white,inf_last,use_rn", [("a", True, True, True), ("b", False, False, True), ("c", True, False, False)])
def test_compositor_fwd_bwd(nsb, tag, white, inf_last, use_rn):
    g = golden("compositor")
    rgb, sig = T(g["rgb"]).requires_grad_(), T(g["sigma"]).requires_grad_()
    rn = T(g["ray_norm"]) if use_rn else None
    comp, w, acc, depth = nsb.volume_render_rays(rgb, sig, T(g["z"]), rn, white, 1e-10, inf_last)
    close(N(comp), g[f"{tag}_comp"]); close(N(w), g[f"{tag}_w"], 1e-4, 1e-7); close(N(acc), g[f"{tag}_acc"])
    close(N(depth), g[f"{tag}_depth"], 1e-4, 1e-5)
    assert comp.shape == (96, 3) and w.shape == (96, 64) and acc.shape == (96, 1) and depth.shape == (96, 1)
    # author's invariant (compare_nerf_repos.py:1199-1204): weights sum to acc
    close(N(w).sum(-1), np.clip(N(acc)[:, 0], 0, 1), 1e-5, 1e-6)
    gi = torch.autograd.grad((comp * T(g[f"{tag}_g_c"])).sum(), [rgb, sig], retain_graph=True)
    close(N(gi[0]), g[f"{tag}_drgb_i"], 1e-4, 1e-6); close(N(gi[1]), g[f"{tag}_dsig_i"], 2e-4, 2e-5)
    gii = torch.autograd.grad((comp * T(g[f"{tag}_g_c"])).sum() + (w * T(g[f"{tag}_g_w"])).sum()
                              + (acc * T(g[f"{tag}_g_a"])).sum() + (depth * T(g[f"{tag}_g_d"])).sum(), [rgb, sig])
    close(N(gii[0]), g[f"{tag}_drgb_ii"], 1e-4, 1e-6); close(N(gii[1]), g[f"{tag}_dsig_ii"], 2e-4, 5e-5)


def test_compositor_ragged_and_long_rays(nsb):
    rng = np.random.default_rng(3)
    # 1..192: one warp per ray; 193..1024: two to four warps per ray; beyond: the strided kernels
    for B, Nn in [(1, 1), (3, 7), (5, 33), (3, 192), (3, 193), (3, 256), (3, 257), (2, 300), (2, 512), (3, 513), (2, 768), (2, 1000), (2, 1024), (2, 1025), (1, 1500)]:
        z = np.sort(rng.uniform(2, 6, (B, Nn)).astype(np.float32), -1)
        rgb = rng.uniform(0, 1, (B, Nn, 3)).astype(np.float32); sig = rng.uniform(0, 3, (B, Nn)).astype(np.float32)
        rn = rng.uniform(1, 1.1, (B, 1)).astype(np.float32)
        comp, w, acc, depth, cache = O.volume_render_rays(rgb, sig, z, rn, True, 1e-10, True, keep=True)
        tr, ts = T(rgb).requires_grad_(), T(sig).requires_grad_()
        c2, w2, a2, d2 = nsb.volume_render_rays(tr, ts, T(z), T(rn), True, 1e-10, True)
        # weights: alpha = 1 - exp(-sigma delta) carries the ABSOLUTE error of exp near 1 (1 ulp = 6e-8), whatever its own size
        close(N(c2), comp); close(N(w2), w, 1e-4, 2.5e-7); close(N(a2), acc); close(N(d2), depth, 1e-4, 1e-5)
        gc = rng.standard_normal((B, 3)).astype(np.float32)
        drgb, dsig = O.volume_render_backward(cache, gc)
        gr = torch.autograd.grad((c2 * T(gc)).sum(), [tr, ts])
        close(N(gr[0]), drgb, 1e-4, 1e-6); close(N(gr[1]), dsig, 5e-4, 2e-5)
    assert nsb.volume_render_rays(torch.zeros((0, 4, 3), device=DEV), torch.zeros((0, 4), device=DEV),
                                  torch.zeros((0, 4), device=DEV))[0].shape == (0, 3)


@pytest.mark.parametrize("Nn", [192, 200, 257, 320, 400, 576, 600, 768, 800, 1024, 1100])
def test_raw_compositor_one_to_four_warps_per_ray(nsb, Nn):
    """The fused raw -> activations -> composite kernels (forward and backward, explicit sigma noise) at sample counts on both
    sides of every kernel boundary (256 / 512 / 768 / 1024) against the oracle's compositor and the chain rule in float64."""
    from nerf_sandbox_b200 import _lib
    L = _lib.lib(); st = _lib.stream()
    rng = np.random.default_rng(Nn)
    B = 5
    raw = rng.standard_normal((B, Nn, 4)).astype(np.float32); raw[..., 3] = raw[..., 3] * 2 + 0.5
    noise = rng.standard_normal((B, Nn)).astype(np.float32)
    z = np.sort(rng.uniform(2, 6, (B, Nn)).astype(np.float32), -1)
    rn = rng.uniform(1, 1.1, (B, 1)).astype(np.float32)
    rgb = 1 / (1 + np.exp(-raw[..., :3].astype(np.float64)))
    pre = raw[..., 3].astype(np.float64) + noise.astype(np.float64) * 0.7
    sig = np.maximum(pre, 0)
    comp, w, acc, depth, cache = O.volume_render_rays(rgb.astype(np.float32), sig.astype(np.float32), z, rn, True, 1e-10, False, keep=True)
    t_raw, t_noise, t_z, t_rn = T(raw.reshape(-1, 4)), T(noise), T(z), T(rn.reshape(-1))
    comp_d = torch.empty(B, 3, device=DEV); w_d = torch.empty(B, Nn, device=DEV); acc_d = torch.empty(B, device=DEV); dep_d = torch.empty(B, device=DEV)
    flags = _lib.WHITE_BKGD | _lib.TRAINING
    _lib.check(L.nsb_composite_raw_fwd(_lib.ptr(t_raw), _lib.ptr(t_noise), 0.7, _lib.ptr(t_z), _lib.ptr(t_rn), _lib.ptr(comp_d), _lib.ptr(w_d),
                                       _lib.ptr(acc_d), _lib.ptr(dep_d), B, Nn, flags, 0, 0, st))
    close(N(comp_d), comp); close(N(w_d), w, 1e-4, 2.5e-7); close(N(acc_d), acc.reshape(-1)); close(N(dep_d), depth.reshape(-1), 1e-4, 1e-5)
    gc = rng.standard_normal((B, 3)).astype(np.float32)
    drgb, dsig = O.volume_render_backward(cache, gc)
    d_ref = np.concatenate([drgb * rgb * (1 - rgb), (dsig * (pre > 0))[..., None]], -1)
    d_raw = torch.empty(B * Nn, 4, device=DEV)
    _lib.check(L.nsb_composite_raw_bwd(_lib.ptr(t_raw), _lib.ptr(t_noise), 0.7, _lib.ptr(t_z), _lib.ptr(t_rn), _lib.ptr(T(gc)), _lib.ptr(d_raw),
                                       B, Nn, flags, 0, 0, st))
    got = N(d_raw).reshape(B, Nn, 4)
    close(got[..., :3], d_ref[..., :3], 2e-4, 1e-6); close(got[..., 3], d_ref[..., 3], 5e-4, 2e-5)


# ---------------------------------------------------------------------------------------------------- K1 MLP (fp32 mode)
def test_mlp_forward_backward_fp32(nsb):
    g = golden("mlp")
    net, p = load_nerf(g["seed"], g["sigma_bias"])
    out = net(T(g["enc_pos"]), T(g["enc_dir"]))
    close(N(out), g["out"], 1e-4, 2e-5)
    out.backward(T(g["d_out"]))
    flat = torch.cat([q.grad.reshape(-1) for q in net.parameters()])
    ref = g["grad_samples"]
    assert np.abs(N(flat)[g["grad_idx"]] - ref).max() <= 2e-4 * np.abs(ref).max()
    norms = np.array([float(q.grad.norm()) for q in net.parameters()])
    close(norms, g["grad_norms"], 2e-4, 1e-6)
    with torch.no_grad():       # eval path (no stash): the fp16-split tensor-core forward, fp32-accurate, same bar vs the reference
        close(N(net(T(g["enc_pos"]), T(g["enc_dir"]))), g["out"], 1e-4, 2e-5)
    with pytest.raises(RuntimeError):
        net(T(g["enc_pos"])[:, :60], T(g["enc_dir"]))


@pytest.mark.parametrize("tag", ["train", "eval", "bare"])
def test_nerf_forward_pass_fp32(nsb, tag):
    g = golden("forward_pass")
    net, _ = load_nerf(g["seed"], g["sigma_bias"])
    pe, de = nsb.get_vanilla_nerf_encoders()
    kw = dict(train=dict(ray_norms=T(g["rays_d_marching_norm"]), viewdirs_world_unit=T(g["viewdirs"]), raw_noise=T(g["noise"]),
                         raw_noise_std=1.0, training=True, infinite_last_bin=True, white_bkgd=True),
              eval=dict(ray_norms=T(g["rays_d_marching_norm"]), viewdirs_world_unit=T(g["viewdirs"]),
                        infinite_last_bin=False, white_bkgd=False),
              bare=dict(ray_norms=None, viewdirs_world_unit=None, infinite_last_bin=True, white_bkgd=True, mlp_chunk=100))[tag]
    comp, w, acc, depth = nsb.nerf_forward_pass(T(g["rays_o_marching"]), T(g["rays_d_marching_unit"]), T(g["z"]),
                                                pos_enc=pe.to(DEV), dir_enc=de.to(DEV), nerf=net, **kw)
    close(N(comp), g[f"{tag}_comp"]); close(N(w), g[f"{tag}_w"], 1e-4, 1e-6)
    close(N(acc), g[f"{tag}_acc"]); close(N(depth), g[f"{tag}_depth"], 1e-4, 1e-5)
    assert w.shape == (24, 64) and acc.shape == (24, 1) and depth.shape == (24, 1)


def _train_inputs(g):
    batch = {k: T(g[k]) for k in ("rays_o_marching", "rays_d_marching_unit", "rays_d_marching_norm", "rays_d_world_unit", "rgb")}
    draws = dict(U=T(g["U"]), u_fine=T(g["u_fine"]), noise_c=T(g["noise_c"]), noise_f=T(g["noise_f"]))
    return batch, draws


def _make_trainer(nsb, g, mode="fp32"):
    tr = nsb.VanillaTrainer(DEV, nc=int(g["nc"]), nf=int(g["nf"]), near=2.0, far=6.0, mode=mode)
    for net, seed in ((tr.nerf_c, g["seed_c"]), (tr.nerf_f, g["seed_f"])):
        p = O.init_params(np.random.default_rng(int(seed)), sigma_bias=float(g["sigma_bias"]))
        net.load_state_dict({k: T(v) for k, v in p.items()})
    return tr


def test_train_step_fp32_matches_reference(nsb):
    g = golden("train_step")
    tr = _make_trainer(nsb, g)
    batch, draws = _train_inputs(g)
    out = tr._train_step(batch, draws)                   # reference contract: autograd loss
    assert abs(float(out["loss"]) - float(g["loss"])) <= 1e-4 * float(g["loss"])
    assert abs(float(out["psnr"]) - float(g["psnr"])) <= 1e-3
    close(N(out["comp_c"]), g["comp_c"]); close(N(out["comp_f"]), g["comp_f"])
    out["loss"].backward()
    for tag, net in (("c", tr.nerf_c), ("f", tr.nerf_f)):
        flat = N(torch.cat([q.grad.reshape(-1) for q in net.parameters()]))
        ref = g[f"grad_samples_{tag}"]
        assert np.abs(flat[g["grad_idx"]] - ref).max() <= 2e-3 * np.abs(ref).max()
        norms = np.array([float(q.grad.norm()) for q in net.parameters()])
        close(norms, g[f"grad_norms_{tag}"], 5e-3 if tag == "f" else 1e-3, 1e-7)   # fp32 noise floor, see oracle test
    # torch Adam over the module parameters works unchanged (trainer.py:383-386) and re-packs the weights
    opt = torch.optim.Adam(tr.parameters(), lr=5e-4)
    opt.step()
    l2 = float(tr._train_step(batch, draws)["loss"])
    assert l2 < float(out["loss"])


def test_fused_step_and_adam_fp32(nsb):
    g = golden("train_step")
    tr = _make_trainer(nsb, g)
    batch, draws = _train_inputs(g)
    p0 = tr.nerf_c.flat_params().clone()
    sc = tr.step(batch, draws)
    assert abs(float(sc[0]) - float(g["loss"])) <= 1e-4 * float(g["loss"])
    # Adam step 1 fed with (almost) the reference's grads: compare where |g| is well above the eps knee
    idx = g["grad_idx"]
    ref_g = g["grad_samples_c"]
    well = np.abs(ref_g) > 1e-6
    p1 = N(tr.nerf_c.flat_params())[idx]
    close(p1[well], g["adam_c"][well], 0, 2e-6)
    assert np.abs(p1 - N(p0)[idx]).max() <= 5.01e-4
    losses = [float(tr.step(batch, draws)[0]) for _ in range(5)]
    assert losses[-1] < float(g["loss"])                   # it trains
    sd = tr.state_dict()
    assert tuple(sd["nerf_c"]["mlp.4.weight"].shape) == (256, 319) and float(sd["opt"]["state"][0]["step"]) == 6.0
    assert tuple(sd["opt"]["state"][8]["exp_avg"].shape) == (256, 319) and len(sd["opt"]["state"]) == 48     # mlp.4.weight


@pytest.mark.parametrize("tag,ilb,nf", [("fine", False, 128), ("fine_inf", True, 128), ("coarse_only", False, 0)])
def test_eval_tile_fp32(nsb, tag, ilb, nf):
    g = golden("eval_tile")
    nc_, _ = load_nerf(g["seed_c"], g["sigma_bias"]); nf_, _ = load_nerf(g["seed_f"], g["sigma_bias"])
    pe, de = nsb.get_vanilla_nerf_encoders()
    H, W = int(g["H"]), int(g["W"])
    r = nsb.render_image_chunked(T(g["rays_o_marching"]), T(g["rays_d_marching_unit"]), T(g["rays_d_marching_norm"]), H, W,
                                 2.0, 6.0, pe.to(DEV), de.to(DEV), nc_, nf_, 64, nf, True, torch.device(DEV), eval_chunk=20,
                                 viewdirs_world_unit=T(g["rays_d_world_unit"]), infinite_last_bin=ilb)
    assert r["rgb"].shape == (H, W, 3) and r["acc"].shape == (H, W, 1) and r["depth"].shape == (H, W, 1)
    close(N(r["rgb"]), g[f"{tag}_rgb"]); close(N(r["acc"]), g[f"{tag}_acc"]); close(N(r["depth"]), g[f"{tag}_depth"], 1e-4, 1e-4)


# ---------------------------------------------------------------------------------------------------- full-size properties
def test_full_size_properties_fp32(nsb):
    """BASELINE cfg2 shape (1024 rays, 64+128) with in-kernel Philox: size-independent invariants."""
    from nerf_sandbox_b200 import _lib
    tr = nsb.VanillaTrainer(DEV, sigma_bias=0.3, seed=1)
    rays = O.synthetic_rays(np.random.default_rng(2), 1024)
    batch = {k: T(v) for k, v in rays.items()}
    s1 = N(tr._fwd_bwd(batch)[0]).copy(); g1 = tr.grads_f.clone()
    s2 = N(tr._fwd_bwd(batch)[0]).copy()
    assert s1[0] == pytest.approx(s2[0], rel=1e-5)                       # same (seed, step) -> same Philox draws
    assert np.isfinite(s1).all() and abs(s1[0] - (s1[2] + s1[3])) < 1e-6
    assert torch.isfinite(g1).all() and float(g1.abs().max()) > 0
    # linearity of the backward in grad_scale
    tr._fwd_bwd(batch, grad_scale=0.5)
    close(N(tr.grads_f), 0.5 * N(g1), 2e-3, 1e-9)
    # eval: 4096 rays, fine z sorted, acc in [0,1], rgb in [0,1]; coarse-only equals a 1-pass render
    o, d, rn = batch["rays_o_marching"].repeat(4, 1), batch["rays_d_marching_unit"].repeat(4, 1), batch["rays_d_marching_norm"].repeat(4, 1).reshape(-1)
    rgb, acc, depth = nsb.render_rays(o, d, rn, d, tr.nerf_c, tr.nerf_f, near=2.0, far=6.0, nc=64, nf=128, white_bkgd=True)
    assert float(rgb.min()) >= 0 and float(rgb.max()) <= 1 and float(acc.min()) >= 0 and float(acc.max()) <= 1
    assert torch.equal(rgb[:1024], rgb[1024:2048])                        # deterministic and tile-independent
    assert float(depth.min()) >= 0 and float(depth.max()) <= 6.0 + 1e-3


def test_reference_style_step_through_module_boundary_fp32(nsb):
    """Trainer._train_step written exactly like the reference (trainer.py:899-1006) but against this package's
    by-name callables -- the drop-in seam -- with autograd flowing through nerf_forward_pass."""
    import torch.nn.functional as F
    g = golden("train_step")
    tr = _make_trainer(nsb, g)
    batch, draws = _train_inputs(g)
    pe, de = tr.pos_enc, tr.dir_enc
    B, nc, nf = int(g["B"]), int(g["nc"]), int(g["nf"])
    t = torch.linspace(0.0, 1.0, steps=nc, device=DEV)
    zc = (2.0 * (1.0 - t) + 6.0 * t).expand(B, nc).contiguous()
    mids = 0.5 * (zc[:, 1:] + zc[:, :-1])
    lower = torch.cat([zc[:, :1], mids], -1); upper = torch.cat([mids, zc[:, -1:]], -1)
    zc = torch.sort(lower + (upper - lower) * draws["U"], -1).values
    kw = dict(pos_enc=pe, dir_enc=de, white_bkgd=True, ray_norms=batch["rays_d_marching_norm"],
              viewdirs_world_unit=batch["rays_d_world_unit"], sigma_activation="relu", raw_noise_std=1.0, training=True,
              infinite_last_bin=True)
    comp_c, w_c, _, _ = nsb.nerf_forward_pass(batch["rays_o_marching"], batch["rays_d_marching_unit"], zc, nerf=tr.nerf_c,
                                              raw_noise=draws["noise_c"], **kw)
    bins_mid = 0.5 * (zc[:, 1:] + zc[:, :-1]); wb = 0.5 * (w_c[:, 1:] + w_c[:, :-1]).detach() + 1e-5
    zf = nsb.sample_pdf(bins_mid, wb, nf, deterministic=False, u=draws["u_fine"])
    z_all = torch.sort(torch.cat([zc, zf], -1), -1).values
    comp_f, _, _, _ = nsb.nerf_forward_pass(batch["rays_o_marching"], batch["rays_d_marching_unit"], z_all, nerf=tr.nerf_f,
                                            raw_noise=draws["noise_f"], **kw)
    loss = F.mse_loss(comp_c.clamp(0, 1), batch["rgb"]) + F.mse_loss(comp_f.clamp(0, 1), batch["rgb"])
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-4 * float(g["loss"])
    loss.backward()
    for tag, net in (("c", tr.nerf_c), ("f", tr.nerf_f)):
        norms = np.array([float(q.grad.norm()) for q in net.parameters()])
        close(norms, g[f"grad_norms_{tag}"], 5e-3 if tag == "f" else 1e-3, 1e-7)


def test_camera_rays_and_render_pose(nsb):
    g = golden("rays")
    for name in g["names"]:
        name = str(name)
        px = g.get(f"{name}_px")
        out = nsb.get_camera_rays(int(g[f"{name}_H"]), int(g[f"{name}_W"]), g[f"{name}_K"], g[f"{name}_c2w"], device=DEV,
                                  convention=str(g[f"{name}_conv"]), pixel_center=bool(g[f"{name}_pc"]), as_ndc=bool(g[f"{name}_ndc"]),
                                  near_plane=float(g[f"{name}_near"]), pixels_xy=px)
        assert len(out) == 6
        for i, a in enumerate(out):
            ref = g[f"{name}_out{i}"]
            assert tuple(a.shape) == ref.shape
            close(N(a), ref, 2e-6, 2e-6)
    with pytest.raises(ValueError):
        nsb.get_camera_rays(4, 4, np.eye(3, dtype=np.float32), np.eye(4, dtype=np.float32), device=DEV, convention="nope")
    with pytest.raises(ValueError):
        nsb.get_camera_rays(4, 4, np.eye(2, dtype=np.float32), np.eye(4, dtype=np.float32), device=DEV)
    # render_pose == render_image_chunked on the generated rays (world and NDC marching)
    nc_, _ = load_nerf(31, 1.0); nf_, _ = load_nerf(32, 1.0)
    pe, de = nsb.get_vanilla_nerf_encoders()
    K, c2w = g["blender_K"], g["blender_c2w"]
    r = nsb.render_pose(c2w, 12, 10, K, 2.0, 6.0, pe.to(DEV), de.to(DEV), nc_, nf_, DEV, nc_eval=64, nf_eval=128, eval_chunk=50)
    w = nsb.get_camera_rays(12, 10, K, c2w, device=DEV, pixel_center=True)
    ref = O.render_rays_eval({k: N(v) for k, v in nc_.state_dict().items()}, {k: N(v) for k, v in nf_.state_dict().items()},
                             N(w[0]), N(w[1]), N(w[2]), N(w[1]), near=2.0, far=6.0, nc=64, nf=128)
    close(N(r["rgb"]).reshape(-1, 3), ref["rgb"]); close(N(r["depth"]).reshape(-1, 1), ref["depth"], 1e-4, 1e-4)
    r2 = nsb.render_pose(g["llff_ndc_c2w"], 12, 16, g["llff_ndc_K"], 1.0, 6.0, pe, de, nc_, nf_, DEV, use_ndc=True)
    assert r2["rgb"].shape == (12, 16, 3) and torch.isfinite(r2["rgb"]).all() and torch.isfinite(r2["depth"]).all()


def test_device_pixel_sampler_matches_reference_semantics(nsb):
    """RandomPixelRaySampler (data/samplers.py:134-290) on device: given the pixels it drew, rgb (white composite) and
    rays must equal the reference's computation; draws respect the precrop window and cover it uniformly."""
    from types import SimpleNamespace
    rng = np.random.default_rng(3)
    H, W, F = 40, 50, 3
    g = golden("rays")
    frames = []
    for f in range(F):
        img = rng.uniform(0, 1, (H, W, 4)).astype(np.float32)
        K = np.array([[60.0, 0, W / 2], [0, 61.0, H / 2], [0, 0, 1]], dtype=np.float32)
        c2w = g["blender_c2w"].copy(); c2w[:3, 3] += f
        frames.append(SimpleNamespace(image=img, K=K, c2w=c2w))
    scene = SimpleNamespace(frames=frames, white_bkgd=True)
    for single, ndc in ((True, False), (False, True)):
        smp = nsb.RandomPixelRaySampler(scene, rays_per_batch=4096, device=DEV, sample_from_single_frame=single,
                                        precrop_iters=1, precrop_frac=0.5, as_ndc=ndc, near_plane=1.0, seed=5)
        b = smp.next_batch(with_pixels=True)
        px, fid = N(b["pixels_xy"]).astype(int), N(b["frame_ids"]).astype(int)
        assert px[:, 0].min() >= W // 4 and px[:, 0].max() < 3 * W // 4 + 1 and px[:, 1].min() >= H // 4 and px[:, 1].max() < 3 * H // 4 + 1
        assert (len(np.unique(fid)) == 1) == single
        assert len(np.unique(px[:, 0])) >= (W // 2) - 1          # covers the crop window
        for f in np.unique(fid):
            m = fid == f
            pix = frames[f].image[px[m, 1], px[m, 0]]
            np.testing.assert_allclose(N(b["rgb"])[m], pix[:, :3] * pix[:, 3:4] + (1 - pix[:, 3:4]), rtol=0, atol=1e-6)
            ref = O.camera_rays(H, W, frames[f].K, frames[f].c2w, pixel_center=True, as_ndc=ndc, near_plane=1.0,
                                pixels_xy=px[m].astype(np.float32))
            for key, r in zip(("rays_o_world", "rays_d_world_unit", "rays_d_world_norm", "rays_o_marching", "rays_d_marching_unit",
                               "rays_d_marching_norm"), ref):
                close(N(b[key])[m], r, 2e-6, 2e-6)
        b2 = smp.next_batch(with_pixels=True)                     # precrop over: full frame
        assert N(b2["pixels_xy"])[:, 0].max() > 3 * W // 4 and not torch.equal(b2["pixels_xy"], b["pixels_xy"])
    # feeds the trainer directly (device-resident loop, no host sync)
    tr = nsb.VanillaTrainer(DEV, mode="fp32", sigma_bias=0.3)
    smp = nsb.RandomPixelRaySampler(scene, rays_per_batch=256, device=DEV, sample_from_single_frame=True)
    sc = tr.step(smp.next_batch())
    assert torch.isfinite(sc).all()


def test_llff_ndc_shape_train_step_fp32(nsb):
    """BASELINE configs[3] shape (LLFF fern 504x378, NDC marching rays, z in [0,1], viewdirs = world dirs) against the oracle."""
    rng = np.random.default_rng(17)
    H, W, f = 378, 504, 407.6
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32)
    c2w = np.array([[1, 0, 0, 0.05], [0, 1, 0, -0.03], [0, 0, 1, 0.02]], dtype=np.float32)
    px = np.stack([rng.integers(0, W, 40), rng.integers(0, H, 40)], -1).astype(np.float32)
    r = O.camera_rays(H, W, K, c2w, pixel_center=True, as_ndc=True, near_plane=1.0, pixels_xy=px)
    assert abs(float(np.median(r[5])) - 2.0) < 0.5                        # NDC rays span z in [0,1]: |d_ndc| ~ 2
    B, nc, nf = 40, 64, 128
    batch = dict(rays_o_marching=r[3], rays_d_marching_unit=r[4], rays_d_marching_norm=r[5], rays_d_world_unit=r[1],
                 rgb=rng.uniform(0, 1, (B, 3)).astype(np.float32))
    draws = dict(U=rng.uniform(0, 1, (B, nc)).astype(np.float32), u_fine=rng.uniform(0, 1, (B, nf)).astype(np.float32),
                 noise_c=rng.standard_normal(B * nc).astype(np.float32), noise_f=rng.standard_normal(B * (nc + nf)).astype(np.float32))
    tr = nsb.VanillaTrainer(DEV, nc=nc, nf=nf, near=0.0, far=1.0, mode="fp32", sigma_bias=2.0)
    pc = {k: N(v) for k, v in tr.nerf_c.state_dict().items()}; pf = {k: N(v) for k, v in tr.nerf_f.state_dict().items()}
    ref = O.train_step(pc, pf, batch, near=0.0, far=1.0, nc=nc, nf=nf, **draws)
    out = tr._train_step({k: T(v) for k, v in batch.items()}, {k: T(v) for k, v in draws.items()})
    assert abs(float(out["loss"].detach()) - float(ref["loss"])) <= 1e-4 * float(ref["loss"])
    close(N(out["comp_f"]), ref["comp_f"]); close(N(out["comp_c"]), ref["comp_c"])
    out["loss"].backward()
    for tag, net in (("c", tr.nerf_c), ("f", tr.nerf_f)):
        got = N(torch.cat([q.grad.reshape(-1) for q in net.parameters()]))
        want = O.flatten_params(ref[f"grads_{tag}"])
        assert np.linalg.norm(got - want) <= (2e-2 if tag == "f" else 2e-3) * np.linalg.norm(want)


def test_frame_outputs_match_reference_image_path():
    """uint8 images as save_rgb_png / save_gray_png write them (render_utils.py:28-47), depth normalisation of
    validation_renderer.py:491-492 and _compute_psnr (:171-196)."""
    import nerf_sandbox_b200 as nsb
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(11)
    H, W = 37, 53
    rgb = torch.rand((H, W, 3), device=dev, generator=g) * 1.3 - 0.15           # values outside [0,1] exercise the clamps
    acc = torch.rand((H, W, 1), device=dev, generator=g) * 1.2 - 0.1
    depth = torch.rand((H, W, 1), device=dev, generator=g) * 6 + 1
    gt = torch.rand((H, W, 3), device=dev, generator=g)
    mask = (torch.rand((H, W, 1), device=dev, generator=g) > 0.3).float()
    out = nsb.frame_outputs({"rgb": rgb, "acc": acc, "depth": depth}, near=2.0, far=6.0, gt_rgb=gt, mask=mask)
    u8 = lambda t: (t.clamp(0, 1).cpu().numpy() * 255.0 + 0.5).astype(np.uint8)
    assert np.array_equal(out["rgb"].cpu().numpy(), u8(rgb))
    assert np.array_equal(out["opacity"].cpu().numpy(), u8(acc.squeeze(-1)))
    dn = ((depth.squeeze(-1) - 2.0) / (6.0 - 2.0 + 1e-8)).clamp(0, 1)
    assert np.array_equal(out["depth"].cpu().numpy(), u8(dn))            # byte-exact: the kernel divides like the reference
    for near, far in ((0.1, 7.3), (1.2345, 3.21)):                       # near/far that are not fp32-representable doubles
        o2 = nsb.frame_outputs({"rgb": rgb, "acc": acc, "depth": depth}, near=near, far=far)
        dn2 = ((depth.squeeze(-1) - near) / (far - near + 1e-8)).clamp(0, 1)
        assert np.array_equal(o2["depth"].cpu().numpy(), u8(dn2))
    pred, gtc = rgb.clamp(0, 1), gt.clamp(0, 1)
    mse = (((pred - gtc) ** 2) * mask).sum() / (mask.sum() * 3).clamp_min(1e-8)
    psnr = -10.0 * torch.log10(mse.clamp_min(1e-10))
    assert abs(float(out["psnr"]) - float(psnr)) <= 1e-4 * abs(float(psnr))
    assert abs(nsb.compute_psnr(rgb, gt) - float(-10.0 * torch.log10(torch.nn.functional.mse_loss(pred, gtc)))) <= 1e-3
    nd = nsb.frame_outputs({"rgb": rgb, "acc": acc, "depth": depth / 7.0}, near=0.0, far=1.0, use_ndc=True)
    assert np.array_equal(nd["depth"].cpu().numpy(), u8((depth / 7.0).squeeze(-1)))


@pytest.mark.parametrize("training", [True, False])
def test_softplus_sigma_fused_matches_torch_composition(nsb, training):
    """sigma_activation='softplus' (render_utils.py:243-244) fused into the compositor kernels against the generic
    composition (our NeRF.forward + torch.sigmoid / F.softplus + volume_render_rays), values and parameter gradients."""
    g = golden("forward_pass")
    net, _ = load_nerf(g["seed"], g["sigma_bias"])
    pe, de = nsb.get_vanilla_nerf_encoders()
    pe, de = pe.to(DEV), de.to(DEV)

    class Foreign(torch.nn.Module):            # same network, but not recognised as the fused triplet
        def __init__(self, inner):
            super().__init__(); self.inner = inner
        def forward(self, a, b):
            return self.inner(a, b)

    kw = dict(ray_norms=T(g["rays_d_marching_norm"]), viewdirs_world_unit=T(g["viewdirs"]), white_bkgd=True, infinite_last_bin=True,
              sigma_activation="softplus", training=training, raw_noise_std=1.0 if training else 0.0,
              raw_noise=T(g["noise"]) if training else None)
    args = (T(g["rays_o_marching"]), T(g["rays_d_marching_unit"]), T(g["z"]))
    tgt = torch.rand((24, 3), device=DEV, generator=torch.Generator(device=DEV).manual_seed(0))
    outs, grads = [], []
    for nerf in (net, Foreign(net)):
        net.zero_grad()
        comp, w, acc, depth = nsb.nerf_forward_pass(*args, pos_enc=pe, dir_enc=de, nerf=nerf, **kw)
        ((comp - tgt) ** 2).mean().backward()
        outs.append([N(comp), N(w), N(acc), N(depth)])
        grads.append(torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone())
    for a, b in zip(*outs):
        close(a, b, 1e-4, 1e-6)
    # softplus keeps every sample alive (unlike relu), so the two differ from the relu goldens
    assert np.abs(outs[0][0] - g["train_comp" if training else "eval_comp"]).max() > 1e-4 or not training
    ga, gb = N(grads[0]).astype(np.float64), N(grads[1]).astype(np.float64)
    assert np.linalg.norm(ga - gb) <= 2e-4 * np.linalg.norm(gb)


def test_resample_in_kernel_draws_are_sorted_uniform_order_statistics(nsb):
    """Without explicit u the resampler draws the sorted uniforms directly (normalised partial sums of exponentials = the
    order statistics of iid uniforms).  Check against the draw-then-sort path with explicit iid u on the same PDF: same
    pooled distribution of fine samples, same mean of every order statistic; merged rows sorted and complete."""
    from nerf_sandbox_b200 import _lib
    L = _lib.lib()
    B, Nc, Nf = 8192, 64, 128
    rng = np.random.default_rng(3)
    zc1 = O.stratified_z(2.0, 6.0, Nc, rng.uniform(0, 1, (1, Nc)).astype(np.float32))
    w1 = (np.exp(-0.5 * ((zc1 - 3.7) / 0.35) ** 2) + 0.02).astype(np.float32)          # one peaked PDF shared by all rays
    zc, w = T(np.repeat(zc1, B, 0)), T(np.repeat(w1, B, 0))
    outs = []
    for u in (None, T(rng.uniform(0, 1, (B, Nf)).astype(np.float32))):
        z_all = torch.empty((B, Nc + Nf), device=DEV); z_f = torch.empty((B, Nf), device=DEV)
        _lib.check(L.nsb_resample_merge(_lib.ptr(zc), _lib.ptr(w), _lib.ptr(u), _lib.ptr(z_all), _lib.ptr(z_f), B, Nc, Nf, 0, 123, 7,
                                        _lib.stream()), "resample")
        za, zf = N(z_all), np.sort(N(z_f), axis=1)
        assert (np.diff(za, axis=1) >= 0).all()
        assert np.array_equal(np.sort(np.concatenate([N(zc), N(z_f)], 1), axis=1), za)
        outs.append(zf)
    a, b = outs
    assert not np.array_equal(a[0], a[1])                                              # rays draw independently
    ha, _ = np.histogram(a, bins=64, range=(2.0, 6.0)); hb, _ = np.histogram(b, bins=64, range=(2.0, 6.0))
    assert np.abs(ha - hb).max() / a.size < 2e-3                                       # pooled distribution
    se = np.sqrt(a.var(0) / B + b.var(0) / B)                                          # every order statistic: same mean ...
    assert (np.abs(a.mean(0) - b.mean(0)) / se).max() < 5.0
    assert np.abs(a.std(0) / b.std(0) - 1.0).max() < 0.06                              # ... and spread (sd of a sample sd ~ 0.8 %)
