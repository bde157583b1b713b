"""BASELINE configs[2]: full 800x800 eval frame (64 + 128 samples, RGB / depth / opacity) over an eval_chunk sweep
{2048, 8192, 16384 (the reference CLI default, scripts/train_nerf.py:159), 65536, full}, through render_image_chunked
(the reference's signature).  One JSON object per line -> profiles/r2_eval_chunk_sweep.jsonl"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_sandbox_b200 as nsb
import bench
dev = torch.device("cuda", 0)
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
tr = nsb.VanillaTrainer(dev, mode=mode, seed=0, sigma_bias=0.3)
o, d, rn = (torch.from_numpy(a).to(dev) for a in bench.frame_rays(np.random.default_rng(0)))
H = W = 800
FLOP = H * W * (64 + 192) * 1_186_816
ref = None
for chunk in (2048, 8192, 16384, 65536, H * W):
    f = lambda: nsb.render_image_chunked(o, d, rn, H, W, 2.0, 6.0, tr.pos_enc, tr.dir_enc, tr.nerf_c, tr.nerf_f, 64, 128, True, dev,
                                         eval_chunk=chunk, viewdirs_world_unit=d)
    out = f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        out = f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if ref is None:
        ref = out["rgb"].clone()
    print(json.dumps({"mode": mode, "eval_chunk": chunk, "launch_sets": -(-H * W // chunk), "ms_per_frame": round(ms, 3), "frames_per_s": round(1e3 / ms, 4),
                      "mlp_tflops": round(FLOP / (ms * 1e-3) / 1e12, 1), "max_abs_diff_vs_chunk2048": float((out["rgb"] - ref).abs().max())}), flush=True)
