"""Per-role cycle accounting of the tc forward kernel (CTA 0) via the debug hook."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_sandbox_b200 as nsb
from nerf_sandbox_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
N = 192; dev = "cuda"; L = _lib.lib()
STASH = len(sys.argv) > 2 and sys.argv[2] == "stash"
fn = L.nsb_debug_tc_layer; fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p] * 8 + [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
net = nsb.NeRF(63, 27, mode="bf16").to(dev)
o = torch.randn(B, 3, device=dev); d = torch.nn.functional.normalize(torch.randn(B, 3, device=dev), dim=-1)
z = torch.sort(torch.rand(B, N, device=dev) * 4 + 2, -1).values.contiguous(); rn = torch.ones(B, device=dev)
raw = torch.empty(B * N, 4, device=dev); cyc = torch.zeros(32 + (B * N // 128 * 684032 // 8 if STASH else 0), dtype=torch.int64, device=dev)
for _ in range(2):
    _lib.check(fn(_lib.ptr(o), _lib.ptr(d), _lib.ptr(z), _lib.ptr(rn), _lib.ptr(d), _lib.ptr(net.packed()), _lib.ptr(raw), cyc.data_ptr(), -2 if STASH else -1, B, N, _lib.stream()))
torch.cuda.synchronize()
c = cyc[:32].cpu().tolist()
pairs = (B * N // 128 + 1) // 2; per_cta = -(-pairs // 148)
print(f"pairs/CTA {per_cta}; layers {per_cta*10}")
print(f"producer: wait_empty {c[0]} of {c[1]} ({100*c[0]/max(c[1],1):.1f}%)")
print(f"mma: wait_in {c[2]} ({100*c[2]/max(c[4],1):.1f}%) wait_full {c[3]} ({100*c[3]/max(c[4],1):.1f}%) total {c[4]}  -> per pair-layer {c[4]/(per_cta*10):.0f} cyc")
print(f"mma (leader of a CTA pair): wait peer_full {c[13]} ({100*c[13]/max(c[4],1):.1f}%) wait peer_in {c[14]} ({100*c[14]/max(c[4],1):.1f}%)")
print(f"peer CTA: producer wait_empty {c[16]} of {c[17]} ({100*c[16]/max(c[17],1):.1f}%); relay wait full {c[18]} ({100*c[18]/max(c[20],1):.1f}%) wait in {c[19]} ({100*c[19]/max(c[20],1):.1f}%) of {c[20]}")
print(f"local load latency when waited: {c[21]/max(c[22],1):.0f} cyc over {c[22]} samples")
for t in (0, 1):
    wa, wb, tc_, tot = c[5 + 4 * t:9 + 4 * t]
    print(f"epilogue {'AB'[t]}: wait_acc {100*wa/max(tot,1):.1f}%  group barriers {100*wb/max(tot,1):.1f}%  column loops {100*tc_/max(tot,1):.1f}%  total {tot}; columns per layer {tc_/(per_cta*10):.0f} cyc")
