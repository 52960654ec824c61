"""Per-step timeline of CTA 0 of the attention backward kernel (debug). usage: python tools/attn_timeline.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "space-filling-curves-for-vision-transformers_b200"))
import torch
from sfcvit import ops, _lib
lib = _lib.load()
B, H, N = 256, 12, 196
D = H * 64
qkv = torch.randn(B * N, 3 * D, device="cuda").bfloat16()
out, lse = ops.attn_fwd(qkv, B, H, N, drop_p=0.1, drop_seed=5)
dout = torch.randn(B * N, D, device="cuda").bfloat16()
for _ in range(3):
    ops.attn_bwd(qkv, out, dout, lse, B, H, N, drop_p=0.1, drop_seed=5)
dbg = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
lib.sfc_debug_set_timeline.argtypes = [ctypes.c_void_p]
lib.sfc_debug_set_timeline(dbg.data_ptr())
ops.attn_bwd(qkv, out, dout, lse, B, H, N, drop_p=0.1, drop_seed=5)
torch.cuda.synchronize()
lib.sfc_debug_set_timeline(None)
t = dbg.cpu().view(64, 16)
t0 = int(t[4, 0])
print("step | WG: start s_full(wait) phaseA_end readout_end dp_wait_end phaseB_end | MMA: loop_start s_issued pds_seen grads_issued dp_issued   (cycles rel. to step 4 start)")
for s in range(4, 24):
    r = [int(x) - t0 for x in t[s]]
    print(s, "| WG", r[0], r[1], r[2], r[3], r[4], r[5], "fence", r[6], "arrive", r[7], "| MMA", r[8], r[9], r[12], r[11], r[10])
d = t[8:40, 0][1:] - t[8:40, 0][:-1]
print("mean cycles per step:", float(d.float().mean()))
