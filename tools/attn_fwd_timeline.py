"""Per-item timeline of CTA 0 of the attention FORWARD kernel (debug build: SFC_ATTN_TIMELINE=1 python .../build.py, then
SFC_ATTN_TIMELINE=1 python tools/attn_fwd_timeline.py [drop%])."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "space-filling-curves-for-vision-transformers_b200"))
import torch
from sfcvit import ops, _lib
lib = _lib.load(build_if_missing=False)
drop = (float(sys.argv[1]) if len(sys.argv) > 1 else 10.0) / 100.0
B, H, N = 256, 12, 196
D = H * 64
qkv = torch.randn(B * N, 3 * D, device="cuda").bfloat16()
for _ in range(3):
    ops.attn_fwd(qkv, B, H, N, drop_p=drop, drop_seed=5)
dbg = torch.zeros(32 * 32, dtype=torch.int64, device="cuda")
lib.sfc_debug_set_timeline.argtypes = [ctypes.c_void_p]
lib.sfc_debug_set_timeline(dbg.data_ptr())
ops.attn_fwd(qkv, B, H, N, drop_p=drop, drop_seed=5)
torch.cuda.synchronize()
lib.sfc_debug_set_timeline(None)
t = dbg.cpu().view(32, 32)
t0 = int(t[2, 0])
print("item | tile A: wait_s s_full max_done exp_done arrived o_full out_done | tile B: same | MMA: pv0 s0 pv1 s1 issued   (cycles rel. to item 2)")
for g in range(2, 16):
    r = [int(x) - t0 for x in t[g]]
    print(g, "| A", *r[0:7], "| B", *r[8:15], "| MMA", *r[16:20])
d = (t[3:18, 0] - t[2:17, 0]).float()
print("mean cycles per item:", float(d.mean()))
