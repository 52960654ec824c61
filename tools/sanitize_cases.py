"""One small invocation of every kernel family through the C ABI (the workload tools/sanitize.sh runs under compute-sanitizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "space-filling-curves-for-vision-transformers_b200"), os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
import kernel_selftest as ks
import gemm_selftest
from sfcvit import ops

res = []
for curve in ("hilbert", "z", "peano", "moore"):
    perm, inv = ops.curve_perm(curve, 14, 14, "cuda")
    res.append(("curve_" + curve, int(perm.sum()) == 196 * 195 // 2))
res.append(("attn_196", ks.check_attn(1, 2, 196)["ok"]))
res.append(("attn_196_drop", ks.check_attn(1, 2, 196, 0.1)["ok"]))
res.append(("attn_576", ks.check_attn(1, 1, 576)["ok"]))
res.append(("patch_tmem", ks.check_patch(3, 3, 224, 16, 1, 768, "hilbert", "fp32")["ok"]))
res.append(("patch_smem", ks.check_patch(5, 3, 32, 2, 4, 256, "peano", "fp32")["ok"]))
res.append(("ln", ks.check_ln(333, 192)["ok"]))
res.append(("adamw", ks.check_adamw(20003)["ok"]))
for epi in ("bias_relu", "bias_res", "relu_mask_res", "fp32"):
    res.append(("gemm_" + epi, gemm_selftest.run(392, 768, 256, False, False, epi)["ok"]))
res.append(("gemm_wgrad_split", gemm_selftest.run(768, 256, 2048, True, True, "fp32", 4)["ok"]))
dy = torch.randn(1568, 768, device="cuda").bfloat16(); x = torch.randn(1568, 256, device="cuda").bfloat16()
dw, db = ops.wgrad(dy, x, torch.bfloat16, want_db=True)
res.append(("wgrad_colsum", bool(((db.float() - dy.float().sum(0)).norm() / dy.float().sum(0).norm()) < 5e-3)))
lg = torch.randn(64, 1000, device="cuda").bfloat16(); tg = torch.softmax(torch.randn(64, 1000, device="cuda"), -1)
loss, lse, ts = ops.softce_fwd(lg, tg)
ops.softce_bwd(lg, tg, lse, ts, torch.ones(1, device="cuda"))
res.append(("softce", bool(torch.isfinite(loss).all())))
torch.cuda.synchronize()
print(res)
assert all(ok for _, ok in res), res
