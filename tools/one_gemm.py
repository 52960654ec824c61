"""Runs one GEMM shape a few times (profiling target). usage: one_gemm.py M N K [bias] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "space-filling-curves-for-vision-transformers_b200"))
import torch
from sfcvit import ops
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
use_bias = len(sys.argv) > 4 and sys.argv[4] == "bias"
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
b = torch.randn(N, device="cuda").bfloat16() if use_bias else None
for _ in range(iters):
    y = ops.gemm(x, w, bias=b)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
