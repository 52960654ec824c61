"""Times (CUDA events) the attention forward kernel alone at a given shape — the probe `ncu --set full` is pointed at.
usage: python tools/attn_probe.py [B] [H] [N] [iters] [dropout percent]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "space-filling-curves-for-vision-transformers_b200"))
from sfcvit import ops  # noqa: E402


def main(B=4, H=12, N=4096, iters=5, drop=0):
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(0)
    qkvs = [torch.randn(B * N, 3 * D, generator=g, device="cuda", dtype=torch.float32).to(torch.bfloat16) for _ in range(3)]
    for q in qkvs:
        ops.attn_fwd(q, B, H, N, drop_p=drop / 100.0, drop_seed=7)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ops.attn_fwd(qkvs[i % 3], B, H, N, drop_p=drop / 100.0, drop_seed=7)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 4.0 * B * N * N * D
    print(f"attn_fwd B={B} H={H} N={N} drop={drop}%: {ms:.4f} ms  {fl / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
