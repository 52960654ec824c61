"""Times (CUDA events) the attention backward kernel alone — the probe `ncu --set full` is pointed at.
usage: python tools/attn_bwd_probe.py [B] [H] [N] [iters] [dropout percent]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "space-filling-curves-for-vision-transformers_b200"))
from sfcvit import ops  # noqa: E402


def main(B=256, H=12, N=196, iters=10, drop=10):
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B * N, 3 * D, generator=g, device="cuda").to(torch.bfloat16)
    dout = torch.randn(B * N, D, generator=g, device="cuda").to(torch.bfloat16)
    out, lse = ops.attn_fwd(qkv, B, H, N, drop_p=drop / 100.0, drop_seed=7)
    for _ in range(3):
        ops.attn_bwd(qkv, out, dout, lse, B, H, N, drop_p=drop / 100.0, drop_seed=7)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.attn_bwd(qkv, out, dout, lse, B, H, N, drop_p=drop / 100.0, drop_seed=7)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"attn_bwd B={B} H={H} N={N} drop={drop}%: {ms:.4f} ms  {10.0 * B * N * N * D / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
