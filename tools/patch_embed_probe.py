"""Times the fused curve-gather patch embed alone on one configuration — the probe `ncu --set full` is pointed at.
usage: python tools/patch_embed_probe.py [img] [patch] [D] [batch] [iters] [bf16]   (default: ViT-Tiny/4 on 32 px, B 8192)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "space-filling-curves-for-vision-transformers_b200"))
from sfcvit import functional as SF  # noqa: E402
from sfcvit import ops  # noqa: E402


def main(img=32, p=4, D=192, B=8192, iters=20, bf16=0):
    dev = torch.device("cuda:0")
    n = img // p
    perm, _ = ops.curve_perm("hilbert", n, n, dev)
    g = torch.Generator(device=dev).manual_seed(0)
    w = (torch.randn(D, 3 * p * p, generator=g, device=dev) * 0.02).to(torch.bfloat16)
    wk = SF.kernel_weight(w, 3, p, 1, "p1p2c")
    bias = torch.zeros(D, dtype=torch.bfloat16, device=dev)
    esz = 2 if bf16 else 4
    nbuf = max(2, min(16, int(400e6 // (B * 3 * img * img * esz)) + 1))
    xs = [torch.randn(B, 3, img, img, generator=g, device=dev) for _ in range(nbuf)]
    if bf16:
        xs = [x.bfloat16() for x in xs]
    out = torch.empty(B, n * n, D, dtype=torch.bfloat16, device=dev)
    for i in range(3):
        ops.patch_embed_fwd(xs[i % nbuf], perm, wk, bias, p, 1, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ops.patch_embed_fwd(xs[i % nbuf], perm, wk, bias, p, 1, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nbytes = B * 3 * img * img * esz + B * n * n * D * 2
    print(f"patch_embed img={img} p={p} D={D} B={B} {'bf16' if bf16 else 'fp32'}: {ms:.4f} ms  {nbytes / ms / 1e6:.0f} GB/s algorithmic "
          f"({nbytes / 1e6:.0f} MB)")


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
