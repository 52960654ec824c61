"""Times (CUDA events) the generic-head-dimension attention (csrc/attention_generic.cu) forward and backward at main.py's
shape (B 512, 4 heads, 64 tokens, head_dim 192 by default), and cross-checks the warp-MMA kernels against the
CUDA-core kernels of the same file when run twice (SFC_ATTN_NO_MMA=1 selects the latter): the checksums printed must
agree to bf16 rounding.
usage: python tools/attn_generic_probe.py [B] [H] [N] [dh] [iters] [dropout percent]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "space-filling-curves-for-vision-transformers_b200"))
from sfcvit import ops  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main(B=512, H=4, N=64, dh=192, iters=20, drop=10):
    D = H * dh
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B * N, 3 * D, generator=g, device="cuda").to(torch.bfloat16)
    dout = torch.randn(B * N, D, generator=g, device="cuda").to(torch.bfloat16)
    p = drop / 100.0
    out, lse = ops.attn_fwd(qkv, B, H, N, drop_p=p, drop_seed=7)
    dqkv = ops.attn_bwd(qkv, out, dout, lse, B, H, N, drop_p=p, drop_seed=7)
    f = timed(lambda: ops.attn_fwd(qkv, B, H, N, drop_p=p, drop_seed=7), iters)
    bw = timed(lambda: ops.attn_bwd(qkv, out, dout, lse, B, H, N, drop_p=p, drop_seed=7), iters)
    path = "cuda-core" if os.environ.get("SFC_ATTN_NO_MMA") == "1" else "warp-mma"
    fl = B * H * N * N * dh
    print(f"[{path}] B={B} H={H} N={N} dh={dh} drop={drop}%: fwd {f:.4f} ms ({4.0 * fl / f / 1e9:.1f} TFLOP/s)  "
          f"bwd {bw:.4f} ms ({10.0 * fl / bw / 1e9:.1f} TFLOP/s)")
    print(f"[{path}] checksums: out {float(out.float().abs().sum()):.6e}  lse {float(lse.sum()):.6e}  "
          f"dq {float(dqkv[:, :D].float().abs().sum()):.6e}  dk {float(dqkv[:, D:2 * D].float().abs().sum()):.6e}  "
          f"dv {float(dqkv[:, 2 * D:].float().abs().sum()):.6e}")


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
