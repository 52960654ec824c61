"""Times the pixel-level (p = 1) curve tokenizers — SURVEY.md §8f row 2 — next to the patch tokenizer of the same token
shape. usage: python tools/pixel_tok_probe.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "space-filling-curves-for-vision-transformers_b200"))


def timeit(fn, iters=10):
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


def main(B=256):
    from src.tokenizers._1D.hilbert_embedding1D import HilbertEmbedding1D
    from src.tokenizers._1D.morton_embedding1D import MortonEmbedding1D
    from src.tokenizers.multiscale.multi_hilbert import SFCEmbedding1D
    dev = torch.device("cuda")
    x = torch.randn(B, 3, 224, 224, device=dev)
    x32 = torch.randn(4 * B, 3, 32, 32, device=dev)
    cases = [("SFCEmbedding1D(224, p16) patch tokens [TMEM kernel]", SFCEmbedding1D(224, 16, 1, 3, 768), x),
             ("HilbertEmbedding1D(224, 256 px/token) pixel tokens", HilbertEmbedding1D(224, 256, 3, 768), x),
             ("MortonEmbedding1D(224, 256 px/token) pixel tokens", MortonEmbedding1D(224, 256, 3, 768), x),
             ("HilbertEmbedding1D(32, 16 px/token) CIFAR pixel tokens", HilbertEmbedding1D(32, 16, 3, 192), x32)]
    with torch.no_grad():
        for name, tok, inp in cases:
            tok = tok.to(dev).to(torch.bfloat16)
            ms = timeit(lambda: tok(inp))
            nbytes = inp.numel() * 4 + inp.shape[0] * tok.n_patches * tok.embed_dim * 2
            print(f"{name}: {ms:.4f} ms  {inp.shape[0] / ms:.0f} img/ms  {nbytes / ms / 1e6:.0f} GB/s algorithmic")


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])


def two_pass(B=256):
    """Alternative for p = 1: materialise the curve-ordered bf16 im2col once (sfc_patch_gather) + one K3 GEMM."""
    from sfcvit import functional as SF, ops
    from src.tokenizers._1D.hilbert_embedding1D import HilbertEmbedding1D
    dev = torch.device("cuda")
    tok = HilbertEmbedding1D(224, 256, 3, 768).to(dev).to(torch.bfloat16)
    x = torch.randn(B, 3, 224, 224, device=dev)
    perm = tok._perm32(dev)
    wk = SF.kernel_weight(tok.proj.weight, 3, 1, 256, "p1p2c")
    bias = tok.proj.bias.detach()
    with torch.no_grad():
        ms_g = timeit(lambda: ops.patch_gather(x, perm, 1, 256))
        A = ops.patch_gather(x, perm, 1, 256)
        ms_m = timeit(lambda: ops.gemm(A, wk, bias=bias))
        ref = tok(x).reshape(-1, 768).float()
        got = ops.gemm(A, wk, bias=bias).float()
    print(f"two-pass: gather {ms_g:.4f} ms + GEMM {ms_m:.4f} ms; max |diff| vs fused {float((got - ref).abs().max()):.3e}")


def byte_input(B=256):
    """The ViT-B/16 patch tokenizer fed fp32 NCHW, bf16 NCHW and the decoded bytes (uint8 NHWC)."""
    from src.tokenizers.multiscale.multi_hilbert import SFCEmbedding1D
    dev = torch.device("cuda")
    tok = SFCEmbedding1D(224, 16, 1, 3, 768).to(dev).to(torch.bfloat16)
    tok.set_uint8_normalization((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    xs = [torch.randn(B, 3, 224, 224, device=dev) for _ in range(4)]
    ins = {"fp32 NCHW": xs, "bf16 NCHW": [x.to(torch.bfloat16) for x in xs],
           "uint8 NHWC": [torch.randint(0, 256, (B, 224, 224, 3), device=dev, dtype=torch.uint8) for _ in range(4)]}
    with torch.no_grad():
        for name, arr in ins.items():
            it = iter(range(10 ** 9))
            ms = timeit(lambda: tok(arr[next(it) % 4]), iters=20)
            nbytes = arr[0].numel() * arr[0].element_size() + B * 196 * 768 * 2
            print(f"patch embed ViT-B/16, {name} input: {ms:.4f} ms  {nbytes / ms / 1e6:.0f} GB/s algorithmic ({nbytes / B} B / image)")


if __name__ == "__main__" and os.environ.get("TWO_PASS"):
    two_pass()
if __name__ == "__main__" and os.environ.get("BYTE_INPUT"):
    byte_input()
