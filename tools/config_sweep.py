"""Sanity + throughput of the other BASELINE.json configurations through the public API (no graphs, small step counts).
usage: python tools/config_sweep.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "space-filling-curves-for-vision-transformers_b200"))
import torch
from src.models.vit import VisionTransformer
from src.tokenizers.multiscale.multi_hilbert import SFCEmbedding1D
from src.tokenizers.multiscale.multi_peano import SFCEmbedding1D as PeanoSFC
from src.curves.space_filling_curves import peano_curve, hilbert_curve, z_curve

dev = torch.device("cuda")

def build(img, D, depth, heads, mlp, classes, curve):
    torch.manual_seed(42)
    prev = torch.get_default_dtype(); torch.set_default_dtype(torch.bfloat16)
    try:
        tok = SFCEmbedding1D(img, 16, 1, 3, D, curve_fn=curve)
        m = VisionTransformer(patch_embed=tok, depth=depth, n_heads=heads, mlp_dim=mlp, num_classes=classes).to(dev)
    finally:
        torch.set_default_dtype(prev)
    return m

def timed(fn, n):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def run(name, img, D, depth, heads, mlp, B, curve, train):
    m = build(img, D, depth, heads, mlp, 1000, curve)
    x = torch.randn(B, 3, img, img, device=dev)
    if train:
        m.train()
        tgt = torch.nn.functional.one_hot(torch.randint(0, 1000, (B,), device=dev), 1000).float()
        def step():
            for p in m.parameters(): p.grad = None
            out = m(x)
            loss = torch.sum(-tgt * torch.nn.functional.log_softmax(out.float(), dim=-1), dim=-1).mean()
            loss.backward()
            return loss
        ms = timed(step, 3)
        l = float(step())
        ok = all(torch.isfinite(p.grad.float()).all() for p in m.parameters() if p.grad is not None)
    else:
        m.eval()
        with torch.no_grad():
            ms = timed(lambda: m(x), 3)
            out = m(x)
        l = float(out.float().abs().mean()); ok = bool(torch.isfinite(out.float()).all())
    print(json.dumps(dict(config=name, batch=B, tokens=(img // 16) ** 2, mode="train fwd+bwd" if train else "inference fwd",
                          ms=round(ms, 3), images_per_s=round(B / ms * 1e3, 1), finite=ok, probe=round(l, 4))), flush=True)
    del m; torch.cuda.empty_cache()

run("vit_s16_224 gen-Hilbert 14x14 inference", 224, 384, 12, 6, 1536, 256, hilbert_curve, False)
run("vit_b16_224 Morton training", 224, 768, 12, 12, 3072, 128, z_curve, True)
run("vit_l16_384 Peano (27->24) training", 384, 1024, 24, 16, 4096, 32, peano_curve, True)
run("vit_l16_384 Hilbert (32->24) training", 384, 1024, 24, 16, 4096, 32, hilbert_curve, True)
run("vit_b16_1024 Hilbert 64x64 (4096 tokens) inference", 1024, 768, 12, 12, 3072, 4, hilbert_curve, False)
