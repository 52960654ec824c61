"""Parity at the SHAPES of BASELINE.json's other configurations (the bench line is ViT-B/16 224; these are the parity
cases): the widths, head counts, grid sizes and curves are the real ones, only the depth / batch are reduced so that
the fp32 CPU oracle finishes in seconds.

* configs[1] ViT-S/16 224 px, embed-and-prune Hilbert on 14 x 14, bf16 inference (all 12 layers);
* configs[3] ViT-L/16 384 px, Peano (27 -> 24 grid, 576 tokens) and Hilbert (32 -> 24), training step;
* configs[4] ViT-B/16 1024 px, Hilbert on 64 x 64 (4096 tokens), long-sequence inference; plus the attention kernel
  alone at N = 1024 / 4096 (forward and backward) against fp32 softmax attention.

Tolerances as in test_gpu_models.py (logits rel-L2 <= 2e-2 vs the fp32 oracle, per-parameter gradient rel-L2 <= 1e-1);
with bf16 PARAMETERS the oracle is given the same bf16-rounded weights, so only the arithmetic differs."""
import os
import sys

import pytest
import torch

import cases
from oracle import model as om

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

CURVE_FN = {"hilbert": "hilbert_curve", "peano": "peano_curve", "z": "z_curve", "moore": "moore_curve"}


def _pair(img, patch, D, depth, heads, mlp, classes, curve):
    from src.curves import space_filling_curves as sc
    from src.models.vit import VisionTransformer
    from src.tokenizers.multiscale.multi_hilbert import SFCEmbedding1D
    torch.manual_seed(cases.INIT_SEED)
    tok_o = om.SFCEmbedding1D(img, patch, 1, 3, D, curve)
    tok_o.n_patches = (img // patch) ** 2
    o = om.zero_dropout(om.VisionTransformer(tok_o, depth=depth, n_heads=heads, mlp_dim=mlp, num_classes=classes))
    torch.manual_seed(cases.INIT_SEED)
    tok_s = SFCEmbedding1D(img, patch, 1, 3, D, curve_fn=getattr(sc, CURVE_FN[curve]))
    s = om.zero_dropout(VisionTransformer(patch_embed=tok_s, depth=depth, n_heads=heads, mlp_dim=mlp, num_classes=classes))
    assert torch.equal(tok_o.sfc_indices, tok_s.sfc_indices.cpu())
    return o, s


@pytest.mark.parametrize("N", [1024, 4096])
def test_attention_long_sequence(cuda_device, N):
    import kernel_selftest as ks
    r = ks.check_attn(1, 2, N)
    assert r["ok"], r


def test_vit_s16_224_bf16_inference(cuda_device):
    o, s = _pair(224, 16, 384, 12, 6, 1536, 1000, "hilbert")
    s = s.to(cuda_device).to(torch.bfloat16).eval()
    o.load_state_dict({k: v.float().cpu() for k, v in s.state_dict().items()})
    o.eval()
    x = cases.make_input((4, 3, 224, 224))
    with torch.no_grad():
        ls = s(x.to(cuda_device).to(torch.bfloat16))
        lo = o(x.to(torch.bfloat16).float())
    assert ls.dtype == torch.bfloat16 and tuple(ls.shape) == (4, 1000)
    assert cases.rel_l2(ls, lo) < 2e-2, cases.rel_l2(ls, lo)


@pytest.mark.parametrize("curve", ["peano", "hilbert"])
def test_vit_l16_384_training_step(cuda_device, curve):
    o, s = _pair(384, 16, 1024, 2, 16, 4096, 1000, curve)
    s = s.to(cuda_device)
    o.train(); s.train()
    x = cases.make_input((2, 3, 384, 384))
    tgt = cases.make_soft_targets(2, 1000)
    lo = o(x)
    om.soft_target_cross_entropy(lo, tgt).backward()
    ls = s(x.to(cuda_device))
    om.soft_target_cross_entropy(ls.float(), tgt.to(cuda_device)).backward()
    assert cases.rel_l2(ls, lo) < 2e-2, cases.rel_l2(ls, lo)
    po, ps = dict(o.named_parameters()), dict(s.named_parameters())
    for n, p in po.items():
        assert cases.rel_l2(ps[n].grad, p.grad) < 1e-1, (n, cases.rel_l2(ps[n].grad, p.grad))


def test_vit_b16_224_training_step(cuda_device):
    """The HEADLINE shape (BASELINE.json metric / configs[2]): ViT-B/16 224 px, D 768, 12 heads, MLP 3072, generalised
    Hilbert on the 14 x 14 grid, forward + soft-target CE + backward vs the fp32 oracle (reference vit.py:325-385); depth 2
    and batch 4 so that the CPU oracle finishes in seconds. Standing tolerances; plus the global gradient cosine."""
    o, s = _pair(224, 16, 768, 2, 12, 3072, 1000, "hilbert")
    s = s.to(cuda_device)
    o.train(); s.train()
    x = cases.make_input((4, 3, 224, 224))
    tgt = cases.make_soft_targets(4, 1000)
    lo = o(x)
    loss_o = om.soft_target_cross_entropy(lo, tgt)
    loss_o.backward()
    ls = s(x.to(cuda_device))
    loss_s = om.soft_target_cross_entropy(ls.float(), tgt.to(cuda_device))
    loss_s.backward()
    assert cases.rel_l2(ls, lo) < 2e-2, cases.rel_l2(ls, lo)
    assert abs(float(loss_s) - float(loss_o)) < 2e-2 * abs(float(loss_o))
    po, ps = dict(o.named_parameters()), dict(s.named_parameters())
    assert set(po) == set(ps)
    for n, p in po.items():
        assert cases.rel_l2(ps[n].grad, p.grad) < 1e-1, (n, cases.rel_l2(ps[n].grad, p.grad))
    go = torch.cat([p.grad.reshape(-1) for _, p in sorted(po.items())])
    gs = torch.cat([ps[n].grad.float().cpu().reshape(-1) for n, _ in sorted(po.items())])
    cos = float(torch.dot(go, gs) / (go.norm() * gs.norm()))
    assert cos >= 0.999, cos


def test_vit_b16_1024_long_sequence_inference(cuda_device):
    """4096 tokens per image: Hilbert order on the 64 x 64 patch grid, 2 of the 12 layers, head W_seq [1536, 4096, 64]."""
    o, s = _pair(1024, 16, 768, 2, 12, 3072, 1000, "hilbert")
    s = s.to(cuda_device).eval()
    o.eval()
    x = cases.make_input((2, 3, 1024, 1024))
    with torch.no_grad():
        ls = s(x.to(cuda_device))
        lo = o(x)
        assert tuple(ls.shape) == (2, 1000)
        assert cases.rel_l2(ls, lo) < 2e-2, cases.rel_l2(ls, lo)
        # batch independence at this length (bit-exact): image 1 alone gives the same logits
        assert torch.equal(s(x[1:].to(cuda_device)), ls[1:])
