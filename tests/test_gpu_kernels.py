"""Kernel-level parity through the C ABI: attention (K4), fused patch embed (K2), LayerNorm / column sums (K5),
fused clip + AdamW (K6), each against torch fp32 math on the same bf16 inputs (tolerances inside kernel_selftest)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.parametrize("args", [(1, 1, 128), (2, 3, 196), (2, 2, 64), (1, 2, 576), (1, 1, 1), (3, 1, 129)])
def test_attention_fwd_bwd(cuda_device, args):
    import kernel_selftest as ks
    r = ks.check_attn(*args)
    assert r["ok"], r


def test_attention_dropout_mask_consistent(cuda_device):
    import kernel_selftest as ks
    r = ks.check_attn(2, 3, 196, 0.1)
    assert r["ok"], r


def test_attention_rejects_other_head_dims(cuda_device):
    import torch
    from sfcvit import ops
    with pytest.raises(RuntimeError):
        ops.attn_fwd(torch.zeros(16, 3 * 96, dtype=torch.bfloat16, device="cuda"), 1, 3, 16)


@pytest.mark.parametrize("args", [(3, 3, 224, 16, 1, 768, "hilbert", "fp32"), (3, 3, 224, 16, 1, 384, "hilbert", "bf16"),
                                  (5, 3, 32, 4, 1, 192, "z", "fp32"), (5, 3, 32, 1, 16, 256, "hilbert", "fp32"),
                                  (5, 3, 32, 2, 4, 256, "peano", "fp32"), (2, 3, 64, 8, 2, 128, "moore", "bf16"),
                                  (2, 3, 384, 16, 1, 1024, "peano", "fp32")])
def test_patch_embed(cuda_device, args):
    import kernel_selftest as ks
    r = ks.check_patch(*args)
    assert r["ok"], r


@pytest.mark.parametrize("args", [(1000, 768), (333, 192), (77, 1024), (5, 2048)])
def test_layernorm_and_colsum(cuda_device, args):
    import kernel_selftest as ks
    r = ks.check_ln(*args)
    assert r["ok"], r


def test_fused_clip_adamw_matches_torch(cuda_device):
    import kernel_selftest as ks
    r = ks.check_adamw()
    assert r["ok"], r
