"""uint8 NHWC input (the decode / ToDtype / Normalize stage before the path, main.py:174-178, folded
into K2). The checker is the fp32 CPU oracle tokenizer fed the float image the reference pipeline would have produced
from the same bytes: x = (bytes / 255 - mean) / std in NCHW. Tolerances: tokens rel-L2 <= 1e-2 (bf16 weights and
outputs), weight / bias gradients rel-L2 <= 2e-2."""
import pytest
import torch

import cases
from oracle import model as om

pytestmark = pytest.mark.gpu

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def _bytes(shape, seed=3):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)


def _float_image(u8):
    x = u8.float().permute(0, 3, 1, 2) / 255.0
    return (x - torch.tensor(MEAN).view(1, 3, 1, 1)) / torch.tensor(STD).view(1, 3, 1, 1)


@pytest.mark.parametrize("img,p,g,D,curve", [(224, 16, 1, 768, "hilbert"),      # TMEM-resident kernel, ViT-B/16 shape
                                              (64, 8, 1, 96, "peano"),           # shared-memory kernel (D % 128 != 0)
                                              (64, 8, 4, 128, "z")])             # grouped pre-patches
def test_uint8_tokens_and_gradients(cuda_device, img, p, g, D, curve):
    import importlib
    mod = importlib.import_module("src.tokenizers.multiscale.multi_" + {"hilbert": "hilbert", "peano": "peano", "z": "morton"}[curve])
    torch.manual_seed(11)
    o = om.SFCEmbedding1D(img, p, g, 3, D, curve)
    s = mod.SFCEmbedding1D(img, p, g, 3, D)
    s.load_state_dict(o.state_dict())
    s = s.to(cuda_device).set_uint8_normalization(MEAN, STD)
    u8 = _bytes((3, img, img, 3))
    x = _float_image(u8)
    to = o(x)
    ts = s(u8.to(cuda_device))
    assert ts.dtype == torch.bfloat16 and tuple(ts.shape) == tuple(to.shape)
    assert cases.rel_l2(ts, to) < 1e-2, cases.rel_l2(ts, to)
    # the float path of the same module on the normalised image agrees too
    tf = s(x.to(cuda_device))
    assert cases.rel_l2(ts, tf.float().cpu()) < 1e-2
    w = torch.randn(to.shape, generator=torch.Generator().manual_seed(5))
    (to * w).sum().backward()
    (ts.float() * w.to(cuda_device)).sum().backward()
    assert cases.rel_l2(s.proj.weight.grad, o.proj.weight.grad) < 2e-2, cases.rel_l2(s.proj.weight.grad, o.proj.weight.grad)
    assert cases.rel_l2(s.proj.bias.grad, o.proj.bias.grad) < 2e-2


def test_uint8_default_is_unit_scale_and_conv_weight_layout(cuda_device):
    """Without set_uint8_normalization the bytes mean x / 255; the Conv2d-weight tokenizer ([D, C, p, p]) takes the same path."""
    from src.tokenizers._2D.hilbert_embedding import HilbertEmbedding
    torch.manual_seed(2)
    s = HilbertEmbedding(64, 8, 3, 128).to(cuda_device)
    u8 = _bytes((2, 64, 64, 3), seed=9)
    xf = (u8.float() / 255.0).permute(0, 3, 1, 2).contiguous()
    a = s(u8.to(cuda_device))
    b = s(xf.to(cuda_device))
    assert cases.rel_l2(a, b.float().cpu()) < 1e-2
    a.float().square().mean().backward()
    ga = s.proj.weight.grad.clone()
    s.proj.weight.grad = None
    s(xf.to(cuda_device)).float().square().mean().backward()
    assert cases.rel_l2(ga, s.proj.weight.grad.cpu()) < 3e-2


def test_uint8_unsupported_layout_raises(cuda_device):
    from src.tokenizers.multiscale.multi_hilbert import SFCEmbedding1D
    s = SFCEmbedding1D(32, 4, 1, 3, 64).to(cuda_device)            # p * C = 12: a patch row is not a whole number of 8-byte runs
    with pytest.raises(RuntimeError):
        s(_bytes((2, 32, 32, 3)).to(cuda_device))
