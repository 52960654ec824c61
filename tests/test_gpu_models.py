"""Model parity (src.models.vit on the libsfcvit kernels) vs the fp32 CPU oracle on the same seeded weights/inputs.
Stated tolerance (measured drift of the reference's own bf16-autocast path): logits rel-L2 <= 2e-2,
per-parameter gradient rel-L2 <= 1e-1, global gradient cosine >= 0.999. Dropout is zeroed on both sides."""
import pytest
import torch

import cases
from oracle import model as om

pytestmark = pytest.mark.gpu


def _build_pair(name):
    from src.models.vit import VisionTransformer, VisionTransformer1D
    vk, tcase, mkw, batch = cases.MODEL_CASES[name]
    kind, kw, shape = cases.TOKENIZER_CASES[tcase]
    torch.manual_seed(cases.INIT_SEED)
    o = om.zero_dropout(om.build_vit(vk, cases.build_oracle_tokenizer(kind, kw), **mkw))
    torch.manual_seed(cases.INIT_SEED)
    cls = VisionTransformer1D if vk == "vit1d" else VisionTransformer
    s = om.zero_dropout(cls(patch_embed=cases.build_src_tokenizer(kind, kw), **mkw))
    x = cases.make_input((batch,) + tuple(shape[1:]))
    tgt = cases.make_soft_targets(batch, mkw["num_classes"])
    return o, s, x, tgt


@pytest.mark.parametrize("name", sorted(cases.MODEL_CASES))
def test_model_forward_backward_parity(cuda_device, name):
    o, s, x, tgt = _build_pair(name)
    sd_o, sd_s = o.state_dict(), s.state_dict()
    assert list(sd_o) == list(sd_s)                                  # same keys, same order as the reference layout
    for k in sd_o:
        assert torch.equal(sd_o[k], sd_s[k]), k                     # identical init under seed 42
    s = s.to(cuda_device)
    o.train(); s.train()
    lo = o(x)
    loss_o = om.soft_target_cross_entropy(lo, tgt)
    loss_o.backward()
    ls = s(x.to(cuda_device))
    assert ls.dtype == torch.float32 and tuple(ls.shape) == tuple(lo.shape)
    loss_s = om.soft_target_cross_entropy(ls.float(), tgt.to(cuda_device))
    loss_s.backward()
    assert cases.rel_l2(ls, lo) < 2e-2
    assert abs(float(loss_s) - float(loss_o)) < 2e-2
    po, ps = dict(o.named_parameters()), dict(s.named_parameters())
    dot = no = ns = 0.0
    for n, p in po.items():
        if p.grad is None:
            assert ps[n].grad is None, n                              # e.g. mlp_mixer.token_mix* never get gradients
            continue
        g = ps[n].grad
        assert g is not None and g.dtype == ps[n].dtype, n
        assert cases.rel_l2(g, p.grad) < 1e-1, (n, cases.rel_l2(g, p.grad))
        gd, pd = g.detach().double().cpu().flatten(), p.grad.double().flatten()
        dot += float(gd @ pd); no += float(pd @ pd); ns += float(gd @ gd)
    assert dot / (no ** 0.5 * ns ** 0.5) > 0.999


def test_eval_autocast_bf16_params_and_checkpoint(cuda_device):
    """main.py's regime: default dtype bf16 parameters, bf16 autocast, eval(): logits are bf16 and match the oracle."""
    o, s, x, _ = _build_pair("vit1d_hier_morton")
    s.load_state_dict(o.state_dict())
    s = s.to(cuda_device).to(torch.bfloat16).eval()
    o.eval()
    with torch.no_grad(), torch.amp.autocast(device_type="cuda", dtype=torch.bfloat16):
        ls = s(x.to(cuda_device))
    with torch.no_grad():
        lo = o(x)
    assert ls.dtype == torch.bfloat16
    assert cases.rel_l2(ls, lo) < 4e-2                                # bf16 weights + bf16 logits


def test_training_mode_dropout_runs_and_is_seeded(cuda_device):
    from src.models.vit import VisionTransformer
    kind, kw, shape = cases.TOKENIZER_CASES["conv_hilbert_32_p4_d192"]
    torch.manual_seed(0)
    m = VisionTransformer(patch_embed=cases.build_src_tokenizer(kind, kw), depth=2, n_heads=3, mlp_dim=384,
                          num_classes=10).to(cuda_device).train()
    x = cases.make_input(shape).to(cuda_device)
    torch.manual_seed(5); a = m(x)
    torch.manual_seed(5); b = m(x)
    torch.manual_seed(6); c = m(x)
    assert torch.equal(a, b) and not torch.equal(a, c)
    a.float().square().mean().backward()
    for n, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n


@pytest.mark.parametrize("name", sorted(cases.ALTVIT_CASES))
def test_altvit_forward_backward_parity(cuda_device, name):
    """src.models.altvit (pre-norm GELU ViT on the libsfcvit kernels) vs the fp32 CPU oracle, same seed and inputs."""
    from oracle import altvit as oa
    from src.models import altvit as sa
    cls, kw, batch = cases.ALTVIT_CASES[name]
    torch.manual_seed(cases.INIT_SEED)
    o = getattr(oa, cls)(**kw)
    torch.manual_seed(cases.INIT_SEED)
    s = getattr(sa, cls)(**kw)
    sd_o, sd_s = o.state_dict(), s.state_dict()
    assert list(sd_o) == list(sd_s)
    for k in sd_o:
        assert torch.equal(sd_o[k], sd_s[k]), k
    if cls == "HilbertViT":
        assert torch.equal(o.to_patch_embedding.hilbert_indices, s.to_patch_embedding.hilbert_indices)   # curve kernel: bit-exact
    s = s.to(cuda_device)
    o.train(); s.train()
    x = cases.make_input((batch, 3, kw["image_size"], kw["image_size"]))
    tgt = cases.make_soft_targets(batch, kw["num_classes"])
    lo = o(x)
    om.soft_target_cross_entropy(lo, tgt).backward()
    ls = s(x.to(cuda_device))
    assert ls.dtype == torch.float32 and tuple(ls.shape) == tuple(lo.shape)
    om.soft_target_cross_entropy(ls.float(), tgt.to(cuda_device)).backward()
    assert cases.rel_l2(ls, lo) < 2e-2
    po, ps = dict(o.named_parameters()), dict(s.named_parameters())
    dot = no = ns = 0.0
    for n, p in po.items():
        g = ps[n].grad
        assert g is not None and g.dtype == ps[n].dtype, n
        assert cases.rel_l2(g, p.grad) < 1e-1, (n, cases.rel_l2(g, p.grad))
        gd, pd = g.detach().double().cpu().flatten(), p.grad.double().flatten()
        dot += float(gd @ pd); no += float(pd @ pd); ns += float(gd @ gd)
    assert dot / (no ** 0.5 * ns ** 0.5) > 0.999


def test_main_py_configuration_trains(cuda_device):
    """The model the reference driver actually builds (main.py:269-282): HierarchicalMortonEmbedding(32, 3, [16, 4, 1], 256)
    -> VisionTransformer1D(depth 8, 4 heads -> head_dim 192, mlp 512, 10 classes), default dtype bf16 (main.py:157),
    one mixup/cutmix training step through src.training with torch's own AdamW, then an evaluation pass."""
    from src.models.vit import VisionTransformer1D
    from src.tokenizers.multiscale.multi_morton import HierarchicalMortonEmbedding
    from src.training.train import evaluate, train_with_mixup_or_cutmix
    from src.training.losses import SoftTargetCrossEntropy
    torch.manual_seed(42)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        pe = HierarchicalMortonEmbedding(img_size=32, in_channels=3, patch_size_list=[16, 4, 1], embed_dim=256)
        model = VisionTransformer1D(patch_embed=pe, depth=8, n_heads=4, mlp_dim=512, num_classes=10).to(cuda_device)
    finally:
        torch.set_default_dtype(prev)
    model = torch.compile(model, mode="reduce-overhead")             # main.py:284 (the forward runs the kernels eagerly)
    opt = torch.optim.AdamW(model.parameters(), lr=3e-4, weight_decay=0.00005)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
    g = torch.Generator().manual_seed(0)
    data = [(torch.randn(16, 3, 32, 32, generator=g), torch.randint(0, 10, (16,), generator=g)) for _ in range(2)]

    class _Loader(list):
        dataset = list(range(32))

    loss, acc = train_with_mixup_or_cutmix(model, _Loader(data), SoftTargetCrossEntropy(), opt, sched, cuda_device)
    assert loss == loss and 0.5 < loss < 10.0 and 0.0 <= acc <= 1.0
    tl, ta = evaluate(model, _Loader(data), torch.nn.CrossEntropyLoss(), cuda_device)
    assert tl == tl and 0.0 <= ta <= 1.0
    sd = model.state_dict()
    assert any(k.startswith("_orig_mod.encoder.transformer.layers.7.self_attn.in_proj_weight") for k in sd)
