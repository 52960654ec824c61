"""CUDA-graph replay of a training step (src.training.graphs.GraphedStep): same numbers as the eager step with dropout
off, fresh dropout masks on every replay (device-side epoch), forward/backward mask agreement inside one replay."""
import pytest
import torch

import cases
from oracle import model as om

pytestmark = pytest.mark.gpu


def _model(cuda_device, dropout):
    from src.models.vit import VisionTransformer
    vk, tcase, mkw, batch = cases.MODEL_CASES["vit_sfc_hilbert_14x14"]
    kind, kw, shape = cases.TOKENIZER_CASES[tcase]
    torch.manual_seed(cases.INIT_SEED)
    m = VisionTransformer(patch_embed=cases.build_src_tokenizer(kind, kw), **mkw)
    if not dropout:
        m = om.zero_dropout(m)
    x = cases.make_input((batch,) + tuple(shape[1:])).to(cuda_device)
    tgt = cases.make_soft_targets(batch, mkw["num_classes"]).to(cuda_device)
    return m.to(cuda_device).train(), x, tgt


def _crit(logits, tgt):
    return om.soft_target_cross_entropy(logits.float(), tgt)


def test_graph_replay_matches_eager(cuda_device):
    from src.training.graphs import GraphedStep
    m, x, tgt = _model(cuda_device, dropout=False)
    loss_e = _crit(m(x), tgt)
    loss_e.backward()
    ref = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    loss_ref = float(loss_e)
    del loss_e                                           # a live eager autograd graph pins AccumulateGrad nodes to the default stream
    step = GraphedStep(m, _crit, x, tgt)
    for _ in range(2):                                   # replays are idempotent (grads assigned, not accumulated)
        loss_g = step(x, tgt)
    torch.cuda.synchronize()
    assert abs(float(loss_g) - loss_ref) < 1e-6
    got = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(ref)
    for n in ref:
        assert torch.equal(got[n], ref[n]), n            # identical kernels, identical order: bit-equal
    # new data through the static buffers
    x2 = torch.roll(x, 1, 0)
    loss_g2 = float(step(x2, torch.roll(tgt, 1, 0)))
    assert abs(loss_g2 - loss_ref) < 1e-3           # a permutation of the batch: same mean loss


def test_graph_dropout_masks_change_per_replay(cuda_device):
    from src.training.graphs import GraphedStep
    m, x, tgt = _model(cuda_device, dropout=True)
    step = GraphedStep(m, _crit, x, tgt)
    losses = []
    for _ in range(4):
        losses.append(float(step(x, tgt)))
    assert len({round(l, 7) for l in losses}) == 4, losses     # frozen host seeds + device epoch => different masks
    g = [p.grad for p in m.parameters() if p.grad is not None]
    assert all(torch.isfinite(t.float()).all() for t in g)
