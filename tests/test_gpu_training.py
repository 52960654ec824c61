"""The four epoch functions of src/training/train.py (mirror of /root/reference/src/training/train.py:57-178) on the GPU:
* against the ORACLE loop — the reference's loop restated on the fp32 CPU oracle model (same batches, same optimizer) —
  for `train`, `train_with_scheduler` and `train_with_mixup_or_cutmix`: returned (loss, accuracy) and the parameter
  UPDATE of the epoch (SGD: update = -lr * sum of gradients) within the standing bf16 tolerances;
* the lazily captured CUDA-graph steps give the same numbers as eager launches (SFC_TRAIN_GRAPH=0), including a ragged
  last batch; `evaluate` covers every sample."""
import copy
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases
from oracle import model as om

pytestmark = pytest.mark.gpu


def _pair(cuda_device, case="vit_conv_hilbert_tiny"):
    from src.models.vit import VisionTransformer, VisionTransformer1D
    vk, tcase, mkw, _ = cases.MODEL_CASES[case]
    kind, kw, shape = cases.TOKENIZER_CASES[tcase]
    torch.manual_seed(cases.INIT_SEED)
    o = om.zero_dropout(om.build_vit(vk, cases.build_oracle_tokenizer(kind, kw), **mkw))
    torch.manual_seed(cases.INIT_SEED)
    cls = VisionTransformer1D if vk == "vit1d" else VisionTransformer
    s = om.zero_dropout(cls(patch_embed=cases.build_src_tokenizer(kind, kw), **mkw)).to(cuda_device)
    return o, s, shape, mkw["num_classes"]


class _Loader(list):
    def __init__(self, batches):
        super().__init__(batches)
        self.dataset = list(range(sum(b[0].shape[0] for b in batches)))


def _data(shape, classes, sizes=(8, 8, 8), seed=3):
    g = torch.Generator().manual_seed(seed)
    return _Loader([(torch.randn(n, *shape[1:], generator=g), torch.randint(0, classes, (n,), generator=g)) for n in sizes])


def _delta(model, before):
    return {n: (p.detach().float().cpu() - before[n]) for n, p in model.named_parameters()}


def _snapshot(model):
    return {n: p.detach().float().cpu().clone() for n, p in model.named_parameters()}


def _compare_updates(ds, do, key_bias_ok=True):
    num = den = dot = ns = no = 0.0
    for n in do:
        a, b = ds[n], do[n]
        if n.endswith("in_proj_bias"):           # the key third has a mathematically zero gradient
            d = a.numel() // 3
            a, b = torch.cat([a[:d], a[2 * d:]]), torch.cat([b[:d], b[2 * d:]])
        if float(b.norm()) == 0.0:
            assert float(a.norm()) < 1e-6, n
            continue
        assert cases.rel_l2(a, b) < 1.5e-1, (n, cases.rel_l2(a, b))
        dot += float(a.flatten() @ b.flatten()); ns += float(a.flatten() @ a.flatten()); no += float(b.flatten() @ b.flatten())
    assert dot / (ns ** 0.5 * no ** 0.5) > 0.995


def test_train_matches_oracle_loop(cuda_device):
    """`train` (reference :57-77): plain cross-entropy epoch."""
    from src.training.train import train
    o, s, shape, classes = _pair(cuda_device)
    data = _data(shape, classes)
    bo, bs = _snapshot(o), _snapshot(s)
    opt_s = torch.optim.SGD(s.parameters(), lr=0.05)
    loss_s, acc_s = train(s, data, torch.nn.CrossEntropyLoss(), opt_s, cuda_device)
    # the reference loop on the oracle
    opt_o = torch.optim.SGD(o.parameters(), lr=0.05)
    o.train()
    tot, cor = 0.0, 0
    for x, y in data:
        opt_o.zero_grad()
        out = o(x)
        loss = F.cross_entropy(out, y)
        loss.backward()
        opt_o.step()
        tot += float(loss) * x.size(0)
        cor += int((out.argmax(1) == y).sum())
    n = len(data.dataset)
    assert abs(loss_s - tot / n) < 2e-2 * max(1.0, tot / n), (loss_s, tot / n)
    assert abs(acc_s - cor / n) <= 2.0 / n
    _compare_updates(_delta(s, bs), _delta(o, bo))


def test_train_with_scheduler_matches_oracle_loop(cuda_device):
    """`train_with_scheduler` (reference :102-130) with the reference's WarmupCosineScheduler (step() returns the rate)."""
    from src.training.scheduler import WarmupCosineScheduler
    from src.training.train import train_with_scheduler
    o, s, shape, classes = _pair(cuda_device, "vit_sfc_hilbert_14x14")
    data = _data(shape, classes, sizes=(6, 6, 6, 6))
    bo, bs = _snapshot(o), _snapshot(s)
    opt_s = torch.optim.SGD(s.parameters(), lr=0.05)
    sch_s = WarmupCosineScheduler(opt_s, warmup_steps=2, total_steps=8)
    loss_s, acc_s = train_with_scheduler(s, data, torch.nn.CrossEntropyLoss(), opt_s, sch_s, cuda_device)
    opt_o = torch.optim.SGD(o.parameters(), lr=0.05)
    sch_o = WarmupCosineScheduler(opt_o, warmup_steps=2, total_steps=8)
    o.train()
    tot, cor = 0.0, 0
    for x, y in data:
        opt_o.zero_grad()
        out = o(x)
        loss = F.cross_entropy(out, y)
        loss.backward()
        opt_o.step()
        sch_o.step()
        tot += float(loss) * x.size(0)
        cor += int((out.argmax(1) == y).sum())
    n = len(data.dataset)
    assert sch_s.current_step == sch_o.current_step == 4
    assert abs(loss_s - tot / n) < 2e-2 * max(1.0, tot / n), (loss_s, tot / n)
    assert abs(acc_s - cor / n) <= 2.0 / n
    _compare_updates(_delta(s, bs), _delta(o, bo))


def _mix_epoch_oracle(o, data, opt, classes, mixup_alpha=0.2, cutmix_alpha=1.0, mix_prob=0.5):
    """/root/reference/src/training/train.py:133-178 restated on CPU tensors (same host RNG consumption as src.training)."""
    from src.training.train import cutmix_data, mixup_data
    o.train()
    tl = tc = ts = 0.0
    for x, y in data:
        x, y = x.clone(), y.clone()
        if np.random.rand() < mix_prob:
            x, ya, yb, lam = mixup_data(x, y, alpha=mixup_alpha)
        else:
            x, ya, yb, lam = cutmix_data(x, y, alpha=cutmix_alpha)
        opt.zero_grad()
        out = o(x)
        soft = lam * F.one_hot(ya, classes).float() + (1 - lam) * F.one_hot(yb, classes).float()
        loss = om.soft_target_cross_entropy(out, soft)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(o.parameters(), 1.0, foreach=False)
        opt.step()
        pred = out.argmax(1)
        tc += float((lam * (pred == ya).float() + (1 - lam) * (pred == yb).float()).sum())
        tl += float(loss) * x.size(0)
        ts += x.size(0)
    return tl / ts, tc / ts


def test_mixup_cutmix_epoch_matches_oracle_loop(cuda_device, monkeypatch):
    """`train_with_mixup_or_cutmix` (reference :133-178) incl. SoftTargetCrossEntropy (main.py:45-51) and the clip at 1.0.
    The batch permutation of mixup / cutmix is drawn on the input's device, so both sides are fed the same CPU draw."""
    from src.training import train as T
    from src.training.losses import SoftTargetCrossEntropy
    o, s, shape, classes = _pair(cuda_device)
    data = _data(shape, classes)
    perms = [torch.randperm(8, generator=torch.Generator().manual_seed(100 + i)) for i in range(len(data))]
    it = {"k": 0}

    def fake_randperm(n, device=None, **kw):
        p = perms[it["k"] % len(perms)]
        it["k"] += 1
        return p.to(device) if device is not None else p
    monkeypatch.setattr(T.torch, "randperm", fake_randperm)
    bo, bs = _snapshot(o), _snapshot(s)
    opt_s = torch.optim.SGD(s.parameters(), lr=0.05)
    sched = torch.optim.lr_scheduler.LambdaLR(opt_s, lambda k: 1.0)
    np.random.seed(11)
    loss_s, acc_s = T.train_with_mixup_or_cutmix(s, data, SoftTargetCrossEntropy(), opt_s, sched, cuda_device)
    it["k"] = 0
    np.random.seed(11)
    loss_o, acc_o = _mix_epoch_oracle(o, data, torch.optim.SGD(o.parameters(), lr=0.05), classes)
    assert abs(loss_s - loss_o) < 2e-2 * max(1.0, loss_o), (loss_s, loss_o)
    assert abs(acc_s - acc_o) < 0.1
    _compare_updates(_delta(s, bs), _delta(o, bo))


def test_graphed_loops_equal_eager_loops(cuda_device, monkeypatch):
    """The lazily captured step (what replaces torch.compile's CUDA graphs, main.py:284) changes nothing: same epoch
    result and bit-identical parameters as eager launches, with a ragged last batch going through the eager path."""
    from src.training import train as T
    from src.training.losses import SoftTargetCrossEntropy
    _, s1, shape, classes = _pair(cuda_device)
    s2 = copy.deepcopy(s1)
    data = _data(shape, classes, sizes=(8, 8, 8, 5))
    res = []
    for m, flag in ((s1, "1"), (s2, "0")):
        monkeypatch.setenv("SFC_TRAIN_GRAPH", flag)
        np.random.seed(5); torch.manual_seed(5); torch.cuda.manual_seed_all(5)
        opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda k: 1.0)
        res.append(T.train_with_mixup_or_cutmix(m, data, SoftTargetCrossEntropy(), opt, sched, cuda_device))
    assert T._has_graphs(s1) and not T._has_graphs(s2)
    assert abs(res[0][0] - res[1][0]) < 1e-6 and abs(res[0][1] - res[1][1]) < 1e-6, res
    for (n, a), (_, b) in zip(s1.named_parameters(), s2.named_parameters()):
        assert torch.equal(a, b), n


def test_evaluate_covers_every_sample(cuda_device):
    from src.training.train import evaluate
    o, s, shape, classes = _pair(cuda_device)
    data = _data(shape, classes, sizes=(8, 8, 3))
    loss_s, acc_s = evaluate(s, data, torch.nn.CrossEntropyLoss(), cuda_device)
    o.eval()
    tot, cor = 0.0, 0
    with torch.no_grad():
        for x, y in data:
            out = o(x)
            tot += float(F.cross_entropy(out, y)) * x.size(0)
            cor += int((out.argmax(1) == y).sum())
    n = len(data.dataset)
    assert abs(loss_s - tot / n) < 2e-2 * max(1.0, tot / n)
    assert abs(acc_s - cor / n) <= 2.0 / n
