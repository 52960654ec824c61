"""Pins the oracle (oracle/) to the fixtures generated from the live reference (tests/golden/make_golden.py).
CPU only. Integer work is compared exactly; float outputs with a 1e-5 tolerance (the fixtures were produced on
another CPU, and MKL/oneDNN reduction orders differ between hosts)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import cases
from oracle import curves as oc
from oracle import model as om

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HASHES = json.load(open(os.path.join(GOLD, "perm_hashes.json")))


def h16(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<i8").tobytes()).hexdigest()[:16]


def test_survey_known_answers():
    # SURVEY.md §8c known-answer heads, generated from the live reference
    assert oc.flat_perm("hilbert", 8, 8)[:10].tolist() == [0, 1, 9, 8, 16, 24, 25, 17, 18, 26]
    assert oc.flat_perm("hilbert", 14, 14)[:10].tolist() == [0, 14, 15, 1, 2, 3, 17, 16, 30, 31]
    assert oc.flat_perm("z", 4, 4)[:10].tolist() == [12, 8, 13, 9, 4, 0, 5, 1, 14, 10]
    assert oc.flat_perm("peano", 24, 24)[:10].tolist() == [0, 1, 2, 26, 25, 24, 48, 49, 50, 51]
    assert oc.flat_perm("moore", 14, 14)[:10].tolist() == [98, 84, 85, 99, 100, 101, 87, 86, 72, 73]
    assert oc.hilbert2d_flat(8)[:10].tolist() == [0, 8, 9, 1, 2, 3, 11, 10, 18, 19]


@pytest.mark.parametrize("curve", cases.CURVES)
def test_square_perm_hashes(curve):
    for n in cases.PERM_SIZES:
        p = oc.flat_perm(curve, n, n)
        assert h16(p) == HASHES["square"][f"{curve}:{n}"], (curve, n)
        assert np.array_equal(np.sort(p), np.arange(n * n))
        if f"{curve}:{n}" in HASHES["heads"]:
            assert p[:10].tolist() == HASHES["heads"][f"{curve}:{n}"]


@pytest.mark.parametrize("curve", cases.CURVES)
def test_rect_perm_hashes(curve):
    for (w, h) in cases.PERM_RECTS:
        ij = oc.embed_and_prune(curve, w, h)
        assert h16(ij[:, 0] * h + ij[:, 1]) == HASHES["rect"][f"{curve}:{w}x{h}"]


def test_large_and_2d_hashes():
    assert h16(oc.flat_perm("hilbert", 1024, 1024)) == HASHES["square"]["hilbert:1024"]
    for g in cases.HILBERT2D_GRIDS:
        assert h16(oc.hilbert2d_flat(g)) == HASHES["hilbert2d"][str(g)]


@pytest.mark.parametrize("name", sorted(cases.TOKENIZER_CASES))
def test_tokenizer_fixture(name):
    kind, kw, shape = cases.TOKENIZER_CASES[name]
    gold = np.load(os.path.join(GOLD, "tokenizers.npz"))[name + "/out"]
    torch.manual_seed(cases.INIT_SEED)
    m = cases.build_oracle_tokenizer(kind, kw)
    with torch.no_grad():
        y = m(cases.make_input(shape))
    assert tuple(y.shape) == gold.shape
    assert cases.rel_l2(y, torch.from_numpy(gold)) < 1e-5


@pytest.mark.parametrize("name", sorted(cases.MODEL_CASES))
def test_model_fixture(name):
    vk, tcase, mkw, batch = cases.MODEL_CASES[name]
    kind, kw, shape = cases.TOKENIZER_CASES[tcase]
    meta = json.load(open(os.path.join(GOLD, "models.json")))[name]
    gold_logits = np.load(os.path.join(GOLD, "models.npz"))[name + "/logits"]
    torch.manual_seed(cases.INIT_SEED)
    m = om.zero_dropout(om.build_vit(vk, cases.build_oracle_tokenizer(kind, kw), **mkw))
    for k, v in m.state_dict().items():
        if v.dtype.is_floating_point:
            assert abs(float(v.double().abs().sum()) - meta["state_abs_sum"][k]) <= 1e-9 * max(1.0, meta["state_abs_sum"][k]), k
    m.train()
    x = cases.make_input((batch,) + tuple(shape[1:]))
    logits = m(x)
    loss = om.soft_target_cross_entropy(logits, cases.make_soft_targets(batch, mkw["num_classes"]))
    loss.backward()
    assert cases.rel_l2(logits, torch.from_numpy(gold_logits)) < 1e-5
    assert abs(float(loss) - meta["loss"]) < 1e-5
    gs = cases.grad_summary(m)
    assert set(gs) == set(meta["grads"])
    for k, (norm, _s, _f) in meta["grads"].items():
        assert abs(gs[k][0] - norm) <= 1e-4 * max(norm, 1e-6), k


@pytest.mark.parametrize("name", sorted(cases.ALTVIT_CASES))
def test_altvit_fixture(name):
    """oracle/altvit.py vs logits / loss / gradient norms recorded from the live reference's src/models/altvit.py."""
    from oracle import altvit as oa
    cls, kw, batch = cases.ALTVIT_CASES[name]
    meta = json.load(open(os.path.join(GOLD, "models.json")))["altvit/" + name]
    gold_logits = np.load(os.path.join(GOLD, "models.npz"))["altvit/" + name + "/logits"]
    torch.manual_seed(cases.INIT_SEED)
    m = getattr(oa, cls)(**kw)
    for k, v in m.state_dict().items():
        if v.dtype.is_floating_point:
            assert abs(float(v.double().abs().sum()) - meta["state_abs_sum"][k]) <= 1e-9 * max(1.0, meta["state_abs_sum"][k]), k
    m.train()
    logits = m(cases.make_input((batch, 3, kw["image_size"], kw["image_size"])))
    loss = om.soft_target_cross_entropy(logits, cases.make_soft_targets(batch, kw["num_classes"]))
    loss.backward()
    assert cases.rel_l2(logits, torch.from_numpy(gold_logits)) < 1e-5
    assert abs(float(loss.detach()) - meta["loss"]) < 1e-5
    gs = cases.grad_summary(m)
    assert set(gs) == set(meta["grads"])
    for k, (norm, _s, _f) in meta["grads"].items():
        assert abs(gs[k][0] - norm) <= 1e-4 * max(norm, 1e-6), k
