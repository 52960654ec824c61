"""Generates the golden fixtures from the LIVE reference (/root/reference) — run in the build container only:

    python tests/golden/make_golden.py

Writes tests/golden/perm_hashes.json, tokenizers.npz, models.npz/json. While generating it asserts that the oracle
(oracle/) reproduces the reference bit-for-bit on the CPU (same seed -> same init -> same outputs), which is what
pins the oracle.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import cases  # noqa: E402
import ref_import  # noqa: E402
from oracle import curves as oc  # noqa: E402
from oracle import model as om  # noqa: E402
from oracle import altvit as oa  # noqa: E402


def h16(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<i8").tobytes()).hexdigest()[:16]


def main():
    torch.set_num_threads(1)          # deterministic CPU reductions
    ref = ref_import.load_reference()
    C = ref["curves"]
    fns = {"hilbert": C.hilbert_curve, "z": C.z_curve, "peano": C.peano_curve, "moore": C.moore_curve}

    # ---------------- permutations
    hashes = {"square": {}, "rect": {}, "hilbert2d": {}, "heads": {}}
    for name, fn in fns.items():
        for n in cases.PERM_SIZES:
            r = np.array([i * n + j for i, j in C.embed_and_prune_sfc(fn, n, n)], dtype=np.int64)
            assert np.array_equal(r, oc.flat_perm(name, n, n)), (name, n)
            hashes["square"][f"{name}:{n}"] = h16(r)
            if n in (4, 8, 14):
                hashes["heads"][f"{name}:{n}"] = r[:10].tolist()
        for (w, h) in cases.PERM_RECTS:
            r = np.array(C.embed_and_prune_sfc(fn, w, h), dtype=np.int64).reshape(-1, 2)
            assert np.array_equal(r, oc.embed_and_prune(name, w, h)), (name, w, h)
            hashes["rect"][f"{name}:{w}x{h}"] = h16(r[:, 0] * h + r[:, 1])
    r = np.array([i * 1024 + j for i, j in C.embed_and_prune_sfc(C.hilbert_curve, 1024, 1024)], dtype=np.int64)
    assert np.array_equal(r, oc.flat_perm("hilbert", 1024, 1024))
    hashes["square"]["hilbert:1024"] = h16(r)
    for g in cases.HILBERT2D_GRIDS:
        r = ref["hilbert_embedding"].HilbertEmbedding(g * 2, 2, 1, 4).hilbert_indices.numpy()
        assert np.array_equal(r, oc.hilbert2d_flat(g)), g
        hashes["hilbert2d"][str(g)] = h16(r)
    json.dump(hashes, open(os.path.join(HERE, "perm_hashes.json"), "w"), indent=1, sort_keys=True)
    print("perm hashes:", sum(len(v) for v in hashes.values()))

    # ---------------- tokenizers
    tok = {}
    for name, (kind, kw, shape) in cases.TOKENIZER_CASES.items():
        torch.manual_seed(cases.INIT_SEED)
        rm = cases.build_reference_tokenizer(ref, kind, kw)
        torch.manual_seed(cases.INIT_SEED)
        omod = cases.build_oracle_tokenizer(kind, kw)
        sd_r, sd_o = rm.state_dict(), omod.state_dict()
        assert set(sd_r) == set(sd_o), (name, set(sd_r) ^ set(sd_o))
        for k in sd_r:
            assert torch.equal(sd_r[k], sd_o[k]), (name, k)
        x = cases.make_input(shape)
        with torch.no_grad():
            yr, yo = rm(x), omod(x)
        assert torch.equal(yr, yo), (name, (yr - yo).abs().max())
        tok[name + "/out"] = yr.numpy()
        print("tokenizer", name, tuple(yr.shape))
    np.savez_compressed(os.path.join(HERE, "tokenizers.npz"), **tok)

    # ---------------- models (fwd + bwd)
    arrays, meta = {}, {}
    for name, (vk, tcase, mkw, batch) in cases.MODEL_CASES.items():
        kind, kw, shape = cases.TOKENIZER_CASES[tcase]
        torch.manual_seed(cases.INIT_SEED)
        rt = cases.build_reference_tokenizer(ref, kind, kw)
        rcls = ref["vit"].VisionTransformer1D if vk == "vit1d" else ref["vit"].VisionTransformer
        rm = om.zero_dropout(rcls(patch_embed=rt, **mkw))
        torch.manual_seed(cases.INIT_SEED)
        omod = om.zero_dropout(om.build_vit(vk, cases.build_oracle_tokenizer(kind, kw), **mkw))
        sd_r, sd_o = rm.state_dict(), omod.state_dict()
        assert set(sd_r) == set(sd_o), (name, set(sd_r) ^ set(sd_o))
        for k in sd_r:
            assert torch.equal(sd_r[k], sd_o[k]), (name, k)
        x = cases.make_input((batch,) + tuple(shape[1:]))
        tgt = cases.make_soft_targets(batch, mkw["num_classes"])
        res = {}
        for tag, m in (("ref", rm), ("oracle", omod)):
            m.train()
            m.zero_grad()
            logits = m(x)
            loss = om.soft_target_cross_entropy(logits, tgt)
            loss.backward()
            res[tag] = (logits.detach(), float(loss), cases.grad_summary(m))
        assert torch.equal(res["ref"][0], res["oracle"][0]), name
        assert res["ref"][1] == res["oracle"][1], name
        assert res["ref"][2] == res["oracle"][2], name
        arrays[name + "/logits"] = res["ref"][0].numpy()
        meta[name] = {"loss": res["ref"][1], "grads": res["ref"][2],
                      "state_abs_sum": {k: float(v.double().abs().sum()) for k, v in sd_r.items() if v.dtype.is_floating_point}}
        print("model", name, "loss", res["ref"][1], "params with grad", len(res["ref"][2]))
    # ---------------- altvit (pre-norm GELU ViT, src/models/altvit.py)
    for name, (cls, kw, batch) in cases.ALTVIT_CASES.items():
        torch.manual_seed(cases.INIT_SEED)
        rm = getattr(ref["altvit"], cls)(**kw)
        torch.manual_seed(cases.INIT_SEED)
        omod = getattr(oa, cls)(**kw)
        sd_r, sd_o = rm.state_dict(), omod.state_dict()
        assert list(sd_r) == list(sd_o), (name, set(sd_r) ^ set(sd_o))
        for k in sd_r:
            assert torch.equal(sd_r[k], sd_o[k]), (name, k)
        if cls == "HilbertViT":
            assert torch.equal(rm.to_patch_embedding.hilbert_indices, omod.to_patch_embedding.hilbert_indices), name
        x = cases.make_input((batch, 3, kw["image_size"], kw["image_size"]))
        tgt = cases.make_soft_targets(batch, kw["num_classes"])
        res = {}
        for tag, mdl in (("ref", rm), ("oracle", omod)):
            mdl.train()
            mdl.zero_grad()
            logits = mdl(x)
            loss = om.soft_target_cross_entropy(logits, tgt)
            loss.backward()
            res[tag] = (logits.detach(), float(loss), cases.grad_summary(mdl))
        assert torch.equal(res["ref"][0], res["oracle"][0]), name
        assert res["ref"][1] == res["oracle"][1], name
        assert res["ref"][2] == res["oracle"][2], name
        arrays["altvit/" + name + "/logits"] = res["ref"][0].numpy()
        meta["altvit/" + name] = {"loss": res["ref"][1], "grads": res["ref"][2],
                                  "state_abs_sum": {k: float(v.double().abs().sum()) for k, v in sd_r.items() if v.dtype.is_floating_point}}
        print("altvit", name, "loss", res["ref"][1], "params with grad", len(res["ref"][2]))
    np.savez_compressed(os.path.join(HERE, "models.npz"), **arrays)
    json.dump(meta, open(os.path.join(HERE, "models.json"), "w"), indent=1, sort_keys=True)
    print("done")


if __name__ == "__main__":
    main()
