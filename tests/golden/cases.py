"""Golden-case definitions shared by make_golden.py (live reference) and the tests (oracle / CUDA path)."""
import numpy as np
import torch

CURVES = ["hilbert", "z", "peano", "moore"]
PERM_SIZES = list(range(1, 34)) + [64, 81, 96, 224, 243, 384]
PERM_RECTS = [(6, 10), (10, 6), (1, 7), (9, 2), (30, 17), (224, 100)]
HILBERT2D_GRIDS = [1, 2, 4, 8, 16, 32, 14, 24]      # 14 / 24: the reference's non-power-of-two quirk (4^order tokens)

# name -> (builder kind, kwargs, input shape)
TOKENIZER_CASES = {
    "sfc_hilbert_32_p4_g1": ("sfc", dict(curve="hilbert", img_size=32, pre=4, group=1, C=3, D=64), (2, 3, 32, 32)),
    "sfc_morton_28_p2_g4": ("sfc", dict(curve="z", img_size=28, pre=2, group=4, C=3, D=64), (2, 3, 28, 28)),
    "sfc_peano_24_p8_g1": ("sfc", dict(curve="peano", img_size=24, pre=8, group=1, C=3, D=32), (3, 3, 24, 24)),
    "sfc_moore_32_p16_g2": ("sfc", dict(curve="moore", img_size=32, pre=16, group=2, C=3, D=48), (2, 3, 32, 32)),
    "pix_hilbert_16_ps16": ("pix", dict(curve="hilbert", img_size=16, patch_size=16, C=3, D=64), (2, 3, 16, 16)),
    "pix_peano_18_ps12": ("pix", dict(curve="peano", img_size=18, patch_size=12, C=3, D=32), (2, 3, 18, 18)),
    "pix_morton_14_ps4": ("pix", dict(curve="z", img_size=14, patch_size=4, C=1, D=40), (3, 1, 14, 14)),
    "pix_moore_12_ps9": ("pix", dict(curve="moore", img_size=12, patch_size=9, C=3, D=24), (1, 3, 12, 12)),
    "pix_raster_16_ps16": ("pix", dict(curve=None, img_size=16, patch_size=16, C=3, D=64), (2, 3, 16, 16)),
    "conv_hilbert_32_p4": ("conv", dict(hilbert=True, img_size=32, patch_size=4, C=3, D=64), (2, 3, 32, 32)),
    "conv_zigzag_32_p4": ("conv", dict(hilbert=False, img_size=32, patch_size=4, C=3, D=64), (2, 3, 32, 32)),
    "hier_morton_32": ("hier", dict(curve="z", img_size=32, C=3, groups=[16, 4, 1], D=64), (2, 3, 32, 32)),
    "hier_hilbert_16_interp": ("hier", dict(curve="hilbert", img_size=16, C=3, groups=[4, 4], D=32), (2, 3, 16, 16)),
    "hier_raster_32": ("hier", dict(curve=None, img_size=32, C=3, groups=[16, 4, 1], D=32), (2, 3, 32, 32)),
}

# name -> (vit kind, tokenizer case, model kwargs, batch)
MODEL_CASES = {
    "vit1d_hier_morton": ("vit1d", "hier_morton_32", dict(depth=2, n_heads=3, mlp_dim=256, num_classes=10), 4),
    "vit_conv_hilbert_tiny": ("vit", "conv_hilbert_32_p4_d192", dict(depth=2, n_heads=3, mlp_dim=384, num_classes=10), 4),
    "vit_sfc_hilbert_14x14": ("vit", "sfc_hilbert_56_p4_d128", dict(depth=1, n_heads=2, mlp_dim=256, num_classes=7), 3),
}
TOKENIZER_CASES["conv_hilbert_32_p4_d192"] = ("conv", dict(hilbert=True, img_size=32, patch_size=4, C=3, D=192), (4, 3, 32, 32))
# 14 x 14 grid: the "generalised Hilbert" of the north star; single-level hierarchical wrapper supplies n_patches
TOKENIZER_CASES["sfc_hilbert_56_p4_d128"] = ("hier", dict(curve="hilbert", img_size=56, C=3, groups=[1], D=128, pre0=4), (3, 3, 56, 56))

# reference src/models/altvit.py (pre-norm GELU ViT): name -> (class, kwargs, batch)
ALTVIT_CASES = {
    "hilbertvit_32_p4": ("HilbertViT", dict(image_size=32, patch_size=4, num_classes=10, dim=128, depth=2, heads=2, mlp_dim=256), 4),
    "simplevit_28_p4": ("SimpleViT", dict(image_size=28, patch_size=4, num_classes=7, dim=64, depth=1, heads=1, mlp_dim=128), 3),
    "hilbertvit_64_p8_tiny": ("HilbertViT", dict(image_size=64, patch_size=8, num_classes=10, dim=192, depth=1, heads=3, mlp_dim=384), 2),
}

INIT_SEED = 42      # main.py:151
DATA_SEED = 7


def make_input(shape, seed=DATA_SEED):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def make_soft_targets(batch, num_classes, seed=DATA_SEED):
    g = torch.Generator().manual_seed(seed + 1)
    ya = torch.randint(0, num_classes, (batch,), generator=g)
    yb = torch.randint(0, num_classes, (batch,), generator=g)
    lam = 0.3
    t = lam * torch.nn.functional.one_hot(ya, num_classes).float() + (1 - lam) * torch.nn.functional.one_hot(yb, num_classes).float()
    return t


def build_oracle_tokenizer(kind, kw):
    from oracle import model as om
    if kind == "sfc":
        return om.SFCEmbedding1D(kw["img_size"], kw["pre"], kw["group"], kw["C"], kw["D"], kw["curve"])
    if kind == "pix":
        return om.PixelCurveEmbedding1D(kw["img_size"], kw["patch_size"], kw["C"], kw["D"], kw["curve"])
    if kind == "conv":
        return om.ConvPatchEmbedding(kw["img_size"], kw["patch_size"], kw["C"], kw["D"], kw["hilbert"])
    if kind == "hier":
        if "pre0" in kw:      # single level with a non-unit pre-patch (a wrapper the reference API can express via img scaling)
            m = om.HierarchicalEmbedding(kw["img_size"] // kw["pre0"], kw["C"] * kw["pre0"] ** 2, kw["groups"], kw["D"], kw["curve"])
            return _Pre0Wrapper(m, kw["pre0"])
        return om.HierarchicalEmbedding(kw["img_size"], kw["C"], kw["groups"], kw["D"], kw["curve"])
    raise KeyError(kind)


class _Pre0Wrapper(torch.nn.Module):
    """Feeds p0 x p0 pixel blocks as channels so that the reference's hierarchical tokenizer (pre-patch 1 at level 0)
    sees a (H/p0) x (W/p0) 'image': a 14 x 14 patch grid of 4 x 4 patches through the reference's own classes."""

    def __init__(self, inner, p0):
        super().__init__()
        self.inner, self.p0 = inner, p0
        self.embed_dim, self.n_patches = inner.embed_dim, inner.n_patches

    def forward(self, x):
        B, C, H, W = x.shape
        p = self.p0
        x = x.reshape(B, C, H // p, p, W // p, p).permute(0, 3, 5, 1, 2, 4).reshape(B, p * p * C, H // p, W // p)
        return self.inner(x)


def build_reference_tokenizer(ref, kind, kw):
    cur = {"hilbert": "hilbert", "z": "morton", "peano": "peano", "moore": "moore"}
    if kind == "sfc":
        mod = ref["multi_" + cur[kw["curve"]]]
        return mod.SFCEmbedding1D(kw["img_size"], kw["pre"], kw["group"], kw["C"], kw["D"])
    if kind == "pix":
        if kw["curve"] is None:
            return ref["zigzag_embedding1D"].RasterScan1DEmbedding(kw["img_size"], kw["patch_size"], kw["C"], kw["D"])
        cls = {"hilbert": "HilbertEmbedding1D", "z": "MortonEmbedding1D", "peano": "PeanoEmbedding1D", "moore": "MooreEmbedding1D"}[kw["curve"]]
        return getattr(ref[cur[kw["curve"]] + "_embedding1D"], cls)(kw["img_size"], kw["patch_size"], kw["C"], kw["D"])
    if kind == "conv":
        if kw["hilbert"]:
            return ref["hilbert_embedding"].HilbertEmbedding(kw["img_size"], kw["patch_size"], kw["C"], kw["D"])
        return ref["zigzag_embedding"].ZigzagEmbedding(kw["img_size"], kw["patch_size"], kw["C"], kw["D"])
    if kind == "hier":
        if kw["curve"] is None:
            cls = ref["multi_zigzag"].HierarchicalRasterScanEmbedding
        else:
            name = {"hilbert": "HierarchicalHilbertEmbedding", "z": "HierarchicalMortonEmbedding",
                    "peano": "HierarchicalPeanoEmbedding", "moore": "HierarchicalMooreEmbedding"}[kw["curve"]]
            cls = getattr(ref["multi_" + cur[kw["curve"]]], name)
        if "pre0" in kw:
            return _Pre0Wrapper(cls(kw["img_size"] // kw["pre0"], kw["C"] * kw["pre0"] ** 2, kw["groups"], kw["D"]), kw["pre0"])
        return cls(kw["img_size"], kw["C"], kw["groups"], kw["D"])
    raise KeyError(kind)


def grad_summary(model):
    out = {}
    for n, p in model.named_parameters():
        if p.grad is not None:
            g = p.grad.detach().double()
            out[n] = [float(g.norm()), float(g.sum()), float(g.flatten()[0])]
    return out


def build_src_tokenizer(kind, kw):
    """The same case through the B200 mirror of the reference API (src.tokenizers.*)."""
    import importlib
    cur = {"hilbert": "hilbert", "z": "morton", "peano": "peano", "moore": "moore"}
    if kind == "sfc":
        mod = importlib.import_module("src.tokenizers.multiscale.multi_" + cur[kw["curve"]])
        return mod.SFCEmbedding1D(kw["img_size"], kw["pre"], kw["group"], kw["C"], kw["D"])
    if kind == "pix":
        if kw["curve"] is None:
            mod = importlib.import_module("src.tokenizers._1D.zigzag_embedding1D")
            return mod.RasterScan1DEmbedding(kw["img_size"], kw["patch_size"], kw["C"], kw["D"])
        cls = {"hilbert": "HilbertEmbedding1D", "z": "MortonEmbedding1D", "peano": "PeanoEmbedding1D", "moore": "MooreEmbedding1D"}[kw["curve"]]
        mod = importlib.import_module("src.tokenizers._1D." + cur[kw["curve"]] + "_embedding1D")
        return getattr(mod, cls)(kw["img_size"], kw["patch_size"], kw["C"], kw["D"])
    if kind == "conv":
        if kw["hilbert"]:
            return importlib.import_module("src.tokenizers._2D.hilbert_embedding").HilbertEmbedding(kw["img_size"], kw["patch_size"], kw["C"], kw["D"])
        return importlib.import_module("src.tokenizers._2D.zigzag_embedding").ZigzagEmbedding(kw["img_size"], kw["patch_size"], kw["C"], kw["D"])
    if kind == "hier":
        if kw["curve"] is None:
            cls = importlib.import_module("src.tokenizers.multiscale.multi_zigzag").HierarchicalRasterScanEmbedding
        else:
            name = {"hilbert": "HierarchicalHilbertEmbedding", "z": "HierarchicalMortonEmbedding",
                    "peano": "HierarchicalPeanoEmbedding", "moore": "HierarchicalMooreEmbedding"}[kw["curve"]]
            cls = getattr(importlib.import_module("src.tokenizers.multiscale.multi_" + cur[kw["curve"]]), name)
        if "pre0" in kw:
            return _Pre0Wrapper(cls(kw["img_size"] // kw["pre0"], kw["C"] * kw["pre0"] ** 2, kw["groups"], kw["D"]), kw["pre0"])
        return cls(kw["img_size"], kw["C"], kw["groups"], kw["D"])
    raise KeyError(kind)


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
