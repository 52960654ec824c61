"""ORACLE (test infrastructure) — CPU restatement of the reference's tokenizers / models in plain stock PyTorch.

The reference's arithmetic for this path *is* stock torch (nn.TransformerEncoder, nn.MultiheadAttention,
F.scaled_dot_product_attention, nn.LayerNorm, nn.GELU, einops.rearrange; torch un-pinned in the reference, installed
here: 2.11.0), so this restatement uses the same torch ops with the reference's module layout and parameter names
(state_dicts are interchangeable with the reference and with the B200 mirror in `src/`). Curve indices come from the
C oracle (oracle/sfc_oracle.c), never from the product's kernels.

Each class cites the reference code it follows. Pinned against the live reference by tests/golden/make_golden.py
(bit-identical CPU outputs for the same seed and input) and re-checked from the committed fixtures in
tests/test_oracle_golden.py. Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import curves as oc


def _patch_rows(x, p):
    """einops 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' (reference multi_hilbert.py:78) with reshape/permute."""
    B, C, H, W = x.shape
    return x.reshape(B, C, H // p, p, W // p, p).permute(0, 2, 4, 3, 5, 1).reshape(B, (H // p) * (W // p), p * p * C)


class SFCEmbedding1D(nn.Module):
    """reference multiscale/multi_hilbert.py:43-84 (identical copies in multi_morton/peano/moore.py)."""

    def __init__(self, img_size, pre_patch_size, group_patch_size, in_channels, embed_dim, curve="hilbert"):
        super().__init__()
        self.pre_patch_size, self.group_patch_size, self.embed_dim = pre_patch_size, group_patch_size, embed_dim
        self.grid_size = img_size // pre_patch_size
        self.input_dim = in_channels * pre_patch_size * pre_patch_size * group_patch_size
        if curve is not None:
            self.register_buffer("sfc_indices", torch.from_numpy(oc.flat_perm(curve, self.grid_size, self.grid_size)).long())
        else:
            self.sfc_indices = None      # raster variant (multi_zigzag.py:55-95) has no buffer and no reorder
        self.proj = nn.Linear(self.input_dim, embed_dim)

    def forward(self, x):
        B = x.shape[0]
        t = _patch_rows(x, self.pre_patch_size)
        if self.sfc_indices is not None:
            t = t[:, self.sfc_indices]
        t = t.reshape(B, t.shape[1] // self.group_patch_size, -1)
        return self.proj(t)


class HierarchicalEmbedding(nn.Module):
    """reference multiscale/multi_hilbert.py:9-40 (and the morton/peano/moore/zigzag copies)."""

    def __init__(self, img_size, in_channels, patch_size_list, embed_dim, curve="hilbert"):
        super().__init__()
        self.levels = nn.ModuleList()
        pre, pre_list = 1, []
        for g in patch_size_list:
            self.levels.append(SFCEmbedding1D(img_size, pre, g, in_channels, embed_dim, curve))
            pre_list.append(pre)
            pre *= 2
        self.patch_list = [int(((img_size // ps) // np.sqrt(gs)) ** 2) for ps, gs in zip(pre_list, patch_size_list)]
        self.embed_dim = embed_dim * len(patch_size_list)
        self.depth = len(patch_size_list)
        self.n_patches = self.patch_list[0]
        self.fusion = nn.Linear(self.embed_dim, self.embed_dim)

    def forward(self, x):
        patches = [lvl(x) for lvl in self.levels]
        n_tokens = self.patch_list[0]
        for i in range(1, len(patches)):
            patches[i] = F.interpolate(patches[i].transpose(1, 2), size=n_tokens, mode="linear",
                                       align_corners=False).transpose(1, 2)
        return self.fusion(torch.cat(patches, dim=-1))


class PixelCurveEmbedding1D(nn.Module):
    """reference _1D/hilbert_embedding1D.py:9-44 (morton/peano/moore copies differ in curve and buffer name);
    curve=None follows _1D/zigzag_embedding1D.py:5-39 (raster, no buffer)."""

    BUFFER = {"hilbert": "hilbert_indices", "z": "z_indices", "peano": "peano_indices", "moore": "moore_indices"}

    def __init__(self, img_size, patch_size, in_channels, embed_dim, curve="hilbert"):
        super().__init__()
        self.n_patches = (img_size * img_size) // patch_size
        self.input_dim = in_channels * patch_size
        self.embed_dim = embed_dim
        self.curve = curve
        if curve is not None:
            self.register_buffer(self.BUFFER[curve], torch.from_numpy(oc.embed_and_prune(curve, img_size, img_size)).long())
        self.proj = nn.Linear(self.input_dim, embed_dim)

    def forward(self, x):
        B = x.shape[0]
        if self.curve is not None:
            idx = getattr(self, self.BUFFER[self.curve])
            t = x[:, :, idx[:, 0], idx[:, 1]].permute(0, 2, 1)
        else:
            t = x.flatten(2).transpose(1, 2)
        return self.proj(t.reshape(B, self.n_patches, self.input_dim))


class ConvPatchEmbedding(nn.Module):
    """reference _2D/zigzag_embedding.py:5-30 (hilbert=False) and _2D/hilbert_embedding.py:9-92 (hilbert=True: the
    un-transposed Hilbert order of :30-78 applied to the raster token sequence)."""

    def __init__(self, img_size, patch_size, in_channels, embed_dim, hilbert=False):
        super().__init__()
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.embed_dim = embed_dim
        self.n_patches = (img_size // patch_size) ** 2
        self.hilbert_indices = torch.from_numpy(oc.hilbert2d_flat(img_size // patch_size)).long() if hilbert else None

    def forward(self, x):
        t = self.proj(x).flatten(2).transpose(1, 2)
        if self.hilbert_indices is not None:
            t = t[:, self.hilbert_indices, :]
        return t


class TransformerSeqEncoder(nn.Module):
    """reference models/vit.py:177-242: stock nn.TransformerEncoder (post-norm, ReLU, dropout, batch_first)."""

    def __init__(self, input_dim, max_len, n_head, hidden_dim, method, dropout_p=0.1, n_layers=1):
        super().__init__()
        layer = nn.TransformerEncoderLayer(d_model=input_dim, nhead=n_head, dim_feedforward=hidden_dim, dropout=dropout_p,
                                           batch_first=True)
        self.transformer = nn.TransformerEncoder(layer, num_layers=n_layers)
        self.to_patch_embedding = method

    def forward(self, x):
        return self.transformer(x)


class MixerBlock(nn.Module):
    """reference models/vit.py:250-273 (token-mix branch present but unused)."""

    def __init__(self, seq_len, embed_dim, hidden_dim, out_dim):
        super().__init__()
        self.token_mix_ln = nn.LayerNorm(embed_dim)
        self.channel_mix_ln = nn.LayerNorm(embed_dim)
        self.token_mix = nn.Sequential(nn.Linear(seq_len, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, seq_len))
        self.channel_mix = nn.Sequential(nn.Linear(embed_dim, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, out_dim))

    def forward(self, x):
        return x + self.channel_mix(self.channel_mix_ln(x))


class FactorisedLinear(nn.Module):
    """reference models/vit.py:276-292."""

    def __init__(self, seq_len, embed_dim, rank, out_dim):
        super().__init__()
        self.W_emb = nn.Parameter(torch.empty(rank, embed_dim))
        self.W_seq = nn.Parameter(torch.empty(out_dim, seq_len, rank))
        nn.init.xavier_normal_(self.W_emb)
        nn.init.xavier_normal_(self.W_seq)

    def forward(self, x):
        h = torch.einsum("bnd, rd -> bnr", x, self.W_emb)
        return torch.einsum("bnr, onr -> bo", h, self.W_seq)


def make_head(embed_dim, seq_len, rank=64, dropout_p=0.5, num_classes=10):
    """reference models/vit.py:295-319 with n_layers=2, mix=False (the only runnable configuration)."""
    return nn.Sequential(nn.LayerNorm(embed_dim), FactorisedLinear(seq_len, embed_dim, rank, embed_dim * 2), nn.GELU(),
                         nn.Dropout(dropout_p), nn.Linear(embed_dim * 2, num_classes))


class VisionTransformer(nn.Module):
    """reference models/vit.py:325-385."""

    def __init__(self, patch_embed, depth=6, n_heads=4, mlp_dim=256, num_classes=10):
        super().__init__()
        self.patch_embed = patch_embed
        d = patch_embed.embed_dim
        self.encoder = TransformerSeqEncoder(d, patch_embed.n_patches, n_heads, mlp_dim, patch_embed, n_layers=depth)
        self.mlp_head = make_head(d, patch_embed.n_patches, num_classes=num_classes)

    def forward(self, x):
        return self.mlp_head(self.encoder(self.patch_embed(x)))


class VisionTransformer1D(nn.Module):
    """reference models/vit.py:392-458."""

    def __init__(self, patch_embed, depth=6, n_heads=4, mlp_dim=256, num_classes=10):
        super().__init__()
        self.patch_embed = patch_embed
        d = patch_embed.embed_dim
        self.mlp_mixer = MixerBlock(patch_embed.n_patches, d, d * 2, d)
        self.encoder = TransformerSeqEncoder(d, patch_embed.n_patches, n_heads, mlp_dim, patch_embed, n_layers=depth)
        self.mlp_head = make_head(d, patch_embed.n_patches, num_classes=num_classes)

    def forward(self, x):
        return self.mlp_head(self.encoder(self.mlp_mixer(self.patch_embed(x))))


def soft_target_cross_entropy(logits, targets):
    """reference main.py:45-51."""
    return -(targets * F.log_softmax(logits, dim=-1)).sum(dim=-1).mean()


def zero_dropout(model):
    """Parity runs zero every dropout, including nn.MultiheadAttention.dropout (a plain float)."""
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0
    return model


def build_vit(kind, tokenizer, *, depth, n_heads, mlp_dim, num_classes):
    cls = VisionTransformer1D if kind == "vit1d" else VisionTransformer
    return cls(tokenizer, depth=depth, n_heads=n_heads, mlp_dim=mlp_dim, num_classes=num_classes)
