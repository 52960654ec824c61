"""ORACLE (test infrastructure) — CPU restatement of the reference's pre-norm "alt" ViT (src/models/altvit.py) in stock
PyTorch: same module tree / parameter names (state_dicts interchange with the reference and with the B200 mirror in
src/models/altvit.py), einops replaced by reshape / permute, Hilbert order from the C oracle.

Pinned against the live reference by tests/golden/make_golden.py (bit-identical CPU logits, loss and gradient summaries
for the same seed) and re-checked from the committed fixtures by tests/test_oracle_golden.py."""
import math

import torch
import torch.nn as nn

from . import curves as oc


def posemb_sincos_1d(n_pos, dim, temperature=10000.0, dtype=torch.float32):
    """altvit.py:16-41: interleaved sin / cos of position * temperature^(-2i/dim)."""
    position = torch.arange(n_pos, dtype=dtype).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2, dtype=dtype) * (-math.log(temperature) / dim))
    pe = torch.zeros(n_pos, dim, dtype=dtype)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def patch_rows(x, p1, p2):
    """'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' (altvit.py:93-94, :171-172)."""
    B, C, H, W = x.shape
    return x.reshape(B, C, H // p1, p1, W // p2, p2).permute(0, 2, 4, 3, 5, 1).reshape(B, (H // p1) * (W // p2), p1 * p2 * C)


def hilbert_flat_indices(n):
    """altvit.py:68-87: flat index y*n + x along the integer Hilbert recursion == embed_and_prune_sfc(hilbert_curve) order
    (asserted against the live reference in make_golden.py)."""
    return torch.from_numpy(oc.flat_perm("hilbert", n, n)).long()


class HilbertPatchEmbedding(nn.Module):
    """altvit.py:46-99: patches in Hilbert order -> LayerNorm -> Linear -> LayerNorm (power-of-two square grids only)."""

    def __init__(self, *, image_size, patch_size, channels, dim):
        super().__init__()
        self.grid_h = self.grid_w = image_size // patch_size
        assert self.grid_h & (self.grid_h - 1) == 0, "Hilbert curve requires square grid size that is a power of 2."
        patch_dim = channels * patch_size * patch_size
        self.patch_height = self.patch_width = patch_size
        self.channels = channels
        self.layernorm1 = nn.LayerNorm(patch_dim)
        self.linear = nn.Linear(patch_dim, dim)
        self.layernorm2 = nn.LayerNorm(dim)
        self.hilbert_indices = hilbert_flat_indices(self.grid_h)       # plain attribute, as in the reference (:66)

    def forward(self, x):
        x = patch_rows(x, self.patch_height, self.patch_width)[:, self.hilbert_indices.to(x.device)]
        return self.layernorm2(self.linear(self.layernorm1(x)))


class FeedForward(nn.Module):
    """altvit.py:102-113."""

    def __init__(self, dim, hidden_dim):
        super().__init__()
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, dim))

    def forward(self, x):
        return self.net(x)


class Attention(nn.Module):
    """altvit.py:116-142: pre-norm, bias-free QKV / output projections, explicit softmax(QK^T * scale) V."""

    def __init__(self, dim, heads=8, dim_head=64):
        super().__init__()
        inner = dim_head * heads
        self.heads, self.scale = heads, dim_head ** -0.5
        self.norm = nn.LayerNorm(dim)
        self.attend = nn.Softmax(dim=-1)
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)

    def forward(self, x):
        B, N, _ = x.shape
        q, k, v = [t.reshape(B, N, self.heads, -1).permute(0, 2, 1, 3) for t in self.to_qkv(self.norm(x)).chunk(3, dim=-1)]
        attn = self.attend(torch.matmul(q, k.transpose(-1, -2)) * self.scale)
        out = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(B, N, -1)
        return self.to_out(out)


class Transformer(nn.Module):
    """altvit.py:145-160."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.layers = nn.ModuleList([nn.ModuleList([Attention(dim, heads=heads, dim_head=dim_head), FeedForward(dim, mlp_dim)])
                                     for _ in range(depth)])

    def forward(self, x):
        for attn, ff in self.layers:
            x = attn(x) + x
            x = ff(x) + x
        return self.norm(x)


class _Patchify(nn.Module):
    """Parameter-free stand-in for einops' Rearrange layer at index 0 of SimpleViT.to_patch_embedding (:171-172)."""

    def __init__(self, p1, p2):
        super().__init__()
        self.p1, self.p2 = p1, p2

    def forward(self, x):
        return patch_rows(x, self.p1, self.p2)


class SimpleViT(nn.Module):
    """altvit.py:163-205: raster patches, LN -> Linear -> LN, sincos 1-D position embedding, mean pool, linear head."""

    def __init__(self, *, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim, channels=3, dim_head=64):
        super().__init__()
        patch_dim = channels * patch_size * patch_size
        self.to_patch_embedding = nn.Sequential(_Patchify(patch_size, patch_size), nn.LayerNorm(patch_dim),
                                                nn.Linear(patch_dim, dim), nn.LayerNorm(dim))
        self.posemb = posemb_sincos_1d((image_size // patch_size) ** 2, dim)
        self.register_buffer("pos_embedding", self.posemb)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim)
        self.pool = "mean"
        self.to_latent = nn.Identity()
        self.linear_head = nn.Linear(dim, num_classes)

    def forward(self, img):
        x = self.to_patch_embedding(img)
        x = x + self.pos_embedding.to(x.device, dtype=x.dtype)
        return self.linear_head(self.to_latent(self.transformer(x).mean(dim=1)))


def hilbert_posemb(hilbert_indices, dim, T=4, h_param=3.0):
    """altvit.py:236-251: sin / cos of (scale + phase) driven by the Hilbert index of each token."""
    n = hilbert_indices.numel()
    N = int(math.sqrt(n))
    pos = hilbert_indices.to(torch.float32).unsqueeze(1)
    i_ar = torch.arange(dim // 2, dtype=torch.float32).unsqueeze(0)
    two_pi = 2 * math.pi
    arg = (2.0 * i_ar * N ** 2 * pos * two_pi) / (T * n * dim) + h_param * (2.0 * i_ar * pos * two_pi) / dim
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=1).type(torch.float32)


class HilbertViT(nn.Module):
    """altvit.py:208-268."""

    def __init__(self, *, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim, channels=3, dim_head=64, T=4,
                 h_param=3.0):
        super().__init__()
        self.grid_h = self.grid_w = image_size // patch_size
        self.to_patch_embedding = HilbertPatchEmbedding(image_size=image_size, patch_size=patch_size, channels=channels, dim=dim)
        self.register_buffer("pos_embedding", hilbert_posemb(self.to_patch_embedding.hilbert_indices, dim, T, h_param))
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim)
        self.pool = "mean"
        self.to_latent = nn.Identity()
        self.linear_head = nn.Linear(dim, num_classes)

    def forward(self, img):
        x = self.to_patch_embedding(img)
        x = x + self.pos_embedding.to(x.device, dtype=x.dtype)
        return self.linear_head(self.to_latent(self.transformer(x).mean(dim=1)))
