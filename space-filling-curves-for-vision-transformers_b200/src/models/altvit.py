"""Pre-norm "alt" ViT family — mirror of the reference's src/models/altvit.py (same classes, constructor signatures,
parameter / buffer names, state_dict keys and RNG consumption at construction) running on the libsfcvit kernels:
LayerNorm (K5), projections and MLP as tcgen05 GEMMs with bias / GELU / residual epilogues (K3), flash attention (K4),
curve-ordered patch gather (K2's gather kernel; the patch vector is normalised before it is projected, altvit.py:97-99,
so the gather is not fused into the GEMM here). Position-embedding add and the mean pool are tiny element-wise /
reduction steps on [B, N, D] and stay in torch. CUDA only."""
import math

import torch
from torch import nn

from sfcvit import functional as SF
from sfcvit import ops


def pair(t):
    return t if isinstance(t, tuple) else (t, t)


def posemb_sincos_1d(n_pos, dim, temperature: float = 10000.0, dtype=torch.float32):
    """Vaswani sin / cos table [n_pos, dim], sin on even columns (reference altvit.py:16-41)."""
    position = torch.arange(n_pos, dtype=dtype).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2, dtype=dtype) * (-math.log(temperature) / dim))
    pe = torch.zeros(n_pos, dim, dtype=dtype)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def _kernel_order(vec, C, p):
    """LayerNorm affine vector over a patch in the reference's (p1 p2 c) feature order -> the gather's (c p1 p2) order."""
    return vec.reshape(p, p, C).permute(2, 0, 1).reshape(-1)


class _PatchNormProject(nn.Module):
    """Shared forward of both patch embeddings: gather (optionally in curve order) -> LN -> Linear -> LN."""

    def _embed(self, img, perm32, ln1, lin, ln2):
        p, C = self.patch_height, self.channels
        x = SF.patch_rows(img, perm32, p, 1)                                  # [B, N, K], K order (c, p1, p2)
        K = x.shape[-1]
        x = SF.layer_norm(x, _kernel_order(ln1.weight, C, p), _kernel_order(ln1.bias, C, p), ln1.eps)
        w = lin.weight.reshape(lin.weight.shape[0], p, p, C).permute(0, 3, 1, 2).reshape(lin.weight.shape[0], K)
        x = SF.linear(x, w, lin.bias)
        return SF.layer_norm(x, ln2.weight, ln2.bias, ln2.eps)


class HilbertPatchEmbedding(_PatchNormProject):
    """reference altvit.py:46-99 (power-of-two square grids; the order comes from the integer curve kernel K1)."""

    def __init__(self, *, image_size, patch_size, channels, dim):
        super().__init__()
        image_height, image_width = pair(image_size)
        patch_height, patch_width = pair(patch_size)
        self.grid_h = image_height // patch_height
        self.grid_w = image_width // patch_width
        assert self.grid_h == self.grid_w and (self.grid_h & (self.grid_h - 1)) == 0, \
            "Hilbert curve requires square grid size that is a power of 2."
        assert patch_height == patch_width, "square patches only"
        patch_dim = channels * patch_height * patch_width
        self.patch_height, self.patch_width, self.channels = patch_height, patch_width, channels
        self.layernorm1 = nn.LayerNorm(patch_dim)
        self.linear = nn.Linear(patch_dim, dim)
        self.layernorm2 = nn.LayerNorm(dim)
        self.hilbert_indices = self._hilbert_order(self.grid_h)                # plain attribute (not a buffer), as there
        self._perm32 = None

    def _hilbert_order(self, n):
        from src.curves.space_filling_curves import curve_permutation
        perm, _ = curve_permutation("hilbert_curve", n, n)
        return perm.long().cpu()

    def forward(self, x):
        if self._perm32 is None or self._perm32.device != x.device:
            self._perm32 = self.hilbert_indices.to(device=x.device, dtype=torch.int32)
        return self._embed(x, self._perm32, self.layernorm1, self.linear, self.layernorm2)


class FeedForward(nn.Module):
    """reference altvit.py:102-113: LN -> Linear -> GELU -> Linear (the residual is added by the caller; fused here
    through `residual`)."""

    def __init__(self, dim, hidden_dim):
        super().__init__()
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, dim))

    def forward(self, x, residual=None):
        ln, l1, _, l2 = self.net
        h = SF.layer_norm(x, ln.weight, ln.bias, ln.eps)
        h = SF.linear(h, l1.weight, l1.bias, act=SF.ACT_GELU)
        return SF.linear(h, l2.weight, l2.bias, residual=residual)


class Attention(nn.Module):
    """reference altvit.py:116-142: pre-norm, bias-free projections, softmax(QK^T dh^-1/2) V as flash attention."""

    def __init__(self, dim, heads=8, dim_head=64):
        super().__init__()
        inner_dim = dim_head * heads
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.norm = nn.LayerNorm(dim)
        self.attend = nn.Softmax(dim=-1)
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Linear(inner_dim, dim, bias=False)

    def forward(self, x, residual=None):
        h = SF.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        qkv = SF.linear(h, self.to_qkv.weight, None)
        out = SF.attention(qkv, self.heads, self.scale)
        return SF.linear(out, self.to_out.weight, None, residual=residual)


class Transformer(nn.Module):
    """reference altvit.py:145-160."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([Attention(dim, heads=heads, dim_head=dim_head), FeedForward(dim, mlp_dim)]))

    def forward(self, x):
        for attn, ff in self.layers:
            x = attn(x, residual=x)                    # x = attn(x) + x, the add fused into the out-projection epilogue
            x = ff(x, residual=x)
        return SF.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)


class _Patchify(nn.Module):
    """Parameter-free placeholder at index 0 of SimpleViT.to_patch_embedding (einops Rearrange in the reference)."""

    def forward(self, x):
        raise RuntimeError("SimpleViT gathers patches inside its fused embedding; this layer is a state_dict placeholder")


class SimpleViT(_PatchNormProject):
    """reference altvit.py:163-205 (raster patch order)."""

    def __init__(self, *, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim, channels=3, dim_head=64):
        super().__init__()
        image_height, image_width = pair(image_size)
        patch_height, patch_width = pair(patch_size)
        assert image_height % patch_height == 0 and image_width % patch_width == 0, 'Image dimensions must be divisible by the patch size.'
        assert patch_height == patch_width, "square patches only"
        patch_dim = channels * patch_height * patch_width
        self.patch_height, self.patch_width, self.channels = patch_height, patch_width, channels
        self.to_patch_embedding = nn.Sequential(_Patchify(), nn.LayerNorm(patch_dim), nn.Linear(patch_dim, dim), nn.LayerNorm(dim))
        n_pos = (image_height // patch_height) * (image_width // patch_width)
        self.posemb = posemb_sincos_1d(n_pos=n_pos, dim=dim)
        self.register_buffer("pos_embedding", self.posemb)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim)
        self.pool = "mean"
        self.to_latent = nn.Identity()
        self.linear_head = nn.Linear(dim, num_classes)
        self._n_pos, self._perm32 = n_pos, None

    @torch.compiler.disable
    def forward(self, img):
        if self._perm32 is None or self._perm32.device != img.device:
            self._perm32 = torch.arange(self._n_pos, dtype=torch.int32, device=img.device)
        _, ln1, lin, ln2 = self.to_patch_embedding
        with torch.autocast(device_type="cuda", enabled=False):
            x = self._embed(img, self._perm32, ln1, lin, ln2)
            x = x + self.pos_embedding.to(x.device, dtype=x.dtype)
            x = self.transformer(x)
            x = self.to_latent(x.mean(dim=1))
            out = _head(self, x)
        return out.to(_out_dtype(self.linear_head.weight))


def _head(self, x):
    return SF.linear(x, self.linear_head.weight, self.linear_head.bias)


def _out_dtype(ref_param):
    """bf16 under autocast (the reference's train / eval loops) or with bf16 parameters (main.py:157), else fp32."""
    return torch.bfloat16 if torch.is_autocast_enabled() else ref_param.dtype


class HilbertViT(nn.Module):
    """reference altvit.py:208-268."""

    def __init__(self, *, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim, channels=3, dim_head=64, T=4, h_param=3.0):
        super().__init__()
        image_height, image_width = pair(image_size)
        patch_height, patch_width = pair(patch_size)
        assert image_height % patch_height == 0 and image_width % patch_width == 0, 'Image dimensions must be divisible by the patch size.'
        self.grid_h = image_height // patch_height
        self.grid_w = image_width // patch_width
        assert self.grid_h == self.grid_w and (self.grid_h & (self.grid_h - 1)) == 0, \
            "Hilbert embedding requires square grid size that is a power of 2."
        self.to_patch_embedding = HilbertPatchEmbedding(image_size=image_size, patch_size=patch_size, channels=channels, dim=dim)
        hilbert_indices = self.to_patch_embedding.hilbert_indices
        n = hilbert_indices.numel()
        N = int(math.sqrt(n))
        assert N * N == n, "Hilbert indices must form a square grid."
        assert dim % 2 == 0, "Feature dimension must be even."
        pos = hilbert_indices.to(torch.float32).unsqueeze(1)
        i_ar = torch.arange(dim // 2, dtype=torch.float32).unsqueeze(0)
        two_pi = 2 * math.pi
        arg = (2.0 * i_ar * N ** 2 * pos * two_pi) / (T * n * dim) + h_param * (2.0 * i_ar * pos * two_pi) / dim
        self.register_buffer("pos_embedding", torch.cat([torch.sin(arg), torch.cos(arg)], dim=1).type(torch.float32))
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim)
        self.pool = "mean"
        self.to_latent = nn.Identity()
        self.linear_head = nn.Linear(dim, num_classes)

    @torch.compiler.disable
    def forward(self, img):
        with torch.autocast(device_type="cuda", enabled=False):
            x = self.to_patch_embedding(img)
            x = x + self.pos_embedding.to(x.device, dtype=x.dtype)
            x = self.transformer(x)
            x = self.to_latent(x.mean(dim=1))
            out = _head(self, x)
        return out.to(_out_dtype(self.linear_head.weight))
