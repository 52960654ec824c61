"""Raster ("zigzag") 2-D patch tokenizer — mirror of the reference's src/tokenizers/_2D/zigzag_embedding.py:5-30:
Conv2d(kernel = stride = patch) then row-major flatten. Runs as the fused gather+GEMM kernel with the identity
permutation and the conv weight used in place (its (c, p1, p2) order is the kernel's native K order)."""
import torch
import torch.nn as nn

from ..base_patch_embedding import BasePatchEmbedding, CurveGatherEmbedding


class ZigzagEmbedding(BasePatchEmbedding, CurveGatherEmbedding):
    _k_order = "cp1p2"

    def __init__(self, img_size, patch_size, in_channels, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.embed_dim = embed_dim
        self.patch_size = patch_size
        self.n_patches = (img_size // patch_size) ** 2
        self._identity = torch.arange(self.n_patches, dtype=torch.long)

    def _flat_index(self):
        return self._identity

    def forward(self, x):
        H, W, p = x.shape[-2], x.shape[-1], self.patch_size
        if x.dtype == torch.uint8:                  # decoded bytes [B, H, W, C] (set_uint8_normalization): no ragged border
            H, W = x.shape[1], x.shape[2]
            if H % p or W % p:
                x = x[:, : H // p * p, : W // p * p].contiguous()
        elif H % p or W % p:                       # Conv2d(stride=p) silently drops the ragged border
            x = x[..., : H // p * p, : W // p * p].contiguous()
        return self._curve_forward(x, self.proj.weight, self.proj.bias, p, 1)
