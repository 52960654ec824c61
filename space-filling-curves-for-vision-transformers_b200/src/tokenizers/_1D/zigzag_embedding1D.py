"""Raster-scan pixel tokenizer — mirror of the reference's src/tokenizers/_1D/zigzag_embedding1D.py:5-39.
Row-major pixels, `patch_size` consecutive pixels per token: the fused operator with the identity permutation."""
import torch
import torch.nn as nn

from ..base_patch_embedding import BasePatchEmbedding, CurveGatherEmbedding


class RasterScan1DEmbedding(BasePatchEmbedding, CurveGatherEmbedding):
    def __init__(self, img_size, patch_size, in_channels, embed_dim):
        super().__init__()
        self.img_size = img_size
        self.patch_size = patch_size
        self.in_channels = in_channels
        self.embed_dim = embed_dim
        num_pixels = img_size * img_size
        assert num_pixels % patch_size == 0, "Image must be divisible into 1D patches"
        self.n_patches = num_pixels // patch_size
        self.input_dim = patch_size * in_channels
        self.proj = nn.Linear(self.input_dim, embed_dim)
        self._identity = torch.arange(num_pixels, dtype=torch.long)   # plain attribute: the reference has no index buffer

    def _flat_index(self):
        return self._identity

    def forward(self, x):
        return self._curve_forward(x, self.proj.weight, self.proj.bias, 1, self.patch_size)
