"""Hilbert pixel-curve tokenizer — mirror of the reference's src/tokenizers/_1D/hilbert_embedding1D.py:9-44.

The curve runs over *pixels* (img_size x img_size); a token is `patch_size` consecutive curve pixels with features in
(pixel, channel) order, projected by `proj`. Equivalent to the fused operator with pre-patch 1 and group patch_size."""
import torch
import torch.nn as nn

from src.curves.space_filling_curves import embed_and_prune_sfc, hilbert_curve
from ..base_patch_embedding import BasePatchEmbedding, CurveGatherEmbedding


class HilbertEmbedding1D(BasePatchEmbedding, CurveGatherEmbedding):
    def __init__(self, img_size, patch_size, in_channels, embed_dim):
        super().__init__()
        self.img_size = img_size
        self.patch_size = patch_size
        self.n_patches = (img_size * img_size) // patch_size
        self.input_dim = in_channels * patch_size
        self.embed_dim = embed_dim
        # [n*n, 2] (row, col) int64, same name/shape as the reference buffer
        self.register_buffer("hilbert_indices", torch.tensor(embed_and_prune_sfc(hilbert_curve, img_size, img_size), dtype=torch.long))
        self.proj = nn.Linear(self.input_dim, self.embed_dim)

    def _flat_index(self):
        idx = self.hilbert_indices
        cache = getattr(self, "_flat_cache", None)
        if cache is None or cache[0] is not idx or cache[1] != idx._version:
            object.__setattr__(self, "_flat_cache", (idx, idx._version, idx[:, 0] * self.img_size + idx[:, 1]))
        return self._flat_cache[2]

    def forward(self, x):
        """x: [B, C, H, W] -> [B, N, D]"""
        return self._curve_forward(x, self.proj.weight, self.proj.bias, 1, self.patch_size)
