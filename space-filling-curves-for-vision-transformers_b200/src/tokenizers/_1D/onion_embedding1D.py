"""Onion (spiral) pixel tokenizer — mirror of the reference's src/tokenizers/_1D/onion_embedding1D.py:10-76.
The spiral is a host-built permutation fed to the same fused gather+GEMM operator (no curve kernel involved)."""
import torch
import torch.nn as nn

from .._spiral import spiral_cells
from ..base_patch_embedding import BasePatchEmbedding, CurveGatherEmbedding


class OnionEmbedding1D(BasePatchEmbedding, CurveGatherEmbedding):
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
        self._flat = {}

    def onion_indices(self, c, d):
        cells = spiral_cells(c, d)
        return (cells[:, 0], cells[:, 1])

    def _flat_index(self):
        key = self.img_size
        if key not in self._flat:
            cells = spiral_cells(self.img_size, self.img_size)
            self._flat[key] = torch.from_numpy(cells[:, 0] * self.img_size + cells[:, 1])
        return self._flat[key]

    def forward(self, x):
        return self._curve_forward(x, self.proj.weight, self.proj.bias, 1, self.patch_size)
