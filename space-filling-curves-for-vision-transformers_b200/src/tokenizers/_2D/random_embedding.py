"""Random-order 2-D patch tokenizer — mirror of the reference's src/tokenizers/_2D/random_embedding.py:5-37: a fresh
random permutation of the patch tokens on every forward, fed to the same fused gather+GEMM operator."""
import torch
import torch.nn as nn

from ..base_patch_embedding import BasePatchEmbedding, CurveGatherEmbedding


class RandomEmbedding(BasePatchEmbedding, CurveGatherEmbedding):
    _k_order = "cp1p2"

    def __init__(self, img_size, patch_size, in_channels, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.embed_dim = embed_dim
        self.patch_size = patch_size
        self.n_patches = (img_size // patch_size) ** 2
        self._current = None

    def _flat_index(self):
        return self._current

    def forward(self, x):
        self._current = torch.randperm(self.n_patches)
        return self._curve_forward(x, self.proj.weight, self.proj.bias, self.patch_size, 1)
