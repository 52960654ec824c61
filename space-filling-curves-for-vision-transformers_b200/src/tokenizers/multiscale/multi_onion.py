"""Onion (spiral) multiscale tokenizer — mirror of the reference's src/tokenizers/multiscale/multi_onion.py
(HierarchicalOnionEmbedding :8-46, OnionEmbedding1D :49-105). Host-built spiral permutation, same fused kernel."""
import torch

from .._spiral import spiral_cells
from ._hierarchy import GroupedCurveLevel, HierarchicalCurveEmbedding


class OnionEmbedding1D(GroupedCurveLevel):
    index_buffer = "onion_indices"

    def __init__(self, img_size, pre_patch_size, group_patch_size, in_channels, embed_dim):
        super().__init__(img_size, pre_patch_size, group_patch_size, in_channels, embed_dim, None)

    def _build_indices(self, n):
        cells = spiral_cells(n, n)
        return torch.from_numpy(cells[:, 0] * n + cells[:, 1])

    def _onion_indices(self, size):
        return self._build_indices(size)


class HierarchicalOnionEmbedding(HierarchicalCurveEmbedding):
    level_cls = OnionEmbedding1D

    def __init__(self, img_size, in_channels, patch_size_list, embed_dim):
        super().__init__(img_size, in_channels, patch_size_list, embed_dim, None)

    def _make_level(self, img_size, pre, group, in_channels, embed_dim, curve_fn):
        return OnionEmbedding1D(img_size, pre, group, in_channels, embed_dim)
