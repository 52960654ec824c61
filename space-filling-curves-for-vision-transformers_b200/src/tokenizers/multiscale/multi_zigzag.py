"""Raster multiscale tokenizer — mirror of the reference's src/tokenizers/multiscale/multi_zigzag.py
(HierarchicalRasterScanEmbedding :7-52, RasterScan1DGroupedEmbedding :55-95): pre-patches in row-major order, no
index buffer. Same fused kernel with the identity permutation."""
import torch

from ._hierarchy import GroupedCurveLevel, HierarchicalCurveEmbedding


class RasterScan1DGroupedEmbedding(GroupedCurveLevel):
    index_buffer = None          # the reference registers no buffer for raster order

    def __init__(self, img_size, pre_patch_size, group_patch_size, in_channels, embed_dim):
        super().__init__(img_size, pre_patch_size, group_patch_size, in_channels, embed_dim, None)

    def _build_indices(self, n):
        return torch.arange(n * n, dtype=torch.long)


class HierarchicalRasterScanEmbedding(HierarchicalCurveEmbedding):
    level_cls = RasterScan1DGroupedEmbedding

    def __init__(self, img_size, in_channels, patch_size_list, embed_dim):
        super().__init__(img_size, in_channels, patch_size_list, embed_dim, None)

    def _make_level(self, img_size, pre, group, in_channels, embed_dim, curve_fn):
        return RasterScan1DGroupedEmbedding(img_size, pre, group, in_channels, embed_dim)
