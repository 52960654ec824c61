"""Peano multiscale tokenizer — mirror of the reference's src/tokenizers/multiscale/multi_peano.py (HierarchicalPeanoEmbedding :9-40,
SFCEmbedding1D :43-84). Implementation shared in _hierarchy.py; each level is one fused curve-gather + GEMM kernel."""
from src.curves.space_filling_curves import embed_and_prune_sfc, peano_curve
from ._hierarchy import GroupedCurveLevel, HierarchicalCurveEmbedding


class SFCEmbedding1D(GroupedCurveLevel):
    def __init__(self, img_size, pre_patch_size, group_patch_size, in_channels, embed_dim, curve_fn=peano_curve):
        super().__init__(img_size, pre_patch_size, group_patch_size, in_channels, embed_dim, curve_fn)


class HierarchicalPeanoEmbedding(HierarchicalCurveEmbedding):
    level_cls = SFCEmbedding1D

    def __init__(self, img_size, in_channels, patch_size_list, embed_dim, curve_fn=peano_curve):
        super().__init__(img_size, in_channels, patch_size_list, embed_dim, curve_fn)
