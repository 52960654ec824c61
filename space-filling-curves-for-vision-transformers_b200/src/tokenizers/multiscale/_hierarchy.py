"""Shared implementation of the reference's multiscale tokenizers (not a reference module).

``SFCEmbedding1D`` (reference multi_hilbert.py:43-84 and its morton/peano/moore copies) is the canonical fused
operator: pre-patches p x p in curve order, g consecutive ones per token, Linear(g*p*p*C -> D). The hierarchical
wrappers (multi_hilbert.py:9-40) stack L such levels with pre-patch sizes 1, 2, 4, ..., bring the coarser token
streams to the finest length with linear interpolation (identity when the lengths already match, as in main.py's
[16, 4, 1]), concatenate on the feature axis and apply a fusion Linear."""
import numpy as np
import torch
import torch.nn as nn

from sfcvit import functional as SF
from src.curves.space_filling_curves import embed_and_prune_sfc
from ..base_patch_embedding import CurveGatherEmbedding


class GroupedCurveLevel(CurveGatherEmbedding):
    """One level: image -> [B, n_final_patches, embed_dim]."""

    index_buffer = "sfc_indices"

    def __init__(self, img_size, pre_patch_size, group_patch_size, in_channels, embed_dim, curve_fn=None):
        super().__init__()
        assert img_size % pre_patch_size == 0, "Image size must be divisible by pre_patch_size"
        self.img_size = img_size
        self.pre_patch_size = pre_patch_size
        self.group_patch_size = group_patch_size
        self.in_channels = in_channels
        self.embed_dim = embed_dim
        self.curve_fn = curve_fn
        self.grid_size = img_size // pre_patch_size
        self.n_pre_patches = self.grid_size * self.grid_size
        self.n_final_patches = self.n_pre_patches // group_patch_size
        self.n_patches = self.n_final_patches      # superset of the reference: lets a single level feed VisionTransformer
        self.pre_patch_dim = in_channels * pre_patch_size * pre_patch_size
        self.input_dim = self.pre_patch_dim * group_patch_size
        idx = self._build_indices(self.grid_size)
        if self.index_buffer is not None:
            self.register_buffer(self.index_buffer, idx.long())
        else:
            self._identity = idx.long()
        self.proj = nn.Linear(self.input_dim, embed_dim)

    def _build_indices(self, n):
        curve = embed_and_prune_sfc(self.curve_fn, n, n)
        return torch.tensor([r * n + c for r, c in curve], dtype=torch.long)

    def _sfc_indices(self, n):           # reference method name (multi_hilbert.py:68-72)
        return self._build_indices(n)

    def _flat_index(self):
        return getattr(self, self.index_buffer) if self.index_buffer is not None else self._identity

    def forward(self, x):
        return self._curve_forward(x, self.proj.weight, self.proj.bias, self.pre_patch_size, self.group_patch_size)


class HierarchicalCurveEmbedding(nn.Module):
    level_cls = GroupedCurveLevel

    def __init__(self, img_size, in_channels, patch_size_list, embed_dim, curve_fn=None):
        super().__init__()
        self.levels = nn.ModuleList()
        pre, pre_list = 1, []
        for group in patch_size_list:
            self.levels.append(self._make_level(img_size, pre, group, in_channels, embed_dim, curve_fn))
            pre_list.append(pre)
            pre *= 2
        self.patch_list = [int(((img_size // ps) // np.sqrt(gs)) ** 2) for ps, gs in zip(pre_list, patch_size_list)]
        self.embed_dim = embed_dim * len(patch_size_list)
        self.depth = len(patch_size_list)
        self.n_patches = self.patch_list[0]
        self.fusion = nn.Linear(self.embed_dim, self.embed_dim)

    def _make_level(self, img_size, pre, group, in_channels, embed_dim, curve_fn):
        return self.level_cls(img_size, pre, group, in_channels, embed_dim, curve_fn)

    def forward(self, x):
        streams = [level(x) for level in self.levels]
        n_tokens = self.patch_list[0]
        # general-length case: kernel K7 resamples each stream along the token axis straight into
        # its column slice of the concat buffer (equal lengths degenerate to a copy)
        cat = SF.concat_streams(streams, n_tokens)        # raises for per-level widths that are not multiples of 8: no torch fallback
        return SF.linear(cat, self.fusion.weight, self.fusion.bias)
