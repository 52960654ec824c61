"""Hilbert 2-D patch tokenizer — mirror of the reference's src/tokenizers/_2D/hilbert_embedding.py:9-92:
Conv2d(kernel = stride = patch) followed by a reorder of the patch tokens along a Hilbert curve. Note the reference's
own orientation here is the TRANSPOSE of src.curves.hilbert_curve (it skips the final mirror/rotation, :47-78), and
`hilbert_indices` is a plain attribute (not a buffer, :28). Both are kept."""
import numpy as np
import torch
import torch.nn as nn

from src.curves.space_filling_curves import curve_permutation
from ..base_patch_embedding import BasePatchEmbedding, CurveGatherEmbedding


class HilbertEmbedding(BasePatchEmbedding, CurveGatherEmbedding):
    _k_order = "cp1p2"

    def __init__(self, img_size, patch_size, in_channels, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        self.n_patches = (img_size // patch_size) ** 2
        self.grid_size = img_size // patch_size
        self.hilbert_indices = self._get_hilbert_indices(self.grid_size)

    def _get_hilbert_indices(self, grid_size):
        """Flat token indices along the (un-transposed) Hilbert curve of order int(log2(grid)) on the unit square,
        scaled to the grid exactly as the reference does (int(x * grid), :41-44) — for a non-power-of-two grid this
        yields only 4^order tokens, a quirk of the reference that is preserved."""
        order = int(np.log2(grid_size))
        P = 2 ** order
        perm, _ = curve_permutation("hilbert_curve", P, P)
        flat = perm.cpu().to(torch.long)
        x_cell, y_cell = flat % P, flat // P                 # kernel (i, j) is the transpose of the raw recursion's (x, y)
        i2 = ((2 * x_cell + 1) * grid_size) // (2 * P)
        j2 = ((2 * y_cell + 1) * grid_size) // (2 * P)
        return i2 * grid_size + j2

    def hilbert_curve(self, order, size=1.0):
        """Un-transformed Hilbert cell centres on [0, size]^2 (reference method of the same name, :47-78)."""
        P = 2 ** order
        perm, _ = curve_permutation("hilbert_curve", P, P)
        flat = perm.cpu().numpy().astype(np.int64)
        cell = size / P
        return [((int(f % P) + 0.5) * cell, (int(f // P) + 0.5) * cell) for f in flat]

    def _flat_index(self):
        return self.hilbert_indices

    def forward(self, x):
        H, W, p = x.shape[-2], x.shape[-1], self.patch_size
        if x.dtype == torch.uint8:                  # decoded bytes [B, H, W, C] (set_uint8_normalization): no ragged border
            H, W = x.shape[1], x.shape[2]
            if H % p or W % p:
                x = x[:, : H // p * p, : W // p * p].contiguous()
        elif H % p or W % p:
            x = x[..., : H // p * p, : W // p * p].contiguous()
        return self._curve_forward(x, self.proj.weight, self.proj.bias, p, 1)
