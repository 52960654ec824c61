"""Tokenizer base classes.

``BasePatchEmbedding`` keeps the reference contract (src/tokenizers/base_patch_embedding.py:6-21):
``forward(x: [B, C, H, W]) -> [B, N, D]`` plus the attributes ``embed_dim`` / ``n_patches`` the models read.

``CurveGatherEmbedding`` is the one operator behind every curve tokenizer of the reference: pre-patches of size
p x p are serialised along a curve, g consecutive pre-patches form a token, and a Linear projects it. It runs as the
single fused kernel K2 (csrc/patch_embed.cu); the curve permutation comes from kernel K1 at construction time.
"""
from abc import ABC, abstractmethod

import torch
import torch.nn as nn

from sfcvit import functional as SF


class BasePatchEmbedding(nn.Module, ABC):
    @abstractmethod
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [B, C, H, W] -> patch embeddings [B, N, D]."""


class CurveGatherEmbedding(nn.Module):
    """Shared machinery (not a reference class). Subclasses set up ``self.proj`` and an index buffer, then call
    ``_curve_forward``. The int32 device permutation consumed by the kernel is derived from the module's (int64,
    state_dict-visible) index buffer, so loading a reference checkpoint keeps working."""

    _k_order = "p1p2c"       # feature order inside a token of the reference Linear tokenizers (multi_hilbert.py:78)

    def _flat_index(self) -> torch.Tensor:
        raise NotImplementedError

    def _perm32(self, device):
        src = self._flat_index()
        cache = getattr(self, "_perm_cache", None)
        if cache is not None and cache[0] is src and cache[1] == src._version and cache[2].device == device:
            return cache[2]
        perm = src.to(device=device, dtype=torch.int32).contiguous()
        object.__setattr__(self, "_perm_cache", (src, src._version, perm))
        return perm

    _u8_norm = None          # (mean, std) per channel for uint8 input; None = x / 255

    def set_uint8_normalization(self, mean, std):
        """Addition to the reference API: lets ``forward`` take the DECODED image bytes — uint8 ``[B, H, W, C]``, what the
        DataLoader holds before ``ToDtype(float32, scale=True)`` + ``Normalize(mean, std)`` (main.py:174-178). That
        stage is folded into the projection weight / bias, the kernel gathers the bytes themselves (4x fewer image
        bytes over PCIe and out of HBM than the fp32 NCHW tensor)."""
        self._u8_norm = (torch.as_tensor(mean, dtype=torch.float32), torch.as_tensor(std, dtype=torch.float32))
        return self

    def _curve_forward(self, x, weight, bias, pre_patch, group):
        if x.dim() != 4:
            raise ValueError(f"expected [B, C, H, W] (or uint8 [B, H, W, C]), got {tuple(x.shape)}")
        return SF.patch_embed(x, weight, bias, self._perm32(x.device), pre_patch, group, self._k_order,
                              norm=self._u8_norm if x.dtype == torch.uint8 else None)
