"""Vision Transformer variants — mirror of the reference's src/models/vit.py (same class names, constructor
signatures, parameter names/shapes and state_dict keys), executed by the libsfcvit kernels:

  patch_embed (fused curve gather + tcgen05 GEMM)  ->  [MixerBlock]  ->  post-norm ReLU encoder layers
  (tcgen05 GEMMs with fused epilogues, tcgen05 flash attention, warp-shuffle LayerNorm)  ->  factorised head.

Parameter containers reuse torch's own module classes (nn.TransformerEncoderLayer, nn.Linear, nn.LayerNorm) so that
initialisation under a given seed, parameter names and `load_state_dict` from a reference checkpoint are identical;
only `forward` is replaced. There is no stock-PyTorch fallback: on a non-CUDA tensor the kernels raise.
"""
import copy
import math

import torch
import torch.nn as nn

from sfcvit import functional as SF
from src.tokenizers.base_patch_embedding import BasePatchEmbedding


def _out_dtype(ref_param):
    """Logit dtype: bf16 under autocast (reference train/eval loops, train.py:90,155) or with bf16 parameters
    (main.py:157), else the parameter dtype."""
    if torch.is_autocast_enabled():
        return torch.bfloat16
    return ref_param.dtype


class TokenAggregator(nn.Module):
    """Depth-wise separable Conv1d token aggregation (reference vit.py:20-42). Every use in the reference is commented
    out (:362, :381); kept as a plain-torch utility so the name resolves — it is not on the kernel path."""

    def __init__(self, dim: int, k: int = 3, s: int = 1):
        super().__init__()
        self.dw = nn.Conv1d(dim, dim, k, s, padding=k // 2, groups=dim)
        self.pw = nn.Conv1d(dim, dim, 1, 1)
        self.act = nn.GELU()
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        return self.norm(self.act(self.pw(self.dw(x.transpose(1, 2))).transpose(1, 2)))


class SfcEncoderLayer(nn.TransformerEncoderLayer):
    """nn.TransformerEncoderLayer parameters (post-norm, ReLU, batch_first), forward = one fused EncoderLayerFn."""

    def forward(self, src, src_mask=None, src_key_padding_mask=None, is_causal=False, seed=0):
        if src_mask is not None or src_key_padding_mask is not None or is_causal:
            raise NotImplementedError("attention masks are not used by the reference ViTs and are not supported")
        if self.norm_first or self.self_attn.batch_first is not True:
            raise NotImplementedError("only the reference configuration (post-norm, batch_first) is supported")
        drops = (self.self_attn.dropout, self.dropout1.p, self.dropout.p, self.dropout2.p) if self.training else (0.0,) * 4
        return SF.encoder_layer(src, self, self.self_attn.num_heads, self.norm1.eps, drops, seed)


class SfcTransformerEncoder(nn.Module):
    """Stack of SfcEncoderLayer; parameter tree identical to nn.TransformerEncoder (`layers.{i}.…`)."""

    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])   # as torch's _get_clones
        self.num_layers = num_layers

    def forward(self, src, mask=None, src_key_padding_mask=None, is_causal=None):
        if mask is not None or src_key_padding_mask is not None:
            raise NotImplementedError("attention masks are not supported")
        x = src
        seed = SF.new_seed() if self.training else 0
        for i, layer in enumerate(self.layers):
            x = layer(x, seed=seed + 16 * i)
        return x


class TransformerSeqEncoder(nn.Module):
    """Reference vit.py:177-242. No CLS token, no positional embedding (both commented out in the reference)."""

    def __init__(self, input_dim, max_len, n_head, hidden_dim, method, dropout_p=0.1, n_layers=1):
        super().__init__()
        self.max_len = max_len
        self.grid_size = int(math.sqrt(max_len))
        encoder_layer = SfcEncoderLayer(d_model=input_dim, nhead=n_head, dim_feedforward=hidden_dim, dropout=dropout_p,
                                        batch_first=True)
        self.transformer = SfcTransformerEncoder(encoder_layer, num_layers=n_layers)
        self.to_patch_embedding = method

    def forward(self, x):
        return self.transformer(x)


class MixerBlock(nn.Module):
    """Reference vit.py:250-273: x + W2 gelu(W1 LN(x)); the token-mix branch exists as parameters but is unused."""

    def __init__(self, seq_len, embed_dim, hidden_dim, out_dim):
        super().__init__()
        self.token_mix_ln = nn.LayerNorm(embed_dim)
        self.channel_mix_ln = nn.LayerNorm(embed_dim)
        self.token_mix = nn.Sequential(nn.Linear(seq_len, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, seq_len))
        self.channel_mix = nn.Sequential(nn.Linear(embed_dim, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, out_dim))

    def forward(self, x):
        ln = self.channel_mix_ln
        h = SF.layer_norm(x, ln.weight, ln.bias, ln.eps)
        h = SF.linear(h, self.channel_mix[0].weight, self.channel_mix[0].bias, act=SF.ACT_GELU)
        return SF.linear(h, self.channel_mix[2].weight, self.channel_mix[2].bias, residual=x)


class FactorisedLinear(nn.Module):
    """Reference vit.py:276-292: (N*D) -> out_dim factorised as y = W_seq . (X W_emb^T); two plain GEMMs."""

    def __init__(self, seq_len, embed_dim, rank, out_dim):
        super().__init__()
        self.W_emb = nn.Parameter(torch.empty(rank, embed_dim))
        self.W_seq = nn.Parameter(torch.empty(out_dim, seq_len, rank))
        nn.init.xavier_normal_(self.W_emb)
        nn.init.xavier_normal_(self.W_seq)

    def forward(self, x, act=SF.ACT_NONE, drop_p=0.0, seed=0):
        B, N, _ = x.shape
        h = SF.linear(x, self.W_emb)                                    # 'bnd,rd->bnr'
        w2 = self.W_seq.reshape(self.W_seq.shape[0], -1)                 # 'bnr,onr->bo' == [B, N*r] x [o, N*r]^T
        return SF.linear(h.reshape(B, -1), w2, None, act=act, drop_p=drop_p, seed=seed)


class MultiLayerPredictor(nn.Sequential):
    """Reference vit.py:295-319. Same Sequential layout (index-based state_dict keys); the forward fuses
    GELU + Dropout into the epilogue of the preceding GEMM."""

    def __init__(self, embed_dim, seq_len, n_layers=2, rank=64, dropout_p=0.5, num_classes=10, mix=False):
        super().__init__()
        if mix:
            self.append(MixerBlock(seq_len, embed_dim, embed_dim * 2))   # TypeError as in the reference (missing out_dim)
        else:
            self.append(nn.LayerNorm(embed_dim))
        fact_out = embed_dim * 2
        self.append(FactorisedLinear(seq_len, embed_dim, rank, fact_out))
        self.append(nn.GELU())
        self.append(nn.Dropout(dropout_p))
        prev_dim = fact_out
        for _ in range(n_layers - 2):
            next_dim = prev_dim // 2
            self.append(nn.Linear(prev_dim, next_dim))
            self.append(nn.GELU())
            self.append(nn.Dropout(dropout_p))
            prev_dim = next_dim
        self.append(nn.Linear(prev_dim, num_classes))

    def forward(self, x):
        mods = list(self)
        seed = SF.new_seed() if self.training else 0
        i = 0
        out = x
        while i < len(mods):
            m = mods[i]
            nxt_gelu = i + 2 < len(mods) and isinstance(mods[i + 1], nn.GELU) and isinstance(mods[i + 2], nn.Dropout)
            if isinstance(m, nn.LayerNorm):
                out = SF.layer_norm(out, m.weight, m.bias, m.eps)
                i += 1
            elif isinstance(m, (FactorisedLinear, nn.Linear)):
                act, p = SF.ACT_NONE, 0.0
                if nxt_gelu:
                    act = SF.ACT_GELU
                    p = mods[i + 2].p if self.training else 0.0
                if isinstance(m, FactorisedLinear):
                    out = m(out, act=act, drop_p=p, seed=seed + i)
                else:
                    out = SF.linear(out, m.weight, m.bias, act=act, drop_p=p, seed=seed + i)
                i += 3 if nxt_gelu else 1
            else:
                out = m(out)
                i += 1
        return out


class VisionTransformer(nn.Module):
    """Reference vit.py:325-385: patch_embed -> encoder -> head. `embed_dim` is ignored (taken from patch_embed)."""

    def __init__(self, patch_embed: BasePatchEmbedding, embed_dim=128, depth=6, n_heads=4, mlp_dim=256, num_classes=10):
        super().__init__()
        self.patch_embed = patch_embed
        embed_dim = patch_embed.embed_dim
        self.encoder = TransformerSeqEncoder(input_dim=embed_dim, max_len=self.patch_embed.n_patches,
                                             method=self.patch_embed, n_head=n_heads, hidden_dim=mlp_dim, n_layers=depth)
        self.mlp_head = MultiLayerPredictor(embed_dim, self.patch_embed.n_patches, n_layers=2, num_classes=num_classes)

    @torch.compiler.disable
    def forward(self, x):
        with torch.autocast(device_type="cuda", enabled=False):
            t = self.patch_embed(x)
            t = self.encoder(t)
            out = self.mlp_head(t)
        return out.to(_out_dtype(self.mlp_head[-1].weight))


class VisionTransformer1D(nn.Module):
    """Reference vit.py:392-458: patch_embed -> MixerBlock -> encoder -> head (what main.py builds, :276-282)."""

    def __init__(self, patch_embed: BasePatchEmbedding, embed_dim=128, depth=6, n_heads=4, mlp_dim=256, num_classes=10):
        super().__init__()
        self.patch_embed = patch_embed
        embed_dim = patch_embed.embed_dim
        self.mlp_mixer = MixerBlock(seq_len=self.patch_embed.n_patches, embed_dim=embed_dim, hidden_dim=embed_dim * 2,
                                    out_dim=embed_dim)
        self.encoder = TransformerSeqEncoder(input_dim=embed_dim, max_len=self.patch_embed.n_patches, n_head=n_heads,
                                             hidden_dim=mlp_dim, n_layers=depth, method=self.patch_embed)
        self.mlp_head = MultiLayerPredictor(embed_dim, self.patch_embed.n_patches, n_layers=2, dropout_p=0.5,
                                            num_classes=num_classes)

    @torch.compiler.disable
    def forward(self, x):
        with torch.autocast(device_type="cuda", enabled=False):
            t = self.patch_embed(x)
            t = self.mlp_mixer(t)
            t = self.encoder(t)
            out = self.mlp_head(t)
        return out.to(_out_dtype(self.mlp_head[-1].weight))


class HierarchicalVisionTransformer1D(nn.Module):
    """Reference vit.py:465-545. The reference class cannot be constructed (MultiLayerPredictor(mix=True) calls
    MixerBlock without `out_dim`, :300-301) and its forward indexes a tensor as a list (:540-543); the same
    constructor failure is reproduced here rather than inventing semantics the reference never had."""

    def __init__(self, patch_embed: BasePatchEmbedding, embed_dim=128, depth=6, n_heads=4, mlp_dim=256, num_classes=10):
        super().__init__()
        self.patch_embed = patch_embed
        embed_dim = patch_embed.embed_dim
        self.encoder = nn.ModuleList([
            TransformerSeqEncoder(input_dim=embed_dim, max_len=patch_embed.patch_list[i], n_head=n_heads,
                                  hidden_dim=mlp_dim, n_layers=depth, method=self.patch_embed.levels[i])
            for i in range(patch_embed.depth)])
        self.fusion_encoder = TransformerSeqEncoder(input_dim=embed_dim, max_len=patch_embed.n_patches, n_head=n_heads,
                                                    hidden_dim=mlp_dim, n_layers=2, method=self.patch_embed)
        self.mlp_head = MultiLayerPredictor(embed_dim, self.patch_embed.n_patches, n_layers=2, dropout_p=0.5,
                                            num_classes=num_classes, mix=True)

    def forward(self, x):
        raise NotImplementedError("HierarchicalVisionTransformer1D is not runnable in the reference")
