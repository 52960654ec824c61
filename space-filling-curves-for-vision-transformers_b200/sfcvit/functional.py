"""Autograd functions over the libsfcvit kernels.

Activations are bf16 (as under the reference's bf16 autocast, src/training/train.py:90,113,155); parameters may be
fp32 or bf16 (main.py:157 sets the default dtype to bf16) — a bf16 shadow is cached per parameter version, and
gradients are produced directly in the parameter's dtype by the GEMM / reduction epilogues.
"""
import weakref

import torch
from torch.autograd import Function

from . import ops

ACT_NONE, ACT_RELU, ACT_GELU = ops.ACT_NONE, ops.ACT_RELU, ops.ACT_GELU

_shadow = {}

# Parameters updated through raw pointers (FusedAdamW's kernel, CUDA-graph replays) never bump torch's `_version`, so
# the weight-conversion caches below are keyed on (tensor version, WEIGHTS_EPOCH); every such update calls
# bump_weights_epoch().
WEIGHTS_EPOCH = 0


def bump_weights_epoch():
    global WEIGHTS_EPOCH
    WEIGHTS_EPOCH += 1


def clear_weight_caches():
    """Drop every cached weight conversion (before / after a CUDA-graph capture: conversions must be re-run INSIDE the
    graph, and entries created during capture alias graph-private memory)."""
    _shadow.clear()
    _wk_cache.clear()
    _wk_u8_cache.clear()


# ---- gradient destinations: slices of a flat gradient bucket (src/training/optim.py) -------------------------------
# FusedAdamW registers, per parameter, the slice of its flat bucket. A backward function asks grad_out(param) for it and
# lets its wgrad GEMM / reduction kernel write the gradient THERE; autograd then adopts the returned view as `.grad`
# (no packing copy before the all-reduce). A slice is handed out at most once between two release_grad_claims() calls
# and only while `.grad` is None — any other situation (gradient accumulation, a parameter used twice) gets None and
# the ordinary freshly-allocated gradient, which the optimizer copies into the bucket itself.
_grad_buf = {}
_claimed = set()


def register_grad_buffer(param, flat, offset):
    _grad_buf[id(param)] = (weakref.ref(param), flat, int(offset))


def unregister_grad_buffers(params):
    for p in params:
        _grad_buf.pop(id(p), None)
        _claimed.discard(id(p))


def release_grad_claims():
    _claimed.clear()


def grad_out(param):
    if param is None:
        return None
    ent = _grad_buf.get(id(param))
    if ent is None or ent[0]() is not param or param.grad is not None or id(param) in _claimed:
        return None
    _claimed.add(id(param))
    n = param.numel()
    return ent[1][ent[2]:ent[2] + n].view(param.shape)


def as_bf16(t):
    """bf16 view/shadow of a parameter (no copy when it already is bf16). Cached on (tensor, version, weights epoch)."""
    if t is None:
        return None
    if t.dtype == torch.bfloat16:
        return t.detach()
    key = id(t)
    ent = _shadow.get(key)
    ver = (t._version, WEIGHTS_EPOCH)
    if ent is not None and ent[0]() is t and ent[1] == ver and ent[2].device == t.device:
        return ent[2]
    sh = t.detach().to(torch.bfloat16)
    _shadow[key] = (weakref.ref(t), ver, sh)
    if len(_shadow) > 4096:
        for k in [k for k, e in _shadow.items() if e[0]() is None]:
            del _shadow[k]
    return sh


def new_seed():
    """64-bit dropout seed from torch's CPU generator (reproducible under torch.manual_seed; no device sync)."""
    return int(torch.empty((), dtype=torch.int64).random_().item())


def _bf16_act(x):
    if not x.is_cuda:
        raise RuntimeError("sfcvit: CUDA tensors only (the B200 path has no CPU fallback)")
    return x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)


def _pad_cols(t, mult=8):
    n = t.shape[1]
    if n % mult == 0 and t.stride(1) == 1 and t.stride(0) % mult == 0:
        return t
    npad = (n + mult - 1) // mult * mult
    out = t.new_zeros((t.shape[0], npad))
    out[:, :n] = t
    return out


def _usable(buf, shape, dtype):
    return buf if (buf is not None and buf.dtype == dtype and tuple(buf.shape) == tuple(shape)) else None


def linear_backward(dy2, x2, wb, w_dtype, need_dx, need_dw, need_db, *, dx_aux=None, dx_aux_mode=ops.AUX_NONE,
                    dx_alpha=1.0, dx_residual=None, w_param=None, b_param=None):
    """Gradients of y = x W^T + b for dy2 [M,N], x2 [M,K], wb bf16 [N,K].
    dX = epilogue(dY.W) (optionally * relu-mask / gelu' of dx_aux, + dx_residual), dW = dY^T.X, db = colsum(dY).
    w_param / b_param: the nn.Parameters, so that dW / db can be written straight into their flat-bucket slices."""
    N = wb.shape[0]
    dyp = _pad_cols(dy2)                     # TMA needs 16-byte row strides (e.g. num_classes = 10)
    dx = dw = db = None
    if need_dx:
        wbp = wb
        if dyp.shape[1] != N:
            wbp = wb.new_zeros((dyp.shape[1], wb.shape[1]))
            wbp[:N] = wb
        dx = ops.gemm(dyp, wbp, b_mn=True, aux=dx_aux, aux_mode=dx_aux_mode, alpha=dx_alpha, residual=dx_residual)
    if need_dw:
        padded = dyp.shape[1] != N
        dst = _usable(grad_out(w_param), (N, x2.shape[1]), w_dtype) if not padded else None
        fuse_db = need_db and not padded
        dwp, db = ops.wgrad(dyp, x2, w_dtype, dw_out=dst, want_db=fuse_db,
                            db_out=_usable(grad_out(b_param), (N,), w_dtype) if fuse_db else None)
        dw = dwp[:N] if padded else dwp
    if need_db and db is None:
        db = ops.colsum(dy2, w_dtype, out=_usable(grad_out(b_param), (N,), w_dtype))
    return dx, dw, db


class LinearFn(Function):
    """y = dropout(act(x W^T + b)) (+ residual). Generic building block (mixer, head, fusion layers)."""

    @staticmethod
    def forward(ctx, x, w, b, act, residual, drop_p, seed, grad_mode=True):
        xs = x.shape
        K = xs[-1]
        x2 = _bf16_act(x).reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        wb, bb = as_bf16(w), as_bf16(b)
        N = wb.shape[0]
        res2 = None
        if residual is not None:
            res2 = _bf16_act(residual).reshape(-1, N)
            if not res2.is_contiguous():
                res2 = res2.contiguous()
        need_grad = grad_mode and any(ctx.needs_input_grad)   # needs_input_grad ignores no_grad(); the caller passes the mode
        want_pre = act == ACT_GELU and need_grad
        # skinny weight-streaming GEMM (the factorised head at small batch: M = B rows against W_seq [2D, N * 64]): a single
        # pass would leave all but N / 256 SMs idle, so K is split (auto factor) and bias + activation run in the reduce
        skinny = x2.shape[0] <= 128 and K >= 8192 and res2 is None and drop_p == 0.0 and not want_pre
        r = ops.gemm(x2, wb, bias=bb, act=act, residual=res2, want_pre=want_pre, drop_p=drop_p, drop_seed=seed,
                     splits=0 if skinny else 1)
        out, pre = r if want_pre else (r, None)
        ctx.act, ctx.drop_p, ctx.seed, ctx.xs = act, drop_p, seed, xs
        ctx.has_res = residual is not None
        ctx.has_bias = b is not None
        ctx.w_dtype = w.dtype
        ctx.w_param, ctx.b_param = (w if w.is_leaf else None), (b if b is not None and b.is_leaf else None)
        # relu: the (post-dropout) output itself is the mask source, but only when no residual was added on top
        aux = pre if act == ACT_GELU else (out if (act == ACT_RELU and residual is None) else None)
        if act == ACT_RELU and residual is not None and need_grad:
            raise RuntimeError("LinearFn: relu + residual needs the pre-residual output; not used by any reference module")
        ctx.save_for_backward(x2, wb, aux)
        return out.reshape(*xs[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, wb, aux = ctx.saved_tensors
        N = wb.shape[0]
        dy2 = _bf16_act(dy).reshape(-1, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        d_res = dy if ctx.has_res else None
        g = dy2
        if ctx.act == ACT_RELU:
            g = ops.act_bwd(dy2, aux, ops.AUX_RELU_MASK, alpha=ops.drop_inv_keep(ctx.drop_p))
        elif ctx.act == ACT_GELU:
            g = ops.act_bwd(dy2, aux, ops.AUX_GELU_GRAD, drop_p=ctx.drop_p, drop_seed=ctx.seed)
        elif ctx.drop_p > 0:
            g = ops.act_bwd(dy2, None, ops.AUX_NONE, drop_p=ctx.drop_p, drop_seed=ctx.seed)
        dx, dw, db = linear_backward(g, x2, wb, ctx.w_dtype, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                     ctx.has_bias and ctx.needs_input_grad[2], w_param=ctx.w_param, b_param=ctx.b_param)
        if dx is not None:
            dx = dx.reshape(ctx.xs)
        return dx, dw, db, None, d_res, None, None, None


def linear(x, w, b=None, act=ACT_NONE, residual=None, drop_p=0.0, seed=0):
    return LinearFn.apply(x, w, b, act, residual, float(drop_p), int(seed), torch.is_grad_enabled())


class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        xs = x.shape
        x2 = _bf16_act(x).reshape(-1, xs[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        gb = as_bf16(gamma)
        y, mean, rstd = ops.layernorm_fwd(x2, gb, as_bf16(beta), eps)
        ctx.save_for_backward(x2, mean, rstd, gb)
        ctx.xs, ctx.p_dtype = xs, gamma.dtype
        ctx.params = (gamma if gamma.is_leaf else None, beta if beta.is_leaf else None)
        return y.reshape(xs)

    @staticmethod
    def backward(ctx, dy):
        x2, mean, rstd, gb = ctx.saved_tensors
        dy2 = _bf16_act(dy).reshape(x2.shape)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx, dg, db, _, _ = ops.layernorm_bwd(dy2, x2, mean, rstd, gb, ctx.p_dtype, dgamma_out=grad_out(ctx.params[0]),
                                             dbeta_out=grad_out(ctx.params[1]))
        return dx.reshape(ctx.xs), dg, db, None


def layer_norm(x, gamma, beta, eps=1e-5):
    return LayerNormFn.apply(x, gamma, beta, float(eps))


class PatchEmbedFn(Function):
    """K2: tokens = Linear(curve-ordered patches) without materialising the im2col tensor (forward).
    w_ref is the reference-layout weight: 'p1p2c' -> [D, g*p*p*C] Linear weight (multi_hilbert.py:66,78-84),
    'cp1p2' -> Conv2d weight [D, C, p, p] (_2D/hilbert_embedding.py:18-23)."""

    @staticmethod
    def forward(ctx, img, w_ref, bias, perm32, p, g, k_order, pos, norm=None):
        if not img.is_cuda:
            raise RuntimeError("sfcvit: CUDA tensors only (the B200 path has no CPU fallback)")
        u8 = img.dtype == torch.uint8
        if not u8 and img.dtype not in (torch.float32, torch.bfloat16):
            img = img.float()
        img = img.contiguous()
        C = img.shape[3] if u8 else img.shape[1]
        D = w_ref.shape[0]
        K = g * p * p * C
        if u8:
            # decoded bytes, NHWC: tokens = W . ((x / 255 - mean) / std) + b  ==  (W * s) . x + (b - W . t) with per-channel
            # s = 1 / (255 std), t = mean / std — folded into the weight and bias, the kernel reads the bytes as they are
            wk, bias_k, scale_k, shift_k = kernel_weight_u8(w_ref, bias, C, p, g, k_order, norm)
            ctx.u8 = (scale_k, shift_k)
        else:
            wk, bias_k = kernel_weight(w_ref, C, p, g, k_order), as_bf16(bias)
            ctx.u8 = None
        out = ops.patch_embed_fwd(img, perm32, wk, bias_k, p, g, pos=as_bf16(pos))
        ctx.save_for_backward(img, perm32)
        ctx.cfg = (p, g, k_order, C, D, K, w_ref.dtype, tuple(w_ref.shape), bias is not None, pos is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        img, perm32 = ctx.saved_tensors
        p, g, k_order, C, D, K, w_dtype, w_shape, has_bias, has_pos = ctx.cfg
        if ctx.needs_input_grad[0]:
            raise RuntimeError("PatchEmbedFn: gradient w.r.t. the input image is not implemented")
        d2 = _bf16_act(dout).reshape(-1, D)
        if not d2.is_contiguous():
            d2 = d2.contiguous()
        dw = db = dpos = None
        if has_bias and ctx.needs_input_grad[2] or (ctx.u8 is not None and ctx.needs_input_grad[1]):
            db = ops.colsum(d2, w_dtype)
        if ctx.needs_input_grad[1]:
            A = ops.patch_gather(img, perm32, p, g)                       # [M, Kpad] curve-ordered im2col (backward only)
            if ctx.u8 is None:
                dwk = ops.gemm(d2, A, a_mn=True, b_mn=True, out_dtype=w_dtype, splits=0)[:, :K]
                if k_order == "p1p2c":
                    dw = dwk.reshape(D, g, C, p, p).permute(0, 1, 3, 4, 2).reshape(w_shape)
                else:
                    dw = dwk.reshape(w_shape)
            else:
                # A holds the raw bytes (K order q, p1, p2, c): dW[d, k] = s_k * (dOut^T A)[d, k] - t_k * db[d]
                scale_k, shift_k = ctx.u8
                G = ops.gemm(d2, A, a_mn=True, b_mn=True, out_dtype=torch.float32, splits=0)[:, :K]
                dwk = (G * scale_k - db.float()[:, None] * shift_k).to(w_dtype)
                if k_order == "p1p2c":
                    dw = dwk.reshape(w_shape)
                else:                                                    # Conv2d weight [D, C, p, p]
                    dw = dwk.reshape(D, p, p, C).permute(0, 3, 1, 2).reshape(w_shape)
        if not (has_bias and ctx.needs_input_grad[2]):
            db = None
        if has_pos and ctx.needs_input_grad[7]:
            dpos = dout.sum(0)
        return None, dw, db, None, None, None, None, dpos, None


_wk_cache = {}


def kernel_weight(w_ref, C, p, g, k_order):
    """Reference-layout projection weight -> kernel layout bf16 [D, Kpad] with K ordered (q, c, p1, p2), zero padded.
    The permutation is applied to the WEIGHT once per parameter version, never to the image data."""
    key = id(w_ref)
    ent = _wk_cache.get(key)
    if ent is not None and ent[0]() is w_ref and ent[1] == (w_ref._version, WEIGHTS_EPOCH) and ent[2].device == w_ref.device:
        return ent[2]
    D = w_ref.shape[0]
    K = g * p * p * C
    Kpad = ops.patch_embed_kpad(C, p, g)
    w = w_ref.detach()
    if k_order == "p1p2c":
        w = w.reshape(D, g, p, p, C).permute(0, 1, 4, 2, 3).reshape(D, K)
    else:
        assert g == 1
        w = w.reshape(D, K)
    wk = torch.zeros((D, Kpad), dtype=torch.bfloat16, device=w_ref.device)
    wk[:, :K] = w
    _wk_cache[key] = (weakref.ref(w_ref), (w_ref._version, WEIGHTS_EPOCH), wk)
    return wk


_wk_u8_cache = {}


def kernel_weight_u8(w_ref, bias, C, p, g, k_order, norm):
    """uint8 NHWC input: kernel weight bf16 [D, Kpad] with K ordered (q, p1, p2, c) and the per-channel value
    normalisation folded in, the matching bias, and the fp32 per-k scale / shift vectors the backward needs."""
    mean, std = norm if norm is not None else (None, None)
    key = id(w_ref)
    ver = (w_ref._version, bias._version if bias is not None else -1, id(mean), id(std), WEIGHTS_EPOCH)
    ent = _wk_u8_cache.get(key)
    if ent is not None and ent[0]() is w_ref and ent[1] == ver and ent[2][0].device == w_ref.device:
        return ent[2]
    D = w_ref.shape[0]
    K = g * p * p * C
    dev = w_ref.device
    mean_t = torch.zeros(C, device=dev) if mean is None else torch.as_tensor(mean, dtype=torch.float32, device=dev).reshape(C)
    std_t = torch.ones(C, device=dev) if std is None else torch.as_tensor(std, dtype=torch.float32, device=dev).reshape(C)
    w = w_ref.detach().float()
    if k_order == "p1p2c":
        w = w.reshape(D, K)                                               # already (q, p1, p2, c)
    else:
        assert g == 1
        w = w.reshape(D, C, p, p).permute(0, 2, 3, 1).reshape(D, K)
    scale_k = (1.0 / (255.0 * std_t)).repeat(K // C)                      # c is the fastest K index
    shift_k = (mean_t / std_t).repeat(K // C)
    wk = torch.zeros((D, ops.patch_embed_kpad(C, p, g)), dtype=torch.bfloat16, device=dev)
    wk[:, :K] = w * scale_k
    b0 = bias.detach().float() if bias is not None else torch.zeros(D, device=dev)
    bias_k = (b0 - w @ shift_k).to(torch.bfloat16)
    out = (wk, bias_k, scale_k, shift_k)
    _wk_u8_cache[key] = (weakref.ref(w_ref), ver, out)
    return out


def patch_embed(img, w_ref, bias, perm32, p, g, k_order="p1p2c", pos=None, norm=None):
    return PatchEmbedFn.apply(img, w_ref, bias, perm32, int(p), int(g), k_order, pos, norm)


class ConcatStreamsFn(Function):
    """cat([resample(s, n_tokens) for s in streams], dim=-1) for bf16 [B, N_l, D_l] token streams (kernel K7); the
    hierarchical tokenizers' interpolate + cat (reference multi_hilbert.py:30-40)."""

    @staticmethod
    def forward(ctx, n_tokens, *streams):
        B = streams[0].shape[0]
        dims = [s.shape[2] for s in streams]
        out = torch.empty((B, n_tokens, sum(dims)), dtype=torch.bfloat16, device=streams[0].device)
        off = 0
        for s in streams:
            ops.interp_concat_fwd(_bf16_act(s), out, off)
            off += s.shape[2]
        ctx.meta = [(s.shape[1], s.shape[2], s.dtype) for s in streams]
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = _bf16_act(dout)
        grads, off = [], 0
        for i, (Ns, D, dt) in enumerate(ctx.meta):
            grads.append(ops.interp_concat_bwd(dout, off, Ns, D).to(dt) if ctx.needs_input_grad[1 + i] else None)
            off += D
        return (None, *grads)


def concat_streams(streams, n_tokens):
    """Kernel K7. Every stream's feature size must be a multiple of 8 (16-byte vectors; the fusion GEMM that follows
    needs 16-byte row strides as well) — there is no stock-torch fallback."""
    bad = [s.shape[2] for s in streams if s.shape[2] % 8]
    if bad:
        raise RuntimeError(f"hierarchical tokenizer: per-level embed_dim must be a multiple of 8 (got {bad}); the kernel "
                           "path has no fallback")
    return ConcatStreamsFn.apply(n_tokens, *streams)


class EncoderLayerFn(Function):
    """One post-norm ReLU transformer encoder layer (torch.nn.TransformerEncoderLayer defaults used by the reference,
    vit.py:197-206): x1 = LN1(x + drop(out_proj(MHA(x)))), x2 = LN2(x1 + drop(W2 drop(relu(W1 x1 + b1)) + b2)).
    Forward: 4 tcgen05 GEMMs with fused bias/ReLU/dropout/residual epilogues + flash attention + 2 LayerNorm kernels.
    Backward: dgrad GEMMs with fused ReLU-mask / residual epilogues, split-K wgrad GEMMs, column-sum bias gradients."""

    @staticmethod
    def forward(ctx, x, in_w, in_b, out_w, out_b, w1, b1, w2, b2, g1, be1, g2, be2, heads, eps, drops, seed):
        B, N, D = x.shape
        x2 = _bf16_act(x).reshape(B * N, D)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        wi, wo, wf1, wf2 = as_bf16(in_w), as_bf16(out_w), as_bf16(w1), as_bf16(w2)
        g1b, g2b = as_bf16(g1), as_bf16(g2)
        s_attn, s_d1, s_ff, s_d2 = seed, seed + 1, seed + 2, seed + 3
        p_attn, p_d1, p_ff, p_d2 = drops      # attention-prob dropout, dropout1, FFN dropout, dropout2
        qkv = ops.gemm(x2, wi, bias=as_bf16(in_b))
        attn, lse = ops.attn_fwd(qkv, B, heads, N, drop_p=p_attn, drop_seed=s_attn)
        y1 = ops.gemm(attn, wo, bias=as_bf16(out_b), residual=x2, drop_p=p_d1, drop_seed=s_d1)
        x1, mean1, rstd1 = ops.layernorm_fwd(y1, g1b, as_bf16(be1), eps)
        h = ops.gemm(x1, wf1, bias=as_bf16(b1), act=ACT_RELU, drop_p=p_ff, drop_seed=s_ff)
        y2 = ops.gemm(h, wf2, bias=as_bf16(b2), residual=x1, drop_p=p_d2, drop_seed=s_d2)
        out, mean2, rstd2 = ops.layernorm_fwd(y2, g2b, as_bf16(be2), eps)
        ctx.save_for_backward(x2, qkv, attn, lse, y1, mean1, rstd1, x1, h, y2, mean2, rstd2, wi, wo, wf1, wf2, g1b, g2b)
        ctx.cfg = (B, N, D, heads, drops, seed, in_w.dtype)
        ctx.params = tuple(t if (t is not None and t.is_leaf) else None
                           for t in (in_w, in_b, out_w, out_b, w1, b1, w2, b2, g1, be1, g2, be2))
        return out.reshape(B, N, D)

    @staticmethod
    def backward(ctx, dout):
        (x2, qkv, attn, lse, y1, mean1, rstd1, x1, h, y2, mean2, rstd2, wi, wo, wf1, wf2, g1b, g2b) = ctx.saved_tensors
        B, N, D, heads, drops, seed, pdt = ctx.cfg
        p_attn, p_d1, p_ff, p_d2 = drops
        s_attn, s_d1, s_ff, s_d2 = seed, seed + 1, seed + 2, seed + 3
        inv_keep = ops.drop_inv_keep(p_ff)
        d2 = _bf16_act(dout).reshape(B * N, D)
        if not d2.is_contiguous():
            d2 = d2.contiguous()
        # gradient destinations: the parameters' slices of the optimizer's flat bucket when registered (else fresh tensors)
        (o_wi, o_bi, o_wo, o_bo, o_w1, o_b1, o_w2, o_b2, o_g1, o_be1, o_g2, o_be2) = (
            _usable(grad_out(t), t.shape, pdt) if t is not None else None for t in ctx.params)
        # LN2
        # LN2 backward also emits the dropout2-masked gradient and its column sums (= linear2.bias gradient)
        dy2, dg2, dbe2, dy2d, db2 = ops.layernorm_bwd(d2, y2, mean2, rstd2, g2b, pdt, drop_p=p_d2, drop_seed=s_d2, want_colsum=True,
                                                      dgamma_out=o_g2, dbeta_out=o_be2, colsum_out=o_b2)
        if dy2d is None:
            dy2d = dy2
        # linear2 (+ReLU/dropout mask fused into the dgrad epilogue)
        dh = ops.gemm(dy2d, wf2, b_mn=True, aux=h, aux_mode=ops.AUX_RELU_MASK, alpha=inv_keep)
        dw2 = ops.gemm(dy2d, h, a_mn=True, b_mn=True, out=o_w2, out_dtype=pdt, splits=0)
        # linear1 (+ residual gradient of x1 fused)
        dx1 = ops.gemm(dh, wf1, b_mn=True, residual=dy2)
        dw1, db1 = ops.wgrad(dh, x1, pdt, dw_out=o_w1, db_out=o_b1, want_db=True)       # db1 = one more accumulator column
        # LN1
        dy1, dg1, dbe1, dy1d, dbo = ops.layernorm_bwd(dx1, y1, mean1, rstd1, g1b, pdt, drop_p=p_d1, drop_seed=s_d1, want_colsum=True,
                                                      dgamma_out=o_g1, dbeta_out=o_be1, colsum_out=o_bo)
        if dy1d is None:
            dy1d = dy1
        # out_proj
        dattn = ops.gemm(dy1d, wo, b_mn=True)
        dwo = ops.gemm(dy1d, attn, a_mn=True, b_mn=True, out=o_wo, out_dtype=pdt, splits=0)
        # attention
        dqkv = ops.attn_bwd(qkv, attn, dattn, lse, B, heads, N, drop_p=p_attn, drop_seed=s_attn)
        # in_proj (+ residual gradient of x fused)
        dx = ops.gemm(dqkv, wi, b_mn=True, residual=dy1) if ctx.needs_input_grad[0] else None
        dwi, dbi = ops.wgrad(dqkv, x2, pdt, dw_out=o_wi, db_out=o_bi, want_db=True)
        if dx is not None:
            dx = dx.reshape(B, N, D)
        return (dx, dwi, dbi, dwo, dbo, dw1, db1, dw2, db2, dg1, dbe1, dg2, dbe2, None, None, None, None)


def encoder_layer(x, layer, heads, eps, drops, seed):
    """drops = (attention-prob dropout, dropout1, FFN dropout, dropout2) probabilities (all 0 in eval mode)."""
    a = layer.self_attn
    return EncoderLayerFn.apply(x, a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias,
                                layer.linear1.weight, layer.linear1.bias, layer.linear2.weight, layer.linear2.bias,
                                layer.norm1.weight, layer.norm1.bias, layer.norm2.weight, layer.norm2.bias,
                                int(heads), float(eps), tuple(float(d) for d in drops), int(seed))


class AttentionFn(Function):
    """softmax(Q K^T * scale) V on the packed projection qkv [B, N, 3 * H * 64] (K4); no probability tensor is kept."""

    @staticmethod
    def forward(ctx, qkv, heads, scale, drop_p, seed):
        B, N, D3 = qkv.shape
        q2 = _bf16_act(qkv).reshape(B * N, D3)
        if not q2.is_contiguous():
            q2 = q2.contiguous()
        out, lse = ops.attn_fwd(q2, B, heads, N, scale=scale, drop_p=drop_p, drop_seed=seed)
        ctx.save_for_backward(q2, out, lse)
        ctx.cfg = (B, N, heads, scale, drop_p, seed)
        return out.reshape(B, N, D3 // 3)

    @staticmethod
    def backward(ctx, dout):
        q2, out, lse = ctx.saved_tensors
        B, N, heads, scale, drop_p, seed = ctx.cfg
        d2 = _bf16_act(dout).reshape(out.shape)
        if not d2.is_contiguous():
            d2 = d2.contiguous()
        dqkv = ops.attn_bwd(q2, out, d2, lse, B, heads, N, scale=scale, drop_p=drop_p, drop_seed=seed)
        return dqkv.reshape(B, N, -1), None, None, None, None


def attention(qkv, heads, scale=None, drop_p=0.0, seed=0):
    if scale is None:
        scale = (qkv.shape[-1] // 3 // heads) ** -0.5
    return AttentionFn.apply(qkv, int(heads), float(scale), float(drop_p), int(seed))


class PatchRowsFn(Function):
    """Curve-ordered patch vectors [B, n_tokens, g*p*p*C] (bf16, feature order (q, c, p1, p2)) gathered straight from the
    NCHW image by sfc_patch_gather — for tokenizers that normalise the patch vector BEFORE projecting it
    (altvit.py:92-99), where the gather cannot be fused into the projection GEMM. No gradient w.r.t. the image."""

    @staticmethod
    def forward(ctx, img, perm32, p, g):
        if not img.is_cuda:
            raise RuntimeError("sfcvit: CUDA tensors only (the B200 path has no CPU fallback)")
        if img.dtype not in (torch.float32, torch.bfloat16):
            img = img.float()
        img = img.contiguous()
        B, C = img.shape[0], img.shape[1]
        K = g * p * p * C
        A = ops.patch_gather(img, perm32, p, g)
        ntok = perm32.numel() // g
        return A[:, :K].reshape(B, ntok, K) if A.shape[1] != K else A.reshape(B, ntok, K)

    @staticmethod
    def backward(ctx, dout):
        return None, None, None, None


def patch_rows(img, perm32, p, g=1):
    return PatchRowsFn.apply(img, perm32, int(p), int(g))


class SoftTargetCEFn(Function):
    """-(targets * log_softmax(logits.float(), -1)).sum(-1).mean() (reference main.py:45-51) as one kernel each way."""

    @staticmethod
    def forward(ctx, logits, targets):
        if not logits.is_cuda:
            raise RuntimeError("sfcvit: CUDA tensors only (the B200 path has no CPU fallback)")
        x = logits if logits.dtype in (torch.bfloat16, torch.float32) else logits.float()
        x = x.reshape(-1, x.shape[-1])
        if x.stride(-1) != 1:
            x = x.contiguous()
        t = targets.reshape(-1, targets.shape[-1]).to(torch.float32)
        if t.stride(-1) != 1:
            t = t.contiguous()
        loss, lse, tsum = ops.softce_fwd(x, t)
        ctx.save_for_backward(x, t, lse, tsum)
        ctx.shape, ctx.dtype = logits.shape, logits.dtype
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        x, t, lse, tsum = ctx.saved_tensors
        dx = ops.softce_bwd(x, t, lse, tsum, dloss.reshape(1).to(torch.float32).contiguous())
        return dx.reshape(ctx.shape).to(ctx.dtype), None


def soft_target_cross_entropy(logits, targets):
    return SoftTargetCEFn.apply(logits, targets)
