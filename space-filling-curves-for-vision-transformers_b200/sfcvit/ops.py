"""Torch-tensor front-ends of the C ABI (device pointers + current stream). CUDA only."""
import ctypes

import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
AUX_NONE, AUX_RELU_MASK, AUX_GELU_GRAD = 0, 1, 2
CURVE_IDS = {"hilbert_curve": 0, "z_curve": 1, "peano_curve": 2, "moore_curve": 3, "raster_curve": 4,
             "hilbert": 0, "z": 1, "morton": 1, "peano": 2, "moore": 3, "raster": 4}


# ---- instrumentation (bench.py): kernel-launch counter and optional per-GEMM CUDA-event timing -------------
LAUNCHES = 0            # kernels launched through the C ABI by this process
GEMM_PROFILE = None     # when a list: (start_event, end_event, flops) appended per sfc_gemm_bf16 call
PE_PROFILE = None       # when a list: (start_event, end_event, algorithmic_bytes, flops) per sfc_patch_embed_fwd call
ATTN_PROFILE = None     # when a list: (start_event, end_event, flops = 4 B H N^2 dh) per sfc_attn_fwd call


def _count(n):
    global LAUNCHES
    LAUNCHES += n


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("sfcvit ops run on CUDA tensors only (no CPU fallback)")


_ws_cache = {}


def _workspace(nbytes, device):
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def curve_id(curve) -> int:
    name = curve if isinstance(curve, str) else getattr(curve, "__name__", None)
    if name not in CURVE_IDS:
        raise ValueError(f"Unknown SFC: {name}")
    return CURVE_IDS[name]


def curve_perm(curve, w: int, h: int, device="cuda"):
    """K1: (perm, inv) int32 device tensors for the w x h grid (reference embed_and_prune_sfc)."""
    lib = _lib.load()
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("curve_perm runs on a CUDA device only")
    cid = curve_id(curve)
    with torch.cuda.device(device):
        perm = torch.empty(w * h, dtype=torch.int32, device=device)
        inv = torch.empty(w * h, dtype=torch.int32, device=device)
        nbytes = lib.sfc_curve_perm_scratch_bytes(cid, w, h)
        scratch = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=device)
        _lib.check(lib.sfc_curve_perm(cid, w, h, _ptr(perm), _ptr(inv), _ptr(scratch), nbytes, _stream()), "sfc_curve_perm")
    _count(2)
    return perm, inv


def gemm(a, b, *, a_mn=False, b_mn=False, bias=None, residual=None, aux=None, aux_mode=AUX_NONE, act=ACT_NONE,
         alpha=1.0, out=None, out_dtype=torch.bfloat16, want_pre=False, splits=1, accumulate=False, drop_p=0.0,
         drop_seed=0, colsum_out=None):
    """D[M,N] = epilogue(alpha * A.B^T).  a: [M,K] (or [K,M] if a_mn), b: [N,K] (or [K,N] if b_mn); bf16, 2-D,
    inner dimension contiguous. Returns out (and out_pre when want_pre).
    colsum_out: optional [M] tensor receiving sum_k A[m, k] from the same kernel (wgrad: the bias gradient); only when
    gemm_colsum_splits(...) > 0 — pass that value as `splits`."""
    lib = _lib.load()
    _require_cuda(a, b, bias, residual, aux, out)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1
    if a_mn:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, (a.shape, b.shape, a_mn, b_mn)
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype in (torch.bfloat16, torch.float32)
    pre = torch.empty((M, N), dtype=torch.bfloat16, device=a.device) if want_pre else None
    if pre is not None:
        assert out.stride(0) == pre.stride(0)
    ep = _lib.SfcGemmEpilogue()
    for t in (bias, residual, aux):
        assert t is None or (t.dtype == torch.bfloat16 and t.stride(-1) == 1)
    ep.bias = bias.data_ptr() if bias is not None else None
    ep.residual = residual.data_ptr() if residual is not None else None
    ep.aux = aux.data_ptr() if aux is not None else None
    ep.out = out.data_ptr()
    ep.out_pre = pre.data_ptr() if pre is not None else None
    ep.ld_out = out.stride(0)
    ep.ld_res = residual.stride(0) if residual is not None else 0
    ep.ld_aux = aux.stride(0) if aux is not None else 0
    ep.alpha = float(alpha)
    ep.act, ep.aux_mode = int(act), int(aux_mode)
    ep.out_fp32 = 1 if out.dtype == torch.float32 else 0
    ep.accumulate = 1 if accumulate else 0
    ep.drop_p = float(drop_p)
    ep.drop_seed = int(drop_seed) & 0xFFFFFFFFFFFFFFFF
    if colsum_out is not None:
        assert colsum_out.numel() == M and colsum_out.is_contiguous() and colsum_out.dtype in (torch.bfloat16, torch.float32)
        ep.colsum_out = colsum_out.data_ptr()
        ep.colsum_fp32 = 1 if colsum_out.dtype == torch.float32 else 0
    else:
        ep.colsum_out = None
        ep.colsum_fp32 = 0
    if splits is None or splits == 0:
        splits = lib.sfc_gemm_suggest_splits(M, N, K)
    ws, ws_bytes = None, 0
    if splits > 1:
        ws_bytes = lib.sfc_gemm_workspace_bytes(M, N, K, splits)
        ws = _workspace(ws_bytes, a.device)
    prof = GEMM_PROFILE
    with torch.cuda.device(a.device):
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        rc = lib.sfc_gemm_bf16(_ptr(a), int(a_mn), a.stride(0), _ptr(b), int(b_mn), b.stride(0), M, N, K,
                               ctypes.byref(ep), _ptr(ws), ws_bytes, int(splits), _stream())
        if prof is not None:
            e1.record()
            prof.append((e0, e1, 2.0 * M * N * K))
    _lib.check(rc, "sfc_gemm_bf16")
    _count(2 if splits > 1 else 1)
    return (out, pre) if want_pre else out


def wgrad(dy, x, w_dtype, *, dw_out=None, db_out=None, want_db=False):
    """dW[N, K] = dY[M, N]^T . X[M, K] (split-K over the M tokens) and, when want_db, db[N] = sum_m dY[m, :] — as one more
    accumulator column of the SAME kernel when the shape is covered (no second pass over dY), else by sfc_colsum."""
    lib = _lib.load()
    Mtok, N = dy.shape
    K = x.shape[1]
    db = None
    if want_db:
        s = lib.sfc_gemm_colsum_splits(N, K, Mtok, 1, 1)
        if s > 0 and dy.data_ptr() % 16 == 0 and x.data_ptr() % 16 == 0:
            db = _param_out(db_out, N, w_dtype, dy.device)
            dw = gemm(dy, x, a_mn=True, b_mn=True, out=dw_out, out_dtype=w_dtype, splits=s, colsum_out=db)
            return dw, db
        db = colsum(dy, w_dtype, out=db_out)
    dw = gemm(dy, x, a_mn=True, b_mn=True, out=dw_out, out_dtype=w_dtype, splits=0)
    return dw, db


# ------------------------------------------------------------------ K5: LayerNorm / column sums
def layernorm_fwd(x, gamma, beta, eps=1e-5, want_stats=True):
    """x: bf16 [rows, D] contiguous. Returns (y, mean, rstd)."""
    lib = _lib.load()
    _require_cuda(x, gamma, beta)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and gamma.dtype == torch.bfloat16 and beta.dtype == torch.bfloat16
    D = x.shape[-1]
    rows = x.numel() // D
    y = torch.empty_like(x)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device) if want_stats else None
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if want_stats else None
    with torch.cuda.device(x.device):
        _lib.check(lib.sfc_layernorm_fwd(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(mean), _ptr(rstd), rows, D,
                                         float(eps), _stream()), "sfc_layernorm_fwd")
    _count(1)
    return y, mean, rstd


def drop_inv_keep(p):
    """Exact rescale used by the kernels: keep probability is 1 - floor(p * 65536) / 65536 (csrc/gemm_epilogue.cuh)."""
    if p <= 0:
        return 1.0
    import numpy as np
    thr = int(np.float32(p) * np.float32(65536.0))
    return 65536.0 / (65536.0 - thr)


def _param_out(buf, n, dtype, device):
    """`buf` when it is a usable destination for an [n] gradient in `dtype` (a slice of a flat gradient bucket), else a
    fresh tensor."""
    if buf is not None and buf.dtype == dtype and buf.numel() == n and buf.is_contiguous():
        return buf
    return torch.empty(n, dtype=dtype, device=device)


def layernorm_bwd(dy, x, mean, rstd, gamma, param_dtype=torch.bfloat16, *, drop_p=0.0, drop_seed=0, want_colsum=False,
                  dgamma_out=None, dbeta_out=None, colsum_out=None):
    """Returns (dx, dgamma, dbeta, dx_drop, colsum): dx_drop is the dropout-masked copy of dx (None when drop_p == 0),
    colsum the column sums of dx_drop (or dx) in param_dtype (None unless want_colsum). *_out: optional destinations
    (slices of a flat gradient bucket)."""
    lib = _lib.load()
    _require_cuda(dy, x, mean, rstd, gamma)
    assert dy.dtype == torch.bfloat16 and dy.is_contiguous() and x.is_contiguous()
    D = x.shape[-1]
    rows = x.numel() // D
    dx = torch.empty_like(x)
    dxd = torch.empty_like(x) if drop_p > 0 else None
    dgamma = _param_out(dgamma_out, D, param_dtype, x.device)
    dbeta = _param_out(dbeta_out, D, param_dtype, x.device)
    csum = _param_out(colsum_out, D, param_dtype, x.device) if want_colsum else None
    nbytes = lib.sfc_layernorm_bwd_scratch_bytes(rows, D)
    scratch = _workspace(nbytes, x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.sfc_layernorm_bwd(_ptr(dy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(dx), _ptr(dxd),
                                         float(drop_p), int(drop_seed) & 0xFFFFFFFFFFFFFFFF, _ptr(dgamma), _ptr(dbeta),
                                         _ptr(csum), 1 if param_dtype == torch.float32 else 0, 0, _ptr(scratch), nbytes,
                                         rows, D, _stream()), "sfc_layernorm_bwd")
    _count(2)
    return dx, dgamma, dbeta, dxd, csum


def colsum(x, out_dtype=torch.bfloat16, out=None):
    """x: bf16 [rows, N] (row stride arbitrary, inner contiguous) -> [N] column sums (bias gradient)."""
    lib = _lib.load()
    _require_cuda(x)
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1
    rows, N = x.shape
    out = _param_out(out, N, out_dtype, x.device)
    nbytes = lib.sfc_colsum_scratch_bytes(rows, N)
    scratch = _workspace(nbytes, x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.sfc_colsum(_ptr(x), x.stride(0), rows, N, _ptr(out), 1 if out_dtype == torch.float32 else 0, 0,
                                  _ptr(scratch), nbytes, _stream()), "sfc_colsum")
    _count(2)
    return out


# ------------------------------------------------------------------ K7: token-axis linear resampling into a concat slice
def interp_concat_fwd(src, dst, col_off):
    """src bf16 [B, Ns, D] -> dst[:, :, col_off : col_off + D] of the bf16 [B, Nd, Dtot] concat buffer
    (F.interpolate(mode='linear', align_corners=False) along tokens; a copy when Ns == Nd)."""
    lib = _lib.load()
    _require_cuda(src, dst)
    assert src.dtype == dst.dtype == torch.bfloat16 and src.dim() == dst.dim() == 3 and src.shape[0] == dst.shape[0]
    assert src.stride(2) == 1 and dst.stride(2) == 1 and src.stride(0) == src.shape[1] * src.stride(1) and dst.is_contiguous()
    B, Ns, D = src.shape
    with torch.cuda.device(src.device):
        _lib.check(lib.sfc_interp_concat_fwd(_ptr(src), src.stride(1), B, Ns, D, ctypes.c_void_p(dst.data_ptr() + 2 * col_off),
                                             dst.stride(1), dst.shape[1], _stream()), "sfc_interp_concat_fwd")
    _count(1)
    return dst


def interp_concat_bwd(ddst, col_off, Ns, D):
    """Transposed operator: gradient of the [B, Ns, D] stream from the concat buffer's gradient slice."""
    lib = _lib.load()
    _require_cuda(ddst)
    assert ddst.dtype == torch.bfloat16 and ddst.dim() == 3 and ddst.is_contiguous()
    B, Nd, _ = ddst.shape
    dsrc = torch.empty((B, Ns, D), dtype=torch.bfloat16, device=ddst.device)
    with torch.cuda.device(ddst.device):
        _lib.check(lib.sfc_interp_concat_bwd(ctypes.c_void_p(ddst.data_ptr() + 2 * col_off), ddst.stride(1), B, Nd, D, _ptr(dsrc), D, Ns,
                                             _stream()), "sfc_interp_concat_bwd")
    _count(1)
    return dsrc


# ------------------------------------------------------------------ K2: fused patch embed
def patch_embed_kpad(C, p, g):
    return _lib.load().sfc_patch_embed_kpad(C, p, g)


IMG_F32_NCHW, IMG_BF16_NCHW, IMG_U8_NHWC = 0, 1, 2      # include/sfcvit.h SFC_IMG_*


def _img_format(img):
    """(format code, (B, C, H, W)): float images are NCHW, uint8 images are decoded bytes in NHWC."""
    if img.dtype == torch.uint8:
        B, H, W, C = img.shape
        return IMG_U8_NHWC, (B, C, H, W)
    B, C, H, W = img.shape
    return (IMG_BF16_NCHW if img.dtype == torch.bfloat16 else IMG_F32_NCHW), (B, C, H, W)


def patch_embed_fwd(img, perm, wk, bias, p, g, *, pos=None, out=None, col_off=0, rows_per_img=None, tok_off=0):
    """img: [B,C,H,W] fp32/bf16 contiguous, or uint8 [B,H,W,C] (then wk's K order is q,p1,p2,c and the value
    normalisation is folded into wk / bias by the caller); perm int32 [(H/p)*(W/p)]; wk bf16 [D, Kpad] (K order q,c,p1,p2).
    Writes out[b, tok_off + t, col_off : col_off + D]; returns out ([B, rows_per_img, ld] bf16)."""
    lib = _lib.load()
    _require_cuda(img, perm, wk, bias, pos, out)
    assert img.dim() == 4 and img.is_contiguous() and img.dtype in (torch.float32, torch.bfloat16, torch.uint8)
    assert perm.dtype == torch.int32 and wk.dtype == torch.bfloat16 and wk.is_contiguous()
    fmt, (B, C, H, W) = _img_format(img)
    D, Kpad = wk.shape
    assert Kpad == lib.sfc_patch_embed_kpad(C, p, g)
    ntok = perm.numel() // g
    if rows_per_img is None:
        rows_per_img = ntok + tok_off
    if out is None:
        out = torch.empty((B, rows_per_img, D), dtype=torch.bfloat16, device=img.device)
    assert out.dtype == torch.bfloat16 and out.stride(-1) == 1 and out.shape[0] == B and out.shape[1] == rows_per_img
    assert out.stride(0) == rows_per_img * out.stride(1)
    # Pixel-level tokenizers (p = 1, _1D/*_embedding1D.py: a token is g curve-consecutive pixels, K = g * C scalars that
    # are NOT contiguous in the image): the fused kernel re-gathers a tile for each of its D / 256 column tiles through
    # per-element loads behind a permutation lookup (1.08 ms at 224 px, 256 px / token, B = 256). Gathering the
    # curve-ordered bf16 rows ONCE (sfc_patch_gather, consecutive lanes = consecutive curve pixels) and running one K3
    # GEMM on them takes 0.19 + 0.05 ms and gives bit-identical tokens (tools/pixel_tok_probe.py).
    if (p == 1 and g * C >= 256 and fmt != IMG_U8_NHWC and pos is None and tok_off == 0 and rows_per_img == ntok
            and PE_PROFILE is None):
        A = patch_gather(img, perm, p, g)
        out2 = out.as_strided((B * ntok, D), (out.stride(1), 1), out.storage_offset() + col_off)
        gemm(A, wk, bias=bias, out=out2)
        return out
    out_ptr = ctypes.c_void_p(out.data_ptr() + 2 * col_off)
    prof = PE_PROFILE
    with torch.cuda.device(img.device):
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        rc = lib.sfc_patch_embed_fwd(_ptr(img), fmt, B, C, H, W, p, g, _ptr(perm),
                                     perm.numel(), _ptr(wk), _ptr(bias), _ptr(pos), pos.stride(0) if pos is not None else 0, out_ptr,
                                     out.stride(1), D, rows_per_img, tok_off, _stream())
        if prof is not None:
            e1.record()
            prof.append((e0, e1, img.numel() * img.element_size() + B * ntok * D * 2, 2.0 * B * ntok * (g * p * p * C) * D))
    _lib.check(rc, "sfc_patch_embed_fwd")
    _count(1)
    return out


def patch_gather(img, perm, p, g):
    """Curve-ordered im2col A[B*ntok, Kpad] bf16 (backward helper)."""
    lib = _lib.load()
    _require_cuda(img, perm)
    fmt, (B, C, H, W) = _img_format(img)
    Kpad = lib.sfc_patch_embed_kpad(C, p, g)
    ntok = perm.numel() // g
    A = torch.empty((B * ntok, Kpad), dtype=torch.bfloat16, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(lib.sfc_patch_gather(_ptr(img), fmt, B, C, H, W, p, g, _ptr(perm),
                                        perm.numel(), _ptr(A), _stream()), "sfc_patch_gather")
    _count(1)
    return A


# ------------------------------------------------------------------ K4: attention
def attn_fwd(qkv, B, H, N, *, scale=None, drop_p=0.0, drop_seed=0):
    """qkv bf16 [B*N, 3D] contiguous -> (out bf16 [B*N, D], lse fp32 [B, H, N])."""
    lib = _lib.load()
    _require_cuda(qkv)
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.dim() == 2
    D = qkv.shape[1] // 3
    assert qkv.shape[0] == B * N
    if scale is None:
        scale = (D // H) ** -0.5
    out = torch.empty((B * N, D), dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty((B, H, N), dtype=torch.float32, device=qkv.device)
    prof = ATTN_PROFILE
    with torch.cuda.device(qkv.device):
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        rc = lib.sfc_attn_fwd(_ptr(qkv), _ptr(out), _ptr(lse), B, H, N, D, float(scale), float(drop_p),
                              int(drop_seed) & 0xFFFFFFFFFFFFFFFF, _stream())
        if prof is not None:
            e1.record()
            prof.append((e0, e1, 4.0 * B * N * N * D))
    _lib.check(rc, "sfc_attn_fwd")
    _count(1)
    return out, lse


def attn_bwd(qkv, out, dout, lse, B, H, N, *, scale=None, drop_p=0.0, drop_seed=0):
    lib = _lib.load()
    _require_cuda(qkv, out, dout, lse)
    assert dout.dtype == torch.bfloat16 and dout.is_contiguous() and out.is_contiguous() and qkv.is_contiguous()
    D = qkv.shape[1] // 3
    if scale is None:
        scale = (D // H) ** -0.5
    dqkv = torch.empty_like(qkv)
    nbytes = lib.sfc_attn_bwd_scratch_bytes(B, N, D)
    scratch = _workspace(nbytes, qkv.device)
    with torch.cuda.device(qkv.device):
        _lib.check(lib.sfc_attn_bwd(_ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), _ptr(dqkv), _ptr(scratch), nbytes, B, H, N,
                                    D, float(scale), float(drop_p), int(drop_seed) & 0xFFFFFFFFFFFFFFFF, _stream()),
                   "sfc_attn_bwd")
    _count(2)
    return dqkv


# ------------------------------------------------------------------ K6: optimizer
_sumsq_scratch = {}


def grad_sumsq(g, accum):
    """accum[0] += sum(g^2), reduced in a fixed order (deterministic across ranks)."""
    lib = _lib.load()
    _require_cuda(g, accum)
    assert g.is_contiguous() and accum.dtype == torch.float32
    key = (g.device.index, torch.cuda.current_stream().cuda_stream)
    scratch = _sumsq_scratch.get(key)
    if scratch is None:
        scratch = torch.zeros(lib.sfc_grad_sumsq_scratch_bytes(), dtype=torch.uint8, device=g.device)   # zeroed ONCE
        _sumsq_scratch[key] = scratch
    with torch.cuda.device(g.device):
        _lib.check(lib.sfc_grad_sumsq(_ptr(g), 1 if g.dtype == torch.float32 else 0, g.numel(), _ptr(accum), _ptr(scratch),
                                      scratch.numel(), _stream()), "sfc_grad_sumsq")
    _count(1)


def adamw_step(p, g, m, v, *, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, max_norm=0.0, stats=None,
               hyper=None):
    """hyper: optional fp32 device tensor {lr, 1 - beta1^t, sqrt(1 - beta2^t)} overriding lr / step at run time."""
    lib = _lib.load()
    _require_cuda(p, g, m, v, stats, hyper)
    assert p.is_contiguous() and g.is_contiguous() and m.is_contiguous() and v.is_contiguous()
    assert p.dtype == g.dtype and m.dtype == v.dtype
    assert hyper is None or (hyper.dtype == torch.float32 and hyper.numel() >= 3)
    with torch.cuda.device(p.device):
        _lib.check(lib.sfc_adamw_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), 1 if p.dtype == torch.float32 else 0,
                                      1 if m.dtype == torch.float32 else 0, float(lr), float(beta1), float(beta2), float(eps),
                                      float(weight_decay), int(step), float(grad_scale), float(max_norm), _ptr(stats),
                                      _ptr(hyper), _stream()), "sfc_adamw_step")
    _count(1)


def store_f32x4(dst, a, b=0.0, c=0.0, d=0.0):
    """dst[0..3] = (a, b, c, d) in stream order (values travel as kernel arguments)."""
    lib = _lib.load()
    _require_cuda(dst)
    assert dst.dtype == torch.float32 and dst.numel() >= 4 and dst.is_contiguous()
    with torch.cuda.device(dst.device):
        _lib.check(lib.sfc_store_f32x4(_ptr(dst), float(a), float(b), float(c), float(d), _stream()), "sfc_store_f32x4")
    _count(1)


def act_bwd(dy, aux, mode, alpha=1.0, drop_p=0.0, drop_seed=0):
    """alpha * dy * f'(aux) * dropout_mask (bf16, contiguous, numel % 8 == 0)."""
    lib = _lib.load()
    _require_cuda(dy, aux)
    assert dy.is_contiguous() and dy.dtype == torch.bfloat16
    assert aux is None or (aux.is_contiguous() and aux.dtype == torch.bfloat16)
    out = torch.empty_like(dy)
    with torch.cuda.device(dy.device):
        _lib.check(lib.sfc_act_bwd(_ptr(dy), _ptr(aux), _ptr(out), dy.numel(), int(mode), float(alpha), float(drop_p),
                                   int(drop_seed) & 0xFFFFFFFFFFFFFFFF, _stream()), "sfc_act_bwd")
    _count(1)
    return out


# ------------------------------------------------------------------ soft-target cross entropy (main.py:45-51)
_softce_scratch = {}


def softce_fwd(logits, targets):
    """logits bf16 / fp32 [B, C], targets fp32 [B, C] -> (loss fp32 [1] = batch mean, row_lse [B], row_tsum [B])."""
    lib = _lib.load()
    _require_cuda(logits, targets)
    assert logits.dim() == 2 and targets.shape == logits.shape and logits.stride(1) == 1 and targets.stride(1) == 1
    assert logits.dtype in (torch.bfloat16, torch.float32) and targets.dtype == torch.float32
    B, C = logits.shape
    key = (logits.device.index, torch.cuda.current_stream().cuda_stream)
    scratch = _softce_scratch.get(key)
    if scratch is None:
        scratch = torch.zeros(lib.sfc_softce_scratch_bytes(), dtype=torch.uint8, device=logits.device)   # zeroed ONCE
        _softce_scratch[key] = scratch
    loss = torch.empty(1, dtype=torch.float32, device=logits.device)
    lse = torch.empty(B, dtype=torch.float32, device=logits.device)
    tsum = torch.empty(B, dtype=torch.float32, device=logits.device)
    with torch.cuda.device(logits.device):
        _lib.check(lib.sfc_softce_fwd(_ptr(logits), 1 if logits.dtype == torch.float32 else 0, logits.stride(0), _ptr(targets),
                                      targets.stride(0), B, C, _ptr(loss), _ptr(lse), _ptr(tsum), _ptr(scratch), scratch.numel(),
                                      _stream()), "sfc_softce_fwd")
    _count(1)
    return loss, lse, tsum


def softce_bwd(logits, targets, lse, tsum, dloss):
    lib = _lib.load()
    _require_cuda(logits, targets, lse, tsum, dloss)
    assert dloss.dtype == torch.float32 and dloss.numel() == 1
    B, C = logits.shape
    dx = torch.empty_like(logits)
    with torch.cuda.device(logits.device):
        _lib.check(lib.sfc_softce_bwd(_ptr(logits), 1 if logits.dtype == torch.float32 else 0, logits.stride(0), _ptr(targets),
                                      targets.stride(0), _ptr(lse), _ptr(tsum), _ptr(dloss), B, C, _ptr(dx), dx.stride(0), _stream()),
                   "sfc_softce_bwd")
    _count(1)
    return dx
