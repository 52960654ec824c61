"""Torch-tensor front-ends of the C ABI (device pointers + current stream). CUDA only."""
import ctypes

import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
AUX_NONE, AUX_RELU_MASK, AUX_GELU_GRAD = 0, 1, 2
CURVE_IDS = {"hilbert_curve": 0, "z_curve": 1, "peano_curve": 2, "moore_curve": 3, "raster_curve": 4,
             "hilbert": 0, "z": 1, "morton": 1, "peano": 2, "moore": 3, "raster": 4}


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
    return perm, inv


def gemm(a, b, *, a_mn=False, b_mn=False, bias=None, residual=None, aux=None, aux_mode=AUX_NONE, act=ACT_NONE,
         alpha=1.0, out=None, out_dtype=torch.bfloat16, want_pre=False, splits=1, accumulate=False, drop_p=0.0,
         drop_seed=0):
    """D[M,N] = epilogue(alpha * A.B^T).  a: [M,K] (or [K,M] if a_mn), b: [N,K] (or [K,N] if b_mn); bf16, 2-D,
    inner dimension contiguous. Returns out (and out_pre when want_pre)."""
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
    if splits is None or splits == 0:
        splits = lib.sfc_gemm_suggest_splits(M, N, K)
    ws, ws_bytes = None, 0
    if splits > 1:
        ws_bytes = lib.sfc_gemm_workspace_bytes(M, N, K, splits)
        ws = _workspace(ws_bytes, a.device)
    with torch.cuda.device(a.device):
        rc = lib.sfc_gemm_bf16(_ptr(a), int(a_mn), a.stride(0), _ptr(b), int(b_mn), b.stride(0), M, N, K,
                               ctypes.byref(ep), _ptr(ws), ws_bytes, int(splits), _stream())
    _lib.check(rc, "sfc_gemm_bf16")
    return (out, pre) if want_pre else out
