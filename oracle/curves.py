"""ORACLE (test infrastructure): ctypes front-end of oracle/sfc_oracle.c.

Mirrors the reference's curve API (src/curves/space_filling_curves.py):
``embed_and_prune(curve, w, h)`` follows embed_and_prune_sfc (:471-491),
``curve_points`` follows the ``sfc(order, size)`` generators (:74-251),
``hilbert2d_flat`` follows _2D/hilbert_embedding.py:30-78.
"""
import ctypes
import numpy as np

from . import build as _build

CURVE_IDS = {"hilbert": 0, "z": 1, "morton": 1, "peano": 2, "moore": 3, "raster": 4,
             "hilbert_curve": 0, "z_curve": 1, "peano_curve": 2, "moore_curve": 3, "raster_curve": 4}

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(_build.build())
        lib.sfc_oracle_grid_size.restype = ctypes.c_int64
        lib.sfc_oracle_grid_size.argtypes = [ctypes.c_int, ctypes.c_int]
        lib.sfc_oracle_curve_points.restype = ctypes.c_int64
        lib.sfc_oracle_curve_points.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_int64]
        lib.sfc_oracle_embed_and_prune.restype = ctypes.c_int64
        lib.sfc_oracle_embed_and_prune.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
        lib.sfc_oracle_hilbert2d_flat.restype = ctypes.c_int64
        lib.sfc_oracle_hilbert2d_flat.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
        _lib = lib
    return _lib


def curve_id(curve) -> int:
    name = curve if isinstance(curve, str) else getattr(curve, "__name__", str(curve))
    if name not in CURVE_IDS:
        raise ValueError(f"Unknown SFC: {name}")
    return CURVE_IDS[name]


def grid_size(order: int, curve) -> int:
    return int(_load().sfc_oracle_grid_size(curve_id(curve), order))


def curve_points(curve, order: int, size: float = 1.0) -> np.ndarray:
    cid = curve_id(curve)
    P = grid_size(order, curve)
    out = np.empty((P * P, 2), dtype=np.float64)
    n = _load().sfc_oracle_curve_points(cid, order, float(size), out.ctypes.data, P * P)
    assert n == P * P, n
    return out


def embed_and_prune(curve, width: int, height: int) -> np.ndarray:
    """-> int64 [width*height, 2] of (i, j) = (row, col) in curve order."""
    out = np.empty((width * height, 2), dtype=np.int64)
    n = _load().sfc_oracle_embed_and_prune(curve_id(curve), width, height, out.ctypes.data, width * height)
    if n < 0:
        raise ValueError(f"oracle embed_and_prune failed ({n}) for {curve}")
    return out[:n]


def flat_perm(curve, width: int, height: int) -> np.ndarray:
    """Flat token permutation r*height+c as tokenizers build it (multi_hilbert.py:70-71)."""
    if curve_id(curve) == 4:
        return np.arange(width * height, dtype=np.int64)  # raster tokenizers do not reorder
    ij = embed_and_prune(curve, width, height)
    return ij[:, 0] * height + ij[:, 1]


def hilbert2d_flat(grid: int) -> np.ndarray:
    order = int(np.log2(grid))
    n = (2 ** order) ** 2
    out = np.empty(n, dtype=np.int64)
    m = _load().sfc_oracle_hilbert2d_flat(grid, out.ctypes.data, n)
    assert m == n
    return out
