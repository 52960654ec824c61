"""Space-filling curves — API mirror of the reference's src/curves/space_filling_curves.py.

The hot functions (``hilbert_curve``, ``z_curve``, ``peano_curve``, ``moore_curve``, ``embed_and_prune_sfc``) run the
integer sm_100a kernel K1 (csrc/curves.cu, csrc/curve_index.h) through the C ABI ``sfc_curve_perm`` and return the
reference's Python types: ``sfc(order, size) -> list[(x, y)]`` cell centres, ``embed_and_prune_sfc -> list[(i, j)]``
python ints in curve order (reference :74-251, :458-491). There is no CPU fallback for those curves: without a CUDA
device they raise. Callables the kernel does not know (a user's own ``curve_fn``) follow the reference's generic
host recipe (generate, floor, filter). ``onion_curve``, ``raster_curve``, ``block_stitch_sfc``,
``find_hamiltonian_path`` and ``refine_curve_to_hamiltonian`` are init-time host utilities outside the kernel path
(SURVEY.md §2 rows 3-5); they are provided so that every name of the reference module resolves.
"""
import math
import sys
from typing import Callable, List, Tuple

import numpy as np
import torch

_KERNEL_CURVES = {"hilbert_curve": 0, "z_curve": 1, "peano_curve": 2, "moore_curve": 3}


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("src.curves: the curve kernels need a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def curve_permutation(sfc, width: int, height: int, device=None):
    """(perm, inv) int32 device tensors: perm[t] = i*height + j of the t-th cell along the curve pruned to the
    width x height grid. This is what the tokenizers keep resident in HBM."""
    from sfcvit import ops
    name = sfc if isinstance(sfc, str) else sfc.__name__
    return ops.curve_perm(name, int(width), int(height), device or _device())


def _kernel_points(name: str, order: int, size: float):
    base = 3 if name == "peano_curve" else 2
    P = base ** int(order)
    perm, _ = curve_permutation(name, P, P)
    flat = perm.cpu().numpy().astype(np.int64)
    cell = size / P
    xs = (flat // P + 0.5) * cell
    ys = (flat % P + 0.5) * cell
    return list(zip(xs.tolist(), ys.tolist()))


def hilbert_curve(order, size=1.0):
    """Hilbert curve cell centres, reference orientation (:168-202: vector recursion followed by a transpose)."""
    return _kernel_points("hilbert_curve", order, size)


def z_curve(order, size=1.0):
    """Z-order / Morton curve (:134-165): quadrant order TR, TL, BR, BL."""
    return _kernel_points("z_curve", order, size)


def peano_curve(order, size=1.0):
    """Peano curve on a 3^order grid (:74-131)."""
    return _kernel_points("peano_curve", order, size)


def moore_curve(order, size=1.0):
    """Moore curve: four Hilbert sub-curves forming a loop (:205-251)."""
    return _kernel_points("moore_curve", order, size)


def raster_curve(order, size=1.0):
    """Row-major cell centres on a 2^order grid (:254-271)."""
    n = 2 ** int(order)
    cell = size / n
    return [((x + 0.5) * cell, (y + 0.5) * cell) for y in range(n) for x in range(n)]


def onion_curve(order, size=1.0):
    """Spiral ('onion') curve on a (2*order) x (2*order) grid (:9-71): rings from the outside in."""
    j = int(order) * 2
    coords = []
    off = 0
    while j > 2:
        coords += [(off + x, off) for x in range(j)]
        coords += [(off + j - 1, off + y) for y in range(1, j)]
        coords += [(off + x, off + j - 1) for x in range(j - 2, -1, -1)]
        coords += [(off, off + y) for y in range(j - 2, 0, -1)]
        off += 1
        j -= 2
    if j == 2:
        coords += [(off, off), (off + 1, off), (off + 1, off + 1), (off, off + 1)]
    cell = size / (int(order) * 2) if order else size
    return [(x * cell + cell / 2, y * cell + cell / 2) for x, y in coords]


def grid_size(order, sfc):
    """Side length of sfc(order) — dispatches on ``sfc.__name__`` like the reference (:458-468)."""
    name = sfc.__name__
    if name in ("hilbert_curve", "z_curve", "moore_curve"):
        return 2 ** order
    elif name == "peano_curve":
        return 3 ** order
    elif name == "onion_curve":
        return order + (order % 2)
    else:
        raise ValueError(f"Unknown SFC: {name}")


def embed_and_prune_sfc(sfc, width, height):
    """Embed the curve in the smallest padded square covering width x height and keep in-domain cells in curve
    order (:471-491). Returns list[(i, j)] of python ints, i = row < width, j = col < height."""
    name = getattr(sfc, "__name__", None)
    if name in _KERNEL_CURVES and getattr(sys.modules[__name__], name, None) is sfc:
        perm, _ = curve_permutation(name, width, height)
        flat = perm.cpu().numpy().astype(np.int64)
        return list(zip((flat // height).tolist(), (flat % height).tolist()))
    # unknown callable: the reference's generic recipe on the host
    order = 0
    while grid_size(order, sfc) < max(width, height):
        order += 1
    P = grid_size(order, sfc)
    curve = []
    for x, y in sfc(order, size=P):
        i, j = int(np.floor(x)), int(np.floor(y))
        if 0 <= i < width and 0 <= j < height:
            curve.append((i, j))
    return curve


def get_symmetries(B: int) -> List[Callable[[float, float], Tuple[float, float]]]:
    """The 8 dihedral symmetries of a B x B block as (x, y) -> (x', y') maps (:494-510)."""
    return [
        lambda x, y: (x, y), lambda x, y: (y, B - x), lambda x, y: (B - x, B - y), lambda x, y: (B - y, x),
        lambda x, y: (B - x, y), lambda x, y: (y, x), lambda x, y: (x, B - y), lambda x, y: (B - y, B - x),
    ]


def block_stitch_sfc(sfc, width: int, height: int):
    """Greedy power-of-base block decomposition with the best of 8 orientations per block (:513-591).
    Returns (curve, blocked_curve). Init-time host utility (not used by any tokenizer): the four library curves run
    in the C++ routine sfc_block_stitch (csrc/host_curves.cu) on the same integer per-index curve code as kernel K1."""
    name = getattr(sfc, "__name__", None)
    if name in _KERNEL_CURVES and getattr(sys.modules[__name__], name, None) is sfc:
        import ctypes
        from sfcvit import _lib, ops
        lib = _lib.load()
        n = width * height
        out = np.empty((n, 2), dtype=np.int32)
        blen = np.empty(n, dtype=np.int32)
        nb = ctypes.c_int(0)
        got = lib.sfc_block_stitch(ops.curve_id(sfc), width, height, out.ctypes.data, n, blen.ctypes.data, n, ctypes.byref(nb))
        if got != n:
            msg = lib.sfc_last_error()
            raise RuntimeError(f"sfc_block_stitch failed: {msg.decode() if msg else got}")
        curve = [(int(i), int(j)) for i, j in out]
        blocked, off = [], 0
        for k in blen[:nb.value]:
            blocked.append(curve[off:off + int(k)])
            off += int(k)
        return curve, blocked
    return _block_stitch_generic(sfc, width, height)


def _block_stitch_generic(sfc, width: int, height: int):
    """The same recipe for a user-supplied curve callable (host Python, as in the reference)."""
    base = 3 if sfc.__name__ == "peano_curve" else 2
    blocks = []

    def collect(x0, y0, w, h):
        if w <= 0 or h <= 0:
            return
        k = int(np.floor(np.log(min(w, h)) / np.log(base)))
        B = base ** k
        blocks.append((x0, y0, B, k))
        collect(x0 + B, y0, w - B, B)
        collect(x0, y0 + B, w, h - B)

    collect(0, 0, width, height)
    raws = [sfc(k, B) for (_x0, _y0, B, k) in blocks]
    entries = [(math.floor(x0 + raw[0][0]), math.floor(y0 + raw[0][1])) for (x0, y0, _B, _k), raw in zip(blocks, raws)]
    visited, curve, blocked = set(), [], []
    prev_exit = None
    for idx, ((x0, y0, B, k), raw) in enumerate(zip(blocks, raws)):
        nxt = entries[idx + 1] if idx + 1 < len(blocks) else None
        best, best_pts = math.inf, None
        for sym in get_symmetries(B):
            pts = [(x0 + math.floor(sym(x, y)[0]), y0 + math.floor(sym(x, y)[1])) for x, y in raw]
            new_pts = [q for q in pts if q not in visited]
            if not new_pts:
                continue
            score = 0
            if prev_exit is not None:
                score += abs(prev_exit[0] - new_pts[0][0]) + abs(prev_exit[1] - new_pts[0][1])
            if nxt is not None:
                score += abs(new_pts[-1][0] - nxt[0]) + abs(new_pts[-1][1] - nxt[1])
            if score < best:
                best, best_pts = score, new_pts
        for q in best_pts:
            visited.add(q)
            curve.append(q)
        blocked.append(best_pts)
        prev_exit = best_pts[-1]
    return curve, blocked


def find_hamiltonian_path(width, height, adjacency_order=None, diag=False, max_steps=200_000_000):
    """Depth-first Hamiltonian path search on the grid with forced-move and flood-fill pruning (:273-443), as the C++
    routine sfc_hamiltonian_path (csrc/host_curves.cu; same neighbour order and tie-breaking, so the same path).
    `max_steps` bounds the exponential worst case (the reference has no bound); None when no path was found."""
    from sfcvit import _lib
    lib = _lib.load()
    total = width * height
    pr = None
    if adjacency_order:
        pr = np.full(total, total, dtype=np.int32)
        for (x, y), v in adjacency_order.items():
            if 0 <= x < width and 0 <= y < height:
                pr[x * height + y] = v
    out = np.empty((total, 2), dtype=np.int32)
    got = lib.sfc_hamiltonian_path(width, height, pr.ctypes.data if pr is not None else None, 1 if diag else 0, int(max_steps),
                                   out.ctypes.data)
    if got != total:
        return None
    return [(int(i), int(j)) for i, j in out]


def refine_curve_to_hamiltonian(curve, width, height):
    """Use a (pruned) curve as visiting priority to find a true Hamiltonian path (:446-455)."""
    return find_hamiltonian_path(width, height, adjacency_order={pt: i for i, pt in enumerate(curve)})
