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
    Returns (curve, blocked_curve). Host-side, init-time utility (not used by any tokenizer)."""
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


def find_hamiltonian_path(width, height, adjacency_order=None, diag=False):
    """Depth-first Hamiltonian path search on the grid with flood-fill and dead-end pruning (:273-443)."""
    sys.setrecursionlimit(10_000_000)
    total = width * height
    visited = [[False] * height for _ in range(width)]
    path = []
    dirs = [(1, 0), (-1, 0), (0, 1), (0, -1)] + ([(1, 1), (1, -1), (-1, 1), (-1, -1)] if diag else [])
    nbrs_of = {(x, y): [(x + dx, y + dy) for dx, dy in dirs if 0 <= x + dx < width and 0 <= y + dy < height]
               for x in range(width) for y in range(height)}

    def ordered(x, y):
        def key(v):
            is_diag = 1 if abs(v[0] - x) == 1 and abs(v[1] - y) == 1 else 0
            return (is_diag, adjacency_order.get(v, total) if adjacency_order else 0)
        return sorted(nbrs_of[(x, y)], key=key)

    def reachable(sx, sy, remaining):
        stack, seen, cnt = [(sx, sy)], {(sx, sy)}, 0
        while stack:
            x, y = stack.pop()
            cnt += 1
            if cnt >= remaining:
                return True
            for n in nbrs_of[(x, y)]:
                if not visited[n[0]][n[1]] and n not in seen:
                    seen.add(n)
                    stack.append(n)
        return cnt >= remaining

    def dfs(x, y):
        if len(path) == total:
            return True
        cand = [n for n in ordered(x, y) if not visited[n[0]][n[1]]]
        forced, kept = [], []
        for n in cand:
            exits = sum(1 for u in nbrs_of[n] if not visited[u[0]][u[1]] and u != (x, y))
            if exits == 0 and len(path) + 1 < total:
                continue
            if exits == 1:
                forced.append(n)
            kept.append(n)
        for nx, ny in (forced or kept):
            visited[nx][ny] = True
            path.append((nx, ny))
            rem = total - len(path)
            if rem == 0 or reachable(nx, ny, rem):
                if dfs(nx, ny):
                    return True
            visited[nx][ny] = False
            path.pop()
        return False

    starts = [min(adjacency_order, key=adjacency_order.get)] if adjacency_order else \
        [(0, 0), (width - 1, 0), (0, height - 1), (width - 1, height - 1)]
    for sx, sy in starts:
        visited[sx][sy] = True
        path[:] = [(sx, sy)]
        if dfs(sx, sy):
            return path
        visited[sx][sy] = False
    return None


def refine_curve_to_hamiltonian(curve, width, height):
    """Use a (pruned) curve as visiting priority to find a true Hamiltonian path (:446-455)."""
    return find_hamiltonian_path(width, height, adjacency_order={pt: i for i, pt in enumerate(curve)})
