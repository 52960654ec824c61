"""Host-built 'onion' spiral permutation (reference _1D/onion_embedding1D.py:36-54, multi_onion.py:72-90): start at the
bottom-left cell, walk right / up / left / down, turning when the next cell is outside or already visited."""
import numpy as np


def spiral_cells(h: int, w: int) -> np.ndarray:
    top, bottom, left, right = 0, h - 1, 0, w - 1
    cells = []
    while top <= bottom and left <= right:
        cells += [(bottom, c) for c in range(left, right + 1)]                 # right along the bottom row
        cells += [(r, right) for r in range(bottom - 1, top - 1, -1)]          # up the right column
        if top < bottom:
            cells += [(top, c) for c in range(right - 1, left - 1, -1)]        # left along the top row
        if left < right:
            cells += [(r, left) for r in range(top + 1, bottom)]               # down the left column
        top, bottom, left, right = top + 1, bottom - 1, left + 1, right - 1
    return np.asarray(cells, dtype=np.int64).reshape(-1, 2)
