"""Drop-in for ``smash.solver._mw_mask`` (routine/mw_mask.f90:11-55)."""
from __future__ import annotations

import numpy as np

_DCOL = (0, -1, -1, -1, 0, 1, 1, 1)      # mw_mask.f90:28-30
_DROW = (1, 1, 0, -1, -1, -1, 0, 1)


def mask_upstream_cells(row, col, mesh, mask):
    """Marks (row, col) -- 1-based like the Fortran routine -- and every cell that drains into it.  Iterative version of
    the recursive routine: same visiting rule (neighbour i flows in iff its flwdir == i)."""
    flwdir = np.asarray(mesh.flwdir)
    nrow, ncol = flwdir.shape
    stack = [(int(row) - 1, int(col) - 1)]
    while stack:
        r, c = stack.pop()
        mask[r, c] = True
        for i in range(8):
            rr, cc = r + _DROW[i], c + _DCOL[i]
            if 0 <= rr < nrow and 0 <= cc < ncol and flwdir[rr, cc] == i + 1 and not mask[rr, cc]:
                stack.append((rr, cc))
    return mask
