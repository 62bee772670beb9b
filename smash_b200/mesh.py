"""Flow accumulation of a D8 flow-direction raster, the integer contract of the meshing step the solver's cell ordering
rests on (smash/mesh/mw_meshing.f90:111-233: ``fill_nipd``, ``downstream_cell_flwacc``, ``flow_accumulation``).

Host-side restatement without recursion: the reference starts at every cell that no neighbour drains into and walks
downstream, adding the accumulated count to the receiving cell and continuing once all of its in-pointing neighbours
have arrived; two cells that drain into each other (|fd - fd'| = 4, a pit pair) neither exchange nor release each other.
Here the same sums are formed front by front (Kahn's order), which is integer-exact whatever the order."""
from __future__ import annotations

import numpy as np

# mw_meshing.f90:163-164 -- the cell a direction code fd = 1..8 points to
DROW = np.array([-1, -1, 0, 1, 1, 1, 0, -1], dtype=np.int64)
DCOL = np.array([0, 1, 1, 1, 0, -1, -1, -1], dtype=np.int64)


def flow_accumulation_device(flwdir, mask=None):
    """The same integer contract on the GPU (``smash_b200_flow_accumulation``, csrc/pre_kernels.cu): every cell nobody
    drains into walks downstream, the last neighbour to arrive carries the sum on.  Raises without a CUDA device."""
    import ctypes as C

    from . import _lib as L
    fd = np.asfortranarray(flwdir, dtype=np.int32)
    nrow, ncol = fd.shape
    mk = None if mask is None else np.asfortranarray(np.asarray(mask) != 0, dtype=np.int32)
    out = np.zeros((nrow, ncol), dtype=np.int32, order="F")
    L.check(L.lib().smash_b200_flow_accumulation(nrow, ncol, L._ip(fd), L._ip(mk) if mk is not None else None, L._ip(out)))
    return out


def flow_accumulation(flwdir, mask=None):
    """``flwacc`` (int32, 1 on cells without inflow) of ``flwdir`` (codes 1..8, anything else = no direction).  ``mask``
    restricts the computation to a catchment window (cells outside neither give nor receive)."""
    fd = np.asarray(flwdir).astype(np.int64)
    nrow, ncol = fd.shape
    ok = (fd >= 1) & (fd <= 8)
    if mask is not None:
        ok &= np.asarray(mask).astype(bool)
    rows, cols = np.nonzero(ok)
    code = fd[rows, cols] - 1
    r2, c2 = rows + DROW[code], cols + DCOL[code]
    inside = (r2 >= 0) & (r2 < nrow) & (c2 >= 0) & (c2 < ncol)
    src = rows[inside] * ncol + cols[inside]
    dst = r2[inside] * ncol + c2[inside]
    fds, fdd = fd.ravel()[src], fd.ravel()[dst]
    okd = ok.ravel()[dst] if mask is not None else np.ones(dst.size, dtype=bool)
    # an edge carries flow unless the two cells point at each other (mw_meshing.f90:182); the in-pointing count of
    # fill_nipd (:111-152) includes the partner, so a pit cell is never released -- it only receives
    carries = okd & ~(((fdd >= 1) & (fdd <= 8)) & (np.abs(fds - fdd) == 4))
    src, dst = src[carries], dst[carries]
    n = nrow * ncol
    down = np.full(n, -1, dtype=np.int64)
    down[src] = dst
    indeg = np.bincount(dst, minlength=n)
    # pit partners count as in-pointing neighbours of each other: they hold their cell back for ever
    pr, pc = np.nonzero(ok)
    pcode = fd[pr, pc] - 1
    qr, qc = pr + DROW[pcode], pc + DCOL[pcode]
    pin = (qr >= 0) & (qr < nrow) & (qc >= 0) & (qc < ncol)
    a, b = (pr[pin] * ncol + pc[pin]), (qr[pin] * ncol + qc[pin])
    fb = fd.ravel()[b]
    mutual = ((fb >= 1) & (fb <= 8)) & (np.abs(fd.ravel()[a] - fb) == 4) & ok.ravel()[b]
    held = np.zeros(n, dtype=bool)
    held[b[mutual]] = True
    acc = np.ones(n, dtype=np.int64)
    active = ok.ravel() if mask is not None else np.ones(n, dtype=bool)
    front = np.flatnonzero((indeg == 0) & active & ~held & (down >= 0))
    while front.size:
        d = down[front]
        np.add.at(acc, d, acc[front])
        np.subtract.at(indeg, d, 1)
        cand = np.unique(d)
        front = cand[(indeg[cand] == 0) & ~held[cand] & (down[cand] >= 0)]
    return acc.reshape(nrow, ncol).astype(np.int32)
