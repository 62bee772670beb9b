"""Drop-in for ``smash.solver._mw_forcing_statistic`` (routine/mw_forcing_statistic.f90:18-225): catchment means of the
forcing and the precipitation indices of Zoccatelli et al. (2011) / Emmanuel et al. (2015).  Host-side NumPy: these run
once per model build, not per solver call."""
from __future__ import annotations

import numpy as np

from ._mw_mask import mask_upstream_cells
from ._mw_sparse_storage import sparse_vector_to_matrix_r


def _gauge_masks(mesh):
    masks = np.zeros((mesh.nrow, mesh.ncol, mesh.ng), dtype=bool, order="F")
    for g in range(mesh.ng):
        # gauge_pos is 0-based on the Python side (_f90wrap_decorator.py:72-106)
        mask_upstream_cells(int(mesh.gauge_pos[g, 0]) + 1, int(mesh.gauge_pos[g, 1]) + 1, mesh, masks[:, :, g])
    return masks


def _step_matrix(setup, mesh, input_data, name, t):
    if setup.sparse_storage:
        m = np.zeros((mesh.nrow, mesh.ncol), dtype=np.float32, order="F")
        sparse_vector_to_matrix_r(mesh, getattr(input_data, "sparse_" + name)[:, t], m)
        return m
    return getattr(input_data, name)[:, :, t]


def compute_mean_forcing(setup, mesh, input_data):
    """mw_forcing_statistic.f90:18-75: mean_prcp / mean_pet (ng, ntime_step) over the cells upstream of each gauge with a
    non-negative value.  A gauge without any valid cell gives 0/0 = NaN, as in the Fortran."""
    masks = _gauge_masks(mesh)
    for t in range(setup._ntime_step):
        mp, me = _step_matrix(setup, mesh, input_data, "prcp", t), _step_matrix(setup, mesh, input_data, "pet", t)
        for g in range(mesh.ng):
            for mat, dst in ((mp, input_data.mean_prcp), (me, input_data.mean_pet)):
                ok = (mat >= 0) & masks[:, :, g]
                n = int(ok.sum())
                with np.errstate(invalid="ignore", divide="ignore"):
                    dst[g, t] = np.float32(mat[ok].sum(dtype=np.float32)) / np.float32(n)


def compute_mean_forcing_device(setup, mesh, input_data):
    """The same means on the GPU (``smash_b200_compute_mean_forcing``, csrc/pre_kernels.cu): catchment masks by walking
    downstream from every cell, one block per (time step, gauge), float64 sums rounded once.  Raises without a CUDA device."""
    import ctypes as C

    from .. import _lib as L
    pk = L.Packed()
    s, m, i = L.pack_setup(setup, mesh, pk), L.pack_mesh(mesh, setup, pk), L.pack_input(input_data, setup, mesh, pk)
    mp = np.zeros((mesh.ng, setup._ntime_step), dtype=np.float32, order="F")
    me = np.zeros((mesh.ng, setup._ntime_step), dtype=np.float32, order="F")
    L.check(L.lib().smash_b200_compute_mean_forcing(C.byref(s), C.byref(m), C.byref(i), L._fp(mp), L._fp(me)))
    input_data.mean_prcp[...] = mp
    input_data.mean_pet[...] = me


def gauge_masks_device(mesh, setup=None):
    """(nrow, ncol, ng) boolean masks of the cells upstream of every gauge, computed on the GPU."""
    import ctypes as C

    from .. import _lib as L
    pk = L.Packed()
    m = L.pack_mesh(mesh, setup, pk)
    out = np.zeros((mesh.nrow, mesh.ncol, mesh.ng), dtype=np.uint8, order="F")
    L.check(L.lib().smash_b200_gauge_masks(C.byref(m), out.ctypes.data_as(C.POINTER(C.c_uint8))))
    return out.astype(bool)


def _quantile(x, q):
    """mwd_cost.f90 quantile (linear interpolation between order statistics, like numpy's default)."""
    return np.quantile(np.asarray(x, np.float64), q).astype(np.float32)


def compute_prcp_indices(setup, mesh, input_data, prcp_indices):
    """mw_forcing_statistic.f90:77-223: prcp_indices(4, ng, ntime_step) = (std, d1, d2, vg) per gauge and time step;
    entries of steps without precipitation are left untouched."""
    masks = _gauge_masks(mesh)
    flwdst = np.asarray(mesh.flwdst, np.float32)
    qtl = (np.arange(0, 101, 10) / np.float32(100.0)).astype(np.float32)
    nq = len(qtl)
    dst_g, qt_g, wf = [], [], np.zeros((nq, mesh.ng), np.float32)
    for g in range(mesh.ng):
        d = flwdst - flwdst[int(mesh.gauge_pos[g, 0]), int(mesh.gauge_pos[g, 1])]
        dst_g.append(d)
        flat = d[masks[:, :, g]]
        qs = _quantile(flat, qtl)
        qt_g.append(qs)
        wf[0, g] = 1.0
        for j in range(1, nq):
            wf[j, g] = wf[j - 1, g] + np.float32(((flat > qs[j - 1]) & (flat <= qs[j])).sum())
    for t in range(setup._ntime_step):
        mat = np.asarray(_step_matrix(setup, mesh, input_data, "prcp", t), np.float32)
        for g in range(mesh.ng):
            mask = (mat >= 0) & masks[:, :, g]
            n = int(mask.sum())
            if n == 0:
                continue
            minv_n = np.float32(1.0) / np.float32(n)
            p, d = mat[mask], dst_g[g][mask]
            sum_p = p.sum(dtype=np.float32)
            if not sum_p > 0:
                continue
            sum_p2, sum_d, sum_d2 = (p * p).sum(dtype=np.float32), d.sum(dtype=np.float32), (d * d).sum(dtype=np.float32)
            sum_pd, sum_pd2 = (p * d).sum(dtype=np.float32), (p * d * d).sum(dtype=np.float32)
            mean_p = minv_n * sum_p
            p0, p1, p2, g1, g2 = mean_p, minv_n * sum_pd, minv_n * sum_pd2, minv_n * sum_d, minv_n * sum_d2
            pwf = np.zeros(nq, np.float32)
            # the reference indexes the gauge cell with (gauge_pos(j,1), gauge_pos(j,1)) -- row twice (:173); kept
            r = int(mesh.gauge_pos[g, 0])
            pwf[0] = max(np.float32(0.0), mat[r, min(r, mesh.ncol - 1)] / sum_p)
            for k in range(1, nq):
                sel = (dst_g[g] > qt_g[g][k - 1]) & (dst_g[g] <= qt_g[g][k])
                c = int(sel.sum())
                mean_subp = np.float32(0.0) if c == 0 else mat[sel].sum(dtype=np.float32) / np.float32(c)
                pwf[k] = pwf[k - 1] + mean_subp / mean_p * wf[k, g]
            with np.errstate(invalid="ignore", divide="ignore"):
                d1 = p1 / (p0 * g1)
                d2 = (np.float32(1.0) / (g2 - g1 * g1)) * ((p2 / p0) - (p1 / p0) * (p1 / p0))
                std = np.sqrt((minv_n * sum_p2) - (mean_p * mean_p))
                vg = np.abs(pwf / pwf[-1] - wf[:, g] / wf[-1, g]).max()
            prcp_indices[:, g, t] = (std, d1, d2, vg)
