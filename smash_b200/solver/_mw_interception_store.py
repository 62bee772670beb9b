"""Drop-in for ``smash.solver._mw_interception_store`` (routine/mw_interception_store.f90:19-160): the interception
capacity ``ci`` of sub-daily runs is set, cell by cell, to the value in 0.1 .. 4.9 mm whose cumulated sub-daily
interception evaporation is closest to the cumulated daily one (gr_interception, operator/md_gr_operator.f90:20-34)."""
from __future__ import annotations

import numpy as np


def adjust_interception_store(setup, mesh, input_data, parameters, nday, day_index):
    T = setup._ntime_step
    act = (np.asarray(mesh.active_cell) == 1) & (np.asarray(mesh._local_active_cell) == 1)
    shape = (mesh.nrow, mesh.ncol)

    def step(name, t):
        if setup.sparse_storage:
            m = np.zeros(shape, np.float32)
            k = np.asarray(mesh._rowcol_to_ind_sparse)
            m[act] = getattr(input_data, "sparse_" + name)[k[act] - 1, t]
            return m
        return np.asarray(getattr(input_data, name)[:, :, t], np.float32)

    day_index = np.asarray(day_index)
    daily_p = np.zeros(shape + (int(nday),), np.float32)
    daily_e = np.zeros(shape + (int(nday),), np.float32)
    n = 0
    for t in range(T):
        if t > 0 and day_index[t] != day_index[t - 1]:
            n += 1
        daily_p[:, :, n] += step("prcp", t)
        daily_e[:, :, n] += step("pet", t)
    daily_cum = np.minimum(daily_p, daily_e).sum(axis=2, dtype=np.float32)
    cmax = (np.float32(0.1) + np.float32(0.1) * np.arange(int(np.ceil((5.0 - 0.1) / 0.1)), dtype=np.float32)).astype(np.float32)
    diff = np.zeros(shape + (len(cmax),), np.float32)
    for i, ci in enumerate(cmax):
        h = np.zeros(shape, np.float32)
        sub = np.zeros(shape, np.float32)
        for t in range(T):
            prcp, pet = step("prcp", t), step("pet", t)
            ei = np.minimum(pet, prcp + h * ci)                                  # md_gr_operator.f90:28
            pn = np.maximum(np.float32(0.0), prcp - ci * (np.float32(1.0) - h) - ei)
            h = np.where(act, h + (prcp - ei - pn) / ci, h).astype(np.float32)
            sub = np.where(act, sub + ei, sub).astype(np.float32)
        diff[:, :, i] = np.abs(sub - daily_cum)
    best = diff.argmin(axis=2)
    parameters.ci[act] = cmax[best[act]]
