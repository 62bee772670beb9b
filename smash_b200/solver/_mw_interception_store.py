"""Drop-in for ``smash.solver._mw_interception_store`` (routine/mw_interception_store.f90:19-160): the interception
capacity ``ci`` of sub-daily gr-b / gr-c runs is set, cell by cell, to the value in 0.1 .. 4.9 mm whose cumulated sub-daily
interception evaporation is closest to the cumulated daily one (gr_interception, operator/md_gr_operator.f90:20-34).

The search runs on the GPU (``smash_b200_adjust_interception_store``, csrc/pre_kernels.cu: one thread per cell and group of
seven capacities, one pass over the cell's forcing series); there is no host version."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib as L

# device time of the last search, ms (diagnostics / bench)
last_kernel_ms = 0.0


def adjust_interception_store(setup, mesh, input_data, parameters, nday, day_index):
    global last_kernel_ms
    pk = L.Packed()
    s, m = L.pack_setup(setup, mesh, pk), L.pack_mesh(mesh, setup, pk)
    i = L.pack_input(input_data, setup, mesh, pk)
    di = np.ascontiguousarray(day_index, dtype=np.int32)
    if di.shape != (int(setup._ntime_step),):
        raise ValueError(f"day_index has shape {di.shape}, expected ({int(setup._ntime_step)},)")
    ci = np.asfortranarray(parameters.ci, dtype=np.float32)
    out = ci if ci is parameters.ci else ci.copy(order="F")
    ms = C.c_float(0.0)
    L.check(L.lib().smash_b200_adjust_interception_store(C.byref(s), C.byref(m), C.byref(i), int(nday), L._ip(di), L._fp(out), C.byref(ms)))
    last_kernel_ms = float(ms.value)
    if out is not parameters.ci:
        parameters.ci[...] = out
