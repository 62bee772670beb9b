"""Drop-in for ``smash.solver._mw_multiple_run`` (routine/mw_multiple_run.f90:68-119)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib as L


def compute_multiple_run(setup, mesh, input_data, parameters, states, output, sample, ind_parameters_states, res_cost,
                         res_qsim):
    """``sample`` (nvar, ns) float32 F-order; ``ind_parameters_states`` 1-based into the 16+8 stacked fields (as the
    reference's Python caller builds it, multiple_run.py:184-200); ``res_cost`` (ns,) and ``res_qsim`` (ng, T, ns) are
    filled in place (a size-0 ``res_qsim`` skips the hydrographs, mw_multiple_run.f90:113)."""
    pk = L.Packed()
    s, m, i = L.pack_setup(setup, mesh, pk), L.pack_mesh(mesh, setup, pk), L.pack_input(input_data, setup, mesh, pk)
    wb = []
    p, st = L.pack_parameters(parameters, pk), L.pack_states(states, pk)
    o = L.pack_output(output, setup, mesh, pk, wb)
    smp = np.asfortranarray(sample, dtype=np.float32)
    ind = np.ascontiguousarray(ind_parameters_states, dtype=np.int32)
    nvar, ns = smp.shape
    rc = np.zeros(ns, dtype=np.float32)
    want_q = res_qsim is not None and res_qsim.size > 0
    rq = np.zeros((mesh.ng, setup._ntime_step, ns), dtype=np.float32, order="F") if want_q else None
    L.check(L.lib().smash_b200_compute_multiple_run(
        C.byref(s), C.byref(m), C.byref(i), C.byref(p), C.byref(st), C.byref(o), L._fp(smp), L._ip(ind), nvar, ns, L._fp(rc),
        L._fp(rq) if want_q else None))
    res_cost[...] = rc
    if want_q:
        res_qsim[...] = rq
