"""Drop-in for the f90wrap module ``smash.solver._mw_forward`` (forward/mw_forward.f90:18-181).

Same function names, argument order and in-place behaviour as the wrapped Fortran; the work is done by
libsmash_b200.so on the GPU.  ``forward_d`` / ``hyper_forward_d`` (tangent mode) are not provided.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib as L


def _prep(setup, mesh, input_data):
    pk = L.Packed()
    return pk, L.pack_setup(setup, mesh, pk), L.pack_mesh(mesh, setup, pk), L.pack_input(input_data, setup, mesh, pk)


def forward(setup, mesh, input_data, parameters, parameters_bgd, states, states_bgd, output, cost=0.0):
    """mw_forward.f90:18-39.  Callers read ``output.cost``; the value is also returned."""
    pk, s, m, i = _prep(setup, mesh, input_data)
    wb = []
    p, pb = L.pack_parameters(parameters, pk, wb), L.pack_parameters(parameters_bgd, pk)
    st, sb = L.pack_states(states, pk, wb), L.pack_states(states_bgd, pk)
    o = L.pack_output(output, setup, mesh, pk, wb)
    c = C.c_float(0.0)
    L.check(L.lib().smash_b200_forward(C.byref(s), C.byref(m), C.byref(i), C.byref(p), C.byref(pb), C.byref(st),
                                       C.byref(sb), C.byref(o), C.byref(c)))
    L.finish_output(o, output, wb)
    return np.float32(c.value)


def forward_b(setup, mesh, input_data, parameters, parameters_b, parameters_bgd, parameters_bgd_b, states, states_b,
              states_bgd, states_bgd_b, output, output_b, cost=0.0, cost_b=1.0):
    """mw_forward.f90:41-68.  ``parameters_b`` / ``states_b`` receive the gradient of ``cost``."""
    pk, s, m, i = _prep(setup, mesh, input_data)
    wb = []
    p, pb = L.pack_parameters(parameters, pk, wb), L.pack_parameters(parameters_bgd, pk)
    st, sb = L.pack_states(states, pk, wb), L.pack_states(states_bgd, pk)
    p_b, st_b = L.pack_parameters(parameters_b, pk, wb), L.pack_states(states_b, pk, wb)
    o = L.pack_output(output, setup, mesh, pk, wb)
    c, cb = C.c_float(0.0), C.c_float(float(cost_b))
    L.check(L.lib().smash_b200_forward_b(C.byref(s), C.byref(m), C.byref(i), C.byref(p), C.byref(p_b), C.byref(pb),
                                         C.byref(st), C.byref(st_b), C.byref(sb), C.byref(o), C.byref(c), C.byref(cb)))
    L.finish_output(o, output, wb)
    if output_b is not None and getattr(output_b, "qsim", None) is not None:
        output_b.qsim[...] = 0.0  # forward_db.f90:8107
    return np.float32(c.value)


def hyper_forward(setup, mesh, input_data, parameters, hyper_parameters, hyper_parameters_bgd, states, hyper_states,
                  hyper_states_bgd, output, cost=0.0):
    """mw_forward.f90:99-123."""
    pk, s, m, i = _prep(setup, mesh, input_data)
    wb = []
    p, st = L.pack_parameters(parameters, pk, wb), L.pack_states(states, pk, wb)
    hp, hpb = L.pack_parameters(hyper_parameters, pk), L.pack_parameters(hyper_parameters_bgd, pk)
    hs, hsb = L.pack_states(hyper_states, pk), L.pack_states(hyper_states_bgd, pk)
    o = L.pack_output(output, setup, mesh, pk, wb)
    c = C.c_float(0.0)
    L.check(L.lib().smash_b200_hyper_forward(C.byref(s), C.byref(m), C.byref(i), C.byref(p), C.byref(hp), C.byref(hpb),
                                             C.byref(st), C.byref(hs), C.byref(hsb), C.byref(o), C.byref(c)))
    L.finish_output(o, output, wb)
    return np.float32(c.value)


def hyper_forward_b(setup, mesh, input_data, parameters, parameters_b, hyper_parameters, hyper_parameters_b,
                    hyper_parameters_bgd, hyper_parameters_bgd_b, states, states_b, hyper_states, hyper_states_b,
                    hyper_states_bgd, hyper_states_bgd_b, output, output_b, cost=0.0, cost_b=1.0):
    """mw_forward.f90:125-152."""
    pk, s, m, i = _prep(setup, mesh, input_data)
    wb = []
    p, st = L.pack_parameters(parameters, pk, wb), L.pack_states(states, pk, wb)
    hp, hs = L.pack_parameters(hyper_parameters, pk), L.pack_states(hyper_states, pk)
    hp_b, hs_b = L.pack_parameters(hyper_parameters_b, pk, wb), L.pack_states(hyper_states_b, pk, wb)
    o = L.pack_output(output, setup, mesh, pk, wb)
    c, cb = C.c_float(0.0), C.c_float(float(cost_b))
    L.check(L.lib().smash_b200_hyper_forward_b(C.byref(s), C.byref(m), C.byref(i), C.byref(p), C.byref(hp), C.byref(hp_b),
                                               C.byref(st), C.byref(hs), C.byref(hs_b), C.byref(o), C.byref(c), C.byref(cb)))
    L.finish_output(o, output, wb)
    return np.float32(c.value)
