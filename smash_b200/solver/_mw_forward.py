"""Drop-in for the f90wrap module ``smash.solver._mw_forward`` (forward/mw_forward.f90:18-181).

Same function names, argument order and in-place behaviour as the wrapped Fortran; the work is done by
libsmash_b200.so on the GPU.  ``forward_d`` / ``hyper_forward_d`` keep the reference's argument lists but are NOT a
Tapenade tangent: the directional derivative of the cost is formed from central differences of the GPU forward (two
extra forward runs), which is what the reference's only caller -- the scalar product test -- needs.
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


def _directional(run, fields, step):
    """(J(k + h dk) - J(k - h dk)) / (2 h) for the perturbation ``fields`` = [(object, name, direction array)]."""
    base = [(obj, n, np.array(getattr(obj, n), dtype=np.float32, order="F", copy=True)) for obj, n, _ in fields]
    out = []
    try:
        for sign in (+1.0, -1.0):
            for (obj, n, d), (_, _, x0) in zip(fields, base):
                setattr(obj, n, np.asfortranarray((x0.astype(np.float64) + sign * step * np.asarray(d, np.float64)).astype(np.float32)))
            out.append(float(run()))
    finally:
        for obj, n, x0 in base:
            setattr(obj, n, x0)
    return (out[0] - out[1]) / (2.0 * step)


def _planes(obj, dobj):
    return [(obj, n, getattr(dobj, n)) for n in vars(obj)
            if isinstance(getattr(obj, n), np.ndarray) and isinstance(getattr(dobj, n, None), np.ndarray) and np.any(getattr(dobj, n))]


def forward_d(setup, mesh, input_data, parameters, parameters_d, parameters_bgd, parameters_bgd_d, states, states_d,
              states_bgd, states_bgd_d, output, output_d, cost=0.0, cost_d=0.0, eps=0.05):
    """Argument list of mw_forward.f90:70-97.  Returns (cost, cost_d) with cost_d = dJ/dk . (parameters_d, states_d) from
    central differences of ``forward`` with the absolute step ``eps`` (no tangent kernel; the background terms are held
    fixed, as in the reference's scalar product test where *_bgd_d = 0)."""
    fields = _planes(parameters, parameters_d) + _planes(states, states_d)

    def run():
        o = output.copy()
        forward(setup, mesh, input_data, parameters.copy(), parameters_bgd, states.copy(), states_bgd, o)
        return o.cost

    cd = _directional(run, fields, float(eps)) if fields else 0.0
    c = forward(setup, mesh, input_data, parameters, parameters_bgd, states, states_bgd, output)
    if output_d is not None and hasattr(output_d, "cost"):
        output_d.cost = np.float32(cd)
    return np.float32(c), np.float32(cd)


def hyper_forward_d(setup, mesh, input_data, parameters, parameters_d, hyper_parameters, hyper_parameters_d,
                    hyper_parameters_bgd, states, states_d, hyper_states, hyper_states_d, hyper_states_bgd, output, output_d,
                    cost=0.0, cost_d=0.0, eps=1e-3):
    """Argument list of mw_forward.f90:154-181: the same central differences along (hyper_parameters_d, hyper_states_d)."""
    fields = _planes(hyper_parameters, hyper_parameters_d) + _planes(hyper_states, hyper_states_d)

    def run():
        o = output.copy()
        hyper_forward(setup, mesh, input_data, parameters.copy(), hyper_parameters, hyper_parameters_bgd, states.copy(),
                      hyper_states, hyper_states_bgd, o)
        return o.cost

    cd = _directional(run, fields, float(eps)) if fields else 0.0
    c = hyper_forward(setup, mesh, input_data, parameters, hyper_parameters, hyper_parameters_bgd, states, hyper_states,
                      hyper_states_bgd, output)
    if output_d is not None and hasattr(output_d, "cost"):
        output_d.cost = np.float32(cd)
    return np.float32(c), np.float32(cd)
