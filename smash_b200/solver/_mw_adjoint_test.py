"""Drop-in for ``smash.solver._mw_adjoint_test`` (optimize/mw_adjoint_test.f90:26-189).

The reference's scalar product test pairs the Tapenade tangent ``forward_d`` with the adjoint ``forward_b``:
sp1 = <dY*, dY> = cost_b * cost_d and sp2 = <dk*, dk> = sum(parameters_b * parameters_d) with dk = 1 on every parameter
plane.  No tangent kernel exists here; the directional derivative cost_d = dJ/dk . dk is obtained from central
differences of the GPU ``forward`` instead (``forward_d`` below), the adjoint side is the GPU ``forward_b``."""
from __future__ import annotations

import numpy as np

from ._derived_types import ParametersDT, StatesDT
from . import _mw_forward


def forward_d(setup, mesh, input_data, parameters, parameters_d, parameters_bgd, states, states_d, states_bgd, output,
              eps=0.05, solver=None):
    """Directional derivative of the cost along (parameters_d, states_d): the tangent linear model of
    forward/mw_forward.f90:70-97 evaluated by central differences with the absolute step ``eps``.
    Returns (cost, cost_d)."""
    forward = (solver or _mw_forward).forward
    names_p = [n for n in vars(parameters) if isinstance(getattr(parameters, n), np.ndarray)]
    names_s = [n for n in vars(states) if isinstance(getattr(states, n), np.ndarray)]

    def run(sign):
        p, s = parameters.copy(), states.copy()
        for obj, dobj, names in ((p, parameters_d, names_p), (s, states_d, names_s)):
            for n in names:
                d = np.asarray(getattr(dobj, n), np.float64)
                if not d.any():
                    continue
                x = np.asarray(getattr(obj, n), np.float64)
                setattr(obj, n, np.asfortranarray((x + sign * step * d).astype(np.float32)))
        o = output.copy()
        forward(setup, mesh, input_data, p, parameters_bgd, s, states_bgd, o)
        return float(o.cost)

    # J(k + h dk) - J(k - h dk) = 2 h dJ/dk . dk; h is absolute (dk = 1 in the scalar product test): small against every
    # field of the structure, large against the float32 noise of the cost
    step = float(eps)
    cp = run(+1.0)
    cm = run(-1.0)
    o = output.copy()
    forward(setup, mesh, input_data, parameters.copy(), parameters_bgd, states.copy(), states_bgd, o)
    return np.float32(o.cost), np.float32((cp - cm) / (2.0 * step))


def scalar_product_test(setup, mesh, input_data, parameters, states, output, verbose=True, solver=None):
    """mw_adjoint_test.f90:26-105.  Prints the reference's three lines and returns (sp1, sp2)."""
    if verbose:
        print("</> Scalar Product Test")
    parameters_bgd, states_bgd = parameters.copy(), states.copy()
    parameters_d, states_d = ParametersDT(mesh), StatesDT(mesh)
    for n in vars(parameters_d):
        if isinstance(getattr(parameters_d, n), np.ndarray):
            getattr(parameters_d, n)[...] = 1.0                           # set_parameters(mesh, parameters_d, 1._sp)
    for n in vars(states_d):
        if isinstance(getattr(states_d, n), np.ndarray):
            getattr(states_d, n)[...] = 0.0                               # set_states(mesh, states_d, 0._sp)
    if verbose:
        print("    Tangent Linear Model dY  = (dM/dk) (k) . dk")
    _, cost_d = forward_d(setup, mesh, input_data, parameters, parameters_d, parameters_bgd, states, states_d, states_bgd, output,
                          solver=solver)
    if verbose:
        print("    Adjoint Model        dk* = (dM/dk)* (k) . dY*")
    parameters_b, states_b = ParametersDT(mesh), StatesDT(mesh)
    (solver or _mw_forward).forward_b(setup, mesh, input_data, parameters.copy(), parameters_b, parameters_bgd, None, states.copy(), states_b, states_bgd,
              None, output, None, 0.0, 1.0)
    sp1 = float(cost_d)                                                   # cost_b * cost_d, cost_b = 1
    sp2 = float(sum(np.asarray(getattr(parameters_b, n), np.float64).sum() for n in vars(parameters_b)
                    if isinstance(getattr(parameters_b, n), np.ndarray)))
    if verbose:
        print("    <dY*, dY> (sp1) = %12.8f" % sp1)
        print("    <dk*, dk> (sp2) = %12.8f" % sp2)
        print("    Relative Error  = %12.8f" % ((sp1 - sp2) / sp1 if sp1 else float("nan")))
    return sp1, sp2
