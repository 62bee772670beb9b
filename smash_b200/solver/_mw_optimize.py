"""Drop-in for the f90wrap module ``smash.solver._mw_optimize`` (optimize/mw_optimize.f90:53-1177).

The reference's optimisers are host-side drivers around ``forward`` / ``forward_b`` / ``hyper_forward_b``; here
they are the same drivers around the GPU entry points of ``_mw_forward`` (libsmash_b200.so), with the same
names, argument order and in-place behaviour:

* ``optimize_sbs``            mw_optimize.f90:53-294   step-by-step search on a transformed uniform control
* ``optimize_lbfgsb``         mw_optimize.f90:484-676  L-BFGS-B on the normalised distributed control
* ``optimize_hyper_lbfgsb``   mw_optimize.f90:779-958  L-BFGS-B on the hyper-linear / hyper-polynomial control

L-BFGS-B itself is the third-party routine ``setulb`` (Zhu, Byrd, Lu, Nocedal, version 3.0; the reference
vendors it as ``solver/optimize/lbfgsb.f``, here SciPy's reverse-communication build of the same 3.0 code is
called, ``scipy.optimize._lbfgsb.setulb``), driven with the reference's settings ``m = 10``, ``factr = 10``
(``1e6`` for hyper), ``pgtol = 1e-12``, ``maxls = 20`` and the reference's own stopping tests on ``isave(30)``
and ``dsave(13)``.

``solver`` (keyword-only) is the module that provides ``forward`` ... ``hyper_forward_b``; it defaults to the
GPU path and exists so that the test-suite can drive the very same host logic with the CPU oracle as a checker.
"""
from __future__ import annotations

import numpy as np

from . import _mw_forward
from ._derived_types import (GNP, GNS, GPARAMETERS_NAME, GSTATES_NAME, Hyper_ParametersDT, Hyper_StatesDT, OutputDT,
                             ParametersDT, StatesDT)

f32 = np.float32

# setulb task codes of SciPy's C build (scipy/optimize/__lbfgsb.c): task[0]
_NEW_X, _START, _FG, _STOP = 1, 0, 3, 5
_STOP_MAXITER, _STOP_PGTOL = 504, 505


# ------------------------------------------------------------------------------------------ helpers
def _planes(parameters, states):
    """get_parameters + get_states (mwd_parameters_manipulation.f90:59-86, mwd_states_manipulation.f90):
    the 16 + 8 (nrow, ncol) planes in GPARAMETERS_NAME / GSTATES_NAME order (references, not copies)."""
    return [getattr(parameters, n) for n in GPARAMETERS_NAME] + [getattr(states, n) for n in GSTATES_NAME]


def _optim_flags(setup):
    o = setup._optimize
    optim = np.concatenate([np.asarray(o.optim_parameters), np.asarray(o.optim_states)]).astype(np.int64)
    lb = np.concatenate([np.asarray(o.lb_parameters, f32), np.asarray(o.lb_states, f32)])
    ub = np.concatenate([np.asarray(o.ub_parameters, f32), np.asarray(o.ub_states, f32)])
    return optim, lb, ub


def _first_active(mesh):
    """``maxloc(mesh%active_cell)``: first maximum in Fortran (column-major) element order."""
    a = np.asarray(mesh.active_cell)
    k = int(np.argmax(a.ravel(order="F")))
    return k % a.shape[0], k // a.shape[0]


def _active_mask_colmajor(mesh):
    """Flat indices (Fortran order) of the active cells in the order ``do col / do row`` visits them
    (mw_optimize.f90:703-712)."""
    return np.flatnonzero(np.asarray(mesh.active_cell).ravel(order="F") == 1)


def normalize_parameters(setup, mesh, parameters):
    """mwd_parameters_manipulation.f90:154-179"""
    o = setup._optimize
    for i, n in enumerate(GPARAMETERS_NAME):
        lb, ub = f32(o.lb_parameters[i]), f32(o.ub_parameters[i])
        a = getattr(parameters, n)
        a[...] = (a - lb) / (ub - lb)


def normalize_states(setup, mesh, states):
    """mwd_states_manipulation.f90 (twin of normalize_parameters)"""
    o = setup._optimize
    for i, n in enumerate(GSTATES_NAME):
        lb, ub = f32(o.lb_states[i]), f32(o.ub_states[i])
        a = getattr(states, n)
        a[...] = (a - lb) / (ub - lb)


# ------------------------------------------------------------------------------------------ sbs
def _transformation_sbs(x, l, u):
    """mw_optimize.f90:422-450 (real(sp) arithmetic)"""
    x_t = np.empty_like(x)
    for i in range(x.size):
        if l[i] < 0:
            x_t[i] = np.arcsinh(x[i])
        elif l[i] >= 0 and u[i] <= 1:
            x_t[i] = np.log(x[i] / (f32(1) - x[i]))
        else:
            x_t[i] = np.log(x[i])
    return x_t


def _inv_transformation_sbs(x_t, l, u):
    """mw_optimize.f90:453-481"""
    x = np.empty_like(x_t)
    for i in range(x.size):
        if l[i] < 0:
            x[i] = np.sinh(x_t[i])
        elif l[i] >= 0 and u[i] <= 1:
            x[i] = np.exp(x_t[i]) / (f32(1) + np.exp(x_t[i]))
        else:
            x[i] = np.exp(x_t[i])
    return x


def _control_to_var_sbs(setup, mesh, parameters, states, x, idx, mask):
    """mw_optimize.f90:374-419: the control value is written on every active cell of its field."""
    planes = _planes(parameters, states)
    for j, i in enumerate(idx):
        planes[i][mask] = x[j]


def optimize_sbs(setup, mesh, input_data, parameters, states, output, *, solver=None):
    """mw_optimize.f90:53-294.  ``parameters``, ``states`` and ``output`` are updated in place."""
    sv = solver or _mw_forward
    o = setup._optimize
    parameters_bgd, states_bgd = parameters.copy(), states.copy()

    def run():
        return f32(sv.forward(setup, mesh, input_data, parameters, parameters_bgd, states, states_bgd, output))

    cost = run()
    optim, lb, ub = _optim_flags(setup)
    idx = np.flatnonzero(optim > 0)
    n = idx.size
    l, u = lb[idx].astype(f32), ub[idx].astype(f32)                      # bounds_initialise_sbs :296-332
    r0, c0 = _first_active(mesh)
    mask = np.asarray(mesh.active_cell) == 1
    planes = _planes(parameters, states)
    x = np.array([planes[i][r0, c0] for i in idx], dtype=f32)            # var_to_control_sbs :334-372
    x_t, l_t, u_t = _transformation_sbs(x, l, u), _transformation_sbs(l, l, u), _transformation_sbs(u, l, u)

    gx = cost
    ga = gx
    clg = f32(0.7) ** (f32(1.0) / f32(n))
    z_t = x_t.copy()
    sdx = np.zeros(n, f32)
    ddx = f32(0.64)
    dxn = ddx
    ia = iaa = iam = jfa = jfaa = 0
    nfg = 1
    if o.verbose:
        print(f"    At iterate    {0:3d}    nfg = {nfg:5d}    J ={gx:10.6f}    ddx ={ddx:5.2f}")
    maxit = int(o.maxiter) * n
    for it in range(1, maxit + 1):
        if dxn > ddx:
            dxn = ddx
        if ddx > 2:
            ddx = dxn
        for i in range(1, n + 1):
            y_t = x_t.copy()
            for jf in (-1, 1):
                if i == iaa and jf == -jfaa:
                    continue
                if x_t[i - 1] <= l_t[i - 1] and jf < 0:
                    continue
                if x_t[i - 1] >= u_t[i - 1] and jf > 0:
                    continue
                y_t[i - 1] = min(max(f32(x_t[i - 1] + f32(jf) * ddx), l_t[i - 1]), u_t[i - 1])
                y = _inv_transformation_sbs(y_t, l, u)
                _control_to_var_sbs(setup, mesh, parameters, states, y, idx, mask)
                f = run()
                nfg += 1
                if f < gx:
                    z_t = y_t.copy()
                    gx = f
                    ia = i
                    jfa = jf
        iaa, jfaa = ia, jfa
        if ia != 0:
            x_t = z_t.copy()
            x = _inv_transformation_sbs(x_t, l, u)
            _control_to_var_sbs(setup, mesh, parameters, states, x, idx, mask)
            sdx = (clg * sdx).astype(f32)
            # sdx(ia) was already scaled by clg on the line above, as in the reference (:199-200)
            sdx[ia - 1] = (f32(1) - clg) * f32(jfa) * ddx + clg * sdx[ia - 1]
            iam += 1
            if iam > 2 * n:
                ddx = f32(ddx * f32(2))
                iam = 0
            if gx < ga - 2:
                ga = gx
        else:
            ddx = f32(ddx / f32(2))
            iam = 0
        if it > 4 * n:
            y_t = np.minimum(np.maximum((x_t + sdx).astype(f32), l_t), u_t)
            y = _inv_transformation_sbs(y_t, l, u)
            _control_to_var_sbs(setup, mesh, parameters, states, y, idx, mask)
            f = run()
            nfg += 1
            if f < gx:
                gx = f
                jfaa = 0
                x_t = y_t.copy()
                x = _inv_transformation_sbs(x_t, l, u)
                _control_to_var_sbs(setup, mesh, parameters, states, x, idx, mask)
                if gx < ga - 2:
                    ga = gx
        ia = 0
        if it % n == 0 and o.verbose:
            print(f"    At iterate    {it // n:3d}    nfg = {nfg:5d}    J ={gx:10.6f}    ddx ={ddx:5.2f}")
        stop = None
        if ddx < f32(0.01):
            stop = "CONVERGENCE: DDX < 0.01"
        elif it == maxit:
            stop = "STOP: TOTAL NO. OF ITERATION EXCEEDS LIMIT"
        if stop:
            if o.verbose:
                print(f"    {stop}\n")
            _control_to_var_sbs(setup, mesh, parameters, states, x, idx, mask)
            run()
            break
    return nfg


# ------------------------------------------------------------------------------------------ l-bfgs-b
class _Setulb:
    """Reverse-communication state of one L-BFGS-B minimisation (the work arrays of mw_optimize.f90:503-520)."""

    def __init__(self, n, m, factr, pgtol, x, l, u, nbd):
        from scipy.optimize import _lbfgsb
        self._setulb = _lbfgsb.setulb
        it = np.int32
        try:                                              # ILP64 builds of SciPy use 64-bit work integers
            from scipy.optimize._lbfgsb_py import HAS_ILP64
            if HAS_ILP64:
                it = np.int64
        except ImportError:
            pass
        self.m, self.factr, self.pgtol = int(m), float(factr), float(pgtol)
        self.x = np.array(x, dtype=np.float64)
        self.l, self.u = np.array(l, dtype=np.float64), np.array(u, dtype=np.float64)
        self.nbd = np.array(nbd, dtype=it)
        self.f = np.array(0.0, dtype=np.float64)
        self.g = np.zeros(n, dtype=np.float64)
        self.wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
        self.iwa = np.zeros(3 * n, dtype=it)
        self.task, self.ln_task = np.zeros(2, dtype=it), np.zeros(2, dtype=it)
        self.lsave, self.isave, self.dsave = np.zeros(4, dtype=it), np.zeros(44, dtype=it), np.zeros(29, np.float64)

    def step(self):
        self._setulb(self.m, self.x, self.l, self.u, self.nbd, self.f, self.g, self.factr, self.pgtol, self.wa, self.iwa,
                     self.task, self.lsave, self.isave, self.dsave, 20, self.ln_task)
        return int(self.task[0])

    @property
    def iteration(self):       # isave(30)
        return int(self.isave[29])

    @property
    def nfg(self):             # isave(34)
        return int(self.isave[33])

    @property
    def projg(self):           # dsave(13)
        return float(self.dsave[12])


def _drive_lbfgsb(sb: _Setulb, setup, output, control_to_var, fg):
    """The ``do while`` of mw_optimize.f90:568-651 / 851-931.  ``fg()`` returns (cost, gradient)."""
    o = setup._optimize
    first = True
    while True:
        task = sb.step()
        control_to_var(sb.x)
        if task == _FG:
            cost, g = fg()
            sb.f = np.array(float(cost), dtype=np.float64)
            sb.g[...] = g
            if first and o.verbose:
                print(f"    At iterate    {0:3d}    nfg = {1:5d}    J ={float(cost):14.6f}    Jobs ={float(output.cost_jobs):14.6f}"
                      f"    Jreg ={float(output.cost_jreg):14.6f}    |proj g| ={sb.projg:10.6f}")
            first = False
            continue
        if task == _NEW_X:
            if o.verbose:
                print(f"    At iterate    {sb.iteration:3d}    nfg = {sb.nfg:5d}    J ={float(sb.f):14.6f}"
                      f"    Jobs ={float(output.cost_jobs):14.6f}    Jreg ={float(output.cost_jreg):14.6f}"
                      f"    |proj g| ={sb.projg:10.6f}")
            if sb.iteration >= int(o.maxiter):
                msg = "STOP: TOTAL NO. OF ITERATION EXCEEDS LIMIT"
                break
            if sb.projg <= 1e-10 * (1.0 + abs(float(sb.f))):
                msg = "STOP: THE PROJECTED GRADIENT IS SUFFICIENTLY SMALL"
                break
            continue
        if task == _START:
            continue
        msg = {4: "CONVERGENCE", 2: "ABNORMAL_TERMINATION_IN_LNSRCH"}.get(task, f"STOP ({task}, {int(sb.task[1])})")
        break
    if o.verbose:
        print(f"    {msg}\n")
    return msg


def optimize_lbfgsb(setup, mesh, input_data, parameters, states, output, *, solver=None):
    """mw_optimize.f90:484-676.  The control is the normalised value of every optimised field on every active
    cell (field-major, cells in column-major order); ``forward`` / ``forward_b`` denormalise in place
    (``denormalize_forward``), so the fields are re-normalised after each call exactly as the reference does."""
    sv = solver or _mw_forward
    o = setup._optimize
    optim, _, _ = _optim_flags(setup)
    idx = np.flatnonzero(optim > 0)
    act = _active_mask_colmajor(mesh)
    nac = act.size
    n = nac * idx.size
    parameters_b, states_b = ParametersDT(mesh), StatesDT(mesh)
    output_b = None

    normalize_parameters(setup, mesh, parameters)
    normalize_states(setup, mesh, states)
    parameters_bgd, states_bgd = parameters.copy(), states.copy()
    zero_p, zero_s = ParametersDT(mesh), StatesDT(mesh)

    def var_to_control(par, sta):                                         # :678-725
        planes = _planes(par, sta)
        x = np.empty(n, dtype=np.float64)
        for j, i in enumerate(idx):
            x[j * nac:(j + 1) * nac] = np.asarray(planes[i]).ravel(order="F")[act]
        return x

    def control_to_var(x):                                                # :727-777
        planes = _planes(parameters, states)
        for j, i in enumerate(idx):
            p = planes[i]
            flat = np.asarray(p).ravel(order="F").copy()
            flat[act] = x[j * nac:(j + 1) * nac].astype(f32)
            p[...] = flat.reshape(p.shape, order="F")

    x0 = var_to_control(parameters, states)
    o.denormalize_forward = True
    try:
        sv.forward(setup, mesh, input_data, parameters, parameters_bgd, states, states_bgd, output)
        normalize_parameters(setup, mesh, parameters)
        normalize_states(setup, mesh, states)
        output._cost_jobs_initial = f32(output.cost_jobs)
        output._cost_jreg_initial = f32(output.cost_jreg)

        def fg():
            cost = sv.forward_b(setup, mesh, input_data, parameters, parameters_b, parameters_bgd, zero_p, states, states_b,
                                states_bgd, zero_s, output, output_b, 0.0, 1.0)
            normalize_parameters(setup, mesh, parameters)
            normalize_states(setup, mesh, states)
            return cost, var_to_control(parameters_b, states_b)

        sb = _Setulb(n, 10, 1e1, 1e-12, x0, np.zeros(n), np.ones(n), np.full(n, 2))
        msg = _drive_lbfgsb(sb, setup, output, control_to_var, fg)
        sv.forward(setup, mesh, input_data, parameters, parameters_bgd, states, states_bgd, output)
    finally:
        o.denormalize_forward = False
    return msg


def optimize_hyper_lbfgsb(setup, mesh, input_data, parameters, states, output, *, solver=None):
    """mw_optimize.f90:779-958.  Descriptors are normalised to [0, 1] for the duration of the optimisation and
    restored afterwards (:960-999); the control holds ``nhyper`` coefficients per optimised field."""
    return optimize_hyper_lbfgsb_multi([(setup, mesh, input_data, parameters, states, output)], solver=solver)


def optimize_hyper_lbfgsb_multi(catchments, *, solver=None, reduce_sum=None, reduce_minmax=None):
    """Regionalised calibration over several catchments that share one hyper-parameter set: minimises the SUM of the
    catchments' costs with the driver of ``optimize_hyper_lbfgsb`` (one catchment = the reference's subroutine).

    ``catchments`` is a list of ``(setup, mesh, input_data, parameters, states, output)``; every ``setup._optimize`` holds
    the same mapping, control vector, bounds and ``maxiter``, and the catchments start from the same uniform background.
    Descriptors are normalised with the extrema over ALL catchments, so that a coefficient means the same everywhere.
    ``reduce_sum(vector) -> vector`` and ``reduce_minmax(mins, maxs) -> (mins, maxs)`` combine the local values with
    those of other processes (``smash_b200.distributed``: one small all-reduce per evaluation); with them every process
    sees the same cost and gradient and therefore walks the same L-BFGS-B path."""
    sv = solver or _mw_forward
    setup = catchments[0][0]
    o = setup._optimize
    nh = int(o.nhyper)
    nd = int(setup._nd)
    optim, lb, ub = _optim_flags(setup)
    idx = np.flatnonzero(optim > 0)
    n = idx.size * nh

    dmin = np.array([min(c[2].descriptor[:, :, i].min() for c in catchments) for i in range(nd)], dtype=f32)
    dmax = np.array([max(c[2].descriptor[:, :, i].max() for c in catchments) for i in range(nd)], dtype=f32)
    if reduce_minmax is not None:
        dmin, dmax = reduce_minmax(dmin, dmax)
    for c in catchments:                                                  # :960-980
        for i in range(nd):
            c[2].descriptor[:, :, i] = (c[2].descriptor[:, :, i] - dmin[i]) / (dmax[i] - dmin[i])

    try:
        hyper_parameters, hyper_states = Hyper_ParametersDT(setup), Hyper_StatesDT(setup)
        hplanes = _planes(hyper_parameters, hyper_states)
        work = []                                                         # per catchment: gradient holders
        for c in catchments:
            work.append((Hyper_ParametersDT(c[0]), Hyper_StatesDT(c[0]), ParametersDT(c[1]), StatesDT(c[1])))

        # problem_initialise_hyper_lbfgsb :1001-1096
        _, mesh0, _, parameters0, states0, _ = catchments[0]
        r0, c0 = _first_active(mesh0)
        v = np.array([p[r0, c0] for p in _planes(parameters0, states0)], dtype=f32)
        first = np.log(np.maximum(f32(1e-8), v - lb) / np.maximum(f32(1e-8), ub - v)).astype(f32)
        nbd, l, u = np.zeros(n, np.int32), np.zeros(n), np.zeros(n)
        for i in range(GNP + GNS):
            hplanes[i][...] = 0.0
            hplanes[i][0, 0] = first[i]
        k = 0
        for i in idx:
            if str(o.mapping).strip() == "hyper-polynomial":
                for j in range(1, nh):                                    # Fortran j = 1 .. nhyper-1, entry j+1
                    if (j + 1) % 2 == 0:
                        hplanes[i][j, 0] = 0.0
                    else:
                        hplanes[i][j, 0] = 1.0
                        nbd[k + j], l[k + j], u[k + j] = 2, 0.5, 2.0
            k += nh
        hyper_parameters_bgd, hyper_states_bgd = hyper_parameters.copy(), hyper_states.copy()

        def var_to_control(hp, hs):                                       # :1098-1135
            pl = _planes(hp, hs)
            return np.concatenate([np.asarray(pl[i])[:, 0].astype(np.float64) for i in idx])

        def control_to_var(x):                                            # :1137-1177
            for j, i in enumerate(idx):
                hplanes[i][:, 0] = x[j * nh:(j + 1) * nh].astype(f32)

        single = len(catchments) == 1 and reduce_sum is None

        def fg():
            if single:
                (su, me, inp, par, sta, out), (hp_b, hs_b, par_b, sta_b) = catchments[0], work[0]
                cost = sv.hyper_forward_b(su, me, inp, par, par_b, hyper_parameters, hp_b, hyper_parameters_bgd, None, sta,
                                          sta_b, hyper_states, hs_b, hyper_states_bgd, None, out, None, 0.0, 1.0)
                return cost, var_to_control(hp_b, hs_b)
            acc = np.zeros(n + 1, dtype=np.float64)
            for (su, me, inp, par, sta, out), (hp_b, hs_b, par_b, sta_b) in zip(catchments, work):
                cost = sv.hyper_forward_b(su, me, inp, par, par_b, hyper_parameters, hp_b, hyper_parameters_bgd, None, sta,
                                          sta_b, hyper_states, hs_b, hyper_states_bgd, None, out, None, 0.0, 1.0)
                acc[0] += float(cost)
                acc[1:] += var_to_control(hp_b, hs_b)
            if reduce_sum is not None:
                acc = np.asarray(reduce_sum(acc), dtype=np.float64)
            return f32(acc[0]), acc[1:]

        sb = _Setulb(n, 10, 1e6, 1e-12, var_to_control(hyper_parameters, hyper_states), l, u, nbd)
        msg = _drive_lbfgsb(sb, setup, catchments[0][5], control_to_var, fg)
        # hyper_forward maps the hyper control to parameters / states in place (forward.f90:117-121), which is
        # what hyper_parameters_to_parameters / hyper_states_to_states repeat at :953-954
        for su, me, inp, par, sta, out in catchments:
            sv.hyper_forward(su, me, inp, par, hyper_parameters, hyper_parameters_bgd, sta, hyper_states, hyper_states_bgd, out)
    finally:
        for c in catchments:                                              # :982-999
            for i in range(nd):
                c[2].descriptor[:, :, i] = c[2].descriptor[:, :, i] * (dmax[i] - dmin[i]) + dmin[i]
    return msg
