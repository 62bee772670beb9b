"""CPU ORACLE -- test infrastructure, NOT the product.

ctypes front-end of ``oracle/smash_oracle.c`` (a C restatement of the reference's gr-a forward solver, cost
function and Tapenade adjoint; every C function cites the reference file:line it follows).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package; ``smash_b200`` never does.

Parity status: PINNED against the reference's golden file ``smash/tests/baseline.hdf5`` (see
``tests/test_oracle_golden.py`` and ``tests/golden/make_golden.py``).  The reference itself cannot be built
in this environment (no Fortran compiler), so there is no ``oracle/_ref``.

The functions take the same duck-typed objects as the reference's wrapped entry points
(``setup, mesh, input_data, parameters, ...`` with the attribute names of ``smash/solver/derived_type``) and
mutate them in place in the same way.  ``precision="f32"`` is the reference's real kind; ``"f64"`` is the
double-precision referee used for finite-difference / Taylor checks.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GNP, GNS = 16, 8
PARAM_NAMES = ("ci", "cp", "beta", "cft", "cst", "alpha", "exc", "b", "cusl1", "cusl2", "clsl", "ks", "ds", "dsm",
               "ws", "lr")
STATE_NAMES = ("hi", "hp", "hft", "hst", "husl1", "husl2", "hlsl", "hlr")
JOBS_FUN = {"nse": 1, "kge": 2, "kge2": 3, "se": 4, "rmse": 5, "logarithmic": 6, "Crc": 7, "Cfp2": 8, "Cfp10": 9, "Cfp50": 10,
            "Cfp90": 11, "Erc": 12, "Elt": 13, "Epf": 14}
JREG_FUN = {"prior": 1, "smoothing": 2, "hard_smoothing": 3}
MAPPING = {"hyper-linear": 1, "hyper-polynomial": 2}


def build(force: bool = False) -> None:
    """Compile liboracle_f32.so / liboracle_f64.so with oracle/Makefile (gcc)."""
    if force:
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)


_libs = {}


def _lib(precision: str):
    if precision not in _libs:
        path = os.path.join(_HERE, f"liboracle_{precision}.so")
        if not os.path.exists(path):
            build()
        try:
            _libs[precision] = C.CDLL(path)
        except OSError:
            build(force=True)
            _libs[precision] = C.CDLL(path)
    return _libs[precision]


STRUCTURES = {"gr-a": 1, "gr-b": 2, "gr-c": 3, "gr-d": 4, "vic-a": 5}   # OST_* of smash_oracle.h


def _make_struct(real):
    rp, ip = C.POINTER(real), C.POINTER(C.c_int32)

    class OProblem(C.Structure):
        _fields_ = [
            ("ntime_step", C.c_int), ("nd", C.c_int), ("dt", real),
            ("sparse_storage", C.c_int), ("save_qsim_domain", C.c_int), ("save_net_prcp_domain", C.c_int),
            ("nrow", C.c_int), ("ncol", C.c_int), ("ng", C.c_int), ("nac", C.c_int), ("dx", real),
            ("flwdir", ip), ("flwacc", ip), ("active_cell", ip), ("local_active_cell", ip), ("path", ip),
            ("gauge_pos", ip), ("rowcol_to_ind_sparse", ip), ("area", rp),
            ("prcp", rp), ("pet", rp), ("qobs", rp), ("descriptor", rp),
            ("njf", C.c_int), ("jobs_fun", ip), ("wjobs_fun", rp),
            ("njr", C.c_int), ("jreg_fun", ip), ("wjreg_fun", rp), ("wjreg", real),
            ("denormalize_forward", C.c_int), ("optimize_start_step", C.c_int), ("mapping", C.c_int), ("nhyper", C.c_int),
            ("optim_parameters", ip), ("optim_states", ip),
            ("lb_parameters", rp), ("ub_parameters", rp), ("lb_states", rp), ("ub_states", rp), ("wgauge", rp),
            ("mean_prcp", rp), ("mask_event", ip), ("structure", C.c_int),
        ]

    return OProblem


_STRUCTS = {"f32": (_make_struct(C.c_float), np.float32, C.c_float),
            "f64": (_make_struct(C.c_double), np.float64, C.c_double)}


class _Ctx:
    def __init__(self, precision):
        self.S, self.dtype, self.creal = _STRUCTS[precision]
        self.lib = _lib(precision)
        self.sfx = "_" + precision
        self.keep = []

    def r(self, a):
        if a is None:
            return None
        b = np.asfortranarray(a, dtype=self.dtype)
        self.keep.append(b)
        return b.ctypes.data_as(C.POINTER(self.creal))

    def i(self, a):
        if a is None:
            return None
        b = np.asfortranarray(a, dtype=np.int32)
        self.keep.append(b)
        return b.ctypes.data_as(C.POINTER(C.c_int32))

    def fn(self, name):
        return getattr(self.lib, name + self.sfx)


def _problem(ctx: _Ctx, setup, mesh, input_data):
    o = setup._optimize
    P = ctx.S()
    P.structure = STRUCTURES[str(getattr(setup, "structure", "gr-a")).strip()]   # forward.f90:43
    P.ntime_step, P.nd, P.dt = int(setup._ntime_step), int(setup._nd), float(setup.dt)
    P.sparse_storage = int(bool(setup.sparse_storage))
    P.save_qsim_domain, P.save_net_prcp_domain = int(bool(setup.save_qsim_domain)), int(bool(setup.save_net_prcp_domain))
    P.nrow, P.ncol, P.ng, P.nac, P.dx = int(mesh.nrow), int(mesh.ncol), int(mesh.ng), int(mesh.nac), float(mesh.dx)
    P.flwdir, P.flwacc, P.active_cell = ctx.i(mesh.flwdir), ctx.i(mesh.flwacc), ctx.i(mesh.active_cell)
    P.local_active_cell = ctx.i(getattr(mesh, "_local_active_cell", None))
    P.path = ctx.i(np.asarray(mesh.path, dtype=np.int32) + 1)  # Python side is 0-based
    if mesh.ng > 0:
        P.gauge_pos = ctx.i(np.asarray(mesh.gauge_pos, dtype=np.int32) + 1)
        P.area = ctx.r(mesh.area)
        P.qobs = ctx.r(input_data.qobs)
        P.wgauge = ctx.r(o.wgauge)
    if setup.sparse_storage:
        P.rowcol_to_ind_sparse = ctx.i(mesh._rowcol_to_ind_sparse)
        P.prcp, P.pet = ctx.r(input_data.sparse_prcp), ctx.r(input_data.sparse_pet)
    else:
        P.prcp, P.pet = ctx.r(input_data.prcp), ctx.r(input_data.pet)
    if setup._nd > 0:
        P.descriptor = ctx.r(input_data.descriptor)
    jf = [str(x).strip() for x in np.atleast_1d(o.jobs_fun)][: int(o.njf)]
    P.njf = len(jf)
    P.jobs_fun = ctx.i(np.array([JOBS_FUN[x] for x in jf], dtype=np.int32))
    P.wjobs_fun = ctx.r(np.atleast_1d(o.wjobs_fun)[: len(jf)])
    jr = [str(x).strip() for x in np.atleast_1d(o.jreg_fun)][: int(o.njr)]
    P.njr = len(jr)
    P.jreg_fun = ctx.i(np.array([JREG_FUN[x] for x in jr], dtype=np.int32))
    P.wjreg_fun = ctx.r(np.atleast_1d(o.wjreg_fun)[: len(jr)])
    P.wjreg = float(o.wjreg)
    P.denormalize_forward = int(bool(o.denormalize_forward))
    P.optimize_start_step = int(o.optimize_start_step)
    P.mapping = MAPPING.get(str(o.mapping).strip(), 0)
    P.nhyper = int(o.nhyper)
    P.optim_parameters, P.optim_states = ctx.i(o.optim_parameters), ctx.i(o.optim_states)
    P.lb_parameters, P.ub_parameters = ctx.r(o.lb_parameters), ctx.r(o.ub_parameters)
    P.lb_states, P.ub_states = ctx.r(o.lb_states), ctx.r(o.ub_states)
    if any(JOBS_FUN[x] >= 7 for x in jf) and mesh.ng > 0:                   # signature objectives (mwd_cost.f90:117-122)
        P.mean_prcp = ctx.r(input_data.mean_prcp)
        P.mask_event = ctx.i(o.mask_event)
    return P


def _stack(obj, names, shape, dtype):
    a = np.zeros(shape + (len(names),), dtype=dtype, order="F")
    for k, n in enumerate(names):
        v = getattr(obj, n, None)
        if v is not None:
            a[..., k] = v
    return a


def _unstack(a, obj, names):
    for k, n in enumerate(names):
        v = getattr(obj, n, None)
        if v is not None:
            v[...] = a[..., k]


def _ptr(ctx, a):
    return a.ctypes.data_as(C.POINTER(ctx.creal)) if a is not None else None


def forward(setup, mesh, input_data, parameters, parameters_bgd, states, states_bgd, output, precision="f32"):
    """base_forward (forward/forward.f90:1-80)."""
    ctx = _Ctx(precision)
    P = _problem(ctx, setup, mesh, input_data)
    shp = (mesh.nrow, mesh.ncol)
    par, par_bgd = _stack(parameters, PARAM_NAMES, shp, ctx.dtype), _stack(parameters_bgd, PARAM_NAMES, shp, ctx.dtype)
    st, st_bgd = _stack(states, STATE_NAMES, shp, ctx.dtype), _stack(states_bgd, STATE_NAMES, shp, ctx.dtype)
    T = setup._ntime_step
    qsim = np.zeros((max(mesh.ng, 1), T), dtype=ctx.dtype, order="F")
    fst = np.zeros(shp + (GNS,), dtype=ctx.dtype, order="F")
    cost3 = np.zeros(3, dtype=ctx.dtype)
    qdom = netp = None
    if setup.save_qsim_domain:
        qdom = np.full((mesh.nac, T) if setup.sparse_storage else shp + (T,), -99.0, dtype=ctx.dtype, order="F")
    if setup.save_net_prcp_domain:
        netp = np.full((mesh.nac, T) if setup.sparse_storage else shp + (T,), -99.0, dtype=ctx.dtype, order="F")
    rc = ctx.fn("oracle_forward")(C.byref(P), _ptr(ctx, par), _ptr(ctx, par_bgd), _ptr(ctx, st), _ptr(ctx, st_bgd),
                                  _ptr(ctx, qsim), _ptr(ctx, fst), _ptr(ctx, cost3), _ptr(ctx, qdom), _ptr(ctx, netp))
    assert rc == 0
    _unstack(par, parameters, PARAM_NAMES)
    _unstack(st, states, STATE_NAMES)
    _unstack(fst, output.fstates, STATE_NAMES)
    if mesh.ng > 0:
        output.qsim = qsim.astype(ctx.dtype)
    if qdom is not None:
        setattr(output, "sparse_qsim_domain" if setup.sparse_storage else "qsim_domain", qdom)
    if netp is not None:
        setattr(output, "sparse_net_prcp_domain" if setup.sparse_storage else "net_prcp_domain", netp)
    output.cost, output.cost_jobs, output.cost_jreg = (ctx.dtype(x) for x in cost3)
    return output.cost


def forward_b(setup, mesh, input_data, parameters, parameters_b, parameters_bgd, states, states_b, states_bgd, output,
              precision="f32"):
    """BASE_FORWARD_B (forward/forward_db.f90:10648-10936) with cost_b = 1."""
    ctx = _Ctx(precision)
    P = _problem(ctx, setup, mesh, input_data)
    shp = (mesh.nrow, mesh.ncol)
    par, par_bgd = _stack(parameters, PARAM_NAMES, shp, ctx.dtype), _stack(parameters_bgd, PARAM_NAMES, shp, ctx.dtype)
    st, st_bgd = _stack(states, STATE_NAMES, shp, ctx.dtype), _stack(states_bgd, STATE_NAMES, shp, ctx.dtype)
    par_b = np.zeros_like(par)
    st_b = np.zeros_like(st)
    qsim = np.zeros((max(mesh.ng, 1), setup._ntime_step), dtype=ctx.dtype, order="F")
    cost3 = np.zeros(3, dtype=ctx.dtype)
    rc = ctx.fn("oracle_forward_b")(C.byref(P), _ptr(ctx, par), _ptr(ctx, par_b), _ptr(ctx, par_bgd), _ptr(ctx, st),
                                    _ptr(ctx, st_b), _ptr(ctx, st_bgd), _ptr(ctx, qsim), _ptr(ctx, cost3))
    assert rc == 0
    _unstack(par, parameters, PARAM_NAMES)
    _unstack(st, states, STATE_NAMES)
    for k, n in enumerate(PARAM_NAMES):
        setattr(parameters_b, n, np.asfortranarray(par_b[..., k]))
    for k, n in enumerate(STATE_NAMES):
        setattr(states_b, n, np.asfortranarray(st_b[..., k]))
    if mesh.ng > 0:
        output.qsim = qsim
    output.cost, output.cost_jobs, output.cost_jreg = (ctx.dtype(x) for x in cost3)
    return output.cost


def hyper_forward(setup, mesh, input_data, parameters, hyper_parameters, states, hyper_states, output, precision="f32"):
    """base_hyper_forward (forward/forward.f90:82-157)."""
    ctx = _Ctx(precision)
    P = _problem(ctx, setup, mesh, input_data)
    shp = (mesh.nrow, mesh.ncol)
    nh = setup._optimize.nhyper
    par, st = _stack(parameters, PARAM_NAMES, shp, ctx.dtype), _stack(states, STATE_NAMES, shp, ctx.dtype)
    hp, hs = _stack(hyper_parameters, PARAM_NAMES, (nh, 1), ctx.dtype), _stack(hyper_states, STATE_NAMES, (nh, 1), ctx.dtype)
    qsim = np.zeros((max(mesh.ng, 1), setup._ntime_step), dtype=ctx.dtype, order="F")
    fst = np.zeros(shp + (GNS,), dtype=ctx.dtype, order="F")
    cost3 = np.zeros(3, dtype=ctx.dtype)
    rc = ctx.fn("oracle_hyper_forward")(C.byref(P), _ptr(ctx, par), _ptr(ctx, hp), _ptr(ctx, st), _ptr(ctx, hs),
                                        _ptr(ctx, qsim), _ptr(ctx, fst), _ptr(ctx, cost3))
    assert rc == 0
    _unstack(par, parameters, PARAM_NAMES)
    _unstack(st, states, STATE_NAMES)
    _unstack(fst, output.fstates, STATE_NAMES)
    if mesh.ng > 0:
        output.qsim = qsim
    output.cost, output.cost_jobs = ctx.dtype(cost3[0]), ctx.dtype(cost3[1])
    return output.cost


def hyper_forward_b(setup, mesh, input_data, parameters, hyper_parameters, hyper_parameters_b, states, hyper_states,
                    hyper_states_b, output, precision="f32"):
    """BASE_HYPER_FORWARD_B (forward/forward_db.f90:11231-11554) with cost_b = 1."""
    ctx = _Ctx(precision)
    P = _problem(ctx, setup, mesh, input_data)
    shp = (mesh.nrow, mesh.ncol)
    nh = setup._optimize.nhyper
    par, st = _stack(parameters, PARAM_NAMES, shp, ctx.dtype), _stack(states, STATE_NAMES, shp, ctx.dtype)
    hp, hs = _stack(hyper_parameters, PARAM_NAMES, (nh, 1), ctx.dtype), _stack(hyper_states, STATE_NAMES, (nh, 1), ctx.dtype)
    hp_b, hs_b = np.zeros_like(hp), np.zeros_like(hs)
    qsim = np.zeros((max(mesh.ng, 1), setup._ntime_step), dtype=ctx.dtype, order="F")
    cost3 = np.zeros(3, dtype=ctx.dtype)
    rc = ctx.fn("oracle_hyper_forward_b")(C.byref(P), _ptr(ctx, par), _ptr(ctx, hp), _ptr(ctx, hp_b), _ptr(ctx, st),
                                          _ptr(ctx, hs), _ptr(ctx, hs_b), _ptr(ctx, qsim), _ptr(ctx, cost3))
    assert rc == 0
    _unstack(par, parameters, PARAM_NAMES)
    _unstack(st, states, STATE_NAMES)
    for k, n in enumerate(PARAM_NAMES):
        setattr(hyper_parameters_b, n, np.asfortranarray(hp_b[..., k]))
    for k, n in enumerate(STATE_NAMES):
        setattr(hyper_states_b, n, np.asfortranarray(hs_b[..., k]))
    if mesh.ng > 0:
        output.qsim = qsim
    output.cost, output.cost_jobs = ctx.dtype(cost3[0]), ctx.dtype(cost3[1])
    return output.cost


def compute_multiple_run(setup, mesh, input_data, parameters, states, output, sample, ind_parameters_states, res_cost,
                         res_qsim, precision="f32", nthreads=None):
    """compute_multiple_run (routine/mw_multiple_run.f90:68-119); OpenMP over members like the reference."""
    ctx = _Ctx(precision)
    P = _problem(ctx, setup, mesh, input_data)
    shp = (mesh.nrow, mesh.ncol)
    par, st = _stack(parameters, PARAM_NAMES, shp, ctx.dtype), _stack(states, STATE_NAMES, shp, ctx.dtype)
    smp = np.asfortranarray(sample, dtype=ctx.dtype)
    ind = np.ascontiguousarray(ind_parameters_states, dtype=np.int32)
    nvar, ns = smp.shape
    rc_ = np.zeros(ns, dtype=ctx.dtype)
    want_q = res_qsim is not None and res_qsim.size > 0
    rq = np.zeros((mesh.ng, setup._ntime_step, ns), dtype=ctx.dtype, order="F") if want_q else None
    nthreads = int(nthreads or setup._ncpu or 1)
    rc = ctx.fn("oracle_multiple_run")(C.byref(P), _ptr(ctx, par), _ptr(ctx, st), _ptr(ctx, smp),
                                       ind.ctypes.data_as(C.POINTER(C.c_int32)), nvar, ns, _ptr(ctx, rc_), _ptr(ctx, rq),
                                       nthreads)
    assert rc == 0
    res_cost[...] = rc_
    if want_q:
        res_qsim[...] = rq


def adjust_interception_store(setup, mesh, input_data, parameters, nday, day_index, precision="f32"):
    """adjust_interception_store (routine/mw_interception_store.f90:19-160): sets parameters.ci on the computed cells."""
    ctx = _Ctx(precision)
    P = _problem(ctx, setup, mesh, input_data)
    ci = np.asfortranarray(parameters.ci, dtype=ctx.dtype).copy(order="F")
    di = np.ascontiguousarray(day_index, dtype=np.int32)
    rc = ctx.fn("oracle_adjust_interception_store")(C.byref(P), int(nday), di.ctypes.data_as(C.POINTER(C.c_int32)), _ptr(ctx, ci))
    assert rc == 0
    parameters.ci[...] = ci


def nse(x, y, precision="f32"):
    ctx = _Ctx(precision)
    f = ctx.fn("oracle_nse")
    f.restype = ctx.creal
    return f(ctx.r(x), ctx.r(y), len(x))


def kge(x, y, precision="f32"):
    ctx = _Ctx(precision)
    f = ctx.fn("oracle_kge")
    f.restype = ctx.creal
    return f(ctx.r(x), ctx.r(y), len(x))


def compute_jobs(setup, mesh, input_data, qsim, precision="f32", adjoint=False):
    """compute_jobs (optimize/mwd_cost.f90:37-156) on a given hydrograph array (ng, T)."""
    ctx = _Ctx(precision)
    P = _problem(ctx, setup, mesh, input_data)
    f = ctx.fn("oracle_compute_jobs")
    f.restype = ctx.creal
    qb = np.zeros((mesh.ng, setup._ntime_step), dtype=ctx.dtype, order="F") if adjoint else None
    j = f(C.byref(P), ctx.r(qsim), _ptr(ctx, qb))
    return (ctx.dtype(j), qb) if adjoint else ctx.dtype(j)
