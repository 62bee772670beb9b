"""Python-level callers of the solver entry points: what ``smash.Model.run / optimize / multiple_run`` do between
their argument standardisation and the wrapped Fortran (smash/core/model.py:430-900,
smash/core/simulation/_optimize.py:25-466, _standardize.py:118-935, multiple_run.py:150-220).

``model`` is any object carrying the six derived types ``setup, mesh, input_data, parameters, states, output``
(``smash_b200.solver._derived_types``) and a ``copy()`` method.  Reading rasters / building the mesh is out of scope
(SURVEY.md section 8f next-4).  Everything here is host logic; the simulations run on the GPU through
``smash_b200.solver`` (``solver=`` lets the test-suite drive the same logic with the CPU oracle as a checker).
"""
from __future__ import annotations

import numpy as np

from .solver import _mw_forward, _mw_multiple_run, _mw_optimize
from .solver._derived_types import GPARAMETERS_NAME, GSTATES_NAME, Optimize_SetupDT

# smash/core/_constant.py:15-45
STRUCTURE_PARAMETERS = {
    "gr-a": ["cp", "cft", "exc", "lr"],
    "gr-b": ["cp", "cft", "exc", "lr"],
    "gr-c": ["cp", "cft", "cst", "exc", "lr"],
    "gr-d": ["cp", "cft", "lr"],
    "vic-a": ["b", "cusl1", "cusl2", "clsl", "ks", "ds", "dsm", "ws", "lr"],
}
STRUCTURE_STATES = {
    "gr-a": ["hp", "hft", "hlr"],
    "gr-b": ["hi", "hp", "hft", "hlr"],
    "gr-c": ["hi", "hp", "hft", "hst", "hlr"],
    "gr-d": ["hp", "hft", "hlr"],
    "vic-a": ["husl1", "husl2", "hlsl"],
}
MAPPING = ("uniform", "distributed", "hyper-linear", "hyper-polynomial")
JOBS_FUN = ("nse", "kge", "kge2", "se", "rmse", "logarithmic")


def _as_array(x, what):
    if isinstance(x, str):
        return np.array(x, ndmin=1)
    if isinstance(x, (list, tuple, np.ndarray)):
        return np.array(x)
    raise TypeError(f"{what} argument must be str or list-like object")


def _gauge_weights(mesh, input_data, gauge, wgauge, ost_step):
    """_standardize_gauge + _standardize_wgauge (_standardize.py:298-398)."""
    code = np.asarray(mesh.code)
    if isinstance(gauge, str):
        if gauge == "all":
            names = code.copy()
        elif gauge == "downstream":
            names = np.array(code[int(np.argmax(mesh.area))], ndmin=1)
        elif gauge in code:
            names = np.array(gauge, ndmin=1)
        else:
            raise ValueError(f"Unknown gauge alias or code '{gauge}'. Choices: ['all', 'downstream'] or {code}")
    else:
        names = _as_array(gauge, "gauge")
    keep = []
    for name in names:
        if name not in code:
            raise ValueError(f"Unknown gauge code '{name}'. Choices: {code}")
        g = int(np.flatnonzero(code == name)[0])
        if not np.all(input_data.qobs[g, ost_step:] < 0):
            keep.append(name)
    if not keep:
        raise ValueError(f"No available observed discharge for optimization at gauge(s) {names}")
    ind = np.isin(code, keep)
    w = np.zeros(code.size, dtype=np.float32)
    if isinstance(wgauge, str):
        if wgauge == "mean":
            w[ind] = 1 / len(keep)
        elif wgauge == "median":
            w[ind] = -50
        elif wgauge == "area":
            w[ind] = mesh.area[ind] / np.sum(mesh.area[ind])
        elif wgauge == "minv_area":
            w[ind] = (1 / mesh.area[ind]) / np.sum(1 / mesh.area[ind])
        else:
            raise ValueError(f"Unknown wgauge alias '{wgauge}'. Choices: ['mean', 'median', 'area', 'minv_area']")
    else:
        wg = np.asarray(wgauge, dtype=np.float32)
        if wg.size != len(keep):
            raise ValueError(f"Inconsistent size between gauge ({len(keep)}) and wgauge ({wg.size})")
        if np.any(wg < 0):
            raise ValueError(f"wgauge can not receive negative values ({wg})")
        w[ind] = wg
    return w


def _setup_optimize(model, mapping, algorithm, control_vector, bounds, jobs_fun, wjobs_fun, gauge, wgauge, ost_step,
                    verbose):
    """reset_optimize_setup + the _standardize_* chain + the field-by-field copy at the top of _optimize_sbs /
    _optimize_lbfgsb (_optimize.py:56-92, 178-229)."""
    setup, mesh = model.setup, model.mesh
    if mapping not in MAPPING:
        raise ValueError(f"Unknown mapping '{mapping}'. Choices: {list(MAPPING)}")
    if algorithm is None:
        algorithm = "sbs" if mapping == "uniform" else "l-bfgs-b"
    if algorithm not in ("sbs", "l-bfgs-b"):
        raise ValueError(f"algorithm '{algorithm}' is not provided by smash_b200 (choices: 'sbs', 'l-bfgs-b')")
    if mapping == "uniform" and algorithm == "l-bfgs-b" or mapping != "uniform" and algorithm == "sbs":
        raise ValueError(f"algorithm '{algorithm}' can not be combined with mapping '{mapping}'")
    structure = str(setup.structure)
    if control_vector is None:
        cv = np.array(STRUCTURE_PARAMETERS[structure])
    else:
        cv = _as_array(control_vector, "control_vector")
        avail = STRUCTURE_PARAMETERS[structure] + STRUCTURE_STATES[structure]
        for name in cv:
            if name not in avail:
                raise ValueError(f"Unknown parameter or state '{name}' for structure '{structure}' in control_vector. "
                                 f"Choices: {avail}")
    jf = _as_array(jobs_fun, "jobs_fun")
    if "kge" in jf and algorithm == "l-bfgs-b":
        raise ValueError("'kge' objective function can not be used with 'l-bfgs-b' algorithm (non convex function)")
    for name in jf:
        if name not in JOBS_FUN:
            raise ValueError(f"objective function '{name}' is not provided by smash_b200. Choices: {list(JOBS_FUN)}")
    wjf = np.ones(jf.size) / jf.size if wjobs_fun is None else np.asarray(wjobs_fun, dtype=np.float64)
    if wjf.size != jf.size:
        raise ValueError(f"Inconsistent size between jobs_fun ({jf.size}) and wjobs_fun ({wjf.size})")

    o = Optimize_SetupDT(setup._ntime_step, setup._nd, mesh.ng, mapping, jf.size, 0)
    o.algorithm = algorithm
    o.verbose = bool(verbose)
    o.jobs_fun = jf.astype("U20")
    o.wjobs_fun = wjf.astype(np.float32)
    o.optimize_start_step = int(ost_step) + 1
    bnd = np.empty((cv.size, 2), dtype=np.float32)
    for i, name in enumerate(cv):
        if name in GPARAMETERS_NAME:
            k = GPARAMETERS_NAME.index(name)
            bnd[i] = o.lb_parameters[k], o.ub_parameters[k]
        else:
            k = GSTATES_NAME.index(name)
            bnd[i] = o.lb_states[k], o.ub_states[k]
    for name, b in (bounds or {}).items():
        if name not in cv:
            continue
        if not isinstance(b, (np.ndarray, list, tuple)) or len(b) != 2:
            raise ValueError(f"bounds values for '{name}' must be list-like object of length 2")
        i = int(np.flatnonzero(cv == name)[0])
        if b[0] is not None:
            bnd[i, 0] = b[0]
        if b[1] is not None:
            bnd[i, 1] = b[1]
        if bnd[i, 0] >= bnd[i, 1]:
            raise ValueError(f"bounds values for '{name}' are invalid lower bound ({bnd[i, 0]}) is greater than or "
                             f"equal to upper bound ({bnd[i, 1]})")
        field = getattr(model.parameters if name in GPARAMETERS_NAME else model.states, name)
        if np.any(field + 1e-3 < bnd[i, 0]) or np.any(field - 1e-3 > bnd[i, 1]):
            raise ValueError(f"bounds values for '{name}' are invalid, background [{np.min(field)} {np.max(field)}] is "
                             f"outside the bounds {bnd[i]}")
    for i, name in enumerate(cv):
        if name in GPARAMETERS_NAME:
            k = GPARAMETERS_NAME.index(name)
            o.optim_parameters[k] = 1
            o.lb_parameters[k], o.ub_parameters[k] = bnd[i]
        else:
            k = GSTATES_NAME.index(name)
            o.optim_states[k] = 1
            o.lb_states[k], o.ub_states[k] = bnd[i]
    if mesh.ng > 0:
        o.wgauge = _gauge_weights(mesh, model.input_data, gauge, wgauge, int(ost_step))
    setup._optimize = o
    return algorithm, cv


def _control_bounds(o, names):
    """Bounds of the control fields as standardised into ``setup._optimize`` (lb_/ub_parameters, lb_/ub_states)."""
    out = []
    for c in names:
        if c in GPARAMETERS_NAME:
            k = GPARAMETERS_NAME.index(c)
            out.append([o.lb_parameters[k].item(), o.ub_parameters[k].item()])
        else:
            k = GSTATES_NAME.index(c)
            out.append([o.lb_states[k].item(), o.ub_states[k].item()])
    return out


def _compute_wjreg_range(wjreg_opt, nb_wjreg_lcurve):
    """_optimize.py:911-947"""
    lw = np.log10(wjreg_opt)
    nb = nb_wjreg_lcurve - 6
    base = np.array(10 ** np.arange(lw - 0.66, lw + 0.67, 0.33), dtype=np.float32)
    if nb > 0:
        lo = lw - 0.66 - (nb - np.ceil(nb / 2.0))
        hi = lw + 0.66 + 1.0 + (nb - np.floor(nb / 2.0))
        lower = np.array(10 ** np.arange(lo, lw - 0.66), dtype=np.float32)
        upper = np.array(10 ** np.arange(lw + 0.66 + 1.0, hi), dtype=np.float32)
        return np.hstack((lower, base, upper))
    return base


def _compute_best_lcurve_weight(jobs, jreg, wjreg, jobs_min, jobs_max, jreg_min, jreg_max):
    """_optimize.py:950-999: largest distance of the normalised L-curve points below the diagonal."""
    best = None
    if jobs.size > 2 and (jreg_max - jreg_min) > 0.0 and (jobs_max - jobs_min) > 0.0:
        max_distance = 0.0
        distance = np.zeros(jobs.size, dtype=np.float32)
        for i in range(jobs.size):
            xr = (jobs_max - jobs[i]) / (jobs_max - jobs_min)
            yr = (jreg[i] - jreg_min) / (jreg_max - jreg_min)
            if yr < xr:
                if jobs[i] < jobs_max:
                    hyp = (xr ** 2.0 + yr ** 2.0) ** 0.5
                    distance[i] = hyp * np.sin(np.pi * 0.25 - np.arccos(xr / hyp))
                else:
                    distance[i] = 0.0
                if distance[i] >= max_distance:
                    max_distance = distance[i]
                    best = wjreg[i]
            else:
                distance[i] = np.nan
        return distance, best
    return np.empty(0), None


def run(model, inplace=False, solver=None):
    """Model.run (model.py:430-530): one forward run, nse at the downstream gauge as the cost."""
    sv = solver or _mw_forward
    inst = model if inplace else model.copy()
    _setup_optimize(inst, "uniform", "sbs", None, None, "nse", None, "downstream", "mean", 0, False)
    inst.setup._optimize.mapping = "..."
    sv.forward(inst.setup, inst.mesh, inst.input_data, inst.parameters, inst.parameters.copy(), inst.states,
               inst.states.copy(), inst.output)
    return None if inplace else inst


def optimize(model, mapping="uniform", algorithm=None, control_vector=None, bounds=None, jobs_fun="nse", wjobs_fun=None,
             gauge="downstream", wgauge="mean", ost_step=0, options=None, verbose=False, inplace=False, solver=None):
    """Model.optimize (model.py:628-900).  ``ost_step`` is the 0-based first time step of the cost window (the
    reference takes a timestamp ``ost`` and converts it, _optimize.py:82-86).  Returns the optimised copy (and the
    L-curve dictionary when ``options['return_lcurve']``)."""
    inst = model if inplace else model.copy()
    algorithm, cv = _setup_optimize(inst, mapping, algorithm, control_vector, bounds, jobs_fun, wjobs_fun, gauge, wgauge,
                                    ost_step, verbose)
    opts = dict(options or {})
    o = inst.setup._optimize
    o.maxiter = int(opts.pop("maxiter", 100))
    a = (inst.setup, inst.mesh, inst.input_data)
    res = None
    if algorithm == "sbs":
        if opts:
            raise KeyError("Unknown algorithm options: '%s'" % ", ".join(map(str, opts)))
        _mw_optimize.optimize_sbs(*a, inst.parameters, inst.states, inst.output, solver=solver)
    else:
        jreg_fun = opts.pop("jreg_fun", None)
        wjreg = opts.pop("wjreg", 0.01)
        wjreg_fun = opts.pop("wjreg_fun", None)
        auto_wjreg = opts.pop("auto_wjreg", None)
        nb_wjreg_lcurve = int(opts.pop("nb_wjreg_lcurve", 6))
        return_lcurve = bool(opts.pop("return_lcurve", False))
        if opts:
            raise KeyError("Unknown algorithm options: '%s'" % ", ".join(map(str, opts)))
        if jreg_fun is not None:                                          # _standardize_jreg_fun / wjreg_fun
            jr = _as_array(jreg_fun, "jreg_fun")
            for name in jr:
                if name not in ("prior", "smoothing", "hard_smoothing"):
                    raise ValueError(f"Unknown regularization function '{name}'. Choices: ['prior', 'smoothing', 'hard_smoothing']")
            if mapping.startswith("hyper"):
                raise ValueError("Regularization function(s) can not be used with hyper mappings")
            o.njr = jr.size
            o.jreg_fun = jr.astype("U20")
            o.wjreg_fun = np.ones(jr.size, np.float32) if wjreg_fun is None else np.asarray(wjreg_fun, np.float32)
            o.wjreg = np.float32(wjreg)
        else:
            o.wjreg = np.float32(0.0)

        def cycle():
            _mw_optimize.optimize_lbfgsb(*a, inst.parameters, inst.states, inst.output, solver=solver)

        def restore(par, sta):
            inst.parameters, inst.states = par.copy(), sta.copy()

        if mapping.startswith("hyper"):
            _mw_optimize.optimize_hyper_lbfgsb(*a, inst.parameters, inst.states, inst.output, solver=solver)
        elif auto_wjreg == "fast":                                        # _optimize.py:258-296
            par_bgd, sta_bgd = inst.parameters.copy(), inst.states.copy()
            o.wjreg = np.float32(0.0)
            cycle()
            o.wjreg = np.float32((inst.output._cost_jobs_initial - inst.output.cost_jobs) / inst.output.cost_jreg)
            restore(par_bgd, sta_bgd)
            cycle()
        elif auto_wjreg == "lcurve":                                      # _optimize.py:298-451
            par_bgd, sta_bgd = inst.parameters.copy(), inst.states.copy()
            o.wjreg = np.float32(0.0)
            cycle()
            jobs_min, jobs_max = inst.output.cost_jobs, inst.output._cost_jobs_initial
            jreg_min, jreg_max = 0.0, inst.output.cost_jreg
            if (jobs_min / jobs_max) < 0.95 and (jreg_max - jreg_min) > 0.0:
                wjreg_opt = (jobs_max - jobs_min) / jreg_max
                wjreg_range = _compute_wjreg_range(wjreg_opt, nb_wjreg_lcurve)
            else:
                wjreg_opt, wjreg_range = 0.0, np.empty(0)
            n = wjreg_range.size + 1
            cost_arr, jobs_arr, jreg_arr, wj_arr = (np.zeros(n, np.float32) for _ in range(4))
            cost_arr[0], jobs_arr[0], jreg_arr[0], wj_arr[0] = inst.output.cost, inst.output.cost_jobs, \
                inst.output.cost_jreg, o.wjreg
            for i, wj in enumerate(wjreg_range):
                o.wjreg = np.float32(wj)
                restore(par_bgd, sta_bgd)
                cycle()
                cost_arr[i + 1], jobs_arr[i + 1], jreg_arr[i + 1], wj_arr[i + 1] = inst.output.cost, \
                    inst.output.cost_jobs, inst.output.cost_jreg, o.wjreg
            jobs_min, jobs_max = np.min(jobs_arr), np.max(jobs_arr)
            jreg_max, jreg_min = np.max(jreg_arr), np.min(jreg_arr)
            distance, best = _compute_best_lcurve_weight(jobs_arr, jreg_arr, wj_arr, jobs_min, jobs_max, jreg_min, jreg_max)
            lcurve = {"cost_jobs_initial": jobs_max, "cost_jreg_initial": jreg_min, "wjreg_lcurve_opt": best,
                      "wjreg_fast": wjreg_opt, "wjreg": wj_arr, "distance": distance, "cost": cost_arr,
                      "cost_jobs": jobs_arr, "cost_jreg": jreg_arr}
            restore(par_bgd, sta_bgd)
            if best is not None:
                o.wjreg = np.float32(best)
                cycle()
            else:
                sv = solver or _mw_forward
                sv.forward(*a, inst.parameters, inst.parameters.copy(), inst.states, inst.states.copy(), inst.output)
            if return_lcurve:
                res = lcurve
        elif auto_wjreg is None:
            cycle()
        else:
            raise ValueError(f"Unknown auto_wjreg '{auto_wjreg}'. Choices: ['fast', 'lcurve']")
    if res is not None:
        return res if inplace else (inst, res)
    return None if inplace else inst


def multiple_run(model, sample, names, jobs_fun="nse", wjobs_fun=None, gauge="downstream", wgauge="mean", ost_step=0,
                 return_qsim=False, solver=None):
    """Model.multiple_run (model.py:903-1040, multiple_run.py:150-220).  ``sample`` is (n, nvar) with one column per
    entry of ``names`` (what ``SampleResult.to_numpy(axis=-1)`` gives); returns ``cost`` (n,) and, if asked,
    ``qsim`` (ng, T, n)."""
    sv = solver or _mw_multiple_run
    inst = model.copy()
    _setup_optimize(inst, "uniform", "sbs", None, None, jobs_fun, wjobs_fun, gauge, wgauge, ost_step, False)
    inst.setup._optimize.mapping = "..."
    ind = []
    for name in names:
        if name in GPARAMETERS_NAME:
            ind.append(GPARAMETERS_NAME.index(name) + 1)
        elif name in GSTATES_NAME:
            ind.append(len(GPARAMETERS_NAME) + GSTATES_NAME.index(name) + 1)
        else:
            raise ValueError(f"Unknown parameter or state '{name}'")
    smp = np.asfortranarray(np.asarray(sample, dtype=np.float32).T)
    ns = smp.shape[1]
    cost = np.zeros(ns, dtype=np.float32)
    qsim = np.zeros((inst.mesh.ng, inst.setup._ntime_step, ns) if return_qsim else (0, 0, 0), dtype=np.float32, order="F")
    sv.compute_multiple_run(inst.setup, inst.mesh, inst.input_data, inst.parameters, inst.states, inst.output, smp,
                            np.array(ind, dtype=np.int32), cost, qsim)
    return (cost, qsim) if return_qsim else cost


# ------------------------------------------------------------------------------------------ samples / Bayesian estimation
class SampleResult(dict):
    """smash/core/generate_samples.py:22-270: the generated samples, one array per name plus ``_<name>`` (the density of
    each drawn value), ``generator``, ``n_sample`` and ``_problem``."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    __setattr__ = dict.__setitem__

    def to_numpy(self, axis=0):
        return np.stack([self[k] for k in self._problem["names"]], axis=axis)

    def slice(self, n, start=0):
        if start < 0 or n > self.n_sample or start >= n:
            raise ValueError("invalid slice")
        d = {k: (v[start:n] if isinstance(v, np.ndarray) and v.shape == (self.n_sample,) else v) for k, v in self.items()}
        d["n_sample"] = n - start
        return SampleResult(d)

    def iterslice(self, by=1):
        for start in range(0, self.n_sample, by):
            yield self.slice(min(start + by, self.n_sample), start)


def get_bound_constraints(model, states=False):
    """Model.get_bound_constraints / _get_bound_constraints (generate_samples.py:389-414)."""
    o = model.setup._optimize
    names = (STRUCTURE_STATES if states else STRUCTURE_PARAMETERS)[str(model.setup.structure)]
    bounds = []
    for name in names:
        if name in GSTATES_NAME:
            k = GSTATES_NAME.index(name)
            bounds.append([o.lb_states[k].item(), o.ub_states[k].item()])
        else:
            k = GPARAMETERS_NAME.index(name)
            bounds.append([o.lb_parameters[k].item(), o.ub_parameters[k].item()])
    return {"num_vars": len(names), "names": list(names), "bounds": bounds}


def generate_samples(problem, generator="uniform", n=1000, random_state=None, mean=None, coef_std=None):
    """smash.generate_samples (generate_samples.py:273-383): one ``np.random`` draw per name, in the order of
    ``problem['names']``, after seeding the legacy global generator."""
    generator = str(generator).lower()
    if generator not in ("uniform", "normal", "gaussian"):
        raise ValueError(f"Unknown generator '{generator}': Choices: ['uniform', 'normal', 'gaussian']")
    ret = {"generator": generator, "n_sample": n, "_problem": dict(problem)}
    if random_state is not None:
        np.random.seed(random_state)
    for i, p in enumerate(problem["names"]):
        low, upp = problem["bounds"][i]
        if generator == "uniform":
            ret[p] = np.random.uniform(low, upp, n)
            ret["_" + p] = np.ones(n) / (upp - low)
        else:
            from scipy.stats import truncnorm
            mu = (low + upp) / 2 if not mean or mean.get(p) is None else mean[p]
            sd = (upp - low) / (3 if coef_std is None else coef_std)
            tn = truncnorm((low - mu) / sd, (upp - mu) / sd, loc=mu, scale=sd)
            ret[p] = tn.rvs(size=n)
            ret["_" + p] = tn.pdf(ret[p])
    return SampleResult(ret)


class BayesResult(dict):
    """smash/core/simulation/bayes_optimize.py:23-76"""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    __setattr__ = dict.__setitem__


def _compute_mean_U(U, J, rho, alpha, mask):
    """bayes_optimize.py:433-454.  ``alpha`` enters as a Python float so that the float32 cost array keeps its type, as
    under the NumPy the reference's golden file was made with."""
    L = np.exp(-(2 ** float(alpha)) * (J / min(J) - 1) ** 2)
    Lrho = L * rho
    C = np.sum(Lrho, axis=2)
    U_alp = 1 / C * np.sum(U * Lrho, axis=2)
    varU = 1 / C * np.sum((U - U_alp[..., np.newaxis]) ** 2 * Lrho, axis=2)
    varU = np.mean(varU[mask])
    Uinf = np.mean(U, axis=2)
    D_alp = np.mean(np.square(U_alp - Uinf)[mask]) / varU
    return U_alp, varU, D_alp


def _bayes(model, sample, alpha, n, random_state, de_bw_method, de_weights, mapping, algorithm, control_vector, bounds, jobs_fun,
           wjobs_fun, gauge, wgauge, ost_step, options, inplace, return_br, solver, mr_solver):
    """_bayes_computation (bayes_optimize.py:79-185).  ``algorithm is None``: Bayesian estimation -- the direct simulations
    of the sample are ONE ``compute_multiple_run`` call (all members in one GPU launch); otherwise every member is the
    first guess of an optimisation."""
    inst = model if inplace else model.copy()
    sv = solver or _mw_forward
    names = None
    if algorithm is None:
        _setup_optimize(inst, "uniform", "sbs", None, None, jobs_fun, wjobs_fun, gauge, wgauge, ost_step, False)
        inst.setup._optimize.mapping = "..."
        if sample is None:
            sample = generate_samples(get_bound_constraints(inst, states=False), "uniform", n, random_state)
        elif not isinstance(sample, SampleResult):
            raise TypeError("sample must be a SampleResult object or None")
    else:
        algorithm, cv = _setup_optimize(inst, mapping, algorithm, control_vector, bounds, jobs_fun, wjobs_fun, gauge, wgauge,
                                        ost_step, False)
        if sample is None:
            bnd = _control_bounds(inst.setup._optimize, cv)
            sample = generate_samples({"num_vars": len(cv), "names": list(cv), "bounds": bnd}, "uniform", n, random_state)
        elif set(sample._problem["names"]) != set(cv):
            raise ValueError(f"Problem names ({sample._problem['names']}) and control vectors ({cv}) must have the same elements")
    if isinstance(alpha, (range, np.ndarray, tuple)):
        alpha = list(alpha)
    elif not isinstance(alpha, (int, float, list)):
        raise TypeError("alpha must be numerical or list-like object")
    names = list(sample._problem["names"])
    ns = sample.n_sample
    shape = (inst.mesh.nrow, inst.mesh.ncol)
    a = (inst.setup, inst.mesh, inst.input_data)

    def field(name):
        return getattr(inst.parameters if name in GPARAMETERS_NAME else inst.states, name)

    # ---- _multi_simu (:308-390)
    prior, density = {}, {}
    if algorithm is None:
        smp = np.asfortranarray(np.stack([np.asarray(sample[p], dtype=np.float32) for p in names]))
        ind = [GPARAMETERS_NAME.index(p) + 1 if p in GPARAMETERS_NAME else len(GPARAMETERS_NAME) + GSTATES_NAME.index(p) + 1
               for p in names]
        cost = np.zeros(ns, dtype=np.float32)
        (mr_solver or _mw_multiple_run).compute_multiple_run(*a, inst.parameters, inst.states, inst.output, smp,
                                                             np.array(ind, dtype=np.int32), cost, np.zeros((0, 0, 0), np.float32))
        for k, p in enumerate(names):
            prior[p] = np.broadcast_to(smp[k][None, None, :], shape + (ns,)).astype(np.float32)
    else:
        opts = dict(options or {})
        inst.setup._optimize.maxiter = int(opts.pop("maxiter", 100))
        if opts:
            raise KeyError("Unknown algorithm options: '%s'" % ", ".join(map(str, opts)))
        cost = np.zeros(ns, dtype=np.float32)
        stack = {p: [] for p in names}
        for i in range(ns):
            run_i = inst.copy()
            run_i.setup._optimize = inst.setup._optimize.copy()
            for p in names:
                getattr(run_i.parameters if p in GPARAMETERS_NAME else run_i.states, p)[...] = sample[p][i]
            drv = _mw_optimize.optimize_sbs if algorithm == "sbs" else (
                _mw_optimize.optimize_hyper_lbfgsb if mapping.startswith("hyper") else _mw_optimize.optimize_lbfgsb)
            drv(run_i.setup, run_i.mesh, run_i.input_data, run_i.parameters, run_i.states, run_i.output, solver=solver)
            cost[i] = run_i.output.cost
            for p in names:
                stack[p].append(np.copy(getattr(run_i.parameters if p in GPARAMETERS_NAME else run_i.states, p)))
        for p in names:
            prior[p] = np.dstack(stack[p])
    prior["cost"] = cost
    ret_data = {"cost": np.array(cost)}
    for p in names:
        ret_data[p] = sample[p]
        density[p] = np.ones(prior[p].shape)

    # ---- _compute_density (:392-430)
    mask = np.where(np.asarray(inst.mesh.active_cell) == 1)
    x, y = mask
    for p in names:
        if algorithm == "l-bfgs-b":
            from scipy.stats import gaussian_kde
            for xi, yi in zip(x, y):
                density[p][xi, yi] = gaussian_kde(prior[p][xi, yi], bw_method=de_bw_method, weights=de_weights)(prior[p][xi, yi])
        elif isinstance(algorithm, str):
            from scipy.stats import gaussian_kde
            u_dis = np.mean(prior[p][mask], axis=0)
            density[p][x, y] = gaussian_kde(u_dis, bw_method=de_bw_method, weights=de_weights)(u_dis)
        else:
            density[p][x, y] = sample["_" + p]

    # ---- _compute_param / _lcurve_compute_param (:457-566)
    def compute_param(al):
        D, var = [], {}
        for p in names:
            u, v, d = _compute_mean_U(prior[p], prior["cost"], density[p], al, mask)
            setattr(inst.parameters if p in GPARAMETERS_NAME else inst.states, p, np.asfortranarray(u, dtype=np.float32))
            D.append(d)
            var[p] = v
        sv.forward(*a, inst.parameters, inst.parameters.copy(), inst.states, inst.states.copy(), inst.output)
        return inst.output.cost, np.mean(D), var

    lcurve = {}
    if isinstance(alpha, list):
        costs, Ds, vars_ = [], [], []
        for al in alpha:
            c, d, v = compute_param(al)
            costs.append(c); Ds.append(d); vars_.append(v)
        cs = (costs - np.min(costs)) / (np.max(costs) - np.min(costs))
        ds = (Ds - np.min(Ds)) / (np.max(Ds) - np.min(Ds))
        alpha_opt = alpha[int(np.argmin(np.square(cs) + np.square(ds)))]
        compute_param(alpha_opt)
        lcurve = {"alpha": alpha, "alpha_opt": alpha_opt, "mahal_dist": Ds, "cost": costs, "var": vars_}
    else:
        compute_param(alpha)
    br = BayesResult(data=ret_data, lcurve=lcurve)
    if return_br:
        return br if inplace else (inst, br)
    return None if inplace else inst


def bayes_estimate(model, sample=None, alpha=4, n=1000, random_state=None, jobs_fun="nse", wjobs_fun=None, gauge="downstream",
                   wgauge="mean", ost_step=0, inplace=False, return_br=False, solver=None, mr_solver=None):
    """Model.bayes_estimate (model.py:849-1010)."""
    return _bayes(model, sample, alpha, n, random_state, None, None, None, None, None, None, jobs_fun, wjobs_fun, gauge, wgauge,
                  ost_step, None, inplace, return_br, solver, mr_solver)


def bayes_optimize(model, sample=None, alpha=4, n=1000, random_state=None, de_bw_method=None, de_weights=None, mapping="uniform",
                   algorithm=None, control_vector=None, bounds=None, jobs_fun="nse", wjobs_fun=None, gauge="downstream",
                   wgauge="mean", ost_step=0, options=None, inplace=False, return_br=False, solver=None):
    """Model.bayes_optimize (model.py:1012-1200)."""
    if algorithm is None:
        algorithm = "sbs" if mapping == "uniform" else "l-bfgs-b"
    return _bayes(model, sample, alpha, n, random_state, de_bw_method, de_weights, mapping, algorithm, control_vector, bounds,
                  jobs_fun, wjobs_fun, gauge, wgauge, ost_step, options, inplace, return_br, solver, None)


def ann_optimize(model, net=None, optimizer="adam", learning_rate=0.003, control_vector=None, bounds=None, jobs_fun="nse",
                 wjobs_fun=None, gauge="downstream", wgauge="mean", ost_step=0, epochs=400, early_stopping=False,
                 random_state=None, verbose=False, inplace=False, return_net=False, solver=None, device_net=False):
    """Model.ann_optimize (model.py:1223-1420, simulation/_ann_optimize.py:20-254): a network maps the normalised
    descriptors of the active cells to the control fields; every epoch is one ``forward_b`` on the GPU.  ``device_net``: the
    network's Dense / Activation layers run on the GPU's tensor cores too (net.DeviceChain), for domain-sized inputs."""
    from .net import Net
    inst = model if inplace else model.copy()
    _, cv = _setup_optimize(inst, "uniform", "sbs", control_vector, bounds, jobs_fun, wjobs_fun, gauge, wgauge, ost_step, verbose)
    o = inst.setup._optimize
    o.mapping = "..."
    bnd = np.array(_control_bounds(o, cv), dtype=np.float32)
    parameters_bgd, states_bgd = inst.parameters.copy(), inst.states.copy()
    desc = inst.input_data.descriptor
    nd = int(inst.setup._nd)
    active, inactive = np.where(np.asarray(inst.mesh.active_cell) == 1), np.where(np.asarray(inst.mesh.active_cell) == 0)
    dmin = np.array([np.amin(desc[..., i]) for i in range(nd)], dtype=np.float32)
    dmax = np.array([np.amax(desc[..., i]) for i in range(nd)], dtype=np.float32)
    norm = np.empty_like(desc)
    for i in range(nd):                                                   # _normalize_descriptor (_optimize.py:801-811)
        norm[..., i] = (desc[..., i] - dmin[i]) / (dmax[i] - dmin[i])
    x_train, x_inactive = norm[active], norm[inactive]
    if net is None:                                                       # auto-graph (_ann_optimize.py:140-180)
        net = Net()
        n_neurons = round(np.sqrt(len(x_train) * nd) * 2 / 3)
        net.add("dense", {"input_shape": (nd,), "neurons": n_neurons, "kernel_initializer": "glorot_uniform"})
        net.add("activation", {"name": "relu"})
        net.add("dense", {"neurons": round(n_neurons / 2), "kernel_initializer": "glorot_uniform"})
        net.add("activation", {"name": "relu"})
        net.add("dense", {"neurons": cv.size, "kernel_initializer": "glorot_uniform"})
        net.add("activation", {"name": "sigmoid"})
        net.add("scale", {"bounds": bnd})
        net.compile(optimizer=optimizer, random_state=random_state, options={"learning_rate": learning_rate})
    elif not isinstance(net, Net):
        raise ValueError(f"Unknown network {net}")
    elif not net.layers:
        raise ValueError("The graph has not been set yet")
    else:
        if net.layers[0].input_shape[0] != nd:
            raise ValueError(f"Inconsistent value between the number of input layer ({net.layers[0].input_shape}) and the "
                             f"number of descriptors ({nd})")
        if net.layers[-1].output_shape()[0] != cv.size:
            raise ValueError(f"Inconsistent value between the number of output layer ({net.layers[-1].output_shape()}) and "
                             f"the number of control vectors ({cv.size})")
    net._fit_d2p(x_train, inst, cv, active, parameters_bgd, states_bgd, epochs, early_stopping, verbose, solver=solver,
                 device=device_net)
    y = net._predict(x_inactive)                                         # predicted maps on the inactive cells
    for i, name in enumerate(cv):
        getattr(inst.parameters if name in GPARAMETERS_NAME else inst.states, name)[inactive] = y[:, i]
    if return_net:
        return net if inplace else (inst, net)
    return None if inplace else inst
