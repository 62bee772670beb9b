"""Validates the oracle's restatement of the Tapenade adjoint (oracle_forward_b / oracle_hyper_forward_b) with the
checks the reference itself uses (optimize/mw_adjoint_test.f90): directional derivatives against central finite
differences of the forward code in double precision, and float32-vs-float64 agreement of the gradient."""
import numpy as np
import pytest

import cases
import oracle
from smash_b200.solver._derived_types import Hyper_ParametersDT, Hyper_StatesDT, ParametersDT, StatesDT

T = 168


def make(jobs=("nse",), **opt):
    m = cases.cance(T=T)
    cases.set_optimize(m.setup, m.mesh, jobs_fun=jobs, **opt)
    rng = np.random.default_rng(3)
    act = m.mesh.active_cell == 1
    for name, lo, hi in (("cp", 50, 600), ("cft", 50, 800), ("exc", -5, 5), ("lr", 1, 30)):
        getattr(m.parameters, name)[act] = rng.uniform(lo, hi, int(act.sum())).astype(np.float32)
    return m


def cost(m, precision="f64"):
    c = m.copy()
    return float(oracle.forward(c.setup, c.mesh, c.input_data, c.parameters, m.parameters.copy(), c.states, m.states.copy(),
                                c.output, precision=precision))


def grad(m, precision="f64"):
    c = m.copy()
    pb, sb = ParametersDT(m.mesh), StatesDT(m.mesh)
    oracle.forward_b(c.setup, c.mesh, c.input_data, c.parameters, pb, m.parameters.copy(), c.states, sb, m.states.copy(),
                     c.output, precision=precision)
    return pb, sb


@pytest.mark.parametrize("jobs", [("nse",), ("kge",), ("nse", "kge2")])
def test_adjoint_vs_finite_differences(jobs):
    m = make(jobs)
    pb, sb = grad(m)
    rng = np.random.default_rng(7)
    act = (m.mesh.active_cell == 1)
    for obj, g, name, eps in ((m.parameters, pb, "cp", 1e-4), (m.parameters, pb, "cft", 1e-4), (m.parameters, pb, "exc", 1e-5),
                              (m.parameters, pb, "lr", 1e-5), (m.states, sb, "hp", 1e-7), (m.states, sb, "hft", 1e-7)):
        direction = rng.standard_normal(act.shape) * act
        base = np.array(getattr(obj, name), dtype=np.float64)
        vals = []
        for sgn in (+1, -1):
            c = m.copy()
            tgt = c.parameters if obj is m.parameters else c.states
            setattr(tgt, name, np.asfortranarray(base + sgn * eps * direction))
            vals.append(float(oracle.forward(c.setup, c.mesh, c.input_data, c.parameters, m.parameters.copy(), c.states,
                                             m.states.copy(), c.output, precision="f64")))
        fd = (vals[0] - vals[1]) / (2 * eps)
        ad = float((np.asarray(getattr(g, name), np.float64) * direction).sum())
        assert np.isclose(fd, ad, rtol=2e-4, atol=1e-10), (name, fd, ad)
    assert not np.any(pb.cp[~act])


def test_adjoint_with_regularisation_and_normalisation():
    # prior + smoothing on normalised controls, denormalize_forward (the VDA setting, mw_optimize.f90:547-561)
    m = make(("nse",), jreg_fun=("prior", "smoothing"), wjreg_fun=[1.0, 0.5], wjreg=1e-3, denormalize_forward=True)
    o = m.setup._optimize
    o.optim_parameters[[1, 3, 6, 15]] = 1
    bgd = m.parameters.copy()
    for i, name in enumerate(oracle.PARAM_NAMES):       # controls live in [0,1]
        setattr(m.parameters, name, np.asfortranarray((getattr(m.parameters, name) - o.lb_parameters[i]) / (o.ub_parameters[i] - o.lb_parameters[i])))
        setattr(bgd, name, np.asfortranarray((getattr(bgd, name) - o.lb_parameters[i]) / (o.ub_parameters[i] - o.lb_parameters[i]) + 0.01))
    for i, name in enumerate(oracle.STATE_NAMES):
        setattr(m.states, name, np.asfortranarray((getattr(m.states, name) - o.lb_states[i]) / (o.ub_states[i] - o.lb_states[i])))

    def J(mm):
        c = mm.copy()
        return float(oracle.forward(c.setup, c.mesh, c.input_data, c.parameters, bgd, c.states, mm.states.copy(), c.output, precision="f64"))

    c = m.copy()
    pb, sb = ParametersDT(m.mesh), StatesDT(m.mesh)
    oracle.forward_b(c.setup, c.mesh, c.input_data, c.parameters, pb, bgd, c.states, sb, m.states.copy(), c.output, precision="f64")
    assert c.output.cost_jreg > 0
    rng = np.random.default_rng(11)
    act = m.mesh.active_cell == 1
    for name in ("cp", "cft", "lr"):
        direction = rng.standard_normal(act.shape) * act
        base = np.array(getattr(m.parameters, name), np.float64)
        eps = 1e-6
        vals = []
        for sgn in (+1, -1):
            mm = m.copy()
            setattr(mm.parameters, name, np.asfortranarray(base + sgn * eps * direction))
            vals.append(J(mm))
        fd = (vals[0] - vals[1]) / (2 * eps)
        ad = float((np.asarray(getattr(pb, name), np.float64) * direction).sum())
        assert np.isclose(fd, ad, rtol=5e-4, atol=1e-9), (name, fd, ad)


def test_f32_gradient_close_to_f64():
    m = make(("nse",))
    p64, s64 = grad(m, "f64")
    p32, s32 = grad(m, "f32")
    for g32, g64, names in ((p32, p64, ("cp", "cft", "exc", "lr")), (s32, s64, ("hp", "hft", "hlr"))):
        for n in names:
            a, b = np.asarray(getattr(g32, n), np.float64), np.asarray(getattr(g64, n), np.float64)
            # float32 rounding noise of the reverse sweep itself: a few 1e-3 of the largest entry
            assert np.abs(a - b).max() <= 1e-2 * np.abs(b).max() + 1e-12, n


@pytest.mark.parametrize("mapping", ["hyper-linear", "hyper-polynomial"])
def test_hyper_adjoint_vs_finite_differences(mapping):
    m = cases.cance(T=T)
    cases.normalize_descriptor(m)
    cases.set_optimize(m.setup, m.mesh, jobs_fun=("nse",), mapping=mapping)
    nh = m.setup._optimize.nhyper
    hp, hs = Hyper_ParametersDT(m.setup), Hyper_StatesDT(m.setup)
    o = m.setup._optimize
    rng = np.random.default_rng(5)
    defaults = dict(cp=200.0, cft=500.0, exc=0.0, lr=5.0, hp=0.01, hft=0.01, hlr=1e-6)
    for names, obj, lb, ub in ((oracle.PARAM_NAMES, hp, o.lb_parameters, o.ub_parameters), (oracle.STATE_NAMES, hs, o.lb_states, o.ub_states)):
        for i, n in enumerate(names):
            h = np.zeros((nh, 1), np.float64, order="F")
            x = (defaults.get(n, 0.5 * (lb[i] + ub[i])) - lb[i]) / (ub[i] - lb[i])
            x = min(max(x, 1e-6), 1 - 1e-6)
            h[0, 0] = np.log(x / (1 - x))                 # logit of the current value (mw_optimize.f90:1043-1046)
            if n in ("cp", "cft", "lr", "exc"):
                if mapping == "hyper-linear":
                    h[1:, 0] = rng.uniform(-0.5, 0.5, nh - 1)
                else:
                    h[1::2, 0] = rng.uniform(-0.5, 0.5, (nh - 1) // 2)
                    h[2::2, 0] = rng.uniform(0.6, 1.8, (nh - 1) // 2)
            elif mapping == "hyper-polynomial":
                h[2::2, 0] = 1.0
            setattr(obj, n, h)

    def J(hp_):
        c = m.copy()
        return float(oracle.hyper_forward(c.setup, c.mesh, c.input_data, c.parameters, hp_, c.states, hs, c.output, precision="f64"))

    c = m.copy()
    hpb, hsb = Hyper_ParametersDT(m.setup), Hyper_StatesDT(m.setup)
    oracle.hyper_forward_b(c.setup, c.mesh, c.input_data, c.parameters, hp, hpb, c.states, hs, hsb, c.output, precision="f64")
    for n in ("cp", "cft", "lr"):
        for k in range(nh):
            eps = 1e-6
            vals = []
            for sgn in (+1, -1):
                h2 = hp.copy()
                a = np.array(getattr(h2, n), np.float64)
                a[k, 0] += sgn * eps
                setattr(h2, n, np.asfortranarray(a))
                vals.append(J(h2))
            fd = (vals[0] - vals[1]) / (2 * eps)
            assert np.isclose(fd, float(getattr(hpb, n)[k, 0]), rtol=5e-4, atol=1e-9), (n, k, fd, float(getattr(hpb, n)[k, 0]))


def test_gr_d_adjoint_vs_finite_differences():
    # GR_D_FORWARD_B (forward_db.f90:9604-9797) restated as the gr-d mode of the taped loop: directional derivatives against
    # central differences of the INDEPENDENT forward restatement (structure_forward), double precision
    m = make(("nse",))
    m.setup.structure = "gr-d"
    pb, sb = grad(m)
    rng = np.random.default_rng(9)
    act = (m.mesh.active_cell == 1)
    for obj, g, name, eps in ((m.parameters, pb, "cp", 1e-4), (m.parameters, pb, "cft", 1e-4), (m.parameters, pb, "lr", 1e-5),
                              (m.states, sb, "hp", 1e-7), (m.states, sb, "hft", 1e-7)):
        direction = rng.standard_normal(act.shape) * act
        base = np.array(getattr(obj, name), dtype=np.float64)
        vals = []
        for sgn in (+1, -1):
            c = m.copy()
            tgt = c.parameters if obj is m.parameters else c.states
            setattr(tgt, name, np.asfortranarray(base + sgn * eps * direction))
            vals.append(float(oracle.forward(c.setup, c.mesh, c.input_data, c.parameters, m.parameters.copy(), c.states,
                                             m.states.copy(), c.output, precision="f64")))
        fd = (vals[0] - vals[1]) / (2 * eps)
        ad = float((np.asarray(getattr(g, name), np.float64) * direction).sum())
        assert np.isclose(fd, ad, rtol=2e-4, atol=1e-10), (name, fd, ad)
    assert not np.any(pb.exc)                                                 # exc is not a parameter of gr-d
    # the cost the adjoint run reports is the forward run's
    c = m.copy()
    pb2, sb2 = ParametersDT(m.mesh), StatesDT(m.mesh)
    cb = oracle.forward_b(c.setup, c.mesh, c.input_data, c.parameters, pb2, m.parameters.copy(), c.states, sb2, m.states.copy(), c.output,
                          precision="f64")
    assert cb is None or np.isclose(float(c.output.cost), cost(m), rtol=1e-12)


def test_other_adjoints_are_refused():
    for s in ("gr-b", "gr-c", "vic-a"):
        m = make(("nse",))
        m.setup.structure = s
        with pytest.raises(AssertionError):
            grad(m)
