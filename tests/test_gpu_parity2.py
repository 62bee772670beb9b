"""More GPU parity cases through the C ABI: hyper mappings, regularised / normalised VDA gradient, forcing gaps,
domain outputs, ensemble chunking, France-scale gradient with pit pairs, and size-independent properties."""
import ctypes as C

import numpy as np
import pytest

import cases
import oracle
import smash_b200
from smash_b200 import _lib as L
from smash_b200.solver._derived_types import Hyper_ParametersDT, Hyper_StatesDT, ParametersDT, StatesDT
from test_gpu_parity import check_grad, close_q, random_fields

pytestmark = pytest.mark.gpu


def hyper_setup(mapping, T=1440):
    m = cases.cance(T=T)
    cases.normalize_descriptor(m)
    cases.set_optimize(m.setup, m.mesh, jobs_fun=("nse",), mapping=mapping)
    o = m.setup._optimize
    nh = o.nhyper
    hp, hs = Hyper_ParametersDT(m.setup), Hyper_StatesDT(m.setup)
    rng = np.random.default_rng(5)
    defaults = dict(cp=200.0, cft=500.0, exc=0.0, lr=5.0, hp=0.01, hft=0.01, hlr=1e-6)
    for names, obj, lb, ub in ((L.PARAM_NAMES, hp, o.lb_parameters, o.ub_parameters), (L.STATE_NAMES, hs, o.lb_states, o.ub_states)):
        for i, n in enumerate(names):
            h = np.zeros((nh, 1), np.float32, order="F")
            x = (defaults.get(n, 0.5 * (float(lb[i]) + float(ub[i]))) - float(lb[i])) / (float(ub[i]) - float(lb[i]))
            x = min(max(x, 1e-6), 1 - 1e-6)
            h[0, 0] = np.log(x / (1 - x))
            if n in ("cp", "cft", "lr", "exc"):
                if mapping == "hyper-linear":
                    h[1:, 0] = rng.uniform(-0.5, 0.5, nh - 1)
                else:
                    h[1::2, 0] = rng.uniform(-0.5, 0.5, (nh - 1) // 2)
                    h[2::2, 0] = rng.uniform(0.6, 1.8, (nh - 1) // 2)
            elif mapping == "hyper-polynomial":
                h[2::2, 0] = 1.0
            setattr(obj, n, h)
    return m, hp, hs


@pytest.mark.parametrize("mapping", ["hyper-linear", "hyper-polynomial"])
def test_hyper_forward_and_adjoint(mapping):
    a, hp, hs = hyper_setup(mapping)
    b = a.copy()
    smash_b200.hyper_forward(a.setup, a.mesh, a.input_data, a.parameters, hp, hp.copy(), a.states, hs, hs.copy(), a.output)
    oracle.hyper_forward(b.setup, b.mesh, b.input_data, b.parameters, hp, b.states, hs, b.output)
    assert close_q(a.output.qsim, b.output.qsim)
    assert abs(float(a.output.cost) - float(b.output.cost)) < 1e-5 * max(1.0, abs(float(b.output.cost)))
    # the mapping rewrites every field over the whole rectangle (mwd_parameters_manipulation.f90:326-358)
    # device expf against glibc's: one unit in the last place of the sigmoid, i.e. 6e-8 of the field's range ub - lb
    o = a.setup._optimize
    for n in ("cp", "cft", "exc", "lr", "ci", "beta"):
        i = list(L.PARAM_NAMES).index(n)
        span = float(o.ub_parameters[i] - o.lb_parameters[i])
        assert np.allclose(getattr(a.parameters, n), getattr(b.parameters, n), rtol=2e-6, atol=5e-7 * span), n
    # states keep their final values (forward.f90:145: no restore in the hyper path)
    assert np.allclose(a.states.hp, b.states.hp, rtol=1e-4, atol=1e-7)
    a2, b2 = a.copy(), b.copy()
    ga, gsa, gb, gsb = (Hyper_ParametersDT(a.setup), Hyper_StatesDT(a.setup), Hyper_ParametersDT(a.setup), Hyper_StatesDT(a.setup))
    smash_b200.hyper_forward_b(a2.setup, a2.mesh, a2.input_data, a2.parameters, None, hp, ga, None, None, a2.states, None, hs, gsa,
                               None, None, a2.output, None)
    oracle.hyper_forward_b(b2.setup, b2.mesh, b2.input_data, b2.parameters, hp, gb, b2.states, hs, gsb, b2.output)
    for n in ("cp", "cft", "exc", "lr"):
        x, y = np.asarray(getattr(ga, n), np.float64).ravel(), np.asarray(getattr(gb, n), np.float64).ravel()
        assert np.allclose(x, y, rtol=2e-3, atol=2e-3 * np.abs(y).max()), (n, x, y)
    for n in ("hp", "hft", "hlr"):
        x, y = np.asarray(getattr(gsa, n), np.float64).ravel(), np.asarray(getattr(gsb, n), np.float64).ravel()
        assert np.allclose(x, y, rtol=2e-3, atol=2e-3 * np.abs(y).max() + 1e-12), (n, x, y)


def _hyper_objects(m, mapping, seed=5):
    o = m.setup._optimize
    nh = o.nhyper
    hp, hs = Hyper_ParametersDT(m.setup), Hyper_StatesDT(m.setup)
    rng = np.random.default_rng(seed)
    defaults = dict(cp=200.0, cft=500.0, exc=0.0, lr=5.0, hp=0.01, hft=0.01, hlr=1e-6)
    for names, obj, lb, ub in ((L.PARAM_NAMES, hp, o.lb_parameters, o.ub_parameters), (L.STATE_NAMES, hs, o.lb_states, o.ub_states)):
        for i, n in enumerate(names):
            h = np.zeros((nh, 1), np.float32, order="F")
            x = (defaults.get(n, 0.5 * (float(lb[i]) + float(ub[i]))) - float(lb[i])) / (float(ub[i]) - float(lb[i]))
            x = min(max(x, 1e-6), 1 - 1e-6)
            h[0, 0] = np.log(x / (1 - x))
            if n in ("cp", "cft", "lr", "exc"):
                if mapping == "hyper-linear":
                    h[1:, 0] = rng.uniform(-0.3, 0.3, nh - 1)
                else:
                    h[1::2, 0] = rng.uniform(-0.3, 0.3, (nh - 1) // 2)
                    h[2::2, 0] = rng.uniform(0.6, 1.8, (nh - 1) // 2)
            elif mapping == "hyper-polynomial":
                h[2::2, 0] = 1.0
            setattr(obj, n, h)
    return hp, hs


def test_hyper_gradient_france_scale_on_device():
    # Regionalisation step at France scale (906 044 cells, nd = 6, hyper-polynomial: 13 coefficients per field): the mapping
    # kernel writes the plan's field planes, the reductions of HYPER_*_B (forward_db.f90:1434-1537) run on the gradient
    # planes in place -- no rectangle leaves the device.  Checked against the oracle's hyper_forward_b (T = 24) and timed
    # with CUDA events: mapping + reductions must stay below 1 ms on top of the sweeps.
    m = cases.france(T=24, ngauge=4, nd=6)
    cases.set_optimize(m.setup, m.mesh, jobs_fun=("nse",), mapping="hyper-polynomial", gauge="all")
    hp, hs = _hyper_objects(m, "hyper-polynomial")
    lib = L.lib()
    pk = L.Packed()
    s_, m_, i_ = L.pack_setup(m.setup, m.mesh, pk), L.pack_mesh(m.mesh, m.setup, pk), L.pack_input(m.input_data, m.setup, m.mesh, pk)
    p_, st_ = L.pack_parameters(m.parameters, pk), L.pack_states(m.states, pk)
    hp_, hs_ = L.pack_parameters(hp, pk), L.pack_states(hs, pk)
    plan = C.c_void_p()
    L.check(lib.smash_b200_plan_create(C.byref(s_), C.byref(m_), 1, C.byref(plan)))
    try:
        L.check(lib.smash_b200_plan_set_forcing(plan, C.byref(s_), C.byref(i_)))
        L.check(lib.smash_b200_plan_set_fields(plan, C.byref(p_), C.byref(st_), None, None, 0))
        nh = m.setup._optimize.nhyper
        hb = np.zeros((7, nh), np.float32)
        ms = (C.c_float * 4)()
        for _ in range(3):
            L.check(lib.smash_b200_plan_run_hyper_gradient(plan, C.byref(s_), C.byref(i_), C.byref(hp_), C.byref(hs_), L._fp(hb), ms))
        print("hyper step device times (ms): mapping %.3f forward %.3f reverse %.3f reductions %.3f" % tuple(ms))
        assert ms[0] + ms[3] < 1.0, tuple(ms)
        gp, gs = ParametersDT(m.mesh), StatesDT(m.mesh)
        for n in ("cp", "cft", "exc", "lr"):
            getattr(gp, n)[...] = 0.0                                    # only the computed cells are written
        for n in ("hp", "hft", "hlr"):
            getattr(gs, n)[...] = 0.0
        gp_, gs_ = L.pack_parameters(gp, pk), L.pack_states(gs, pk)
        L.check(lib.smash_b200_plan_get_gradient(plan, C.byref(gp_), C.byref(gs_)))
    finally:
        lib.smash_b200_plan_destroy(plan)
    # (a) the reductions against a float64 evaluation of HYPER_*_B on the gradient planes the sweeps left on the device
    o = m.setup._optimize
    d = np.asarray(m.input_data.descriptor, np.float64)
    for f, n in enumerate(("cp", "cft", "exc", "lr", "hp", "hft", "hlr")):
        src, grd = (hp, gp) if f < 4 else (hs, gs)
        i = list(L.PARAM_NAMES if f < 4 else L.STATE_NAMES).index(n)
        lb, ub = (o.lb_parameters, o.ub_parameters) if f < 4 else (o.lb_states, o.ub_states)
        h = np.asarray(getattr(src, n), np.float64).ravel()
        z = h[0] + sum(h[2 * j - 1] * d[..., j - 1] ** h[2 * j] for j in range(1, 7) if h[2 * j - 1] != 0.0)
        e = np.exp(-z)
        g = np.asarray(getattr(grd, n), np.float64) * float(ub[i] - lb[i]) * e / (1.0 + e) ** 2
        ref = np.zeros(nh)
        ref[0] = g.sum()
        for j in range(1, 7):
            pw = d[..., j - 1] ** h[2 * j]
            ref[2 * j - 1] = (pw * g).sum()
            pos = d[..., j - 1] > 0
            ref[2 * j] = (pw[pos] * np.log(d[..., j - 1][pos]) * h[2 * j - 1] * g[pos]).sum()
        assert np.allclose(hb[f], ref, rtol=1e-4, atol=1e-5 * np.abs(ref).max() + 1e-30), (n, hb[f], ref)
    # (b) end to end against the oracle's hyper_forward_b.  cft is left out: over 24 steps its per-cell gradient (3e-10) sits
    # at the float32 noise floor of the sweeps (cp: 2e-5), so its sum over 1.4e5 cells is not comparable
    b = m.copy()
    gb, gsb = Hyper_ParametersDT(m.setup), Hyper_StatesDT(m.setup)
    oracle.hyper_forward_b(b.setup, b.mesh, b.input_data, b.parameters, hp, gb, b.states, hs, gsb, b.output)
    for f, n in ((0, "cp"), (2, "exc"), (3, "lr")):
        x, y = hb[f].astype(np.float64), np.asarray(getattr(gb, n), np.float64).ravel()
        assert np.allclose(x, y, rtol=2e-3, atol=2e-3 * np.abs(y).max() + 1e-12), (n, x, y)


def test_vda_gradient_with_regularisation_and_normalisation():
    # optimize_lbfgsb setting (mw_optimize.f90:547-561): normalised controls, denormalize_forward, prior + smoothing
    def make():
        m = cases.cance(T=480)
        cases.set_optimize(m.setup, m.mesh, jobs_fun=("nse",), jreg_fun=("prior", "smoothing"), wjreg_fun=[1.0, 0.5], wjreg=1e-3,
                           denormalize_forward=True)
        random_fields(m)
        o = m.setup._optimize
        o.optim_parameters[[1, 3, 6, 15]] = 1
        bgd = m.parameters.copy()
        for i, n in enumerate(L.PARAM_NAMES):
            rngw = o.ub_parameters[i] - o.lb_parameters[i]
            setattr(m.parameters, n, np.asfortranarray((getattr(m.parameters, n) - o.lb_parameters[i]) / rngw))
            setattr(bgd, n, np.asfortranarray((getattr(bgd, n) - o.lb_parameters[i]) / rngw + np.float32(0.01)))
        for i, n in enumerate(L.STATE_NAMES):
            setattr(m.states, n, np.asfortranarray((getattr(m.states, n) - o.lb_states[i]) / (o.ub_states[i] - o.lb_states[i])))
        return m, bgd

    (a, bgd_a), (b, bgd_b) = make(), make()
    pa, sa, pb, sb = ParametersDT(a.mesh), StatesDT(a.mesh), ParametersDT(b.mesh), StatesDT(b.mesh)
    sbgd = a.states.copy()
    smash_b200.forward_b(a.setup, a.mesh, a.input_data, a.parameters, pa, bgd_a, None, a.states, sa, sbgd, None, a.output, None)
    oracle.forward_b(b.setup, b.mesh, b.input_data, b.parameters, pb, bgd_b, b.states, sb, sbgd, b.output)
    assert float(b.output.cost_jreg) > 0
    # Jreg: the device adds the 3 136 squared terms in float64 and lands on the float64 oracle (2e-6); the reference's
    # sequential float32 sum (mwd_cost.f90:1210-1216) is 1.6e-5 away from both
    assert np.isclose(float(a.output.cost_jreg), float(b.output.cost_jreg), rtol=1e-4)
    (c, bgd_c) = make()
    oracle.forward_b(c.setup, c.mesh, c.input_data, c.parameters, ParametersDT(c.mesh), bgd_c, c.states, StatesDT(c.mesh), sbgd,
                     c.output, precision="f64")
    assert np.isclose(float(a.output.cost_jreg), float(c.output.cost_jreg), rtol=2e-6)
    assert abs(float(a.output.cost) - float(b.output.cost)) < 1e-5 * max(1.0, abs(float(b.output.cost)))
    check_grad(pa, pb, ("cp", "cft", "exc", "lr"))
    check_grad(sa, sb, ("hp", "hft", "hlr"))
    # parameters / states come back denormalised (forward_db.f90:10697-10703)
    assert np.allclose(a.parameters.cp, b.parameters.cp, rtol=1e-6)
    # the plain forward with the same setting: cost identical, round-trip of the controls as in compute_cost
    (a, bgd_a), (b, bgd_b) = make(), make()
    smash_b200.forward(a.setup, a.mesh, a.input_data, a.parameters, bgd_a, a.states, sbgd, a.output)
    oracle.forward(b.setup, b.mesh, b.input_data, b.parameters, bgd_b, b.states, sbgd, b.output)
    assert abs(float(a.output.cost) - float(b.output.cost)) < 1e-5 * max(1.0, abs(float(b.output.cost)))
    assert np.array_equal(a.parameters.cft, b.parameters.cft)


def with_gaps(T=360, frac=0.02):
    m = cases.cance(sparse=True, T=T)
    prcp, pet = cases.synthetic_forcing(m.mesh.nac, T, seed=3, gap_fraction=frac)
    pet = np.array(pet, order="F")
    pet[::7, 5::50] = -99.0                         # PET gaps too (md_forward_structure.f90:106)
    m.input_data.sparse_prcp, m.input_data.sparse_pet = np.asfortranarray(prcp), pet
    random_fields(m)
    return m


def test_forcing_gaps_forward_and_gradient():
    a, b = with_gaps(), with_gaps()
    assert (a.input_data.sparse_prcp < 0).any()
    smash_b200.forward(a.setup, a.mesh, a.input_data, a.parameters, a.parameters.copy(), a.states, a.states.copy(), a.output)
    oracle.forward(b.setup, b.mesh, b.input_data, b.parameters, b.parameters.copy(), b.states, b.states.copy(), b.output)
    assert close_q(a.output.qsim, b.output.qsim)
    pa, sa, pb, sb = ParametersDT(a.mesh), StatesDT(a.mesh), ParametersDT(b.mesh), StatesDT(b.mesh)
    smash_b200.forward_b(a.setup, a.mesh, a.input_data, a.parameters, pa, a.parameters.copy(), None, a.states, sa, a.states.copy(),
                         None, a.output, None)
    oracle.forward_b(b.setup, b.mesh, b.input_data, b.parameters, pb, b.parameters.copy(), b.states, sb, b.states.copy(), b.output)
    check_grad(pa, pb, ("cp", "cft", "exc", "lr"))
    check_grad(sa, sb, ("hp", "hft", "hlr"))


@pytest.mark.parametrize("sparse", [False, True])
def test_domain_outputs(sparse):
    def make():
        m = cases.cance(sparse=sparse, T=96)
        m.setup.save_qsim_domain = True
        m.setup.save_net_prcp_domain = True
        from smash_b200.solver._derived_types import OutputDT
        m.output = OutputDT(m.setup, m.mesh)
        return m
    a, b = make(), make()
    smash_b200.forward(a.setup, a.mesh, a.input_data, a.parameters, a.parameters.copy(), a.states, a.states.copy(), a.output)
    oracle.forward(b.setup, b.mesh, b.input_data, b.parameters, b.parameters.copy(), b.states, b.states.copy(), b.output)
    for name in (("sparse_qsim_domain", "sparse_net_prcp_domain") if sparse else ("qsim_domain", "net_prcp_domain")):
        x, y = getattr(a.output, name), getattr(b.output, name)
        assert x.shape == y.shape
        assert close_q(x, y), name
    if not sparse:   # inactive cells keep the -99 fill (mwd_output.f90:80-100), gauges read the domain discharge
        assert np.all(a.output.qsim_domain[a.mesh.active_cell == 0] == -99.0)
        g = a.mesh.gauge_pos
        assert np.array_equal(a.output.qsim_domain[g[0, 0], g[0, 1], :], a.output.qsim[0])


def test_multiple_run_chunked_equals_single_launch(golden):
    m = cases.cance(T=240)
    rng = np.random.RandomState(1)
    ns = 37
    smp = np.asfortranarray(np.stack([rng.uniform(lo, hi, ns) for lo, hi in [(1, 1e3), (1, 1e3), (-50, 50), (1, 1e3)]]).astype(np.float32))
    out = []
    for budget in (16384, 1):            # 1 MB per launch forces several chunks
        L.lib().smash_b200_set_option(b"member_budget_mb", budget)
        cost = np.zeros(ns, np.float32)
        qsim = np.zeros((3, 240, ns), np.float32, order="F")
        smash_b200.compute_multiple_run(m.setup, m.mesh, m.input_data, m.parameters, m.states, m.output, smp,
                                        cases.IND_CP_CFT_EXC_LR, cost, qsim)
        out.append((cost.copy(), qsim.copy()))
    L.lib().smash_b200_set_option(b"member_budget_mb", 16384)
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    # a member equals a plain forward run with the same uniform parameters (test_simu.py:57-73)
    k = 5
    f = m.copy()
    for j, n in enumerate(("cp", "cft", "exc", "lr")):
        getattr(f.parameters, n)[...] = smp[j, k]
    smash_b200.forward(f.setup, f.mesh, f.input_data, f.parameters, f.parameters.copy(), f.states, f.states.copy(), f.output)
    # (bit-identical when both calls run on the same engine; ensembles default to the fused engine, single runs to the
    # split engine, whose routing scan rounds differently in the last place)
    assert np.allclose(f.output.qsim, out[0][1][:, :, k], rtol=1e-5, atol=1e-9)
    assert np.isclose(float(f.output.cost), out[0][0][k], rtol=1e-6)


def test_france_window_gradient():
    # 300x300 window of the France mesh (387 blocks, cross-block flags in both sweeps), 4 synthetic gauges on the
    # largest rivers, NSE with equal gauge weights
    def make():
        m = cases.france(T=96, sub=(400, 700, 400, 700), ngauge=4)
        m.setup.save_qsim_domain = False
        random_fields(m, seed=2)
        return m
    a, b = make(), make()
    pa, sa, pb, sb = ParametersDT(a.mesh), StatesDT(a.mesh), ParametersDT(b.mesh), StatesDT(b.mesh)
    smash_b200.forward_b(a.setup, a.mesh, a.input_data, a.parameters, pa, a.parameters.copy(), None, a.states, sa, a.states.copy(),
                         None, a.output, None)
    oracle.forward_b(b.setup, b.mesh, b.input_data, b.parameters, pb, b.parameters.copy(), b.states, sb, b.states.copy(), b.output)
    assert close_q(a.output.qsim, b.output.qsim)
    assert np.isclose(float(a.output.cost), float(b.output.cost), rtol=5e-4)
    check_grad(pa, pb, ("cp", "cft", "exc", "lr"))
    check_grad(sa, sb, ("hp", "hft", "hlr"))


def test_france_full_gradient_pit_pairs_checksum():
    # whole France mesh (906 044 cells, 50 pit pairs): forward+adjoint against the oracle on 24 steps, 3 gauges
    def make():
        m = cases.france(T=24, ngauge=3)
        m.setup.save_qsim_domain = False
        return m
    a, b = make(), make()
    pa, sa, pb, sb = ParametersDT(a.mesh), StatesDT(a.mesh), ParametersDT(b.mesh), StatesDT(b.mesh)
    smash_b200.forward_b(a.setup, a.mesh, a.input_data, a.parameters, pa, a.parameters.copy(), None, a.states, sa, a.states.copy(),
                         None, a.output, None)
    oracle.forward_b(b.setup, b.mesh, b.input_data, b.parameters, pb, b.parameters.copy(), b.states, sb, b.states.copy(), b.output)
    assert close_q(a.output.qsim, b.output.qsim)
    # over 24 steps the transfer-store fields (cft, hft) have no measurable influence (gradient ~1e-10, pure rounding
    # noise in any float32 implementation): only the well-conditioned fields are compared
    check_grad(pa, pb, ("cp", "lr"))
    check_grad(sa, sb, ("hp", "hlr"))
    inactive = a.mesh.active_cell == 0
    assert not np.any(pa.cp[inactive])


def test_linearity_of_gradient_in_cost_b():
    # forward_b is linear in the seed cost_b (BASE_FORWARD_B): doubling the seed doubles the gradient
    m = cases.cance(T=240)
    random_fields(m)
    g = []
    for seed in (1.0, 2.0):
        c = m.copy()
        pb, sb = ParametersDT(m.mesh), StatesDT(m.mesh)
        smash_b200.forward_b(c.setup, c.mesh, c.input_data, c.parameters, pb, c.parameters.copy(), None, c.states, sb,
                             c.states.copy(), None, c.output, None, 0.0, seed)
        g.append(pb.cp.copy())
    assert np.allclose(2.0 * g[0], g[1], rtol=1e-5, atol=1e-12)


def test_run_to_run_determinism():
    a, b = cases.cance(T=240), cases.cance(T=240)
    for m in (a, b):
        smash_b200.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(), m.output)
    assert np.array_equal(a.output.qsim, b.output.qsim)


def test_scalar_product_test_on_gpu():
    # the reference's scalar product test (optimize/mw_adjoint_test.f90:26-105) through the shim module, both sides on the
    # GPU: <dY*, dY> from central differences of forward along dk = 1, <dk*, dk> from forward_b
    from smash_b200.solver import _mw_adjoint_test as A
    c = cases.cance(T=480)
    cases.set_optimize(c.setup, c.mesh, jobs_fun=("nse",))
    random_fields(c)
    sp1, sp2 = A.scalar_product_test(c.setup, c.mesh, c.input_data, c.parameters, c.states, c.output, verbose=False)
    print("scalar product test: sp1 = %.6e sp2 = %.6e rel %.2e" % (sp1, sp2, abs(sp1 - sp2) / abs(sp1)))
    assert sp1 != 0.0 and abs(sp1 - sp2) <= 2e-2 * abs(sp1), (sp1, sp2)


def test_forward_d_shim_matches_the_adjoint():
    # _mw_forward.forward_d (reference argument list, mw_forward.f90:70-97): cost_d along a direction that mixes two fields
    # against <forward_b, direction>
    from smash_b200.solver import _mw_forward as F
    c = cases.cance(T=480)
    cases.set_optimize(c.setup, c.mesh, jobs_fun=("nse",))
    random_fields(c)
    pd, sd = ParametersDT(c.mesh), StatesDT(c.mesh)
    for obj in (pd, sd):
        for n in vars(obj):
            if isinstance(getattr(obj, n), np.ndarray):
                getattr(obj, n)[...] = 0.0
    pd.cp[...] = 1.0
    pd.lr[...] = 0.5
    cost, cost_d = F.forward_d(c.setup, c.mesh, c.input_data, c.parameters, pd, c.parameters.copy(), None, c.states, sd,
                               c.states.copy(), None, c.output, None)
    pb, sb = ParametersDT(c.mesh), StatesDT(c.mesh)
    F.forward_b(c.setup, c.mesh, c.input_data, c.parameters.copy(), pb, c.parameters.copy(), None, c.states.copy(), sb,
                c.states.copy(), None, c.output, None)
    ref = float((np.asarray(pb.cp, np.float64) * pd.cp).sum() + (np.asarray(pb.lr, np.float64) * pd.lr).sum())
    print("forward_d: cost %.6f cost_d %.6e <forward_b, d> %.6e" % (float(cost), float(cost_d), ref))
    assert ref != 0.0 and abs(float(cost_d) - ref) <= 2e-2 * abs(ref)


def test_gpu_adjoint_against_gpu_finite_differences():
    # CUDA-only self-check of forward_b: for every control field, 5 random directions d on the active cells; the
    # directional derivative <forward_b, d> against (J(x + h d) - J(x - h d)) / 2h of the CUDA forward.  Everything is
    # float32, so the difference quotient carries the rounding noise of J (about 2e-7 J / h): directions whose derivative
    # is below ten times that noise are reported but not judged; the others must agree to 3 %.
    m = cases.cance(T=720)
    cases.set_optimize(m.setup, m.mesh, jobs_fun=("nse",))
    random_fields(m, seed=31)
    act = m.mesh.active_cell == 1
    pg, sg = ParametersDT(m.mesh), StatesDT(m.mesh)
    a = m.copy()
    smash_b200.forward_b(a.setup, a.mesh, a.input_data, a.parameters, pg, a.parameters.copy(), None, a.states, sg,
                         a.states.copy(), None, a.output, None)
    J0 = float(a.output.cost)

    def cost(field, is_state, delta):
        b = m.copy()
        obj = b.states if is_state else b.parameters
        getattr(obj, field)[...] = (np.asarray(getattr(obj, field), np.float64) + delta).astype(np.float32)
        smash_b200.forward(b.setup, b.mesh, b.input_data, b.parameters, b.parameters.copy(), b.states, b.states.copy(), b.output)
        return float(b.output.cost)

    rng = np.random.default_rng(7)
    steps = dict(cp=2.0, cft=2.0, exc=0.05, lr=0.05, hp=2e-3, hft=2e-3, hlr=2e-7)
    judged, worst = 0, 0.0
    for field, h in steps.items():
        is_state = field in ("hp", "hft", "hlr")
        g = np.asarray(getattr(sg if is_state else pg, field), np.float64)
        for k in range(5):
            d = np.zeros(act.shape)
            d[act] = rng.uniform(-1.0, 1.0, int(act.sum()))
            fd = (cost(field, is_state, h * d) - cost(field, is_state, -h * d)) / (2.0 * h)
            ad = float((g * d).sum())
            noise = 2e-7 * abs(J0) / h
            rel = abs(fd - ad) / max(abs(fd), 1e-300)
            print("%-4s dir %d: adjoint %.5e fd %.5e rel %.2e%s" % (field, k, ad, fd, rel, "" if abs(fd) > 10 * noise else "  (below the FD noise floor)"))
            if abs(fd) > 10 * noise:
                judged += 1
                worst = max(worst, rel)
                assert rel <= 3e-2, (field, k, ad, fd)
    assert judged >= 15, judged
    print("judged %d directions, worst relative difference %.2e" % (judged, worst))


def _pit_pair_cells(mesh):
    """Cells of the 2-cycles of the flow-direction graph: A -> B and B -> A."""
    fd = np.asarray(mesh.flwdir)
    drow = np.array([1, 1, 0, -1, -1, -1, 0, 1]); dcol = np.array([0, -1, -1, -1, 0, 1, 1, 1])   # md_routing_operator.f90:29-30
    act = np.asarray(mesh.active_cell) == 1
    rr, cc = np.nonzero(act & (fd >= 1) & (fd <= 8))
    # a cell with flwdir d flows to (row - drow[d-1], col - dcol[d-1])
    tr, tc = rr - drow[fd[rr, cc] - 1], cc - dcol[fd[rr, cc] - 1]
    ok = (tr >= 0) & (tr < mesh.nrow) & (tc >= 0) & (tc < mesh.ncol)
    rr, cc, tr, tc = rr[ok], cc[ok], tr[ok], tc[ok]
    fd2 = fd[tr, tc]
    ok2 = (fd2 >= 1) & (fd2 <= 8) & act[tr, tc]
    rr, cc, tr, tc, fd2 = rr[ok2], cc[ok2], tr[ok2], tc[ok2], fd2[ok2]
    back = (tr - drow[fd2 - 1] == rr) & (tc - dcol[fd2 - 1] == cc)
    return rr[back], cc[back]


@pytest.mark.parametrize("engine", ["row_passes", "subtree"])
def test_france_full_size_against_oracle_on_basins(engine):
    # engine: the row-based passes of the drop-in call (streamed), and the subtree engine that the plan API -- and with it the
    # bench -- runs (here forced into the drop-in call: the forcing goes up in one piece, the series come back through the
    # scatter pass).
    # BASELINE.json's bench configuration at its own size (906 044 cells, T = 720, sparse_qsim_domain): the device series of
    # the whole domain against the CPU oracle on the Loire basin (136 170 cells, the longest rivers of the mesh), three
    # other basins and basins that end in a pit pair -- the oracle computes a basin alone through mesh%local_active_cell
    # (md_forward_structure.f90:88), which is exact because basins exchange nothing.  Reports the measured differences.
    from smash_b200 import distributed as D
    m = cases.france(T=720)
    lib = L.lib()
    if engine == "subtree":
        lib.smash_b200_set_option(b"sub_engine", 1)
        lib.smash_b200_set_option(b"stream", 0)
        lib.smash_b200_clear_cache()
    try:
        smash_b200.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(), m.output)
    finally:
        lib.smash_b200_set_option(b"sub_engine", -1)
        lib.smash_b200_set_option(b"stream", 1)
        lib.smash_b200_clear_cache()
    gpu = m.output.sparse_qsim_domain
    labels, nb = D.basin_labels(m.mesh, m.setup)
    size = np.bincount(labels[labels >= 0], minlength=nb)
    order = np.argsort(-size, kind="stable")
    rng = np.random.default_rng(3)
    mid = [int(b) for b in rng.choice(order[5:400], 3, replace=False)]
    pr, pc = _pit_pair_cells(m.mesh)
    pit = sorted({int(labels[r, c]) for r, c in zip(pr, pc) if labels[r, c] >= 0}, key=lambda b: -size[b])[:6]
    assert len(pr) >= 100 and pit, (len(pr), pit)                          # the France mesh holds 50 computed pit pairs
    chosen = [int(order[0])] + mid + [b for b in pit if b != int(order[0])]
    mask = np.isin(labels, chosen)
    k = m.mesh._rowcol_to_ind_sparse
    c = cases.france(T=720)
    c.mesh._local_active_cell = np.asfortranarray(mask.astype(np.int32))
    oracle.forward(c.setup, c.mesh, c.input_data, c.parameters, c.parameters.copy(), c.states, c.states.copy(), c.output)
    rows = k[mask] - 1
    a, b = np.asarray(gpu[rows], np.float64), np.asarray(c.output.sparse_qsim_domain[rows], np.float64)
    err = np.abs(a - b)
    big = np.abs(b) > 1e-3
    print("France T=720 vs oracle on %d basins (%d cells, %d of them pit-pair cells): max abs %.3e, max rel (|q| > 1e-3) %.3e, "
          "median rel %.3e, q max %.3e" % (len(chosen), mask.sum(), int(mask[pr, pc].sum()), err.max(), (err[big] / np.abs(b[big])).max(),
                                           np.median(err[big] / np.abs(b[big])), b.max()))
    assert b.max() > 100.0                                                  # the Loire carries real discharge
    assert np.all(err <= 1e-4 + 2e-3 * np.abs(b)), float((err - 2e-3 * np.abs(b)).max())
    prow = k[pr, pc][mask[pr, pc]] - 1
    assert len(prow) > 0
    ep = np.abs(np.asarray(gpu[prow], np.float64) - c.output.sparse_qsim_domain[prow])
    assert np.all(ep <= 1e-4 + 2e-3 * np.abs(c.output.sparse_qsim_domain[prow]))


@pytest.mark.parametrize("T", [24, 168, 1440])
def test_math0_against_the_relative_gate(T):
    # math = 0 (IEEE division / sqrt, libm tanhf, the reference's statement order, no FMA contraction; fused engine) against
    # the float32 oracle at the relative gate SURVEY 8(c) proposed, |dq| <= floor + 1e-4 |q|.  The relative part holds over
    # the whole 1440 steps.  The absolute floor cannot be the proposed 1e-6 m3/s: the reference's transfer formula
    # q = (ht_imd - ht) ct cancels (md_gr_operator.f90:104-106), so last-place differences between CUDA's and glibc's
    # tanhf / powf leave ~1e-6 of noise per cell, which the 383 cells of the catchment add up at the gauge (measured: 4e-6
    # after a week, 7e-6 after 1440 steps; the float32 oracle is itself 3e-4 relative away from the float64 one).  Gate:
    # 1e-5 + 1e-4 |q| for every horizon, measured values printed.
    lib = L.lib()
    lib.smash_b200_set_option(b"math", 0)
    lib.smash_b200_clear_cache()
    try:
        a, b = cases.cance(T=T), cases.cance(T=T)
        for m in (a, b):
            random_fields(m, seed=41)
        smash_b200.forward(a.setup, a.mesh, a.input_data, a.parameters, a.parameters.copy(), a.states, a.states.copy(), a.output)
        oracle.forward(b.setup, b.mesh, b.input_data, b.parameters, b.parameters.copy(), b.states, b.states.copy(), b.output)
    finally:
        lib.smash_b200_set_option(b"math", 1)
        lib.smash_b200_clear_cache()
    qa, qb = np.asarray(a.output.qsim, np.float64), np.asarray(b.output.qsim, np.float64)
    excess = np.abs(qa - qb) - 1e-4 * np.abs(qb)
    print("math = 0, T = %d: max(|dq| - 1e-4 |q|) = %.3e, max |dq| = %.3e" % (T, excess.max(), np.abs(qa - qb).max()))
    assert np.all(np.abs(qa - qb) <= 1e-5 + 1e-4 * np.abs(qb)), float(excess.max())


def test_signature_objectives_against_oracle():
    # signature-based objectives (mwd_cost.f90:770-970): continuous (Crc, flow percentiles) and event-based (Erc, Elt, Epf)
    # with an event mask of two events, mixed with nse; the device cost kernel against the oracle, gauge by gauge
    from smash_b200.solver._mw_forcing_statistic import compute_mean_forcing
    names = ["nse", "Crc", "Cfp2", "Cfp10", "Cfp50", "Cfp90", "Erc", "Elt", "Epf"]
    for k, gauge in enumerate(("downstream", "all")):
        a, b = cases.cance(), cases.cance()
        for m in (a, b):
            random_fields(m, seed=51)
            cases.set_optimize(m.setup, m.mesh, jobs_fun=tuple(names), wjobs_fun=[1, 2, 2, 2, 2, 2, 2, 2, 2], gauge=gauge)
            me = m.setup._optimize.mask_event
            me[:, 100:220] = 1
            me[:, 700:900] = 2
            compute_mean_forcing(m.setup, m.mesh, m.input_data)
        smash_b200.forward(a.setup, a.mesh, a.input_data, a.parameters, a.parameters.copy(), a.states, a.states.copy(), a.output)
        oracle.forward(b.setup, b.mesh, b.input_data, b.parameters, b.parameters.copy(), b.states, b.states.copy(), b.output)
        print("signature cost (%s): gpu %.7f oracle %.7f" % (gauge, float(a.output.cost), float(b.output.cost)))
        assert float(b.output.cost) > 0.5
        assert abs(float(a.output.cost) - float(b.output.cost)) <= 2e-3 * abs(float(b.output.cost))
    # one signature at a time on the downstream gauge
    for name in names[1:]:
        a, b = cases.cance(T=960), cases.cance(T=960)
        for m in (a, b):
            cases.set_optimize(m.setup, m.mesh, jobs_fun=(name,))
            m.setup._optimize.mask_event[:, 100:220] = 1
            m.setup._optimize.mask_event[:, 700:900] = 2
            compute_mean_forcing(m.setup, m.mesh, m.input_data)
        smash_b200.forward(a.setup, a.mesh, a.input_data, a.parameters, a.parameters.copy(), a.states, a.states.copy(), a.output)
        oracle.forward(b.setup, b.mesh, b.input_data, b.parameters, b.parameters.copy(), b.states, b.states.copy(), b.output)
        assert abs(float(a.output.cost) - float(b.output.cost)) <= 1e-5 + 2e-3 * abs(float(b.output.cost)), (name, float(a.output.cost), float(b.output.cost))
    # the adjoint of a signature objective is not provided: loud failure
    pb, sb = ParametersDT(a.mesh), StatesDT(a.mesh)
    with pytest.raises(RuntimeError, match="signature"):
        smash_b200.forward_b(a.setup, a.mesh, a.input_data, a.parameters, pb, a.parameters.copy(), None, a.states, sb, a.states.copy(),
                             None, a.output, None)
