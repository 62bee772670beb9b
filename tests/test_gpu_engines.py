"""The two engines behind the ABI against each other, the plan API's per-kernel timers, and a
full-size property check of the routing pass (BASELINE.json's France configuration, T = 720)."""
import ctypes as C

import numpy as np
import pytest

import cases
import smash_b200
from smash_b200 import _lib as L
from smash_b200.solver._derived_types import ParametersDT, StatesDT
from test_gpu_parity import check_grad, random_fields

pytestmark = pytest.mark.gpu


def _with_engine(engine, fn):
    lib = L.lib()
    lib.smash_b200_set_option(b"engine", engine)
    try:
        return fn()
    finally:
        lib.smash_b200_set_option(b"engine", -1)
        lib.smash_b200_clear_cache()


def test_split_and_fused_engines_agree():
    # 300 x 300 window of France, 96 steps, 4 gauges: discharge, cost and gradient of the split engine (reservoir pass +
    # routing scan) against the fused tick wavefront, which evaluates the routing recurrence strictly sequentially
    def run():
        m = cases.france(T=96, sub=(400, 700, 400, 700), ngauge=4)
        random_fields(m, seed=3)
        pb, sb = ParametersDT(m.mesh), StatesDT(m.mesh)
        smash_b200.forward_b(m.setup, m.mesh, m.input_data, m.parameters, pb, m.parameters.copy(), None, m.states, sb,
                             m.states.copy(), None, m.output, None)
        return m, pb, sb
    a, pa, sa = _with_engine(1, run)
    b, pb, sb = _with_engine(0, run)
    qa, qb = np.asarray(a.output.qsim, np.float64), np.asarray(b.output.qsim, np.float64)
    assert np.all(np.abs(qa - qb) <= 1e-6 + 1e-4 * np.abs(qb)), float(np.abs(qa - qb).max())
    assert np.isclose(float(a.output.cost), float(b.output.cost), rtol=1e-4)
    check_grad(pa, pb, ("cp", "cft", "exc", "lr"))
    check_grad(sa, sb, ("hp", "hft", "hlr"))


def test_streamed_forward_agrees():
    # large host-resident sparse forcing: the ABI forward streams 256-step windows (upload, kernels, download overlapped).
    # Same run with streaming off (one 640-step window): discharge, domain series, final states and cost must agree to the
    # last places (the routing scan groups the time steps differently), and the second call reuses the device forcing.
    lib = L.lib()
    def run(stream, version=0):
        lib.smash_b200_set_option(b"stream", stream)
        lib.smash_b200_set_option(b"stream_min_mb", 0)
        try:
            m = cases.france(T=600, sub=(400, 560, 400, 560), ngauge=3)
            random_fields(m, seed=7)
            m.setup.save_net_prcp_domain = True
            m.output = type(m.output)(m.setup, m.mesh)
            m.input_data._forcing_version = version
            smash_b200.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(),
                               m.output)
            first = np.array(m.output.sparse_qsim_domain, copy=True)
            if version:
                m.output.sparse_qsim_domain[...] = -1.0
                smash_b200.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states,
                                   m.states.copy(), m.output)
                assert np.array_equal(first, m.output.sparse_qsim_domain)
            # the gradient through the same cached plan (256-step routing windows when streaming is on)
            m.grad = (ParametersDT(m.mesh), StatesDT(m.mesh))
            g = cases.france(T=600, sub=(400, 560, 400, 560), ngauge=3)
            random_fields(g, seed=7)
            smash_b200.forward_b(g.setup, g.mesh, g.input_data, g.parameters, m.grad[0], g.parameters.copy(), None, g.states,
                                 m.grad[1], g.states.copy(), None, g.output, None)
            m.grad_cost = float(g.output.cost)
            return m
        finally:
            lib.smash_b200_set_option(b"stream", 1)
            lib.smash_b200_set_option(b"stream_min_mb", 256)
            lib.smash_b200_clear_cache()
    a, b = run(1), run(0)
    assert np.isclose(a.grad_cost, b.grad_cost, rtol=1e-5)
    check_grad(a.grad[0], b.grad[0], ("cp", "cft", "exc", "lr"))
    check_grad(a.grad[1], b.grad[1], ("hp", "hft", "hlr"))
    assert a.mesh.nac % 4 == 0, a.mesh.nac                                   # else the streamed path is not taken
    run(1, version=77)
    for x, y in ((a.output.qsim, b.output.qsim), (a.output.sparse_qsim_domain, b.output.sparse_qsim_domain),
                 (a.output.sparse_net_prcp_domain, b.output.sparse_net_prcp_domain), (a.output.fstates.hlr, b.output.fstates.hlr),
                 (a.output.fstates.hp, b.output.fstates.hp)):
        x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
        assert np.all(np.abs(x - y) <= 1e-7 + 1e-5 * np.abs(y)), float(np.abs(x - y).max())
    assert np.isclose(float(a.output.cost), float(b.output.cost), rtol=1e-5)
    import oracle
    c = cases.france(T=600, sub=(400, 560, 400, 560), ngauge=3)
    random_fields(c, seed=7)
    oracle.forward(c.setup, c.mesh, c.input_data, c.parameters, c.parameters.copy(), c.states, c.states.copy(), c.output)
    qa, qc = np.asarray(a.output.sparse_qsim_domain, np.float64), np.asarray(c.output.sparse_qsim_domain, np.float64)
    assert np.all(np.abs(qa - qc) <= 1e-4 + 2e-3 * np.abs(qc)), float(np.abs(qa - qc).max())


def _plan(m, members=1):
    lib = L.lib()
    pk = L.Packed()
    s_, m_, i_ = L.pack_setup(m.setup, m.mesh, pk), L.pack_mesh(m.mesh, m.setup, pk), L.pack_input(m.input_data, m.setup, m.mesh, pk)
    p_, st_ = L.pack_parameters(m.parameters, pk), L.pack_states(m.states, pk)
    plan = C.c_void_p()
    L.check(lib.smash_b200_plan_create(C.byref(s_), C.byref(m_), members, C.byref(plan)))
    L.check(lib.smash_b200_plan_set_forcing(plan, C.byref(s_), C.byref(i_)))
    L.check(lib.smash_b200_plan_set_fields(plan, C.byref(p_), C.byref(st_), None, None, 0))
    return plan, pk


def test_plan_kernel_times_and_stats():
    lib = L.lib()
    m = cases.cance(sparse=True, T=240)
    plan, keep = _plan(m)
    try:
        if lib.smash_b200_plan_stat(plan, b"engine") != 1.0:                     # default: split engine (math = 1)
            pytest.skip("the fused engine was selected (math = 0 or SMASH_B200_ENGINE=0)")
        routed, src = lib.smash_b200_plan_stat(plan, b"routed_cells"), lib.smash_b200_plan_stat(plan, b"source_cells")
        assert routed + src == 383 and src == 194                                # Cance: 194 cells with flwacc == 1
        assert lib.smash_b200_plan_stat(plan, b"inflow_edges") == 382            # a tree with one outlet
        ms = C.c_float(0)
        L.check(lib.smash_b200_plan_run_forward(plan, C.byref(ms)))
        kt = (C.c_float * 5)()
        L.check(lib.smash_b200_plan_kernel_times(plan, kt))
        assert kt[0] > 0 and kt[1] > 0 and kt[3] < 0 and kt[4] < 0             # forward kernels ran, adjoint ones did not
        assert kt[0] + kt[1] <= ms.value * 1.05
        f, r = C.c_float(0), C.c_float(0)
        L.check(lib.smash_b200_plan_run_gradient(plan, C.byref(f), C.byref(r)))
        L.check(lib.smash_b200_plan_kernel_times(plan, kt))
        assert kt[3] > 0 and kt[4] > 0
    finally:
        lib.smash_b200_plan_destroy(plan)


def test_france_full_size_engines_and_routing_properties():
    # BASELINE.json's France configuration at full size (906 044 cells, T = 720): the CPU oracle needs minutes here, so the
    # run is checked (a) by the checksum of the whole domain discharge (sum over 6.5e8 cell-steps, double accumulation)
    # of the split engine against the fused engine -- two independent implementations, each pinned to the oracle at
    # smaller sizes -- and (b) through a property of linear_routing (md_routing_operator.f90:62-79): a shorter routing
    # time lr drains the routing stores earlier, so more water has passed through the cells by the end of the run.
    lib = L.lib()
    m = cases.france(T=720)

    def checksum(engine, lr):
        m.parameters.lr[...] = lr
        lib.smash_b200_set_option(b"engine", engine)
        try:
            plan, keep = _plan(m)
            try:
                assert lib.smash_b200_plan_stat(plan, b"engine") == float(engine)
                ms = C.c_float(0)
                L.check(lib.smash_b200_plan_run_forward(plan, C.byref(ms)))
                chk = C.c_double(0)
                L.check(lib.smash_b200_plan_checksum(plan, C.byref(chk)))
                return chk.value
            finally:
                lib.smash_b200_plan_destroy(plan)
        finally:
            lib.smash_b200_set_option(b"engine", -1)

    split, fused, fast = checksum(1, 5.0), checksum(0, 5.0), checksum(1, 0.05)
    assert np.isfinite([split, fused, fast]).all() and split > 0
    assert abs(split - fused) <= 2e-6 * fused, (split, fused)
    assert fast > split


def test_multiple_run_on_split_engine_lane_per_member(golden):
    # ensembles run on the split engine by default (ensemble_engine = -1 -> 1 with math = 1): reservoir pass + routing
    # with lane = member (route_members_kernel); checked against the reference's golden ensemble (smash/tests/baseline.hdf5
    # multiple_run.*, atol 1e-4 as in test_simu.py:53) and against the fused engine (ensemble_engine = 0)
    lib = L.lib()
    m = cases.cance()
    rng = np.random.RandomState(99)
    ns = 10
    smp = np.asfortranarray(np.stack([rng.uniform(lo, hi, ns) for lo, hi in [(1e-6, 1e3), (1e-6, 1e3), (-50, 50), (1e-6, 1e3)]]).astype(np.float32))
    res = {}
    for eng in (0, 1):
        lib.smash_b200_set_option(b"ensemble_engine", eng)
        try:
            cost = np.zeros(ns, np.float32)
            qsim = np.zeros((3, m.setup._ntime_step, ns), np.float32, order="F")
            smash_b200.compute_multiple_run(m.setup, m.mesh, m.input_data, m.parameters, m.states, m.output, smp,
                                            cases.IND_CP_CFT_EXC_LR, cost, qsim)
            res[eng] = (cost, qsim)
        finally:
            lib.smash_b200_set_option(b"ensemble_engine", -1)
            lib.smash_b200_clear_cache()
    assert np.allclose(res[1][1], golden["multiple_run.qsim"], atol=1e-4)
    assert np.allclose(res[1][0], golden["multiple_run.cost"], atol=1e-4, rtol=1e-5)
    assert np.allclose(res[1][1], res[0][1], rtol=2e-3, atol=1e-4)


def test_domain_split_by_basin_equals_full_run():
    # the basins of one domain computed separately (what each rank of forward_sharded_by_basin does, here one
    # after the other on one GPU) give the series of the undivided run: basins exchange nothing
    from smash_b200 import distributed as D
    full = cases.france(T=96, sub=(300, 700, 300, 700), ngauge=3)
    random_fields(full, seed=9)
    smash_b200.forward(full.setup, full.mesh, full.input_data, full.parameters, full.parameters.copy(), full.states,
                       full.states.copy(), full.output)
    masks, load = D.basin_masks(full.mesh, 3, full.setup)
    assert load.sum() == full.mesh.nac and load.min() > 0
    got = np.zeros_like(full.output.sparse_qsim_domain)
    k = full.mesh._rowcol_to_ind_sparse
    for mk in masks:
        part = cases.france(T=96, sub=(300, 700, 300, 700), ngauge=3)
        random_fields(part, seed=9)
        part.mesh._local_active_cell = mk
        smash_b200.forward(part.setup, part.mesh, part.input_data, part.parameters, part.parameters.copy(), part.states,
                           part.states.copy(), part.output)
        own = np.zeros(full.mesh.nac, dtype=bool)
        own[k[mk == 1] - 1] = True
        got[own] = part.output.sparse_qsim_domain[own]
    a, b = np.asarray(got, np.float64), np.asarray(full.output.sparse_qsim_domain, np.float64)
    assert np.all(np.abs(a - b) <= 1e-9 + 1e-6 * np.abs(b)), float(np.abs(a - b).max())
    L.lib().smash_b200_clear_cache()


def _run_window(m, opts):
    lib = L.lib()
    for k, v in opts.items():
        lib.smash_b200_set_option(k.encode(), v)
    lib.smash_b200_clear_cache()
    try:
        smash_b200.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(), m.output)
        return m
    finally:
        for k in opts:
            lib.smash_b200_set_option(k.encode(), {"tick_pass": 0, "shallow_acc": 32, "tick_variant": 8, "tick_min_cells": 65536,
                                                   "tick_ctas_per_sm": 0, "sub_engine": -1, "sub_min_cells": 65536, "sub_scatter": 0}[k])
        lib.smash_b200_clear_cache()


@pytest.mark.parametrize("opts", [{"shallow_acc": 32}, {"shallow_acc": 3}, {"shallow_acc": 100000, "tick_variant": 6},
                                  {"shallow_acc": 200, "tick_variant": 4}, {"shallow_acc": 1, "tick_ctas_per_sm": 2}])
def test_tick_pass_agrees_with_row_passes(opts):
    # forward runs of large domains: the tick pass (reservoirs + routing of every cell in one kernel, 8 steps at a time:
    # tile tickets route the shallow cells, reach tickets the deep chains, discharge blocks handed from producer to consumer)
    # against the row-based passes (tick_pass = 0).  T = 100 ends in a partial window; shallow_acc moves the class boundary
    # from "every gathering cell sits in a reach" to "everything the tile rounds allow is shallow"; the crop holds pit pairs.
    def model():
        m = cases.france(T=100, sub=(250, 900, 250, 900), ngauge=4)
        random_fields(m, seed=11)
        m.setup.save_net_prcp_domain = True
        m.output = type(m.output)(m.setup, m.mesh)
        return m
    o = dict(opts)
    o["tick_min_cells"] = 1000
    o["tick_pass"] = 1
    a = _run_window(model(), o)
    b = _run_window(model(), {"tick_pass": 0})
    assert a.mesh.nac > 100000
    for name, x, y in (("qsim", a.output.qsim, b.output.qsim), ("qdom", a.output.sparse_qsim_domain, b.output.sparse_qsim_domain),
                       ("netp", a.output.sparse_net_prcp_domain, b.output.sparse_net_prcp_domain),
                       ("hlr", a.output.fstates.hlr, b.output.fstates.hlr), ("hp", a.output.fstates.hp, b.output.fstates.hp),
                       ("hft", a.output.fstates.hft, b.output.fstates.hft)):
        x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
        assert np.all(np.abs(x - y) <= 1e-7 + 1e-5 * np.abs(y)), (name, float(np.abs(x - y).max()))
    assert np.isclose(float(a.output.cost), float(b.output.cost), rtol=1e-5)


@pytest.mark.parametrize("acc", [16, 2])
def test_tick_pass_against_oracle(acc):
    import oracle
    m = cases.france(T=50, sub=(400, 700, 400, 700), ngauge=3)
    random_fields(m, seed=13)
    c = m.copy()
    c.output = type(m.output)(m.setup, m.mesh)
    a = _run_window(m, {"tick_pass": 1, "tick_min_cells": 1000, "shallow_acc": acc})
    oracle.forward(c.setup, c.mesh, c.input_data, c.parameters, c.parameters.copy(), c.states, c.states.copy(), c.output)
    qa, qc = np.asarray(a.output.sparse_qsim_domain, np.float64), np.asarray(c.output.sparse_qsim_domain, np.float64)
    assert np.all(np.abs(qa - qc) <= 1e-4 + 2e-3 * np.abs(qc)), float(np.abs(qa - qc).max())
    assert np.all(np.abs(np.asarray(a.output.qsim, np.float64) - c.output.qsim) <= 1e-4 + 2e-3 * np.abs(c.output.qsim))
    for name in ("hp", "hft", "hlr"):
        x, y = np.asarray(getattr(a.output.fstates, name), np.float64), np.asarray(getattr(c.output.fstates, name), np.float64)
        assert np.all(np.abs(x - y) <= 1e-5 + 2e-3 * np.abs(y)), (name, float(np.abs(x - y).max()))


def test_tick_pass_cance_against_oracle():
    # the whole Cance catchment through the tick pass (383 cells: 12 tiles and a few reaches), T = 1440, against the oracle
    import oracle
    m = cases.cance(sparse=True, T=1440)
    random_fields(m, seed=14)
    m.setup.save_qsim_domain = True
    m.output = type(m.output)(m.setup, m.mesh)
    c = m.copy()
    c.output = type(m.output)(m.setup, m.mesh)
    a = _run_window(m, {"tick_pass": 1, "tick_min_cells": 0, "shallow_acc": 8})
    oracle.forward(c.setup, c.mesh, c.input_data, c.parameters, c.parameters.copy(), c.states, c.states.copy(), c.output)
    qa, qc = np.asarray(a.output.sparse_qsim_domain, np.float64), np.asarray(c.output.sparse_qsim_domain, np.float64)
    assert np.all(np.abs(qa - qc) <= 1e-4 + 2e-3 * np.abs(qc)), float(np.abs(qa - qc).max())
    assert np.all(np.abs(np.asarray(a.output.qsim, np.float64) - c.output.qsim) <= 1e-4 + 2e-3 * np.abs(c.output.qsim))


def _gradient_with(opts, make):
    lib = L.lib()
    defaults = {"adjoint_checkpoint": -1, "stream_min_mb": 256, "tape_budget_mb": 16384}
    for k, v in opts.items():
        lib.smash_b200_set_option(k.encode(), v)
    lib.smash_b200_clear_cache()
    try:
        m = make()
        pb, sb = ParametersDT(m.mesh), StatesDT(m.mesh)
        smash_b200.forward_b(m.setup, m.mesh, m.input_data, m.parameters, pb, m.parameters.copy(), None, m.states, sb,
                             m.states.copy(), None, m.output, None)
        return m, pb, sb
    finally:
        for k in opts:
            lib.smash_b200_set_option(k.encode(), defaults[k])
        lib.smash_b200_clear_cache()


@pytest.mark.parametrize("case", ["france", "cance"])
def test_checkpointed_adjoint_equals_store_all(case):
    # adjoint_checkpoint = 1: the reverse sweep runs window by window (256 steps) from the states kept at every window start;
    # each window is replayed with the tape on right before its reverse sweep, so the tape holds one window instead of the
    # whole run.  Same windows, same arithmetic: the gradient must equal the store-all one (stream_min_mb = 0 gives the
    # store-all run the same 256-step routing windows).  France window: 600 steps = 2 full windows + 88 steps, pit pairs
    # across window boundaries included (the 700 x 700 crop holds several of the mesh's 2-cycles).
    if case == "france":
        def make():
            m = cases.france(T=600, sub=(200, 900, 200, 900), ngauge=3)
            random_fields(m, seed=21)
            return m
    else:
        def make():
            m = cases.cance(sparse=True, T=1440)
            random_fields(m, seed=22)
            return m
    a, pa, sa = _gradient_with({"adjoint_checkpoint": 1, "stream_min_mb": 0}, make)
    b, pb, sb = _gradient_with({"adjoint_checkpoint": 0, "stream_min_mb": 0}, make)
    assert np.isclose(float(a.output.cost), float(b.output.cost), rtol=1e-6)
    for g1, g2, names in ((pa, pb, ("cp", "cft", "exc", "lr")), (sa, sb, ("hp", "hft", "hlr"))):
        for n in names:
            x, y = np.asarray(getattr(g1, n), np.float64), np.asarray(getattr(g2, n), np.float64)
            scale = np.abs(y).max()
            assert np.abs(x - y).max() <= 1e-5 * scale + 1e-30, (n, float(np.abs(x - y).max()), float(scale))
    if case == "cance":
        import oracle
        c = make()
        pc, sc = ParametersDT(c.mesh), StatesDT(c.mesh)
        oracle.forward_b(c.setup, c.mesh, c.input_data, c.parameters, pc, c.parameters.copy(), c.states, sc, c.states.copy(), c.output)
        check_grad(pa, pc, ("cp", "cft", "exc", "lr"))
        check_grad(sa, sc, ("hp", "hft", "hlr"))


def test_sub_engine_agrees_with_row_passes():
    # the subtree engine (sub_kernels.cu: the engine's own cell order, routing inside the warp as a wavefront, only subtree
    # roots hand their series to other tiles) against the row-based passes on a 650 x 650 France crop with pit pairs; T = 100
    # ends in a partial window
    def model():
        m = cases.france(T=100, sub=(250, 900, 250, 900), ngauge=4)
        random_fields(m, seed=11)
        m.setup.save_net_prcp_domain = True
        m.output = type(m.output)(m.setup, m.mesh)
        return m
    a = _run_window(model(), {"sub_engine": 1, "sub_min_cells": 1000})
    b = _run_window(model(), {"sub_engine": 0})
    assert a.mesh.nac > 100000
    for name, x, y in (("qsim", a.output.qsim, b.output.qsim), ("qdom", a.output.sparse_qsim_domain, b.output.sparse_qsim_domain),
                       ("netp", a.output.sparse_net_prcp_domain, b.output.sparse_net_prcp_domain),
                       ("hlr", a.output.fstates.hlr, b.output.fstates.hlr), ("hp", a.output.fstates.hp, b.output.fstates.hp),
                       ("hft", a.output.fstates.hft, b.output.fstates.hft)):
        x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
        assert np.all(np.abs(x - y) <= 1e-7 + 1e-5 * np.abs(y)), (name, float(np.abs(x - y).max()))
    assert np.isclose(float(a.output.cost), float(b.output.cost), rtol=1e-5)


@pytest.mark.parametrize("case", ["france", "cance"])
def test_sub_engine_against_oracle(case):
    import oracle
    if case == "france":
        m = cases.france(T=50, sub=(400, 700, 400, 700), ngauge=3)
    else:
        m = cases.cance(sparse=True, T=1440)
        m.setup.save_qsim_domain = True
        m.output = type(m.output)(m.setup, m.mesh)
    random_fields(m, seed=13)
    c = m.copy()
    c.output = type(m.output)(m.setup, m.mesh)
    a = _run_window(m, {"sub_engine": 1, "sub_min_cells": 0})
    oracle.forward(c.setup, c.mesh, c.input_data, c.parameters, c.parameters.copy(), c.states, c.states.copy(), c.output)
    qa, qc = np.asarray(a.output.sparse_qsim_domain, np.float64), np.asarray(c.output.sparse_qsim_domain, np.float64)
    assert np.all(np.abs(qa - qc) <= 1e-4 + 2e-3 * np.abs(qc)), float(np.abs(qa - qc).max())
    assert np.all(np.abs(np.asarray(a.output.qsim, np.float64) - c.output.qsim) <= 1e-4 + 2e-3 * np.abs(c.output.qsim))
    for name in ("hp", "hft", "hlr"):
        x, y = np.asarray(getattr(a.output.fstates, name), np.float64), np.asarray(getattr(c.output.fstates, name), np.float64)
        assert np.all(np.abs(x - y) <= 1e-5 + 2e-3 * np.abs(y)), (name, float(np.abs(x - y).max()))
