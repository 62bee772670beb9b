"""GPU parity of the structures other than gr-a (gr-b, gr-c, gr-d, vic-a; md_forward_structure.f90:216-931): the CUDA reservoir pass
of struct_kernels.cu + the shared routing passes, through the C ABI, against the CPU oracle's restatement on the same inputs.

The reference ships no golden vector of these structures (smash/tests/baseline.hdf5 holds gr-a only), so the oracle side of this file
is UNPINNED: agreement here means "the device arithmetic is the oracle's restatement", not "is the reference's".
Tolerances: those of test_gpu_parity.py (discharge |d| <= 1e-4 + 2e-3 |ref|, cost abs 1e-5, final stores rtol 1e-4)."""
import numpy as np
import pytest

import cases
import oracle
import smash_b200
from smash_b200.simulation import STRUCTURE_STATES
from smash_b200.solver._derived_types import OutputDT, ParametersDT, StatesDT

pytestmark = pytest.mark.gpu

STRUCTURES = ("gr-b", "gr-c", "gr-d", "vic-a")
PLANES = {   # fields drawn at random per active cell: name, low, high
    "gr-b": (("ci", 0.5, 5), ("cp", 50, 600), ("cft", 50, 800), ("exc", -5, 5), ("lr", 1, 30)),
    "gr-c": (("ci", 0.5, 5), ("cp", 50, 600), ("cft", 50, 800), ("cst", 100, 2000), ("exc", -5, 5), ("lr", 1, 30)),
    "gr-d": (("cp", 50, 600), ("cft", 50, 800), ("lr", 1, 30)),
    "vic-a": (("b", 0.05, 1.0), ("cusl1", 20, 300), ("cusl2", 100, 800), ("clsl", 500, 1800), ("ks", 1, 40), ("ds", 0.01, 0.5),
              ("dsm", 0.1, 5), ("ws", 0.3, 0.95), ("lr", 1, 30)),
}


def close_q(a, b, c=None):
    """|a - b| <= 1e-4 + 2e-3 |b| -- or, where the float64 build `c` of the oracle is given, `a` within twice the float32 build's own
    distance to it (plus the same tolerance).  The power-law stores (gr_transfer md_gr_operator.f90:104-106, vic_interflow
    md_vic_operator.f90:130-132) form (h - h') * c with h' = h (1 - 1e-8 .. 1e-4): float32 keeps zero to four digits of that
    difference, the same rounding hits every cell that shares the parameter values and adds up along the rivers, so on low flows the
    reference arithmetic is only reproducible to its own float32-vs-float64 distance (measured: up to 1 % of a 0.07 m3/s flow for
    vic-a with default parameters, tools/diag_struct.py)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    ok = np.abs(a - b) <= 1e-4 + 2e-3 * np.abs(b)
    if c is not None:
        c = np.asarray(c, np.float64)
        ok |= np.abs(a - c) <= 1e-4 + 2e-3 * np.abs(c) + 2.0 * np.abs(b - c)
    return bool(np.all(ok))


def randomize(m, structure, seed=3):
    rng = np.random.default_rng(seed)
    act = m.mesh.active_cell == 1
    for name, lo, hi in PLANES[structure]:
        getattr(m.parameters, name)[act] = rng.uniform(lo, hi, int(act.sum())).astype(np.float32)
    for name in STRUCTURE_STATES[structure]:
        if name != "hlr":
            getattr(m.states, name)[act] = rng.uniform(0.05, 0.6, int(act.sum())).astype(np.float32)


def pair(make, structure, random=True, **opt):
    a, b = make(), make()
    for m in (a, b):
        m.setup.structure = structure
        if opt:
            cases.set_optimize(m.setup, m.mesh, **opt)
        if random:
            randomize(m, structure)
    smash_b200.forward(a.setup, a.mesh, a.input_data, a.parameters, a.parameters.copy(), a.states, a.states.copy(), a.output)
    oracle.forward(b.setup, b.mesh, b.input_data, b.parameters, b.parameters.copy(), b.states, b.states.copy(), b.output)
    return a, b


def exact(make, structure, random=True):
    """The float64 build of the oracle on the same case (referee of close_q)."""
    c = make()
    c.setup.structure = structure
    if random:
        randomize(c, structure)
    oracle.forward(c.setup, c.mesh, c.input_data, c.parameters, c.parameters.copy(), c.states, c.states.copy(), c.output, precision="f64")
    return c


def check_states(a, b, structure):
    for n in set(STRUCTURE_STATES[structure]) | {"hlr"}:
        x, y = getattr(a.output.fstates, n), getattr(b.output.fstates, n)
        # vic-a: the float32 and float64 builds of the oracle already differ by 5e-6 .. 1e-5 in these stores after 1440 steps
        assert np.allclose(x, y, rtol=1e-4, atol=2e-6 if structure != "vic-a" else 2e-5), (n, float(np.abs(x - y).max()))
        assert np.array_equal(getattr(a.states, n), getattr(b.states, n)), n       # states restored (forward.f90:72)


@pytest.mark.parametrize("structure", STRUCTURES)
@pytest.mark.parametrize("random", [False, True])
def test_forward_cance_vs_oracle(structure, random):
    a, b = pair(cases.cance, structure, random=random, jobs_fun=("nse", "kge"))
    c = exact(cases.cance, structure, random=random)
    assert np.abs(b.output.qsim).max() > 1.0                                       # the case produces a flood
    assert close_q(a.output.qsim, b.output.qsim, c.output.qsim), float(np.abs(a.output.qsim - b.output.qsim).max())
    assert abs(float(a.output.cost) - float(b.output.cost)) < (1e-5 if structure != "vic-a" else 2e-4)
    check_states(a, b, structure)


def test_vic_a_reference_order_arithmetic():
    # option math = 0: IEEE division / sqrt and the reference's own (cancelling) form of the power-law stores -- the float32 build of
    # the oracle is then met with the plain tolerance, without the float64 referee
    lib = smash_b200._lib.lib()
    lib.smash_b200_set_option(b"math", 0)
    try:
        a, b = pair(cases.cance, "vic-a", random=True)
    finally:
        lib.smash_b200_set_option(b"math", 1)
    assert close_q(a.output.qsim, b.output.qsim), float(np.abs(a.output.qsim - b.output.qsim).max())
    assert abs(float(a.output.cost) - float(b.output.cost)) < 2e-5
    check_states(a, b, "vic-a")


@pytest.mark.parametrize("structure", ["gr-c", "vic-a"])
def test_forward_cance_sparse_domain_outputs(structure):
    def make():
        m = cases.cance(sparse=True, T=480)
        m.setup.save_qsim_domain = True
        m.setup.save_net_prcp_domain = True
        m.output = OutputDT(m.setup, m.mesh)
        return m
    a, b = pair(make, structure)
    c = exact(make, structure)
    assert close_q(a.output.sparse_qsim_domain, b.output.sparse_qsim_domain, c.output.sparse_qsim_domain)
    assert close_q(a.output.sparse_net_prcp_domain, b.output.sparse_net_prcp_domain, c.output.sparse_net_prcp_domain)
    assert close_q(a.output.qsim, b.output.qsim, c.output.qsim)


@pytest.mark.parametrize("structure", STRUCTURES)
def test_forward_forcing_gaps(structure):
    # missing forcing (prcp or pet < 0, md_forward_structure.f90:288 etc.): the production part is skipped, the transfer stores empty
    def make():
        m = cases.cance(T=240)
        rng = np.random.default_rng(11)
        hole = rng.random(m.input_data.prcp.shape[2]) < 0.08
        m.input_data.prcp[..., hole] = -99.0
        m.input_data.pet[..., rng.random(m.input_data.pet.shape[2]) < 0.05] = -99.0
        return m
    a, b = pair(make, structure)
    c = exact(make, structure)
    assert close_q(a.output.qsim, b.output.qsim, c.output.qsim), float(np.abs(a.output.qsim - b.output.qsim).max())
    check_states(a, b, structure)


@pytest.mark.parametrize("structure", ["gr-b", "gr-d", "vic-a"])
def test_forward_france_crop_vs_oracle(structure):
    # a France crop with pit pairs and long rivers: the routing passes behind the new reservoir pass, sparse storage
    def make():
        return cases.france(T=24, sub=(300, 700, 300, 700))
    a, b = pair(make, structure)
    qa, qb = a.output.sparse_qsim_domain, b.output.sparse_qsim_domain
    assert qa.shape == qb.shape and qa.shape[0] > 50000
    assert close_q(qa, qb, exact(make, structure).output.sparse_qsim_domain), float(np.abs(qa - qb).max())
    check_states(a, b, structure)


def test_multiple_run_gr_c_vs_oracle():
    # an ensemble over planes gr-a does not have (ci, cst) next to cp and lr (mw_multiple_run.f90:40-65)
    m = cases.cance(T=480)
    m.setup.structure = "gr-c"
    names = ("ci", "cp", "cst", "lr")
    ind = np.array([1 + smash_b200.solver._derived_types.GPARAMETERS_NAME.index(n) for n in names], np.int32)
    rng = np.random.default_rng(5)
    smp = np.asfortranarray(np.stack([rng.uniform(0.5, 5, 6), rng.uniform(50, 600, 6), rng.uniform(100, 2000, 6), rng.uniform(1, 30, 6)])
                            .astype(np.float32))
    cost = np.zeros(6, np.float32)
    qsim = np.zeros((3, 480, 6), np.float32, order="F")
    smash_b200.compute_multiple_run(m.setup, m.mesh, m.input_data, m.parameters, m.states, m.output, smp, ind, cost, qsim)
    want_cost = np.zeros(6, np.float32)
    want_q = np.zeros((3, 480, 6), np.float32, order="F")
    oracle.compute_multiple_run(m.setup, m.mesh, m.input_data, m.parameters, m.states, m.output, smp, ind, want_cost, want_q)
    ref_cost, ref_q = np.zeros(6, np.float64), np.zeros((3, 480, 6), np.float64, order="F")
    oracle.compute_multiple_run(m.setup, m.mesh, m.input_data, m.parameters, m.states, m.output, smp, ind, ref_cost, ref_q, precision="f64")
    assert np.allclose(cost, want_cost, atol=1e-4, rtol=1e-5)
    assert close_q(qsim, want_q, ref_q)


@pytest.mark.parametrize("jobs", [("nse",), ("kge",)])
def test_gr_d_gradient_vs_oracle(jobs):
    # gr-d runs on gr-a's kernels with the shares 1 / 0 and no exchange (SplitArgs::grd): forward tape + reverse sweep against the
    # oracle's restatement of GR_D_FORWARD_B (forward_db.f90:9604-9797), tolerances of test_gpu_parity.py
    from test_gpu_parity import check_grad, gradients, random_fields
    a, b = cases.cance(), cases.cance()
    for m in (a, b):
        m.setup.structure = "gr-d"
        cases.set_optimize(m.setup, m.mesh, jobs_fun=jobs, gauge="all")
        random_fields(m)
    pa, sa = gradients(a, "gpu")
    pb, sb = gradients(b, "cpu")
    check_grad(pa, pb, ("cp", "cft", "lr"))
    check_grad(sa, sb, ("hp", "hft", "hlr"))
    assert not np.any(pa.exc) and not np.any(pb.exc)                           # exc is not a parameter of gr-d
    assert abs(float(a.output.cost) - float(b.output.cost)) < 1e-5


@pytest.mark.parametrize("structure", ["gr-b", "gr-c", "vic-a"])
def test_adjoint_is_refused(structure):
    # GR_A_FORWARD_B and GR_D_FORWARD_B are built: for the others the library says so instead of returning a gr-a gradient
    m = cases.cance(T=48)
    m.setup.structure = structure
    pb, sb = ParametersDT(m.mesh), StatesDT(m.mesh)
    with pytest.raises(RuntimeError, match="gr-a and gr-d only"):
        smash_b200.forward_b(m.setup, m.mesh, m.input_data, m.parameters, pb, m.parameters.copy(), None, m.states, sb, m.states.copy(),
                             None, m.output, None)


def test_structures_differ():
    # the five structures are five different models: no silent gr-a run behind another name
    q = {}
    for s in ("gr-a",) + STRUCTURES:
        m = cases.cance()
        m.setup.structure = s
        m.parameters.ci[...] = 2.0
        m.parameters.exc[...] = -0.5
        smash_b200.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(), m.output)
        q[s] = np.array(m.output.qsim)
    keys = list(q)
    for i in range(len(keys)):
        for j in range(i + 1, len(keys)):
            assert np.abs(q[keys[i]] - q[keys[j]]).max() > 0.5, (keys[i], keys[j])
