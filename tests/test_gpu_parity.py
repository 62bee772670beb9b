"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes shim in smash_b200.solver), against the
CPU oracle on the same inputs and against the committed golden vectors.

Stated tolerances:
  discharge  |d| <= 1e-4 + 2e-3*|ref|  (atol = the reference's own test tolerance, smash/tests/core/test_simu.py:53;
             the relative part is twice the measured float32-vs-float64 distance of the reference arithmetic itself);
             against the golden file the reference's own np.allclose(atol=1e-4) is used unchanged.  The float32 model itself is only
             reproducible to ~7e-4 relative on Cance: that is the distance between the f32 and f64 builds of the oracle
             (tools/diag_parity.py), i.e. the rounding noise of the reference's real kind through 1440 nonlinear steps.
  cost       abs 1e-5
  gradients  rel inf-norm 2e-3 per field and cosine >= 0.9995 against the Tapenade restatement (the f32 and f64 builds of
             the restatement differ from each other by up to 7e-4 / cosine 0.99995 on the same cases, tools/diag_grad.py)."""
import numpy as np
import pytest

import cases
import oracle
import smash_b200
from smash_b200.solver._derived_types import ParametersDT, StatesDT

pytestmark = pytest.mark.gpu


def close_q(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return bool(np.all(np.abs(a - b) <= 1e-4 + 2e-3 * np.abs(b)))


def run_both(T=None, sparse=False, **opt):
    a, b = cases.cance(sparse=sparse, T=T), cases.cance(sparse=sparse, T=T)
    for m in (a, b):
        cases.set_optimize(m.setup, m.mesh, **opt)
    smash_b200.forward(a.setup, a.mesh, a.input_data, a.parameters, a.parameters.copy(), a.states, a.states.copy(), a.output)
    oracle.forward(b.setup, b.mesh, b.input_data, b.parameters, b.parameters.copy(), b.states, b.states.copy(), b.output)
    return a, b


@pytest.mark.parametrize("sparse", [False, True])
def test_forward_cance_vs_oracle(sparse):
    a, b = run_both(sparse=sparse, jobs_fun=("nse",))
    assert close_q(a.output.qsim, b.output.qsim)
    assert abs(float(a.output.cost) - float(b.output.cost)) < 1e-5
    for n in ("hp", "hft", "hlr"):
        assert np.allclose(getattr(a.output.fstates, n), getattr(b.output.fstates, n), rtol=1e-4, atol=1e-7), n
        assert np.array_equal(getattr(a.states, n), getattr(b.states, n)), n      # states restored


def test_forward_cance_vs_golden(golden):
    a, _ = run_both(jobs_fun=())
    got = cases.output_cost(a, oracle.nse, oracle.kge)
    assert np.allclose(got, golden["run.cost"], atol=1e-5)


@pytest.mark.parametrize("jobs", [("kge",), ("nse", "kge"), ("kge2",), ("se",), ("rmse",), ("logarithmic",)])
def test_cost_functions(jobs):
    # the cost kernel in isolation: oracle compute_jobs evaluated on the GPU's own hydrographs
    a, b = run_both(T=480, jobs_fun=jobs, gauge="all")
    want = oracle.compute_jobs(a.setup, a.mesh, a.input_data, a.output.qsim)
    assert np.isclose(float(a.output.cost), float(want), rtol=2e-6, atol=1e-6)
    if jobs != ("logarithmic",):   # sum of x*log(y/x)^2 is dominated by near-zero flows: not comparable across runs
        assert np.isclose(float(a.output.cost), float(b.output.cost), rtol=2e-4, atol=1e-5)


def test_cost_median_gauges():
    a, b = run_both(T=480, jobs_fun=("nse",), wgauge=[-1.0, -1.0, -1.0])
    assert np.isclose(float(a.output.cost), float(b.output.cost), rtol=2e-5, atol=1e-5)


def test_multiple_run_vs_golden(golden):
    m = cases.cance()
    smp = golden["samples.cp_cft_exc_lr"].astype(np.float32)
    cost = np.zeros(10, np.float32)
    qsim = np.zeros((3, 1440, 10), np.float32, order="F")
    smash_b200.compute_multiple_run(m.setup, m.mesh, m.input_data, m.parameters, m.states, m.output, smp,
                                    cases.IND_CP_CFT_EXC_LR, cost, qsim)
    assert np.allclose(cost, golden["multiple_run.cost"], atol=1e-4, rtol=1e-5)
    assert np.allclose(qsim, golden["multiple_run.qsim"], atol=1e-4, rtol=1e-4)


def gradients(model, backend, **kw):
    pb, sb = ParametersDT(model.mesh), StatesDT(model.mesh)
    if backend == "gpu":
        smash_b200.forward_b(model.setup, model.mesh, model.input_data, model.parameters, pb, model.parameters.copy(), None,
                             model.states, sb, model.states.copy(), None, model.output, None)
    else:
        oracle.forward_b(model.setup, model.mesh, model.input_data, model.parameters, pb, model.parameters.copy(),
                         model.states, sb, model.states.copy(), model.output, **kw)
    return pb, sb


def random_fields(m, seed=1):
    rng = np.random.default_rng(seed)
    act = m.mesh.active_cell == 1
    for name, lo, hi in (("cp", 50, 600), ("cft", 50, 800), ("exc", -5, 5), ("lr", 1, 30)):
        f = getattr(m.parameters, name)
        f[act] = rng.uniform(lo, hi, int(act.sum())).astype(np.float32)


def check_grad(ga, gb, names):
    for n in names:
        x, y = np.asarray(getattr(ga, n), np.float64), np.asarray(getattr(gb, n), np.float64)
        scale = np.abs(y).max()
        assert np.abs(x - y).max() <= 2e-3 * scale + 1e-12, (n, np.abs(x - y).max(), scale)
        if scale > 0:
            cos = (x * y).sum() / np.sqrt((x * x).sum() * (y * y).sum())
            assert cos >= 0.9995, (n, cos)


@pytest.mark.parametrize("jobs", [("nse",), ("kge",)])
def test_gradient_cance_vs_oracle(jobs):
    a, b = cases.cance(T=720), cases.cance(T=720)
    for m in (a, b):
        cases.set_optimize(m.setup, m.mesh, jobs_fun=jobs)
        random_fields(m)
    pa, sa = gradients(a, "gpu")
    pb, sb = gradients(b, "cpu")
    assert abs(float(a.output.cost) - float(b.output.cost)) < 1e-5
    check_grad(pa, pb, ("cp", "cft", "exc", "lr"))
    check_grad(sa, sb, ("hp", "hft", "hlr"))
    for n in ("ci", "beta", "cst", "alpha"):
        assert not np.any(getattr(pa, n))
    inactive = a.mesh.active_cell == 0
    assert not np.any(pa.cp[inactive])


def test_forward_france_vs_oracle():
    # France 1 km mesh (906 044 cells, 50 pit pairs, 3 540 blocks with cross-block flags), 24 synthetic steps
    a, b = cases.france(T=24), cases.france(T=24)
    smash_b200.forward(a.setup, a.mesh, a.input_data, a.parameters, a.parameters.copy(), a.states, a.states.copy(), a.output)
    oracle.forward(b.setup, b.mesh, b.input_data, b.parameters, b.parameters.copy(), b.states, b.states.copy(), b.output)
    qa, qb = a.output.sparse_qsim_domain, b.output.sparse_qsim_domain
    assert qa.shape == qb.shape
    assert close_q(qa, qb), float(np.abs(qa - qb).max())
    for n in ("hp", "hft", "hlr"):
        assert np.allclose(getattr(a.output.fstates, n), getattr(b.output.fstates, n), rtol=1e-4, atol=1e-7), n
