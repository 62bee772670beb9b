"""Pins the CPU oracle (oracle/smash_oracle.c) to the reference's own golden file smash/tests/baseline.hdf5
(subset committed as tests/golden/cance_golden.npz by tests/golden/make_golden.py).  Tolerances are the ones the
reference's tests use: atol=1e-6 for costs (smash/tests/core/test_simu.py:25), atol=1e-4 for ensemble hydrographs
and costs (:53)."""
import numpy as np

import cases
import oracle


def test_run_cost(golden):
    # generic_run, test_simu.py:12-25: Model.run() leaves njf = 0 -> cost 0; then 1-NSE and KGE per gauge
    m = cases.cance()
    cases.set_optimize(m.setup, m.mesh, jobs_fun=())
    oracle.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(), m.output)
    got = cases.output_cost(m, oracle.nse, oracle.kge)
    assert np.allclose(got, golden["run.cost"], atol=1e-6)
    # states are restored, final states saved (forward.f90:71-72)
    assert np.all(m.states.hp == np.float32(0.01))
    assert not np.all(m.output.fstates.hp == np.float32(0.01))


def test_run_docstring_values():
    # Model.run docstring (smash/core/model.py:476-477): first / last qsim of the downstream gauge
    m = cases.cance()
    oracle.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(), m.output)
    q = m.output.qsim[0]
    assert np.allclose(q[:3], [1.9826449e-03, 1.3466686e-07, 6.7618025e-12], rtol=1e-5)
    assert np.allclose(q[-3:], [2.0916510e01, 2.0762346e01, 2.0610489e01], rtol=1e-4)


def _multiple_run(m, smp, nthreads=4):
    ns = smp.shape[1]
    cost = np.zeros(ns, np.float32)
    qsim = np.zeros((m.mesh.ng, m.setup._ntime_step, ns), np.float32, order="F")
    oracle.compute_multiple_run(m.setup, m.mesh, m.input_data, m.parameters, m.states, m.output, smp, cases.IND_CP_CFT_EXC_LR,
                                cost, qsim, nthreads=nthreads)
    return cost, qsim


def test_multiple_run(golden):
    # generic_multiple_run, test_simu.py:28-53
    m = cases.cance()
    smp = golden["samples.cp_cft_exc_lr"].astype(np.float32)
    assert np.allclose(smp[:, 0], [672.2786, 769.7930, -28.83132, 144.0110], rtol=1e-6)
    cost, qsim = _multiple_run(m, smp)
    assert np.allclose(cost, golden["multiple_run.cost"], atol=1e-4)
    assert np.allclose(qsim, golden["multiple_run.qsim"], atol=1e-4)
    for i in range(5):
        c, q = _multiple_run(m, smp[:, 2 * i:2 * i + 2])
        assert np.allclose(c, golden[f"mutiple_run.slc_{i + 1}.cost"], atol=1e-4)
        assert np.allclose(q, golden[f"mutiple_run.slc_{i + 1}.qsim"], atol=1e-4)


def test_sparse_equals_dense():
    # sparse_storage only changes where forcing is read from (md_forward_structure.f90:94-104)
    a, b = cases.cance(sparse=False, T=240), cases.cance(sparse=True, T=240)
    for m in (a, b):
        oracle.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(), m.output)
    assert np.array_equal(a.output.qsim, b.output.qsim)
    assert a.output.cost == b.output.cost
