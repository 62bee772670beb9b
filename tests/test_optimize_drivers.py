"""Host-side optimiser drivers (smash_b200/solver/_mw_optimize.py, smash_b200/simulation.py) against the reference's
golden ``optimize.*`` values of smash/tests/baseline.hdf5 (generic_optimize, smash/tests/core/test_simu.py:77-189).

CPU tests drive the drivers with the oracle as the solver (tests/oracle_solver.py): they pin the *driver logic*
(control vector layout, transformations, L-BFGS-B settings, stopping rules, auto-wjreg) to the golden file.  The GPU
tests run the same cases through libsmash_b200.so.  Tolerance: the reference compares at atol = 1e-6 with its own
Fortran arithmetic; the float32 oracle under these drivers reproduces the costs to <= 1.8e-6 and the optimised
maps to ~1e-3 relative (the line search amplifies last-place differences), which is what is asserted here."""
import numpy as np
import pytest

import cases
import oracle
import oracle_solver
from smash_b200 import simulation

COST_ATOL = 5e-6      # CPU: driver logic on the float32 oracle (measured <= 1.8e-6)
COST_ATOL_GPU = 2e-5  # GPU: device arithmetic differs from the oracle in the last places (measured <= 6.9e-6 on B200)

CASES = {
    "optimize.uniform_sbs.cost": dict(mapping="uniform", algorithm="sbs", options={"maxiter": 1}),
    "optimize.distributed_l-bfgs-b.cost": dict(mapping="distributed", algorithm="l-bfgs-b", options={"maxiter": 1}),
    "optimize.hyper-linear_l-bfgs-b.cost": dict(mapping="hyper-linear", algorithm="l-bfgs-b", options={"maxiter": 1}),
    "optimize.hyper-polynomial_l-bfgs-b.cost": dict(mapping="hyper-polynomial", algorithm="l-bfgs-b",
                                                    options={"maxiter": 1}),
    "optimize.uniform_sbs_mtg.cost": dict(gauge="all", wgauge="median", options={"maxiter": 1}),
    "optimize.distributed_l-bfgs-b_reg_fast.cost": dict(
        mapping="distributed", control_vector=["cp", "cft", "lr"],
        options={"maxiter": 2, "jreg_fun": ["prior", "smoothing"], "wjreg_fun": [1.0, 2.0], "auto_wjreg": "fast"}),
    "optimize.distributed_l-bfgs-b_reg_lcurve.cost": dict(
        mapping="distributed", control_vector=["cp", "cft", "lr"],
        options={"maxiter": 2, "jreg_fun": ["prior", "smoothing"], "wjreg_fun": [1.0, 2.0], "auto_wjreg": "lcurve",
                 "nb_wjreg_lcurve": 8}),
}


def _cost(inst):
    return cases.output_cost(inst, oracle.nse, oracle.kge)


def _check_case(key, golden, solver):
    inst = simulation.optimize(cases.cance(), solver=solver, **CASES[key])
    got, want = _cost(inst), golden[key]
    atol = COST_ATOL if solver is not None else COST_ATOL_GPU
    print(key, "max |cost - golden| =", np.abs(got - want).max())
    assert np.allclose(got, want, atol=atol), (key, got, want)
    return inst


@pytest.mark.parametrize("key", list(CASES))
def test_optimize_golden_cpu(key, golden):
    _check_case(key, golden, oracle_solver)


def test_optimize_states_and_bounds_cpu(golden):
    # states in the control vector (test_simu.py:119-128)
    m = cases.cance()
    m.states.hlr[...] = 1.0
    inst = simulation.optimize(m, control_vector=["cp", "hlr"], options={"maxiter": 1}, solver=oracle_solver)
    assert np.allclose(_cost(inst), golden["optimize.uniform_sbs_states.cost"], atol=COST_ATOL)
    assert np.allclose(inst.states.hlr, golden["optimize.uniform_sbs_states.hlr"], rtol=1e-5)
    # user bounds (test_simu.py:131-142)
    inst = simulation.optimize(cases.cance(), mapping="distributed", algorithm="l-bfgs-b", control_vector=["cp", "cft"],
                               bounds={"cp": [1, 300]}, options={"maxiter": 1}, solver=oracle_solver)
    assert np.allclose(_cost(inst), golden["optimize.distributed_l-bfgs-b_bounds.cost"], atol=COST_ATOL)
    assert np.allclose(inst.parameters.cp, golden["optimize.distributed_l-bfgs-b_bounds.cp"], rtol=2e-3)
    assert np.allclose(inst.parameters.cft, golden["optimize.distributed_l-bfgs-b_bounds.cft"], rtol=2e-3)


def test_sbs_leaves_other_fields_alone_cpu():
    # test_simu.py:196-211: a distributed prior of a field that is not optimised survives optimize_sbs
    m = cases.cance(T=240)
    m.parameters.cft = np.asfortranarray(np.random.default_rng(0).random(m.parameters.cft.shape, dtype=np.float32) + 500)
    inst = simulation.optimize(m, control_vector="cp", options={"maxiter": 1}, solver=oracle_solver)
    assert np.array_equal(m.parameters.cft, inst.parameters.cft)
    assert not np.array_equal(m.parameters.cp, inst.parameters.cp)


def test_multiple_run_equals_forward_cpu():
    # test_simu.py:55-74: multiple_run of one member = optimize with maxiter 0 on the same uniform parameters
    m = cases.cance(T=240)
    smp = cases.golden("cance_golden.npz")["samples.cp_cft_exc_lr"].T[:2]
    cost, qsim = simulation.multiple_run(m, smp, ["cp", "cft", "exc", "lr"], return_qsim=True, solver=oracle_solver)
    for k in range(2):
        inst = m.copy()
        for name, v in zip(["cp", "cft", "exc", "lr"], smp[k]):
            getattr(inst.parameters, name)[...] = v
        inst = simulation.optimize(inst, options={"maxiter": 0}, solver=oracle_solver)
        assert np.allclose(inst.output.cost, cost[k], atol=1e-4)
        assert np.allclose(inst.output.qsim, qsim[..., k], atol=1e-4)


def test_argument_errors():
    m = cases.cance(T=24)
    with pytest.raises(ValueError):
        simulation.optimize(m, mapping="nope", solver=oracle_solver)
    with pytest.raises(ValueError):
        simulation.optimize(m, mapping="distributed", jobs_fun="kge", solver=oracle_solver)
    with pytest.raises(ValueError):
        simulation.optimize(m, control_vector=["cst"], solver=oracle_solver)
    with pytest.raises(ValueError):
        simulation.optimize(m, bounds={"cp": [300, 1]}, solver=oracle_solver)
    with pytest.raises(KeyError):
        simulation.optimize(m, options={"maxiter": 1, "bogus": 2}, solver=oracle_solver)


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("key", list(CASES))
def test_optimize_golden_gpu(key, golden):
    _check_case(key, golden, None)


@pytest.mark.gpu
def test_optimize_bounds_maps_gpu(golden):
    inst = simulation.optimize(cases.cance(), mapping="distributed", algorithm="l-bfgs-b", control_vector=["cp", "cft"],
                               bounds={"cp": [1, 300]}, options={"maxiter": 1})
    assert np.allclose(_cost(inst), golden["optimize.distributed_l-bfgs-b_bounds.cost"], atol=COST_ATOL_GPU)
    assert np.allclose(inst.parameters.cp, golden["optimize.distributed_l-bfgs-b_bounds.cp"], rtol=2e-3)
    assert np.allclose(inst.parameters.cft, golden["optimize.distributed_l-bfgs-b_bounds.cft"], rtol=2e-3)


@pytest.mark.gpu
def test_vda_converges_gpu():
    """BASELINE.json configs[1]: distributed-mapping variational calibration on Cance driven by the adjoint
    gradient.  30 L-BFGS-B iterations must take 1 - NSE well below the uniform first guess, and the GPU-driven
    optimisation must follow the oracle-driven one (same driver, same settings)."""
    g = simulation.optimize(cases.cance(), mapping="distributed", options={"maxiter": 30})
    first = simulation.run(cases.cance())
    j0 = float(first.output.cost)                       # 1 - NSE at the downstream gauge, uniform first guess
    assert g.output.cost < 0.5 * j0
    c = simulation.optimize(cases.cance(), mapping="distributed", options={"maxiter": 5})
    r = simulation.optimize(cases.cance(), mapping="distributed", options={"maxiter": 5}, solver=oracle_solver)
    assert abs(float(c.output.cost) - float(r.output.cost)) < 5e-3


# ------------------------------------------------------------------------------------------ Bayesian estimation
def test_generate_samples_golden(golden):
    # generate_samples(problem, n=10, random_state=99) of generic_multiple_run (test_simu.py:28-31)
    m = cases.cance(T=24)
    sr = simulation.generate_samples(simulation.get_bound_constraints(m), n=10, random_state=99)
    assert np.allclose(sr.to_numpy(axis=0), golden["samples.cp_cft_exc_lr"], rtol=1e-12)
    assert [s.n_sample for s in sr.iterslice(4)] == [4, 4, 2]


def _bayes_estimate(golden, solver, mr_solver, atol):
    # generic_bayes_estimate (test_simu.py:214-237)
    inst, br = simulation.bayes_estimate(cases.cance(), alpha=np.linspace(-1, 5, 10), n=5, return_br=True, random_state=11,
                                         solver=solver, mr_solver=mr_solver)
    got = np.array(br.lcurve["cost"])
    print("bayes_estimate.br_cost max diff", np.abs(got - golden["bayes_estimate.br_cost"]).max())
    assert np.allclose(got, golden["bayes_estimate.br_cost"], atol=atol)
    assert np.allclose(_cost(inst), golden["bayes_estimate.cost"], atol=atol)


def test_bayes_estimate_golden_cpu(golden):
    _bayes_estimate(golden, oracle_solver, oracle_solver, COST_ATOL)


def test_bayes_optimize_golden_cpu(golden):
    # generic_bayes_optimize (test_simu.py:240-268)
    inst, br = simulation.bayes_optimize(cases.cance(), alpha=np.linspace(-1, 5, 10), n=5, mapping="distributed",
                                         algorithm="l-bfgs-b", options={"maxiter": 1}, return_br=True, random_state=11,
                                         solver=oracle_solver)
    got = np.array(br.lcurve["cost"])
    print("bayes_optimize.br_cost max diff", np.abs(got - golden["bayes_optimize.br_cost"]).max())
    assert np.allclose(got, golden["bayes_optimize.br_cost"], atol=2e-5)
    assert np.allclose(_cost(inst), golden["bayes_optimize.cost"], atol=2e-5)


@pytest.mark.gpu
def test_bayes_estimate_golden_gpu(golden):
    _bayes_estimate(golden, None, None, COST_ATOL_GPU)


# ------------------------------------------------------------------------------------------ ANN mapping
def _ann_1(golden, solver, atol):
    # generic_ann_optimize_1 (test_simu.py:271-289): auto-graph 2 -> 18 -> 9 -> 4, Adam
    inst, net = simulation.ann_optimize(cases.cance(), epochs=5, learning_rate=0.001, return_net=True, random_state=11,
                                        solver=solver)
    assert [getattr(l, "neurons", None) for l in net.layers if hasattr(l, "neurons")] == [18, 9, 4] and net.n_params() == 265
    loss = np.array(net.history["loss_train"])
    print("ann_optimize_1.loss max diff", np.abs(loss - golden["ann_optimize_1.loss"]).max())
    assert np.allclose(loss, golden["ann_optimize_1.loss"], atol=atol)
    assert np.allclose(_cost(inst), golden["ann_optimize_1.cost"], atol=atol)


def _ann_2(golden, solver, atol):
    # generic_ann_optimize_2 (test_simu.py:292-330): user graph, SGD with momentum
    from smash_b200.net import Net
    m = cases.cance()
    problem = simulation.get_bound_constraints(m, states=False)
    net = Net()
    net.add(layer="dense", options={"input_shape": (2,), "neurons": 16})
    net.add(layer="activation", options={"name": "relu"})
    net.add(layer="dense", options={"neurons": 8})
    net.add(layer="activation", options={"name": "relu"})
    net.add(layer="dense", options={"neurons": problem["num_vars"]})
    net.add(layer="activation", options={"name": "sigmoid"})
    net.add(layer="scale", options={"bounds": problem["bounds"]})
    net.compile(optimizer="sgd", options={"learning_rate": 0.01, "momentum": 0.001}, random_state=11)
    inst = simulation.ann_optimize(m, net=net, epochs=5, solver=solver)
    loss = np.array(net.history["loss_train"])
    print("ann_optimize_2.loss max diff", np.abs(loss - golden["ann_optimize_2.loss"]).max())
    assert np.allclose(loss, golden["ann_optimize_2.loss"], atol=atol)
    assert np.allclose(_cost(inst), golden["ann_optimize_2.cost"], atol=atol)


def test_ann_optimize_golden_cpu(golden):
    _ann_1(golden, oracle_solver, COST_ATOL)
    _ann_2(golden, oracle_solver, COST_ATOL)


@pytest.mark.gpu
def test_ann_optimize_golden_gpu(golden):
    _ann_1(golden, None, COST_ATOL_GPU)
    _ann_2(golden, None, COST_ATOL_GPU)


def test_gen_samples_uniform_and_normal_golden(golden):
    # generic_gen_samples (smash/tests/core/test_gen_samples.py:9-35)
    m = cases.cance(T=24)
    problem = simulation.get_bound_constraints(m)
    uni = simulation.generate_samples(problem, generator="uniform", n=20, random_state=11).to_numpy(axis=-1)
    nor = simulation.generate_samples(problem, generator="normal", n=20,
                                      mean={problem["names"][1]: 1 / 3 * np.mean(problem["bounds"][1])}, coef_std=2,
                                      random_state=11).to_numpy(axis=-1)
    assert np.allclose(uni, golden["gen_samples.uni"], atol=1e-6)
    assert np.allclose(nor, golden["gen_samples.nor"], atol=1e-6)


def test_net_init_golden(golden):
    # generic_net_init (smash/tests/core/test_net.py:9-66): graph, He / Glorot weights of the legacy global generator
    from smash_b200.net import Net
    net = Net()
    for i in range(4):
        if i == 0:
            net.add(layer="dense", options={"input_shape": (6,), "neurons": 16, "kernel_initializer": "he_uniform"})
        else:
            net.add(layer="dense", options={"neurons": round(16 * (4 - i) / 4), "kernel_initializer": "he_uniform"})
        net.add(layer="activation", options={"name": "relu"})
        net.add(layer="dropout", options={"drop_rate": 0.1})
    net.add(layer="dense", options={"neurons": 2, "kernel_initializer": "glorot_uniform"})
    net.add(layer="activation", options={"name": "sigmoid"})
    net.compile(optimizer="adam", options={"learning_rate": 0.002, "b1": 0.8, "b2": 0.99}, random_state=11)
    graph = np.array([l.layer_name() for l in net.layers]).astype("S")
    assert np.array_equal(graph, golden["net_init.graph"])
    for i in range(4):
        layer = net.layers[3 * i]
        assert np.allclose(layer.weight, golden[f"net_init.weight_layer_{i + 1}"], atol=1e-6)
        assert np.allclose(layer.bias, golden[f"net_init.bias_layer_{i + 1}"], atol=1e-6)


# ---- the other structures under the derivative-free driver (their adjoint is not built: l-bfgs-b is refused) ---------------------

OTHER_STRUCTURES = ("gr-b", "gr-c", "gr-d", "vic-a")


def _sbs_other(structure, solver):
    m = cases.cance(T=480)
    m.setup.structure = structure
    m.parameters.ci[...] = 2.0
    before = m.copy()
    run = oracle_solver if solver is not None else __import__("smash_b200")
    run.forward(before.setup, before.mesh, before.input_data, before.parameters, before.parameters.copy(), before.states,
                before.states.copy(), before.output)
    inst = simulation.optimize(m, algorithm="sbs", options={"maxiter": 2}, solver=solver)
    return before, inst


@pytest.mark.parametrize("structure", OTHER_STRUCTURES)
def test_sbs_other_structures_cpu(structure):
    # default control vector = STRUCTURE_PARAMETERS[structure] (_constant.py:13-19), uniform mapping, two sbs iterations
    before, inst = _sbs_other(structure, oracle_solver)
    assert float(inst.output.cost) < float(before.output.cost)
    moved = [n for n in simulation.STRUCTURE_PARAMETERS[structure]
             if not np.array_equal(getattr(inst.parameters, n), getattr(before.parameters, n))]
    assert moved, "no parameter of the control vector moved"
    untouched = set(simulation.GPARAMETERS_NAME) - set(simulation.STRUCTURE_PARAMETERS[structure])
    for n in untouched:
        assert np.array_equal(getattr(inst.parameters, n), getattr(before.parameters, n)), n


@pytest.mark.gpu
@pytest.mark.parametrize("structure", OTHER_STRUCTURES)
def test_sbs_other_structures_gpu(structure):
    # the same search through libsmash_b200.so: same path through the parameter space, same cost to the device tolerance
    _, cpu = _sbs_other(structure, oracle_solver)
    _, gpu = _sbs_other(structure, None)
    assert abs(float(gpu.output.cost) - float(cpu.output.cost)) < (5e-4 if structure == "vic-a" else COST_ATOL_GPU * 5)
    for n in simulation.STRUCTURE_PARAMETERS[structure]:
        a, b = getattr(gpu.parameters, n), getattr(cpu.parameters, n)
        assert np.allclose(a, b, rtol=2e-2, atol=1e-3), (n, float(a.max()), float(b.max()))


@pytest.mark.gpu
def test_lbfgsb_is_refused_for_other_structures():
    m = cases.cance(T=48)
    m.setup.structure = "gr-b"
    with pytest.raises(RuntimeError, match="gr-a and gr-d only"):
        simulation.optimize(m, mapping="distributed", algorithm="l-bfgs-b", options={"maxiter": 1})


def _lbfgsb_gr_d(solver):
    m = cases.cance()
    m.setup.structure = "gr-d"
    return simulation.optimize(m, mapping="distributed", algorithm="l-bfgs-b", options={"maxiter": 2}, solver=solver)


def test_lbfgsb_gr_d_cpu():
    # gr-d has an adjoint (GR_D_FORWARD_B): the variational driver runs; default control vector cp, cft, lr
    m = cases.cance()
    m.setup.structure = "gr-d"
    oracle_solver.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(), m.output)
    inst = _lbfgsb_gr_d(oracle_solver)
    assert float(inst.output.cost) < float(m.output.cost)
    assert np.array_equal(inst.parameters.exc, m.parameters.exc)


@pytest.mark.gpu
def test_lbfgsb_gr_d_gpu():
    cpu, gpu = _lbfgsb_gr_d(oracle_solver), _lbfgsb_gr_d(None)
    assert abs(float(gpu.output.cost) - float(cpu.output.cost)) < 1e-4
    for n in ("cp", "cft", "lr"):
        a, b = getattr(gpu.parameters, n), getattr(cpu.parameters, n)
        assert np.allclose(a, b, rtol=2e-2, atol=1e-2), n
