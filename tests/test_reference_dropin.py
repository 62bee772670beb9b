"""The shim package ``smash_b200.solver`` under the reference's own, unmodified Python layer.

``smash.solver`` (f90wrap-generated in the reference, absent from its tree) is replaced by ``smash_b200.solver`` in
``sys.modules``; ``smash.core.model.Model`` is then imported from /root/reference (read-only, this container only) and driven
through ``Model.run`` / ``multiple_run`` / ``optimize`` on the Cance inputs, against the golden values of
``smash/tests/baseline.hdf5``.  There is no GPU here, so the entry points that compute (``forward`` ...) are backed by the CPU
oracle (tests/oracle_solver.py): the test pins the *surface* of the shim -- module names, constructors, attribute
semantics, in-place behaviour -- which is what decides whether the reference's callers run unchanged.  Third-party
packages the reference imports but this image lacks (osgeo, h5py, SALib, terminaltables) are stubbed: none of them is
on the exercised path."""
import functools
import importlib
import os
import sys
import types

import numpy as np
import pytest

import cases
import oracle_solver

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "smash", "core")), reason="reference tree not present")

SHIMS = ["_mwd_setup", "_mwd_mesh", "_mwd_input_data", "_mwd_parameters", "_mwd_states", "_mwd_output", "_mw_forward",
         "_mw_multiple_run", "_mw_optimize", "_mw_adjoint_test", "_mw_sparse_storage", "_mw_forcing_statistic",
         "_mw_interception_store", "_mw_mask", "_mw_derived_type_copy", "_mw_derived_type_update"]


@pytest.fixture(scope="module")
def ref_smash():
    saved = dict(sys.modules)
    for name, attrs in (("osgeo", {}), ("osgeo.gdal", {}), ("h5py", {}), ("SALib", {"ProblemSpec": object}),
                        ("terminaltables", {"AsciiTable": object})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
    sys.modules["osgeo"].gdal = sys.modules["osgeo.gdal"]
    pkg = types.ModuleType("smash")
    pkg.__path__ = [os.path.join(REF, "smash")]                     # sub-packages import from the reference, __init__ is skipped
    sys.modules["smash"] = pkg
    sol = types.ModuleType("smash.solver")
    sol.__path__ = []
    sys.modules["smash.solver"] = sol
    for name in SHIMS:
        mod = importlib.import_module("smash_b200.solver." + name)
        if name in ("_mw_forward", "_mw_multiple_run", "_mw_optimize", "_mw_adjoint_test"):
            clone = types.ModuleType("smash.solver." + name)         # same names, the oracle as the engine (no GPU here)
            clone.__dict__.update({k: v for k, v in vars(mod).items() if not k.startswith("__")})
            if name == "_mw_optimize":
                for fn in ("optimize_sbs", "optimize_lbfgsb", "optimize_hyper_lbfgsb"):
                    setattr(clone, fn, functools.partial(getattr(mod, fn), solver=oracle_solver))
            elif name == "_mw_adjoint_test":
                clone.scalar_product_test = functools.partial(mod.scalar_product_test, solver=oracle_solver)
            else:
                for fn in ("forward", "forward_b", "hyper_forward", "hyper_forward_b", "compute_multiple_run"):
                    if hasattr(clone, fn):
                        setattr(clone, fn, getattr(oracle_solver, fn))
            mod = clone
        sys.modules["smash.solver." + name] = mod
        setattr(sol, name, mod)
    try:
        yield importlib.import_module("smash.core.model")
    finally:
        for k in list(sys.modules):
            if k not in saved:
                del sys.modules[k]
        sys.modules.update(saved)


def cance_model(ref):
    """``smash.Model(*load_dataset("Cance"))`` without the raster readers: same setup / mesh dictionaries, forcing filled
    from the arrays the readers would produce (tests/golden/cance_inputs.npz)."""
    d = cases.golden("cance_inputs.npz")
    g = cases.golden("cance_golden.npz")
    setup = dict(structure="gr-a", dt=3600, start_time="2014-09-15 00:00", end_time="2014-11-14 00:00", read_qobs=False,
                 read_prcp=False, read_pet=False, read_descriptor=False, descriptor_name=["slope", "dd"])
    mesh = dict(dx=float(d["dx"]), nrow=int(d["nrow"]), ncol=int(d["ncol"]), ng=int(d["ng"]), nac=int(d["nac"]),
                xmin=float(np.ravel(g["mesh_io.xmin"])[0]), ymax=float(np.ravel(g["mesh_io.ymax"])[0]), flwdir=d["flwdir"], flwacc=d["flwacc"],
                flwdst=d["flwdst"], active_cell=d["active_cell"], path=d["path"], gauge_pos=d["gauge_pos"], area=d["area"],
                code=np.array([str(c) for c in d["code"]]))
    m = ref.Model(setup, mesh)
    c = cases.cance()
    m.input_data.qobs = c.input_data.qobs
    m.input_data.prcp = c.input_data.prcp
    m.input_data.pet = c.input_data.pet
    m.input_data.descriptor = c.input_data.descriptor
    sys.modules["smash.solver._mw_forcing_statistic"].compute_mean_forcing(m.setup, m.mesh, m.input_data)
    return m


def test_model_build_and_run(ref_smash, golden):
    m = cance_model(ref_smash)
    assert m.setup._ntime_step == 1440 and isinstance(m.setup._ntime_step, int)
    assert np.all(m.parameters.lr == np.float32(5.0)) and m.parameters.lr.shape == (28, 28)      # _build_model.py:257
    m.run(inplace=True)
    c = cases.cance()
    import oracle
    oracle.forward(c.setup, c.mesh, c.input_data, c.parameters, c.parameters.copy(), c.states, c.states.copy(), c.output)
    assert np.allclose(m.output.qsim, c.output.qsim, atol=1e-6)
    assert np.allclose(m.output.qsim[0, [0, -3, -2, -1]], [1.9826449e-03, 2.0916510e+01, 2.0762346e+01, 2.0610489e+01],
                       rtol=2e-6)                                                                  # Model.run docstring, model.py:476-477
    assert repr(m).endswith("Last update: Forward Run")
    assert float(m.output.cost) == 0.0                                                             # run() leaves njf = 0


def test_model_multiple_run(ref_smash, golden):
    m = cance_model(ref_smash)
    gs = importlib.import_module("smash.core.generate_samples")
    problem = {"num_vars": 4, "names": ["cp", "cft", "exc", "lr"], "bounds": [[1e-6, 1e3], [1e-6, 1e3], [-50, 50], [1e-6, 1e3]]}
    sample = gs.generate_samples(problem, n=10, random_state=99)
    res = m.multiple_run(sample, ncpu=1, return_qsim=True, verbose=False)
    assert np.allclose(res.cost, golden["multiple_run.cost"], atol=1e-4)
    assert np.allclose(res.qsim, golden["multiple_run.qsim"], atol=1e-4)


@pytest.mark.parametrize("key,kw", [
    ("optimize.uniform_sbs.cost", dict(mapping="uniform", algorithm="sbs", options={"maxiter": 1})),
    ("optimize.distributed_l-bfgs-b.cost", dict(mapping="distributed", algorithm="l-bfgs-b", options={"maxiter": 1})),
    ("optimize.hyper-linear_l-bfgs-b.cost", dict(mapping="hyper-linear", algorithm="l-bfgs-b", options={"maxiter": 1})),
    ("optimize.hyper-polynomial_l-bfgs-b.cost", dict(mapping="hyper-polynomial", algorithm="l-bfgs-b", options={"maxiter": 1})),
    ("optimize.uniform_sbs_mtg.cost", dict(gauge="all", wgauge="median", options={"maxiter": 1})),
    ("optimize.uniform_nelder-mead.cost",
     dict(mapping="uniform", algorithm="nelder-mead", jobs_fun=["nse", "Crc", "Cfp10", "Cfp50", "Epf", "Elt"],
          wjobs_fun=[1, 2, 2, 2, 2, 2], event_seg={"peak_quant": 0.99}, options={"maxiter": 10})),
    ("optimize.distributed_l-bfgs-b_reg_fast.cost",
     dict(mapping="distributed", control_vector=["cp", "cft", "lr"],
          options={"maxiter": 2, "jreg_fun": ["prior", "smoothing"], "wjreg_fun": [1.0, 2.0], "auto_wjreg": "fast"})),
])
def test_model_optimize(ref_smash, golden, key, kw):
    # the calls of the reference's own test (smash/tests/core/test_simu.py:76-171) through its own Model.optimize:
    # _standardize_* -> update_optimize_setup_* -> optimize_sbs / optimize_lbfgsb / optimize_hyper_lbfgsb of the shim
    import oracle
    m = cance_model(ref_smash)
    inst = m.optimize(verbose=False, **kw)
    got = cases.output_cost(inst, oracle.nse, oracle.kge)
    print(key, "max |cost - golden| =", np.abs(got - golden[key]).max())
    assert np.allclose(got, golden[key], atol=1e-5), (got, golden[key])
    assert float(m.output.cost) == 0.0 and inst is not m                 # inplace = False worked on a copy


def test_scalar_product_test_module(ref_smash):
    # mw_adjoint_test.scalar_product_test through the shim (directional derivative by central differences of forward
    # against the adjoint): <dY*, dY> and <dk*, dk> agree to float32 finite-difference accuracy
    spt = sys.modules["smash.solver._mw_adjoint_test"]
    c = cases.cance(T=240)
    cases.set_optimize(c.setup, c.mesh, jobs_fun=("nse",))
    sp1, sp2 = spt.scalar_product_test(c.setup, c.mesh, c.input_data, c.parameters, c.states, c.output, verbose=False)
    assert sp1 != 0.0 and abs(sp1 - sp2) <= 2e-2 * abs(sp1), (sp1, sp2)
