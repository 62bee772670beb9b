"""Preprocessing on the device (csrc/pre_kernels.cu): flow accumulation, gauge masks and catchment
means of the forcing against the golden mesh values and the host restatements."""
import numpy as np
import pytest

import cases
import oracle
import smash_b200
from smash_b200.mesh import flow_accumulation, flow_accumulation_device
from smash_b200.solver import _mw_forcing_statistic as FS

pytestmark = pytest.mark.gpu


def test_flow_accumulation_on_device_is_integer_exact(golden):
    # mw_meshing.f90:204-233 on the GPU: the golden flwacc of the Cance window (catchment mask) and of the France mesh
    # (1.24 M cells, 288 pit cells), integer for integer, and the host restatement on a random sub-window
    act = golden["mesh_io.active_cell"] == 1
    fa = flow_accumulation_device(golden["mesh_io.flwdir"], mask=act)
    assert np.array_equal(fa[act], golden["mesh_io.flwacc"][act])
    f = cases.golden("france_mesh.npz")
    fa = flow_accumulation_device(f["flwdir"])
    actf = f["active_cell"] == 1
    assert np.array_equal(fa[actf], f["flwacc"][actf])
    assert int(fa.max()) == 136170
    sub = np.ascontiguousarray(f["flwdir"][300:700, 250:800])
    assert np.array_equal(flow_accumulation_device(sub), flow_accumulation(sub))


@pytest.mark.parametrize("sparse", [False, True])
def test_mean_forcing_on_device(sparse):
    # mw_forcing_statistic.f90:18-75: catchment means per gauge and time step; the device sums in float64 and rounds once,
    # the host restatement sums in float32 like the Fortran: 1e-5 relative
    m = cases.cance(sparse=sparse, T=240)
    rng = np.random.RandomState(5)
    if sparse:                                                            # a few gaps (negative = no data) must be skipped
        m.input_data.sparse_prcp[rng.randint(0, m.mesh.nac, 40), rng.randint(0, 240, 40)] = -99.0
    else:
        m.input_data.prcp[rng.randint(0, m.mesh.nrow, 40), rng.randint(0, m.mesh.ncol, 40), rng.randint(0, 240, 40)] = -99.0
    FS.compute_mean_forcing(m.setup, m.mesh, m.input_data)
    hp, he = m.input_data.mean_prcp.copy(), m.input_data.mean_pet.copy()
    m.input_data.mean_prcp[...] = 0; m.input_data.mean_pet[...] = 0
    FS.compute_mean_forcing_device(m.setup, m.mesh, m.input_data)
    assert np.allclose(m.input_data.mean_prcp, hp, rtol=1e-5, atol=1e-7)
    assert np.allclose(m.input_data.mean_pet, he, rtol=1e-5, atol=1e-7)
    assert float(hp.max()) > 0.1
    masks = FS.gauge_masks_device(m.mesh, m.setup)
    assert np.array_equal(masks, FS._gauge_masks(m.mesh))


# ---- adjust_interception_store (mw_interception_store.f90:19-160): integer-like contract, the chosen capacity per cell ---------

def _days(T, steps_per_day=24):
    return (np.arange(T) // steps_per_day + 1).astype(np.int32)


@pytest.mark.parametrize("sparse", [False, True])
def test_interception_store_cance(sparse):
    from smash_b200.solver import _mw_interception_store as dev
    a, b = cases.cance(sparse=sparse), cases.cance(sparse=sparse)
    nday = 1440 // 24
    dev.adjust_interception_store(a.setup, a.mesh, a.input_data, a.parameters, nday, _days(1440))
    oracle.adjust_interception_store(b.setup, b.mesh, b.input_data, b.parameters, nday, _days(1440))
    ca, cb = np.asarray(a.parameters.ci), np.asarray(b.parameters.ci)
    act = a.mesh.active_cell == 1
    assert np.array_equal(ca, cb)                                              # bit-exact: every statement is one rounded operation
    assert len(np.unique(ca[act])) > 3 and np.all(ca[~act] == np.float32(1e-6))
    assert dev.last_kernel_ms > 0.0


def test_interception_store_france_crop_with_gaps_and_short_last_day():
    from smash_b200.solver import _mw_interception_store as dev

    def make():
        m = cases.france(T=60, sub=(300, 600, 300, 600))
        m.input_data.sparse_prcp[::7, 13] = -99.0                              # gaps go through the same arithmetic (:117 has no test)
        return m
    a, b = make(), make()
    di = _days(60)                                                             # 2 days and a half
    dev.adjust_interception_store(a.setup, a.mesh, a.input_data, a.parameters, 3, di)
    oracle.adjust_interception_store(b.setup, b.mesh, b.input_data, b.parameters, 3, di)
    assert np.array_equal(np.asarray(a.parameters.ci), np.asarray(b.parameters.ci))
    with pytest.raises(RuntimeError, match="nday"):
        dev.adjust_interception_store(a.setup, a.mesh, a.input_data, a.parameters, 2, di)


def test_gr_b_run_with_the_calibrated_capacity():
    # the caller's sequence (_build_model.py: adjust the store, then run): ci from the device search feeds the gr-b forward run
    from smash_b200.solver import _mw_interception_store as dev
    a, b = cases.cance(), cases.cance()
    for m in (a, b):
        m.setup.structure = "gr-b"
    dev.adjust_interception_store(a.setup, a.mesh, a.input_data, a.parameters, 60, _days(1440))
    oracle.adjust_interception_store(b.setup, b.mesh, b.input_data, b.parameters, 60, _days(1440))
    smash_b200.forward(a.setup, a.mesh, a.input_data, a.parameters, a.parameters.copy(), a.states, a.states.copy(), a.output)
    oracle.forward(b.setup, b.mesh, b.input_data, b.parameters, b.parameters.copy(), b.states, b.states.copy(), b.output)
    qa, qb = np.asarray(a.output.qsim, np.float64), np.asarray(b.output.qsim, np.float64)
    assert np.all(np.abs(qa - qb) <= 1e-4 + 2e-3 * np.abs(qb))
