"""Preprocessing on the device (csrc/pre_kernels.cu): flow accumulation, gauge masks and catchment
means of the forcing against the golden mesh values and the host restatements."""
import numpy as np
import pytest

import cases
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
