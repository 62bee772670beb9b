"""CPU checks of the oracle's restatement of gr-b, gr-c, gr-d and vic-a (md_forward_structure.f90:216-931, md_vic_operator.f90).

The reference holds no golden vector of these structures (smash/tests/baseline.hdf5 and test_simu.py run gr-a only) and its Fortran
cannot be compiled here, so this part of the oracle is UNPINNED.  What can be checked without the reference: the reduction to the
golden-pinned gr-a where the structure has one, float32 against float64 builds, and the bookkeeping of forward.f90."""
import numpy as np
import pytest

import cases
import oracle
from smash_b200.simulation import STRUCTURE_PARAMETERS, STRUCTURE_STATES
from smash_b200.solver._derived_types import GSTATES_NAME, ParametersDT, StatesDT

OTHERS = ("gr-b", "gr-c", "gr-d", "vic-a")


def run(structure, T=1440, precision="f32", **planes):
    m = cases.cance(T=T)
    m.setup.structure = structure
    for k, v in planes.items():
        getattr(m.parameters, k)[...] = v
    oracle.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(), m.output,
                   precision=precision)
    return m


def test_structure_tables_match_reference_constants():
    # smash/core/_constant.py:13-33 (vic-a's state list has no hlr there either)
    assert STRUCTURE_PARAMETERS["gr-c"] == ["cp", "cft", "cst", "exc", "lr"]
    assert STRUCTURE_PARAMETERS["vic-a"] == ["b", "cusl1", "cusl2", "clsl", "ks", "ds", "dsm", "ws", "lr"]
    assert STRUCTURE_STATES["gr-b"] == ["hi", "hp", "hft", "hlr"]
    assert STRUCTURE_STATES["vic-a"] == ["husl1", "husl2", "hlsl"]
    assert set(oracle.STRUCTURES) == set(STRUCTURE_PARAMETERS)


def test_gr_b_reduces_to_gr_a_without_interception():
    # ci -> 0 empties the interception store every step: gr_interception (md_gr_operator.f90:20-34) becomes the
    # ei = min(pet, prcp), pn = max(0, prcp - ei) of gr_a_forward (md_forward_structure.f90:112-116), whose run is pinned by the golden file
    a, b = run("gr-a", exc=-1.0), run("gr-b", exc=-1.0, ci=1e-6)
    qa, qb = np.asarray(a.output.qsim), np.asarray(b.output.qsim)
    assert qa.max() > 50
    assert np.abs(qa - qb).max() <= 2e-5 * qa.max()
    assert np.allclose(a.output.fstates.hp, b.output.fstates.hp, atol=2e-6)


@pytest.mark.parametrize("structure", OTHERS)
def test_f32_against_f64_build(structure):
    x, y = run(structure, ci=2.0, exc=-0.5), run(structure, precision="f64", ci=2.0, exc=-0.5)
    qx, qy = np.asarray(x.output.qsim, np.float64), np.asarray(y.output.qsim, np.float64)
    assert np.isfinite(qx).all() and qy.max() > 1.0
    assert np.abs(qx - qy).max() <= 2e-3 * qy.max()


@pytest.mark.parametrize("structure", OTHERS)
def test_forward_bookkeeping(structure):
    # forward.f90:41-72: fstates = the stores after the run, states restored; stores the structure does not own are untouched
    m = run(structure, T=240, ci=2.0)
    own = set(STRUCTURE_STATES[structure]) | {"hlr"}
    act = m.mesh.active_cell == 1
    for n in GSTATES_NAME:
        before, after = getattr(m.states, n), getattr(m.output.fstates, n)
        assert np.array_equal(before, getattr(StatesDT(m.mesh), n)), n
        if n in own:
            assert np.any(after[act] != before[act]), n
        else:
            assert np.array_equal(after, before), n
    assert np.all(np.asarray(m.output.qsim) >= -1e-6)   # (ht_imd - ht) * ct of gr_transfer cancels to a few ulp below zero in float32


def test_five_different_models():
    q = {s: np.asarray(run(s, ci=2.0, exc=-0.5).output.qsim) for s in ("gr-a",) + OTHERS}
    keys = list(q)
    for i in range(len(keys)):
        for j in range(i + 1, len(keys)):
            assert np.abs(q[keys[i]] - q[keys[j]]).max() > 0.5, (keys[i], keys[j])


def test_adjoint_is_gr_a_only():
    m = cases.cance(T=48)
    m.setup.structure = "gr-c"
    pb, sb = ParametersDT(m.mesh), StatesDT(m.mesh)
    with pytest.raises(AssertionError):
        oracle.forward_b(m.setup, m.mesh, m.input_data, m.parameters, pb, m.parameters.copy(), m.states, sb, m.states.copy(), m.output)
