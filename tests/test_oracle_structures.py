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


# ---- adjust_interception_store (mw_interception_store.f90:19-160) ---------------------------------------------------------

def _day_index(T, steps_per_day=24):
    return (np.arange(T) // steps_per_day + 1).astype(np.int32)


def _one_cell_by_hand(prcp, pet, day_index):
    """The routine for one cell in NumPy float32 scalars, written from the Fortran text: daily sums, 49 capacities, minloc."""
    f = np.float32
    days = {}
    for p, e, d in zip(prcp, pet, day_index):
        a, b = days.get(d, (f(0), f(0)))
        days[d] = (f(a + p), f(b + e))
    daily = f(0)
    for d in sorted(days):
        daily = f(daily + min(days[d][0], days[d][1]))
    best, best_diff = None, None
    for i in range(49):
        ci = f(f(0.1) + f(f(i) * f(0.1)))
        h, sub = f(0), f(0)
        for p, e in zip(prcp, pet):
            ei = min(e, f(p + f(h * ci)))
            pn = max(f(0), f(f(p - f(ci * f(f(1) - h))) - ei))
            h = f(h + f(f(f(p - ei) - pn) / ci))
            sub = f(sub + ei)
        diff = abs(f(sub - daily))
        if best is None or diff < best_diff:
            best, best_diff = ci, diff
    return best


def test_interception_store_against_a_hand_restatement():
    m = cases.cance(T=240)
    m.setup.structure = "gr-b"
    di = _day_index(240)
    oracle.adjust_interception_store(m.setup, m.mesh, m.input_data, m.parameters, 10, di)
    ci = np.asarray(m.parameters.ci)
    act = m.mesh.active_cell == 1
    assert np.all(ci[~act] == np.float32(1e-6))                              # untouched outside the computed cells
    assert ci[act].min() >= np.float32(0.1) and ci[act].max() <= np.float32(4.9) + 1e-6
    rr, cc = np.nonzero(act)
    for k in (0, 57, 191, len(rr) - 1):
        r, c = rr[k], cc[k]
        want = _one_cell_by_hand(m.input_data.prcp[r, c, :], m.input_data.pet[r, c, :], di)
        assert ci[r, c] == want, (k, ci[r, c], want)


def test_interception_store_without_rain_takes_the_first_capacity():
    m = cases.cance(T=96, sparse=True)
    m.input_data.sparse_prcp[...] = 0.0
    oracle.adjust_interception_store(m.setup, m.mesh, m.input_data, m.parameters, 4, _day_index(96))
    act = m.mesh.active_cell == 1
    assert np.all(np.asarray(m.parameters.ci)[act] == np.float32(0.1))       # every diff is 0: minloc returns the first
