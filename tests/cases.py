"""Shared builders for the parity tests: the Cance model rebuilt from tests/golden/*.npz (what
``smash.Model(*load_dataset("Cance"))`` holds in the reference, smash/tests/test_define_global_vars.py:9-14),
synthetic forcing and the France mesh."""
from __future__ import annotations

import os

import numpy as np

from smash_b200.solver._derived_types import (Hyper_ParametersDT, Hyper_StatesDT, Input_DataDT, MeshDT, Optimize_SetupDT,
                                              OutputDT, ParametersDT, SetupDT, StatesDT, compute_rowcol_to_ind_sparse)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# smash/core/_constant.py:49-77
RATIO_PET_HOURLY = np.array([0, 0, 0, 0, 0, 0, 0, 0.035, 0.062, 0.079, 0.097, 0.11, 0.117, 0.117, 0.11, 0.097, 0.079, 0.062,
                             0.035, 0, 0, 0, 0, 0], dtype=np.float32)


class Model:
    """The six derived types a reference ``smash.Model`` carries."""

    def __init__(self, setup, mesh, input_data, parameters, states, output):
        self.setup, self.mesh, self.input_data = setup, mesh, input_data
        self.parameters, self.states, self.output = parameters, states, output

    def copy(self):
        return Model(self.setup.copy(), self.mesh, self.input_data, self.parameters.copy(), self.states.copy(),
                     self.output.copy())


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def set_optimize(setup, mesh, jobs_fun=("nse",), wjobs_fun=None, gauge="downstream", wgauge=None, mapping="...",
                 jreg_fun=(), wjreg_fun=None, wjreg=0.0, ost=1, denormalize_forward=False):
    """What _standardize_*_args + reset_optimize_setup (mw_derived_type_update.f90:13-59) leave in setup._optimize."""
    njf, njr = len(jobs_fun), len(jreg_fun)
    o = Optimize_SetupDT(setup._ntime_step, setup._nd, mesh.ng, mapping, njf, njr)
    o.jobs_fun = np.array(list(jobs_fun), dtype="U20")
    o.wjobs_fun = np.full(njf, 1.0 / njf if njf else 0.0, dtype=np.float32) if wjobs_fun is None else np.asarray(wjobs_fun, np.float32)
    o.jreg_fun = np.array(list(jreg_fun), dtype="U20")
    o.wjreg_fun = np.ones(njr, np.float32) if wjreg_fun is None else np.asarray(wjreg_fun, np.float32)
    o.wjreg = np.float32(wjreg)
    o.optimize_start_step = int(ost)
    o.denormalize_forward = bool(denormalize_forward)
    if mesh.ng > 0:
        if wgauge is not None:
            o.wgauge = np.asarray(wgauge, dtype=np.float32)
        elif gauge == "downstream":  # _standardize.py:308-311,362-363: argmax(area), weight 1
            w = np.zeros(mesh.ng, np.float32)
            w[int(np.argmax(mesh.area))] = 1.0
            o.wgauge = w
        elif gauge == "all":
            o.wgauge = np.full(mesh.ng, 1.0 / mesh.ng, np.float32)
    setup._optimize = o
    return o


def cance(sparse=False, T=None, jobs_fun=("nse",)):
    d = golden("cance_inputs.npz")
    nrow, ncol, ng = int(d["nrow"]), int(d["ncol"]), int(d["ng"])
    Tfull = d["prcp"].shape[0]
    T = Tfull if T is None else int(T)
    setup = SetupDT(nd=2, ng=ng)
    setup.structure = "gr-a"
    setup.dt = np.float32(d["dt"])
    setup.sparse_storage = bool(sparse)
    setup._ntime_step = T
    setup.descriptor_name = np.array(["slope", "dd"], dtype="U20")
    mesh = MeshDT(setup, nrow, ncol, ng)
    mesh.dx = np.float32(d["dx"])
    mesh.nac = int(d["nac"])
    mesh.flwdir = np.asfortranarray(d["flwdir"], dtype=np.int32)
    mesh.flwacc = np.asfortranarray(d["flwacc"], dtype=np.int32)
    mesh.active_cell = np.asfortranarray(d["active_cell"], dtype=np.int32)
    mesh.path = np.asfortranarray(d["path"], dtype=np.int32)          # 0-based as stored by save_mesh
    mesh.gauge_pos = np.asfortranarray(d["gauge_pos"], dtype=np.int32)
    mesh.area = d["area"].astype(np.float32)
    mesh.flwdst = np.asfortranarray(d["flwdst"], dtype=np.float32)
    mesh.code = d["code"].astype("U20")
    mesh._local_active_cell = np.asfortranarray(mesh.active_cell.copy())
    compute_rowcol_to_ind_sparse(mesh)
    inp = Input_DataDT(setup, mesh)
    prcp = d["prcp"][:T]                                                # (T,nrow,ncol)
    pet = (d["pet_daily"][d["pet_day"][:T]] * d["pet_ratio"][:T, None, None].astype(np.float64)).astype(np.float32)
    if sparse:
        k = mesh._rowcol_to_ind_sparse
        rr, cc = np.nonzero(k > 0)
        order = np.argsort(k[rr, cc])
        rr, cc = rr[order], cc[order]
        inp.sparse_prcp = np.asfortranarray(prcp[:, rr, cc].T)
        inp.sparse_pet = np.asfortranarray(pet[:, rr, cc].T)
    else:
        inp.prcp = np.asfortranarray(np.moveaxis(prcp, 0, 2))
        inp.pet = np.asfortranarray(np.moveaxis(pet, 0, 2))
    inp.qobs = np.asfortranarray(d["qobs"][:, :T], dtype=np.float32)
    inp.descriptor = np.asfortranarray(d["descriptor"], dtype=np.float32)
    par = ParametersDT(mesh)
    par.lr[...] = np.float32(setup.dt) * np.float32(5.0 / 3600.0)       # _build_model.py:257
    st = StatesDT(mesh)
    out = OutputDT(setup, mesh)
    set_optimize(setup, mesh, jobs_fun=jobs_fun)
    return Model(setup, mesh, inp, par, st, out)


# ind_parameters_states of cp, cft, exc, lr in the 16+8 stacked planes (1-based, multiple_run.py:184-200)
IND_CP_CFT_EXC_LR = np.array([2, 4, 7, 16], dtype=np.int32)


def output_cost(model, nse, kge):
    """smash/tests/core/test_simu.py:319-334"""
    qo, qs = model.input_data.qobs, model.output.qsim
    ret = np.zeros(3 * model.mesh.ng, dtype=np.float32)
    for i in range(model.mesh.ng):
        ret[3 * i:3 * i + 3] = (model.output.cost, nse(qo[i], qs[i]), kge(qo[i], qs[i]))
    return ret


def normalize_descriptor(model):
    """optimize_hyper_lbfgsb normalises descriptors to [0,1] over the domain before mapping (mw_optimize.f90:960-980)."""
    d = model.input_data.descriptor
    out = np.empty_like(d)
    for j in range(d.shape[2]):
        lo, hi = d[..., j].min(), d[..., j].max()
        out[..., j] = (d[..., j] - lo) / (hi - lo)
    model.input_data.descriptor = np.asfortranarray(out)


def synthetic_forcing(nac, T, seed=0, gap_fraction=0.0):
    """prcp = Bernoulli(0.15)*Gamma(0.6, 4.0) mm/h, pet = U(1,4) mm/d * RATIO_PET_HOURLY[hour].
    Returns Fortran-ordered (nac, T) float32 arrays (memory [t][k], mwd_input_data.f90:73-80)."""
    rng = np.random.default_rng(seed)
    prcp = np.empty((T, nac), dtype=np.float32)
    chunks = list(range(0, T, 16))
    seeds = np.random.SeedSequence(seed).spawn(len(chunks))

    def fill(job):                                  # independent stream per 16-step chunk (NumPy RNGs release the GIL)
        t0, ss = job
        r = np.random.default_rng(ss)
        n = min(16, T - t0)
        wet = r.random((n, nac), dtype=np.float32) < np.float32(0.15)
        prcp[t0:t0 + n] = wet * (r.standard_gamma(0.6, (n, nac), dtype=np.float32) * np.float32(4.0))
        if gap_fraction > 0:
            gaps = r.random((n, nac), dtype=np.float32) < np.float32(gap_fraction)
            prcp[t0:t0 + n][gaps] = np.float32(-99.0)

    if T * nac > 50_000_000:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
            list(ex.map(fill, zip(chunks, seeds)))
    else:
        for job in zip(chunks, seeds):
            fill(job)
    pet_day = rng.uniform(1.0, 4.0, nac).astype(np.float32)
    pet = RATIO_PET_HOURLY[np.arange(T) % 24][:, None] * pet_day[None, :]
    return prcp.T, np.ascontiguousarray(pet, dtype=np.float32).T


def france(T=24, seed=0, sub=None, ngauge=0, nd=0, qobs_from_oracle=True):
    """France 1 km mesh (mesh_France.hdf5) with synthetic sparse forcing; `sub=(r0,r1,c0,c1)` crops a window
    (flow directions leaving the window simply drain nowhere, as at the domain edge); `ngauge` puts synthetic gauges
    on the cells with the largest flow accumulation (the shipped France mesh has none)."""
    d = golden("france_mesh.npz")
    flwdir, flwacc, active = d["flwdir"].astype(np.int32), d["flwacc"], d["active_cell"].astype(np.int32)
    path = d["path"].astype(np.int32)
    if sub is not None:
        r0, r1, c0, c1 = sub
        flwdir, flwacc, active = flwdir[r0:r1, c0:c1], flwacc[r0:r1, c0:c1], active[r0:r1, c0:c1]
        keep = (path[0] >= r0) & (path[0] < r1) & (path[1] >= c0) & (path[1] < c1)
        path = path[:, keep] - np.array([[r0], [c0]], dtype=np.int32)
    nrow, ncol = flwdir.shape
    setup = SetupDT(nd=nd, ng=ngauge)
    setup.sparse_storage = True
    setup._ntime_step = int(T)
    setup.save_qsim_domain = True                                       # setup_France.yaml:18
    mesh = MeshDT(setup, nrow, ncol, ngauge)
    mesh.dx = np.float32(d["dx"])
    mesh.flwdir = np.asfortranarray(flwdir)
    mesh.flwacc = np.asfortranarray(flwacc)
    mesh.active_cell = np.asfortranarray(active)
    mesh._local_active_cell = np.asfortranarray(active.copy())
    full = np.full((2, nrow * ncol), -100, dtype=np.int32)
    full[:, :path.shape[1]] = path
    mesh.path = np.asfortranarray(full)
    mesh.nac = int(active.sum())
    compute_rowcol_to_ind_sparse(mesh)
    if ngauge > 0:
        fa = np.where(active == 1, flwacc, 0)
        flat = np.argsort(fa.ravel(), kind="stable")[::-1][:ngauge]
        gr, gc = np.unravel_index(flat, fa.shape)
        mesh.gauge_pos = np.asfortranarray(np.stack([gr, gc], axis=1).astype(np.int32))
        mesh.area = (flwacc[gr, gc].astype(np.float32) * mesh.dx * mesh.dx).astype(np.float32)
    inp = Input_DataDT(setup, mesh)
    inp.sparse_prcp, inp.sparse_pet = synthetic_forcing(mesh.nac, T, seed)
    if nd > 0:                                                          # synthetic descriptors, already normalised to [0, 1]
        inp.descriptor = np.asfortranarray(np.random.default_rng(seed + 5).uniform(0.0, 1.0, (nrow, ncol, nd)).astype(np.float32))
    par = ParametersDT(mesh)
    st = StatesDT(mesh)
    out = OutputDT(setup, mesh)
    set_optimize(setup, mesh, jobs_fun=("nse",) if ngauge else (), gauge="all")
    model = Model(setup, mesh, inp, par, st, out)
    if ngauge > 0 and not qobs_from_oracle:
        # long runs (the oracle would need minutes): positive synthetic observations of a plausible magnitude
        area = flwacc[gr, gc].astype(np.float64) * float(mesh.dx) ** 2
        base = (0.05e-3 / 3600.0) * area                                     # 0.05 mm/h over the drained area, m3/s
        inp.qobs = np.asfortranarray((base[:, None] * np.random.default_rng(seed + 17).gamma(2.0, 0.5, (ngauge, T))).astype(np.float32))
    elif ngauge > 0:
        # "observations" = a run of the CPU oracle with perturbed parameters, +-5 % multiplicative noise
        import oracle
        truth = model.copy()
        truth.setup.save_qsim_domain = False
        truth.parameters.cp[...] = 260.0
        truth.parameters.cft[...] = 380.0
        truth.parameters.lr[...] = 7.0
        oracle.forward(truth.setup, truth.mesh, truth.input_data, truth.parameters, truth.parameters.copy(), truth.states,
                       truth.states.copy(), truth.output)
        noise = 1.0 + 0.05 * np.random.default_rng(seed + 17).uniform(-1, 1, truth.output.qsim.shape)
        inp.qobs = np.asfortranarray((truth.output.qsim * noise).astype(np.float32))
    return model


def from_flwdir(flwdir, T=24, seed=0, ngauge=0, dx=1000.0):
    """A model on an arbitrary D8 raster (codes 1..8, anything else = not a cell): flow accumulation by the host restatement
    of the meshing step, path = stable argsort of flwacc (meshing.py:216-218), sparse synthetic forcing, gauges on the cells
    with the largest flow accumulation."""
    from smash_b200.mesh import flow_accumulation
    flwdir = np.asarray(flwdir, dtype=np.int32)
    nrow, ncol = flwdir.shape
    active = ((flwdir >= 1) & (flwdir <= 8)).astype(np.int32)
    flwacc = flow_accumulation(flwdir, mask=active == 1)
    order = np.argsort(np.where(active == 1, flwacc, np.iinfo(np.int32).max).ravel(), kind="stable")[: int(active.sum())]
    pr, pc = np.unravel_index(order, flwdir.shape)
    setup = SetupDT(nd=0, ng=ngauge)
    setup.sparse_storage = True
    setup._ntime_step = int(T)
    setup.save_qsim_domain = True
    mesh = MeshDT(setup, nrow, ncol, ngauge)
    mesh.dx = np.float32(dx)
    mesh.flwdir = np.asfortranarray(flwdir)
    mesh.flwacc = np.asfortranarray(flwacc)
    mesh.active_cell = np.asfortranarray(active)
    mesh._local_active_cell = np.asfortranarray(active.copy())
    full = np.full((2, nrow * ncol), -100, dtype=np.int32)
    full[0, :pr.size], full[1, :pr.size] = pr, pc
    mesh.path = np.asfortranarray(full)
    mesh.nac = int(active.sum())
    compute_rowcol_to_ind_sparse(mesh)
    if ngauge > 0:
        fa = np.where(active == 1, flwacc, 0)
        flat = np.argsort(fa.ravel(), kind="stable")[::-1][:ngauge]
        gr, gc = np.unravel_index(flat, fa.shape)
        mesh.gauge_pos = np.asfortranarray(np.stack([gr, gc], axis=1).astype(np.int32))
        mesh.area = (flwacc[gr, gc].astype(np.float32) * mesh.dx * mesh.dx).astype(np.float32)
    inp = Input_DataDT(setup, mesh)
    inp.sparse_prcp, inp.sparse_pet = synthetic_forcing(mesh.nac, T, seed)
    par, st, out = ParametersDT(mesh), StatesDT(mesh), OutputDT(setup, mesh)
    set_optimize(setup, mesh, jobs_fun=(), gauge="all")
    return Model(setup, mesh, inp, par, st, out)


def hyper_objects(model):
    return Hyper_ParametersDT(model.setup), Hyper_StatesDT(model.setup)
