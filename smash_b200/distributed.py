"""Multi-GPU plumbing for the places the path shards, one process per GPU:

* ensembles (``compute_multiple_run``, mw_multiple_run.f90:96-117: members are independent) -- contiguous blocks of
  members per rank, no data-path collective; the per-member costs / hydrographs are gathered afterwards;
* multi-catchment regionalised calibration -- every rank evaluates ``hyper_forward_b`` on its own catchment(s) and the
  shared hyper-parameter gradient (a few hundred floats) is summed with ONE all-reduce;
* one large domain split by drainage basin -- no exchange on the data path.

The collective is NCCL over NVLink, called through the library's own communicator (``smash_b200_comm_*``: libnccl.so.2
opened at run time; no PyTorch anywhere in the package).  Every function takes a ``comm`` object with ``rank``,
``world``, ``allreduce(array, op)`` (in place) and ``allgather(array)``; ``None`` means the process-wide ``NcclComm``
built from the launcher's environment (RANK / WORLD_SIZE / LOCAL_RANK, as set by torchrun).  The CPU tests pass a
gloo-backed object with the same four members (tests/gloo_comm.py)."""
from __future__ import annotations

import atexit
import ctypes as C
import os
import time

import numpy as np

PARAM_NAMES = ("ci", "cp", "beta", "cft", "cst", "alpha", "exc", "b", "cusl1", "cusl2", "clsl", "ks", "ds", "dsm", "ws", "lr")
STATE_NAMES = ("hi", "hp", "hft", "hst", "husl1", "husl2", "hlsl", "hlr")

_KIND = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2}
_OP = {"sum": 0, "max": 2, "min": 3}


class NcclComm:
    """One communicator per process: rank 0 creates the NCCL unique id and leaves it in a file named after the launcher's
    process id and port (all ranks of one node share both), the other ranks pick it up."""

    def __init__(self, rank=None, world=None, device=None, id_file=None, timeout_s=120.0):
        from . import _lib as L
        self._L = L
        lib = L.lib()
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else int(rank)
        self.world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else int(world)
        device = int(os.environ.get("LOCAL_RANK", str(self.rank))) if device is None else int(device)
        L.check(lib.smash_b200_set_device(device))
        if id_file is None:
            id_file = os.environ.get("SMASH_B200_COMM_ID_FILE") or os.path.join(
                "/tmp", "smash_b200_comm_%d_%s.id" % (os.getppid(), os.environ.get("MASTER_PORT", "0")))
        uid = C.create_string_buffer(128)
        if self.rank == 0:
            self._check(lib.smash_b200_comm_unique_id(uid))
            tmp = id_file + ".tmp%d" % os.getpid()
            with open(tmp, "wb") as f:
                f.write(uid.raw)
            os.replace(tmp, id_file)
            atexit.register(lambda: os.path.exists(id_file) and os.remove(id_file))
        else:
            t0 = time.time()
            while not (os.path.exists(id_file) and os.path.getsize(id_file) == 128):
                if time.time() - t0 > timeout_s:
                    raise RuntimeError(f"NCCL unique id file {id_file} did not appear")
                time.sleep(0.01)
            with open(id_file, "rb") as f:
                uid.raw = f.read(128)
        self._h = C.c_void_p()
        self._check(lib.smash_b200_comm_create(uid, self.rank, self.world, C.byref(self._h)))

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("smash_b200 communicator: " + self._L.lib().smash_b200_comm_last_error().decode())

    def allreduce(self, a, op="sum"):
        """In place on a C-contiguous float32 / float64 / int32 array."""
        assert a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"]
        self._check(self._L.lib().smash_b200_comm_allreduce(self._h, a.ctypes.data_as(C.c_void_p), a.size, _KIND[a.dtype], _OP[op]))
        return a

    def allgather(self, a):
        """(world, a.size) array, row r = rank r's contribution."""
        a = np.ascontiguousarray(a)
        out = np.empty((self.world, a.size), dtype=a.dtype)
        self._check(self._L.lib().smash_b200_comm_allgather(self._h, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                                            a.size, _KIND[a.dtype]))
        return out

    def barrier(self):
        self.allreduce(np.zeros(1, np.int32))

    def close(self):
        if getattr(self, "_h", None):
            self._L.lib().smash_b200_comm_destroy(self._h)
            self._h = None


_default = None


def default_comm():
    global _default
    if _default is None:
        _default = NcclComm()
    return _default


def member_slice(ns: int, rank: int, world: int) -> slice:
    """Contiguous block of members owned by ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(ns, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def multiple_run_sharded(setup, mesh, input_data, parameters, states, output, sample, ind_parameters_states, res_cost,
                         res_qsim, comm=None, compute=None):
    """``compute_multiple_run`` with the members split over the ranks of ``comm``; every rank ends with the full
    ``res_cost`` (and ``res_qsim`` when it is not size-0)."""
    comm = comm or default_comm()
    if compute is None:
        from .solver._mw_multiple_run import compute_multiple_run as compute
    rank, world = comm.rank, comm.world
    sample = np.asfortranarray(sample, dtype=np.float32)
    ns = sample.shape[1]
    sl = member_slice(ns, rank, world)
    n_loc = sl.stop - sl.start
    want_q = res_qsim is not None and res_qsim.size > 0
    nq = mesh.ng * setup._ntime_step
    cost_loc = np.zeros(n_loc, np.float32)
    q_loc = np.zeros((mesh.ng, setup._ntime_step, n_loc), np.float32, order="F") if want_q else np.zeros((0,), np.float32)
    if n_loc:
        compute(setup, mesh, input_data, parameters, states, output, np.asfortranarray(sample[:, sl]), ind_parameters_states,
                cost_loc, q_loc)
    nmax = -(-ns // world)
    width = 1 + (nq if want_q else 0)
    buf = np.zeros((nmax, width), dtype=np.float32)
    buf[:n_loc, 0] = cost_loc
    if want_q and n_loc:
        buf[:n_loc, 1:] = q_loc.reshape(nq, n_loc, order="F").T
    out = comm.allgather(buf).reshape(world, nmax, width)
    for r in range(world):
        s = member_slice(ns, r, world)
        part = out[r][: s.stop - s.start]
        res_cost[s] = part[:, 0]
        if want_q:
            res_qsim[:, :, s] = part[:, 1:].T.reshape(mesh.ng, setup._ntime_step, s.stop - s.start, order="F")
    return res_cost


def allreduce_shared_gradient(cost, hyper_parameters_b, hyper_states_b, comm=None):
    """Sum over ranks of (cost, d cost / d hyper-parameters, d cost / d hyper-states): the one collective of the
    multi-catchment regionalised calibration.  Updates the *_b objects in place and returns the summed cost."""
    comm = comm or default_comm()
    parts = [np.asarray([cost], np.float32)]
    refs = []
    for obj, names in ((hyper_parameters_b, PARAM_NAMES), (hyper_states_b, STATE_NAMES)):
        for n in names:
            a = getattr(obj, n, None)
            if a is not None:
                refs.append((obj, n, a.shape))
                parts.append(np.asarray(a, np.float32).ravel(order="F"))
    flat = np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)
    comm.allreduce(flat, "sum")
    k = 1
    for obj, n, shp in refs:
        size = int(np.prod(shp))
        getattr(obj, n)[...] = flat[k:k + size].reshape(shp, order="F")
        k += size
    return np.float32(flat[0])


def optimize_hyper_lbfgsb_sharded(catchments, comm=None, solver=None):
    """Regionalised multi-catchment calibration with the catchments spread over the ranks of ``comm``: every rank passes
    its own ``(setup, mesh, input_data, parameters, states, output)`` tuples, the summed cost and shared hyper-parameter
    gradient travel in ONE all-reduce of ``1 + n_control * nhyper`` values per evaluation (plus one min / max all-reduce
    of the descriptor extrema at the start), and every rank walks the same L-BFGS-B path
    (``smash_b200.solver._mw_optimize.optimize_hyper_lbfgsb_multi``)."""
    from .solver._mw_optimize import optimize_hyper_lbfgsb_multi
    comm = comm or default_comm()

    def reduce_sum(vec):
        return comm.allreduce(np.ascontiguousarray(vec, dtype=np.float64), "sum")

    def reduce_minmax(mins, maxs):
        lo, hi = np.ascontiguousarray(mins, dtype=np.float32), np.ascontiguousarray(maxs, dtype=np.float32)
        return comm.allreduce(lo, "min"), comm.allreduce(hi, "max")

    return optimize_hyper_lbfgsb_multi(catchments, solver=solver, reduce_sum=reduce_sum, reduce_minmax=reduce_minmax)


# ---- a single large domain split by drainage basin ("France single run: shards by basin") -----------------
def basin_labels(mesh, setup=None):
    """Label of the drainage basin (connected component of the D8 gather graph, md_routing_operator.f90:37-53) of every
    computed cell, as an (nrow, ncol) int32 array (-1 elsewhere), and the number of basins.  Basins exchange nothing: no
    cell of one gathers a cell of another."""
    import ctypes as C

    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    from . import _lib as L
    pk = L.Packed()
    me = L.pack_mesh(mesh, setup, pk)
    act = (np.asarray(mesh.active_cell) == 1) & (np.asarray(getattr(mesh, "_local_active_cell", mesh.active_cell)) == 1)
    n = int(act.sum())
    info = (C.c_int64 * 8)()
    cell, task, pos, down = (np.zeros(n, np.int32) for _ in range(4))
    L.check(L.lib().smash_b200_mesh_chains(C.byref(me), info, L._ip(cell), L._ip(task), L._ip(pos), L._ip(down)))
    has = down >= 0
    src = np.nonzero(has)[0]
    g = coo_matrix((np.ones(src.size, np.int8), (src, down[has])), shape=(n, n))
    nb, lab = connected_components(g, directed=False)
    out = np.full(mesh.nrow * mesh.ncol, -1, dtype=np.int32)
    out[cell] = lab.astype(np.int32)
    return np.asfortranarray(out.reshape((mesh.nrow, mesh.ncol), order="F")), int(nb)


def basin_masks(mesh, world, setup=None):
    """``world`` masks for ``mesh._local_active_cell`` (mwd_mesh.f90:68, md_forward_structure.f90:88): whole basins are
    assigned to ranks, largest first, each to the least loaded rank.  The largest basin bounds the imbalance."""
    labels, nb = basin_labels(mesh, setup)
    size = np.bincount(labels[labels >= 0], minlength=nb)
    load = np.zeros(world, dtype=np.int64)
    owner = np.zeros(nb, dtype=np.int32)
    for b in np.argsort(-size, kind="stable"):
        r = int(np.argmin(load))
        owner[b] = r
        load[r] += size[b]
    own = np.where(labels >= 0, owner[np.maximum(labels, 0)], -1)
    return [np.asfortranarray((own == r).astype(np.int32)) for r in range(world)], load


def forward_sharded_by_basin(model, comm=None, gather=True, solver=None):
    """One forward run of a large domain with its basins spread over the ranks of ``comm``: every rank computes the cells
    of its basins only (``local_active_cell``), there is no exchange on the data path.  With ``gather`` the domain series
    (``sparse_qsim_domain`` / ``qsim_domain``) and ``qsim`` of all ranks are combined by one all-reduce afterwards."""
    comm = comm or default_comm()
    if solver is None:
        from .solver import _mw_forward as solver
    rank, world = comm.rank, comm.world
    masks, _ = basin_masks(model.mesh, world, model.setup)
    keep = model.mesh._local_active_cell
    model.mesh._local_active_cell = masks[rank]
    if hasattr(model.mesh, "_b200_cache"):
        del model.mesh._b200_cache
    try:
        solver.forward(model.setup, model.mesh, model.input_data, model.parameters, model.parameters.copy(), model.states,
                       model.states.copy(), model.output)
    finally:
        model.mesh._local_active_cell = keep
        if hasattr(model.mesh, "_b200_cache"):
            del model.mesh._b200_cache
    if gather:
        mine = masks[rank] == 1
        for name in ("sparse_qsim_domain", "qsim_domain"):
            a = getattr(model.output, name, None)
            if a is None:
                continue
            if name == "sparse_qsim_domain":
                k = model.mesh._rowcol_to_ind_sparse
                own = np.zeros(model.mesh.nac, dtype=bool)
                own[k[mine] - 1] = True
                part = np.where(own[:, None], a, np.float32(0.0))
            else:
                part = np.where(mine[:, :, None], a, np.float32(0.0))
            t = np.ascontiguousarray(part, dtype=np.float32)
            comm.allreduce(t, "sum")
            a[...] = t
        if model.mesh.ng > 0:
            gp = np.asarray(model.mesh.gauge_pos)
            own_g = mine[gp[:, 0], gp[:, 1]]
            t = np.ascontiguousarray(np.where(own_g[:, None], model.output.qsim, np.float32(0.0)), dtype=np.float32)
            comm.allreduce(t, "sum")
            model.output.qsim[...] = t
    return masks[rank]
