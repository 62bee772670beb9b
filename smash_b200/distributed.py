"""Multi-GPU plumbing for the two places the path shards, one process per GPU:

* ensembles (``compute_multiple_run``, mw_multiple_run.f90:96-117: members are independent) -- contiguous blocks of
  members per rank, no data-path collective; the per-member costs / hydrographs are gathered afterwards;
* multi-catchment regionalised calibration -- every rank evaluates ``hyper_forward_b`` on its own catchment(s) and the
  shared hyper-parameter gradient (a few hundred floats) is summed with ONE all-reduce.

``torch.distributed`` is used for the process group only (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np

PARAM_NAMES = ("ci", "cp", "beta", "cft", "cst", "alpha", "exc", "b", "cusl1", "cusl2", "clsl", "ks", "ds", "dsm", "ws", "lr")
STATE_NAMES = ("hi", "hp", "hft", "hst", "husl1", "husl2", "hlsl", "hlr")


def member_slice(ns: int, rank: int, world: int) -> slice:
    """Contiguous block of members owned by ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(ns, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def _dist():
    import torch.distributed as dist
    return dist


def _device(group=None):
    import torch
    dist = _dist()
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def multiple_run_sharded(setup, mesh, input_data, parameters, states, output, sample, ind_parameters_states, res_cost,
                         res_qsim, group=None, compute=None):
    """``compute_multiple_run`` with the members split over the ranks of ``group``; every rank ends with the full
    ``res_cost`` (and ``res_qsim`` when it is not size-0)."""
    import torch
    dist = _dist()
    if compute is None:
        from .solver._mw_multiple_run import compute_multiple_run as compute
    rank, world = dist.get_rank(group), dist.get_world_size(group)
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
    dev = _device(group)
    nmax = -(-ns // world)
    width = 1 + (nq if want_q else 0)
    buf = torch.zeros((nmax, width), dtype=torch.float32)
    buf[:n_loc, 0] = torch.from_numpy(cost_loc)
    if want_q and n_loc:
        buf[:n_loc, 1:] = torch.from_numpy(np.ascontiguousarray(q_loc.reshape(nq, n_loc, order="F").T))
    buf = buf.to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    for r in range(world):
        s = member_slice(ns, r, world)
        part = out[r][: s.stop - s.start].cpu().numpy()
        res_cost[s] = part[:, 0]
        if want_q:
            res_qsim[:, :, s] = part[:, 1:].T.reshape(mesh.ng, setup._ntime_step, s.stop - s.start, order="F")
    return res_cost


def allreduce_shared_gradient(cost, hyper_parameters_b, hyper_states_b, group=None):
    """Sum over ranks of (cost, d cost / d hyper-parameters, d cost / d hyper-states): the one collective of the
    multi-catchment regionalised calibration.  Updates the *_b objects in place and returns the summed cost."""
    import torch
    dist = _dist()
    parts = [np.asarray([cost], np.float32)]
    refs = []
    for obj, names in ((hyper_parameters_b, PARAM_NAMES), (hyper_states_b, STATE_NAMES)):
        for n in names:
            a = getattr(obj, n, None)
            if a is not None:
                refs.append((obj, n, a.shape))
                parts.append(np.asarray(a, np.float32).ravel(order="F"))
    flat = torch.from_numpy(np.concatenate(parts)).to(_device(group))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat = flat.cpu().numpy()
    k = 1
    for obj, n, shp in refs:
        size = int(np.prod(shp))
        getattr(obj, n)[...] = flat[k:k + size].reshape(shp, order="F")
        k += size
    return np.float32(flat[0])


def optimize_hyper_lbfgsb_sharded(catchments, group=None, solver=None):
    """Regionalised multi-catchment calibration with the catchments spread over the ranks of ``group``: every rank passes
    its own ``(setup, mesh, input_data, parameters, states, output)`` tuples, the summed cost and shared hyper-parameter
    gradient travel in ONE all-reduce of ``1 + n_control * nhyper`` values per evaluation (plus one min / max all-reduce
    of the descriptor extrema at the start), and every rank walks the same L-BFGS-B path
    (``smash_b200.solver._mw_optimize.optimize_hyper_lbfgsb_multi``)."""
    import torch
    from .solver._mw_optimize import optimize_hyper_lbfgsb_multi
    dist = _dist()
    dev = _device(group)

    def reduce_sum(vec):
        t = torch.from_numpy(np.asarray(vec, dtype=np.float64)).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.cpu().numpy()

    def reduce_minmax(mins, maxs):
        lo, hi = torch.from_numpy(np.asarray(mins, np.float32)).to(dev), torch.from_numpy(np.asarray(maxs, np.float32)).to(dev)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
        return lo.cpu().numpy(), hi.cpu().numpy()

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


def forward_sharded_by_basin(model, group=None, gather=True, solver=None):
    """One forward run of a large domain with its basins spread over the ranks of ``group``: every rank computes the cells
    of its basins only (``local_active_cell``), there is no exchange on the data path.  With ``gather`` the domain series
    (``sparse_qsim_domain`` / ``qsim_domain``) and ``qsim`` of all ranks are combined by one all-reduce afterwards."""
    import torch
    dist = _dist()
    if solver is None:
        from .solver import _mw_forward as solver
    rank, world = dist.get_rank(group), dist.get_world_size(group)
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
        dev = _device(group)
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
            t = torch.from_numpy(np.ascontiguousarray(part, dtype=np.float32)).to(dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            a[...] = t.cpu().numpy()
        if model.mesh.ng > 0:
            gp = np.asarray(model.mesh.gauge_pos)
            own_g = mine[gp[:, 0], gp[:, 1]]
            t = torch.from_numpy(np.ascontiguousarray(np.where(own_g[:, None], model.output.qsim, np.float32(0.0)),
                                                      dtype=np.float32)).to(dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            model.output.qsim[...] = t.cpu().numpy()
    return masks[rank]
