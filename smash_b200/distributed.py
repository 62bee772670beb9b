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
