"""world_size-2 gloo test (CPU) of the N>1 path (the product talks NCCL through its own communicator; here a gloo-backed
object with the same interface, tests/gloo_comm.py, stands in): member sharding + gather of the ensemble results, and the single
all-reduce of the shared hyper-parameter gradient.  The per-rank compute is replaced by a deterministic stand-in so
that only the host-side plumbing is exercised here; the GPU arithmetic is covered by the -m gpu tests."""
import os

import numpy as np
import torch.multiprocessing as mp


def _fake_compute(setup, mesh, input_data, parameters, states, output, sample, ind, res_cost, res_qsim):
    res_cost[...] = sample.sum(axis=0)
    if res_qsim.size:
        res_qsim[...] = sample[0][None, None, :] * np.arange(1, mesh.ng * setup._ntime_step + 1).reshape(
            mesh.ng, setup._ntime_step, order="F")[..., None]


def _worker(rank, world, port, ns, ret):
    import torch.distributed as dist

    import cases
    from smash_b200 import distributed as D
    from smash_b200.solver._derived_types import Hyper_ParametersDT, Hyper_StatesDT
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = cases.cance(T=12)
    rng = np.random.RandomState(0)
    sample = np.asfortranarray(rng.uniform(0, 1, (4, ns)).astype(np.float32))
    cost = np.zeros(ns, np.float32)
    qsim = np.zeros((m.mesh.ng, 12, ns), np.float32, order="F")
    from gloo_comm import GlooComm
    comm = GlooComm()
    D.multiple_run_sharded(m.setup, m.mesh, m.input_data, m.parameters, m.states, m.output, sample, cases.IND_CP_CFT_EXC_LR,
                           cost, qsim, comm=comm, compute=_fake_compute)
    cases.set_optimize(m.setup, m.mesh, mapping="hyper-linear")
    hpb, hsb = Hyper_ParametersDT(m.setup), Hyper_StatesDT(m.setup)
    hpb.cp[...] = rank + 1.0
    hsb.hlr[...] = 10.0 * (rank + 1)
    total = D.allreduce_shared_gradient(np.float32(0.5 + rank), hpb, hsb, comm=comm)
    if rank == 0:
        ret["cost"], ret["qsim"], ret["sample"] = cost, qsim, sample
        ret["total"], ret["cp"], ret["hlr"] = float(total), hpb.cp.copy(), hsb.hlr.copy()
    dist.destroy_process_group()


def test_member_slices_cover_everything():
    from smash_b200.distributed import member_slice
    for ns in (0, 1, 7, 10, 4096):
        for world in (1, 2, 3, 8):
            idx = np.concatenate([np.arange(ns)[member_slice(ns, r, world)] for r in range(world)])
            assert np.array_equal(idx, np.arange(ns))
            sizes = [member_slice(ns, r, world).stop - member_slice(ns, r, world).start for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_ensemble_and_allreduce_world2():
    world, ns = 2, 7
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, 29500 + os.getpid() % 1000, ns, ret), nprocs=world, join=True)
        sample = ret["sample"]
        assert np.allclose(ret["cost"], sample.sum(axis=0))
        want = sample[0][None, None, :] * np.arange(1, 3 * 12 + 1).reshape(3, 12, order="F")[..., None]
        assert np.allclose(ret["qsim"], want)
        assert ret["total"] == 0.5 + 1.5
        assert np.all(ret["cp"] == 3.0) and np.all(ret["hlr"] == 30.0)


def _catchment(k):
    """Two distinct 'catchments' for the regionalisation test: the Cance mesh over different periods / observations."""
    import cases
    m = cases.cance(T=240 if k == 0 else 300)
    if k == 1:
        m.input_data.qobs = np.asfortranarray(m.input_data.qobs * np.float32(1.2))
    cases.set_optimize(m.setup, m.mesh, jobs_fun=("nse",), mapping="hyper-polynomial")
    o = m.setup._optimize
    o.optim_parameters[[1, 3, 6, 15]] = 1                                  # cp, cft, exc, lr
    o.maxiter = 3
    o.verbose = False
    return m


def _as_tuple(m):
    return (m.setup, m.mesh, m.input_data, m.parameters, m.states, m.output)


def _regional_worker(rank, world, port, ret):
    import torch.distributed as dist

    import oracle_solver
    from smash_b200 import distributed as D
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _catchment(rank)
    from gloo_comm import GlooComm
    D.optimize_hyper_lbfgsb_sharded([_as_tuple(m)], comm=GlooComm(), solver=oracle_solver)
    ret[rank] = (float(m.output.cost), m.parameters.cp.copy(), m.parameters.lr.copy())
    dist.destroy_process_group()


def test_regionalised_calibration_world2_equals_single_process():
    # configs[4]: catchments spread over ranks, shared hyper-parameters, one all-reduce of (cost, gradient) per evaluation.
    # Two ranks with one catchment each must walk the same L-BFGS-B path as one process holding both.
    import oracle_solver
    from smash_b200.solver import _mw_optimize
    a, b = _catchment(0), _catchment(1)
    d0 = a.input_data.descriptor.copy()
    _mw_optimize.optimize_hyper_lbfgsb_multi([_as_tuple(a), _as_tuple(b)], solver=oracle_solver)
    assert np.allclose(a.input_data.descriptor, d0, rtol=1e-6)              # descriptors restored
    first = _catchment(0)
    oracle_solver.forward(first.setup, first.mesh, first.input_data, first.parameters, first.parameters.copy(), first.states,
                          first.states.copy(), first.output)
    assert float(a.output.cost) < float(first.output.cost)                  # the shared mapping improved catchment 0
    assert np.ptp(a.parameters.cp[a.mesh.active_cell == 1]) > 0             # and it is spatially distributed
    assert np.array_equal(a.parameters.cp, b.parameters.cp)                 # same mapping, same descriptors -> same field
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_regional_worker, args=(2, 29600 + os.getpid() % 1000, ret), nprocs=2, join=True)
        for rank, m in ((0, a), (1, b)):
            cost, cp, lr = ret[rank]
            assert np.isclose(cost, float(m.output.cost), rtol=1e-6)
            assert np.allclose(cp, m.parameters.cp, rtol=1e-6) and np.allclose(lr, m.parameters.lr, rtol=1e-6)


def _basin_worker(rank, world, port, ret):
    import torch.distributed as dist

    import cases
    import oracle_solver
    from smash_b200 import distributed as D
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = cases.france(T=24, sub=(400, 520, 400, 520), ngauge=2)
    from gloo_comm import GlooComm
    mask = D.forward_sharded_by_basin(m, comm=GlooComm(), solver=oracle_solver)
    ret[rank] = (m.output.sparse_qsim_domain.copy(), m.output.qsim.copy(), int(mask.sum()))
    dist.destroy_process_group()


def test_domain_sharded_by_basin_world2_equals_full_run():
    # one domain, its drainage basins spread over the ranks, no exchange on the data path; the gathered series
    # are those of the undivided run
    import cases
    import oracle_solver
    full = cases.france(T=24, sub=(400, 520, 400, 520), ngauge=2)
    oracle_solver.forward(full.setup, full.mesh, full.input_data, full.parameters, full.parameters.copy(), full.states,
                          full.states.copy(), full.output)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_basin_worker, args=(2, 29700 + os.getpid() % 1000, ret), nprocs=2, join=True)
        assert ret[0][2] + ret[1][2] == full.mesh.nac and min(ret[0][2], ret[1][2]) > 0
        for rank in (0, 1):
            assert np.array_equal(ret[rank][0], full.output.sparse_qsim_domain)
            assert np.array_equal(ret[rank][1], full.output.qsim)
