"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/smash_b200.h
declares, fails loudly without a device, and its host-side mesh ordering keeps the integer contract."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cases
from smash_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "smash_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(smash_b200_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    lib = L.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(L.EXPORTS) == declared


def test_version_and_no_device_is_loud():
    lib = L.lib()
    assert b"sm_100a" in lib.smash_b200_version()
    if lib.smash_b200_device_count() == 0:
        import smash_b200
        m = cases.cance(T=24)
        with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
            smash_b200.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(),
                               m.output)
        # the other structures and the interception search are device-only as well: same loud failure, no host path
        m.setup.structure = "vic-a"
        with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
            smash_b200.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(),
                               m.output)
        from smash_b200.solver import _mw_interception_store as ics
        with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
            ics.adjust_interception_store(m.setup, m.mesh, m.input_data, m.parameters, 1, np.ones(24, np.int32))


def test_interception_store_argument_checks():
    # argument errors are reported before the device is needed
    from smash_b200.solver import _mw_interception_store as ics
    m = cases.cance(T=48)
    with pytest.raises(ValueError, match="day_index"):
        ics.adjust_interception_store(m.setup, m.mesh, m.input_data, m.parameters, 2, np.ones(24, np.int32))
    with pytest.raises(RuntimeError, match="nday"):
        ics.adjust_interception_store(m.setup, m.mesh, m.input_data, m.parameters, 1, (np.arange(48) // 24 + 1).astype(np.int32))


def test_unknown_structure_is_refused_on_the_host():
    import smash_b200
    m = cases.cance(T=24)
    m.setup.structure = "gr-z"
    with pytest.raises(ValueError, match="unknown structure"):
        smash_b200.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(), m.output)


def mesh_order(m, block=0):
    pk = L.Packed()
    s, me = L.pack_setup(m.setup, m.mesh, pk), L.pack_mesh(m.mesh, m.setup, pk)
    info = (C.c_int64 * 12)()
    n = int(((m.mesh.active_cell == 1) & (m.mesh._local_active_cell == 1)).sum())
    order, blk, off = (np.zeros(n, np.int32) for _ in range(3))
    L.check(L.lib().smash_b200_mesh_order(C.byref(s), C.byref(me), block, info, L._ip(order), L._ip(blk), L._ip(off)))
    return order, blk, off, list(info)


DROW = np.array([1, 1, 0, -1, -1, -1, 0, 1])   # md_routing_operator.f90:29-30
DCOL = np.array([0, -1, -1, -1, 0, 1, 1, 1])


def check_order(m, order, blk, off, info):
    nrow, ncol = m.mesh.nrow, m.mesh.ncol
    act = ((m.mesh.active_cell == 1) & (m.mesh._local_active_cell == 1)).ravel(order="F")
    assert info[0] == act.sum()
    assert np.array_equal(np.sort(order), np.nonzero(act)[0])            # a permutation of the computed cells
    pos = np.full(nrow * ncol, -1, np.int64)
    pos[order] = np.arange(len(order))
    rank = np.full(nrow * ncol, -1, np.int64)                              # rank in the stored path
    p = m.mesh.path
    ok = (p[0] >= 0) & (p[1] >= 0)
    rank[p[0][ok] + p[1][ok].astype(np.int64) * nrow] = np.arange(ok.sum())
    flwdir, flwacc = m.mesh.flwdir.ravel(order="F"), m.mesh.flwacc.ravel(order="F")
    rows, cols = order % nrow, order // nrow
    d = flwdir[order]
    tr, tc = rows - DROW[d - 1], cols - DCOL[d - 1]                       # downstream cell of every computed cell
    inb = (tr >= 0) & (tr < nrow) & (tc >= 0) & (tc < ncol)
    tgt = np.where(inb, tr + tc * nrow, 0)
    edge = inb & act[tgt] & (flwacc[tgt] > 1)
    src_pos, tgt_pos = np.nonzero(edge)[0], pos[tgt[edge]]
    same_step = rank[order[src_pos]] < rank[tgt[edge]]
    # producers never sit in a later block; inside a block a same-step producer is exactly one tick ahead
    assert np.all(blk[src_pos] <= blk[tgt_pos])
    same_blk = blk[src_pos] == blk[tgt_pos]
    normal = same_blk & same_step
    gap = off[tgt_pos[normal]] - off[src_pos[normal]]
    assert np.all((gap == 1) | (gap == 0))
    assert (gap == 0).sum() == info[6]                                    # only pit pairs share a tick
    lag = same_blk & ~same_step
    assert np.all(off[tgt_pos[lag]] == off[src_pos[lag]])
    assert np.all(blk[src_pos[~same_step]] == blk[tgt_pos[~same_step]]) or info[6] == 0
    assert info[5] == int((~same_blk).sum())
    assert off.max() == info[3]


def test_order_cance():
    m = cases.cance(T=24)
    order, blk, off, info = mesh_order(m)
    check_order(m, order, blk, off, info)
    assert info[1] == 1 and info[2] == 384 and info[6] == 0
    for b in (32, 64, 128):
        check_order(m, *mesh_order(m, b))


def test_order_france():
    m = cases.france(T=24)
    order, blk, off, info = mesh_order(m, 256)
    check_order(m, order, blk, off, info)
    assert info[0] == 906044 and info[6] == 50
    assert info[9] < 200, "chain of dependent blocks should stay short"


def test_mesh_golden_integers(golden):
    # the mesh arrays the solver consumes are the reference's own (smash/tests/io/test_io.py:27-32, test_meshing.py:31-42)
    d = cases.golden("cance_inputs.npz")
    assert np.array_equal(d["flwacc"], golden["mesh_io.flwacc"])
    assert np.array_equal(d["path"], golden["mesh_io.path"])
    assert np.array_equal(d["flwdir"], golden["mesh_io.flwdir"].astype(np.int32))
    assert np.array_equal(d["flwacc"], golden["xy_mesh.flwacc"])
    assert np.array_equal(d["gauge_pos"], golden["mesh_io.gauge_pos"])


def mesh_chains(m):
    pk = L.Packed()
    me = L.pack_mesh(m.mesh, m.setup, pk)
    info = (C.c_int64 * 8)()
    n = int(((m.mesh.active_cell == 1) & (m.mesh._local_active_cell == 1)).sum())
    cell, task, pos, down = (np.zeros(n, np.int32) for _ in range(4))
    L.check(L.lib().smash_b200_mesh_chains(C.byref(me), info, L._ip(cell), L._ip(task), L._ip(pos), L._ip(down)))
    return cell, task, pos, down, list(info)


def check_chains(m, cell, task, pos, down, info):
    """Invariants the routing pass of the split engine relies on (route_graph.hpp)."""
    nrow = m.mesh.nrow
    n = info[0]
    flwacc = m.mesh.flwacc.ravel(order="F")[cell]
    # cell order = the stored path restricted to computed cells
    p = m.mesh.path
    ok = (p[0] >= 0) & (p[1] >= 0)
    flat = p[0][ok] + p[1][ok].astype(np.int64) * nrow
    act = ((m.mesh.active_cell == 1) & (m.mesh._local_active_cell == 1)).ravel(order="F")
    assert np.array_equal(cell, flat[act[flat]])
    # every gathering cell is routed by exactly one task, lone source cells by none
    assert np.all(task[flwacc > 1] >= 0)
    assert info[6] == int((flwacc <= 1).sum())
    assert len(np.unique(task[task >= 0])) == info[1] + info[2]
    has = down >= 0
    src, dst = np.nonzero(has)[0], down[has]
    assert np.all(flwacc[dst] > 1)                                         # md_routing_operator.f90:35
    same = task[src] == task[dst]
    # inside a chain the consumer directly follows its producer; across tasks the producer's task runs earlier
    chain = same & (task[src] < info[1])
    assert np.all(pos[dst[chain]] == pos[src[chain]] + 1)
    cross = ~same & (task[src] >= 0)
    nded = info[7] >> 1                                                    # longest chains: dedicated warps, last chain tasks
    dedicated = (task >= info[1] - nded) & (task < info[1])
    assert np.all((task[src[cross]] < task[dst[cross]]) | dedicated[src[cross]])
    # a producer that feeds another task is the last cell of its own chain
    last = np.zeros(info[1] + info[2], np.int64)
    np.maximum.at(last, task[task >= 0], pos[task >= 0])
    assert np.all(pos[src[cross]] == last[task[src[cross]]])
    # pit pairs: two cells that gather each other, routed together
    pair = same & (task[src] >= info[1])
    assert pair.sum() == 2 * info[2]
    assert np.all(down[dst[pair]] == src[pair])


def test_chains_cance():
    m = cases.cance(T=24)
    out = mesh_chains(m)
    check_chains(m, *out)
    info = out[-1]
    assert info[0] == 383 and info[2] == 0 and info[5] == 31 and info[7] & 1 == 1


def test_chains_france():
    m = cases.france(T=24)
    out = mesh_chains(m)
    check_chains(m, *out)
    info = out[-1]
    assert info[0] == 906044 and info[2] == 50 and info[7] & 1 == 1 and 0 < info[7] >> 1 <= 256
    assert info[3] <= 16 and info[5] <= 820


def tick_schedule(m, shallow_acc, nwarp, nwin):
    pk = L.Packed()
    me = L.pack_mesh(m.mesh, m.setup, pk)
    info = (C.c_int64 * 12)()
    n = int(((m.mesh.active_cell == 1) & (m.mesh._local_active_cell == 1)).sum())
    unit, sigma = np.zeros(n, np.int32), np.zeros(n, np.int32)
    L.check(L.lib().smash_b200_mesh_tick_schedule(C.byref(me), shallow_acc, nwarp, nwin, info, L._ip(unit), L._ip(sigma)))
    return list(info), unit, sigma


@pytest.mark.parametrize("shallow_acc,nwarp", [(4, 7), (32, 64), (1, 3)])
def test_tick_schedule_cance(shallow_acc, nwarp):
    # the ticket order of the tick pass (tick_kernels.cu): every ticket only reads smaller keys and the host replay of the
    # warps' in-order walks completes (no deadlock), whatever the number of warps
    m = cases.cance(T=24)
    info, unit, sigma = tick_schedule(m, shallow_acc, nwarp, 5)
    assert info[0] == 383 and info[10] == 1 and info[9] == (info[1] + info[2]) * 5
    cell, task, pos, down, _ = mesh_chains(m)
    has = (down >= 0) & (unit >= 0) & (unit[np.maximum(down, 0)] >= 0)
    diff = unit[down[has]] != unit[has]
    assert np.all(sigma[down[has]][diff] > sigma[has][diff])                 # a consumer's stage lies above its producer's


def test_tick_schedule_france():
    m = cases.france(T=8)
    info, unit, sigma = tick_schedule(m, 32, 4736, 3)
    assert info[0] == 906044 and info[6] == 100 and info[10] == 1 and info[9] == (info[1] + info[2]) * 3
    cell, task, pos, down, ci = mesh_chains(m)
    assert info[4] + info[5] + info[6] == 906044 - ci[6]                     # every gathering cell has a class
    assert info[3] <= 96 and info[7] <= 24 and info[8] >= 700               # stages stay shallow: the Loire is a pipeline of reaches
    has = (down >= 0) & (unit >= 0) & (unit[np.maximum(down, 0)] >= 0)
    diff = unit[down[has]] != unit[has]
    assert np.all(sigma[down[has]][diff] > sigma[has][diff])


def test_basin_masks_partition_the_domain():
    # a France run shards by basin -- whole basins per rank, nothing gathered across ranks
    from smash_b200 import distributed as D
    m = cases.france(T=8)
    labels, nb = D.basin_labels(m.mesh, m.setup)
    act = m.mesh.active_cell == 1
    assert nb == 3434 and np.bincount(labels[labels >= 0]).max() == 136170  # the largest basin (the Loire) = the largest flwacc
    assert np.all(labels[act] >= 0) and np.all(labels[~act] == -1)
    masks, load = D.basin_masks(m.mesh, 4, m.setup)
    total = sum(mk.astype(np.int64) for mk in masks)
    assert np.array_equal(total, act.astype(np.int64))                       # a partition of the active cells
    cell, task, pos, down, info = mesh_chains(m)
    own = np.full(m.mesh.nrow * m.mesh.ncol, -1)
    for r, mk in enumerate(masks):
        own[np.flatnonzero(mk.ravel(order="F") == 1)] = r
    has = down >= 0
    assert np.array_equal(own[cell[has]], own[cell[down[has]]])              # no gather crosses a rank boundary
    biggest = np.bincount(labels[labels >= 0]).max()
    assert load.max() - load.min() <= biggest and load.sum() == act.sum()


def test_flow_accumulation_bit_exact(golden):
    # the reference's flow accumulation (mw_meshing.f90:204-233) recomputed from the flow directions: integer-exact on the
    # Cance catchment window (golden mesh_io.flwacc / xy_mesh.flwacc) and on the France mesh (288 pit cells)
    from smash_b200.mesh import flow_accumulation
    act = golden["mesh_io.active_cell"] == 1
    fa = flow_accumulation(golden["mesh_io.flwdir"], mask=act)
    assert np.array_equal(fa[act], golden["mesh_io.flwacc"][act])
    assert np.array_equal(fa[act], golden["xy_mesh.flwacc"][act])
    f = cases.golden("france_mesh.npz")
    fa = flow_accumulation(f["flwdir"])
    actf = f["active_cell"] == 1
    assert np.array_equal(fa[actf], f["flwacc"][actf])


def test_path_is_sorted_by_flow_accumulation(golden):
    # the stored cell order (meshing.py:216-218: argsort of flwacc) is ascending in flwacc, so every cell comes after the
    # cells it gathers -- the property the solver's single sweep over `path` relies on (md_forward_structure.f90:82-92)
    for fa, path in ((golden["mesh_io.flwacc"], golden["mesh_io.path"]),
                     (cases.golden("france_mesh.npz")["flwacc"], cases.golden("france_mesh.npz")["path"])):
        ok = (path[0] >= 0) & (path[1] >= 0)
        seq = fa[path[0][ok], path[1][ok]]
        assert np.all(np.diff(seq.astype(np.int64)) >= 0)
