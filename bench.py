#!/usr/bin/env python
"""bench.py -- forward / adjoint solver throughput on B200 (contract: the task statement, section 4).

N = 1 (the configuration BASELINE.json's HBM-roofline metric is quoted on, configs[3]):
  France 1 km flow-direction mesh (906 044 active cells, mesh_France.hdf5 -> tests/golden/france_mesh.npz), T = 720
  synthetic hourly steps (SURVEY.md 8d recipe), gr-a, default parameters, save_qsim_domain as in setup_France.yaml.
  One "step" = one forward run over the whole mesh and all T time steps.
    value     active-cell-timesteps/s, forcing resident in HBM (plan API), kernel launches only, result stays in HBM
    e2e       the same metric through the drop-in call smash_b200.forward(...) with HOST arrays: per step the forcing
              (5.2 GB) goes host->device and the domain discharge (2.6 GB) comes back
    roofline  the WHOLE STEP against SURVEY 8(d): 12 algorithmic bytes per active cell-step (prcp + pet read, q written)
              over the CUDA-event step time; per-kernel rows (time, share of the step, DRAM traffic from the committed
              ncu capture) under roofline.kernels
    gradient  fwd + adjoint gradient on the same mesh (evals/s, fraction of the 40 B per cell-step roofline)
    ensemble  32 768-member Cance ensemble on this one GPU (the sharded headline of N > 1, here for the scaling ratio)
    cpu_baseline  the C oracle (restatement of the Fortran solver, oracle/) on one host core, bounded sample
    parity    measured max errors of the device path against the oracle on small cases (not pass / fail)

N > 1 (torchrun, one process per GPU; BASELINE.json configs[2]): the 32 768-member Cance ensemble
  (compute_multiple_run) with contiguous member blocks per rank -- fixed total work, "scaling": "strong", no data-path
  collective; one all-gather of the costs afterwards.  Secondary first-class keys: the regionalised multi-catchment
  calibration (one all-reduce per evaluation) and one France run split by drainage basin.  The collective is NCCL through
  the library's own communicator (no PyTorch).

--impl reference: the oracle port timed on the host: the France sample on one thread at N = 1 (a single forward run of the
reference is single-threaded), the ensemble sample on all host threads (OpenMP over members, mw_multiple_run.f90:96) at N > 1.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC_FRANCE = "forward active-cell-timesteps/s (France 1km mesh, gr-a); gradient: fwd+adjoint evals/s; % HBM roofline"
METRIC_ENS = "forward active-cell-timesteps/s (32768-member Cance ensemble, compute_multiple_run, members sharded over the GPUs)"
UNIT = "cell-timesteps/s"
ENS_MEMBERS = 32768
CANCE_CELLS, CANCE_T = 383, 1440


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe), every 10 ms."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "10",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            time.sleep(0.25)                      # nvidia-smi needs ~0.2 s before its first line
            self.skip = len(self.rows)            # lines printed before the timed region
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.05)
        self.proc.terminate()
        rows = self.rows[self.skip:] or self.rows[-1:]
        sm = [float(r[0]) for r in rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def ensemble_sample(ns):
    rng = np.random.RandomState(99)
    bounds = [(1e-6, 1e3), (1e-6, 1e3), (-50.0, 50.0), (1e-6, 1e3)]                 # generate_samples.py:357-364
    return np.asfortranarray(np.stack([rng.uniform(lo, hi, ns) for lo, hi in bounds]).astype(np.float32))


# ------------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    import cases
    import oracle
    if args.gpus <= 1:
        Ts = args.ref_steps
        m = cases.france(T=Ts, seed=0)                     # save_qsim_domain = True, as the device arm
        units = m.mesh.nac * Ts
        cores = 1

        def step():
            oracle.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(), m.output)

        metric = METRIC_FRANCE
        workload = (f"France 1km mesh forward gr-a, nac=906044, T={args.T}, save_qsim_domain (reference arm: bounded sample, "
                    f"the first {Ts} time steps per step)")
        sample = (f"France mesh, {Ts} of {args.T} time steps per step ({units} cell-steps), save_qsim_domain, 1 thread (a single "
                  "forward run of the reference is single-threaded)")
    else:
        ns = args.ref_members
        cores = os.cpu_count() or 1
        m = cases.cance()
        smp = ensemble_sample(ns)
        cost, q0 = np.zeros(ns, np.float32), np.zeros((0,), np.float32)
        units = ns * CANCE_CELLS * CANCE_T

        def step():
            oracle.compute_multiple_run(m.setup, m.mesh, m.input_data, m.parameters, m.states, m.output, smp,
                                        cases.IND_CP_CFT_EXC_LR, cost, q0, nthreads=cores)

        metric = METRIC_ENS
        workload = f"Cance {ENS_MEMBERS}-member ensemble (reference arm: bounded sample of {ns} members per step)"
        sample = f"{ns} of {ENS_MEMBERS} members per step ({units} cell-steps), OpenMP over members, {cores} threads"
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = units / dt
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak" if args.gpus <= 1 else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": workload},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# ensemble (configs[2]): device-resident value + e2e through compute_multiple_run
# ------------------------------------------------------------------------------------------------
def bench_ensemble(lib, L, smash_b200, cases, comm, rank, world, steps, warmup, sampler=None):
    from smash_b200 import distributed as sdist
    m = cases.cance()
    m.input_data._forcing_version = 1
    smp = ensemble_sample(ENS_MEMBERS)
    sl = sdist.member_slice(ENS_MEMBERS, rank, world)
    n_loc = sl.stop - sl.start
    units = ENS_MEMBERS * CANCE_CELLS * CANCE_T
    chunk = min(n_loc, 8192)
    pk = L.Packed()
    s_, m_, i_ = L.pack_setup(m.setup, m.mesh, pk), L.pack_mesh(m.mesh, m.setup, pk), L.pack_input(m.input_data, m.setup, m.mesh, pk)
    p_, st_ = L.pack_parameters(m.parameters, pk), L.pack_states(m.states, pk)
    plan = C.c_void_p()
    L.check(lib.smash_b200_plan_create(C.byref(s_), C.byref(m_), chunk, C.byref(plan)))
    L.check(lib.smash_b200_plan_set_forcing(plan, C.byref(s_), C.byref(i_)))
    ind = np.ascontiguousarray(cases.IND_CP_CFT_EXC_LR, dtype=np.int32)
    loc = np.asfortranarray(smp[:, sl])
    chunks = [np.asfortranarray(loc[:, k:k + chunk]) for k in range(0, n_loc, chunk)]
    if chunks and chunks[-1].shape[1] < chunk:                       # pad the last chunk (a plan has a fixed member count)
        pad = np.repeat(chunks[-1][:, -1:], chunk - chunks[-1].shape[1], axis=1)
        chunks[-1] = np.asfortranarray(np.concatenate([chunks[-1], pad], axis=1))
    ms = C.c_float(0.0)
    kt = (C.c_float * 5)()

    def dev_step():
        tot, kk = 0.0, np.zeros(2)
        for ch in chunks:
            L.check(lib.smash_b200_plan_set_fields(plan, C.byref(p_), C.byref(st_), L._fp(ch), L._ip(ind), 4))
            L.check(lib.smash_b200_plan_run_forward(plan, C.byref(ms)))
            tot += ms.value
            L.check(lib.smash_b200_plan_kernel_times(plan, kt))
            kk += [kt[0], kt[1]]
        return tot, kk

    for _ in range(warmup):
        dev_step()
    if sampler:
        sampler.start()
    comm_barrier(comm)
    t0 = time.perf_counter()
    dev_ms, kks = [], []
    for _ in range(steps):
        t, kk = dev_step()
        dev_ms.append(t)
        kks.append(kk)
    comm_barrier(comm)
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    info = (C.c_int64 * 12)()
    lib.smash_b200_plan_info(plan, info)
    launches = int(info[7]) * len(chunks)
    lib.smash_b200_plan_destroy(plan)
    dms = comm_max(comm, float(np.mean(dev_ms)))                     # CUDA-event time of the kernels, max over ranks
    wall = comm_max(comm, wall)
    kk = np.mean(np.array(kks), axis=0)
    # e2e: the drop-in call with host arrays (sample in, costs out; the forcing of the mesh is resident after the first call)
    cost, q0 = np.zeros(ENS_MEMBERS, np.float32), np.zeros((0,), np.float32)

    def e2e_step():
        if world > 1:
            sdist.multiple_run_sharded(m.setup, m.mesh, m.input_data, m.parameters, m.states, m.output, smp,
                                       cases.IND_CP_CFT_EXC_LR, cost, q0, comm=comm)
        else:
            smash_b200.compute_multiple_run(m.setup, m.mesh, m.input_data, m.parameters, m.states, m.output, smp,
                                            cases.IND_CP_CFT_EXC_LR, cost, q0)

    e2e_step()
    comm_barrier(comm)
    ne = max(1, min(3, steps))
    t0 = time.perf_counter()
    for _ in range(ne):
        e2e_step()
    comm_barrier(comm)
    e2e_wall = comm_max(comm, time.perf_counter() - t0) / ne
    return {
        "units": units, "device_ms": dms, "wall_ms_per_step": wall / steps * 1e3, "value": units / (dms * 1e-3),
        "members": ENS_MEMBERS, "members_per_rank": n_loc, "members_per_launch": chunk, "launches_per_step": launches,
        "kernels_ms": {"vertical_forward_kernel": float(kk[0]), "route_members_kernel": float(kk[1])},
        "e2e": {"value": units / e2e_wall, "unit": UNIT, "ms_per_step": e2e_wall * 1e3,
                "h2d_bytes_per_step": int(smp.nbytes // world), "d2h_bytes_per_step": int(4 * n_loc)},
        "cost_checksum": float(np.sum(cost[np.isfinite(cost)], dtype=np.float64)), "clocks": clocks,
    }


def comm_barrier(comm):
    if comm is not None:
        comm.barrier()


def comm_max(comm, x):
    if comm is None:
        return x
    a = np.array([x], dtype=np.float64)
    comm.allreduce(a, "max")
    return float(a[0])


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--T", type=int, default=720, help="time steps of the France workload")
    ap.add_argument("--ref-steps", type=int, default=8, help="France time steps per reference-arm step")
    ap.add_argument("--ref-members", type=int, default=128, help="ensemble members per reference-arm step (N > 1)")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import cases
    import smash_b200
    from smash_b200 import _lib as L
    lib = L.lib()
    if lib.smash_b200_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: smash_b200 has no CPU fallback")
    L.check(lib.smash_b200_set_device(local_rank))
    comm = None
    if world > 1:
        from smash_b200 import distributed as sdist
        comm = sdist.NcclComm(rank=rank, world=world, device=local_rank)
    peak, peak_src = measured_peaks()

    if world > 1:
        bench_sharded(args, lib, L, smash_b200, cases, comm, rank, world, local_rank, peak, peak_src)
        comm.close()
        return
    bench_france(args, lib, L, smash_b200, cases, local_rank, peak, peak_src)


# ------------------------------------------------------------------------------------------------
# N > 1: the sharded ensemble as the headline, regional calibration and basin-split France next to it
# ------------------------------------------------------------------------------------------------
def bench_sharded(args, lib, L, smash_b200, cases, comm, rank, world, local_rank, peak, peak_src):
    from smash_b200 import distributed as sdist
    ens = bench_ensemble(lib, L, smash_b200, cases, comm, rank, world, args.steps, args.warmup, ClockSampler(local_rank))
    out = {
        "metric": METRIC_ENS, "value": ens["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ens["device_ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"Cance {ENS_MEMBERS}-member ensemble, T={CANCE_T}, nac={CANCE_CELLS}, generate_samples(random_state=99) parameter sets",
                   "members_per_rank": ens["members_per_rank"], "members_per_launch": ens["members_per_launch"],
                   "l2": "inputs larger than L2: the members' row buffers (2.4 MB per member) stream through HBM; the shared "
                         "forcing (4.4 MB) is L2-resident by nature of the workload",
                   "parallelism": f"members sharded over {world} ranks, no data-path collective, one all-gather of the costs",
                   "wall_ms_per_step": ens["wall_ms_per_step"], "cost_checksum": ens["cost_checksum"]},
        "roofline": {"bound": "hbm", "achieved": 8.0 * ens["units"] / (ens["device_ms"] * 1e-3) / 1e9 / world, "peak": peak, "unit": "GB/s",
                     "frac": 8.0 * ens["units"] / (ens["device_ms"] * 1e-3) / 1e9 / peak / world, "traffic": None, "peak_source": peak_src,
                     "note": "per GPU, against 8 B per member cell-step (SURVEY 8d); the forcing is shared by the members and "
                             "L2-resident, so this configuration is bound by instruction issue and the row traffic of the split "
                             "passes, not by algorithmic HBM bytes", "kernels_ms": ens["kernels_ms"]},
        "cpu_baseline": None, "e2e": ens["e2e"], "gpu_launches": args.steps * ens["launches_per_step"], "clocks": ens["clocks"],
    }
    if not args.no_extra:
        # configs[4]: regionalised calibration, 8 Cance-sized catchments per rank (different periods / observations), shared
        # hyper-polynomial mapping, ONE all-reduce of (cost, gradient) per evaluation
        def catchments():
            cs = []
            for k in range(8):
                mk = cases.cance(T=1440 - 24 * ((rank * 8 + k) % 5))
                mk.input_data.qobs = np.asfortranarray(mk.input_data.qobs * np.float32(1.0 + 0.01 * (rank * 8 + k)))
                mk.input_data._forcing_version = 100 + rank * 8 + k
                cases.set_optimize(mk.setup, mk.mesh, jobs_fun=("nse",), mapping="hyper-polynomial")
                mk.setup._optimize.optim_parameters[[1, 3, 6, 15]] = 1
                mk.setup._optimize.maxiter = 5
                mk.setup._optimize.verbose = False
                cs.append((mk.setup, mk.mesh, mk.input_data, mk.parameters, mk.states, mk.output))
            return cs
        sdist.optimize_hyper_lbfgsb_sharded(catchments(), comm=comm)
        cs = catchments()
        comm_barrier(comm)
        t0 = time.perf_counter()
        sdist.optimize_hyper_lbfgsb_sharded(cs, comm=comm)
        comm_barrier(comm)
        dtr = comm_max(comm, time.perf_counter() - t0)
        out["regional_calibration"] = {"catchments": 8 * world, "catchments_per_rank": 8, "iterations": 5, "wall_s": dtr,
                                       "s_per_iteration": dtr / 5, "scaling": "weak",
                                       "collective": "NCCL all-reduce of 1 + 4 x 5 float64 per evaluation (library communicator)",
                                       "rank0_cost": float(cs[0][5].cost)}
        # ONE France run with its 3 434 drainage basins spread over the ranks (whole basins per rank, no exchange)
        model = cases.france(T=args.T, seed=0)
        units = int(model.mesh.nac) * args.T
        masks, load = sdist.basin_masks(model.mesh, world, model.setup)
        model.mesh._local_active_cell = masks[rank]
        pkb = L.Packed()
        sb_, mb_, ib_ = L.pack_setup(model.setup, model.mesh, pkb), L.pack_mesh(model.mesh, model.setup, pkb), L.pack_input(model.input_data, model.setup, model.mesh, pkb)
        pb_, stb_ = L.pack_parameters(model.parameters, pkb), L.pack_states(model.states, pkb)
        planb = C.c_void_p()
        L.check(lib.smash_b200_plan_create(C.byref(sb_), C.byref(mb_), 1, C.byref(planb)))
        L.check(lib.smash_b200_plan_set_forcing(planb, C.byref(sb_), C.byref(ib_)))
        L.check(lib.smash_b200_plan_set_fields(planb, C.byref(pb_), C.byref(stb_), None, None, 0))
        msb = C.c_float(0.0)
        for _ in range(3):
            L.check(lib.smash_b200_plan_run_forward(planb, C.byref(msb)))
        comm_barrier(comm)
        nrep, tot = 20, 0.0
        for _ in range(nrep):
            L.check(lib.smash_b200_plan_run_forward(planb, C.byref(msb)))
            tot += msb.value
        dtb = comm_max(comm, tot / nrep)
        lib.smash_b200_plan_destroy(planb)
        out["france_split_by_basin"] = {"cell_timesteps_per_s": units / (dtb * 1e-3), "device_ms_per_run": dtb, "scaling": "strong",
                                        "cells_per_rank": [int(x) for x in load], "collective": "none",
                                        "frac_of_12B_roofline_all_gpus": 12.0 * units / (dtb * 1e-3) / 1e9 / (peak * world)}
    if rank == 0:
        print(json.dumps(out))


# ------------------------------------------------------------------------------------------------
# N = 1: France forward (headline), gradient, ensemble, e2e, CPU baseline, parity numbers
# ------------------------------------------------------------------------------------------------
def bench_france(args, lib, L, smash_b200, cases, local_rank, peak, peak_src):
    T = args.T
    t0 = time.time()
    model = cases.france(T=T, seed=0)
    t_build = time.time() - t0
    nac = int(model.mesh.nac)
    units = nac * T
    pk = L.Packed()
    s_, m_, i_ = L.pack_setup(model.setup, model.mesh, pk), L.pack_mesh(model.mesh, model.setup, pk), L.pack_input(model.input_data, model.setup, model.mesh, pk)
    p_, st_ = L.pack_parameters(model.parameters, pk), L.pack_states(model.states, pk)
    plan = C.c_void_p()
    L.check(lib.smash_b200_plan_create(C.byref(s_), C.byref(m_), 1, C.byref(plan)))
    L.check(lib.smash_b200_plan_set_forcing(plan, C.byref(s_), C.byref(i_)))
    L.check(lib.smash_b200_plan_set_fields(plan, C.byref(p_), C.byref(st_), None, None, 0))
    info = (C.c_int64 * 12)()
    lib.smash_b200_plan_info(plan, info)
    ms = C.c_float(0.0)
    kt = (C.c_float * 5)()
    for _ in range(args.warmup):
        L.check(lib.smash_b200_plan_run_forward(plan, C.byref(ms)))
    sampler = ClockSampler(local_rank)
    sampler.start()
    kernel_ms, per_kernel = [], []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        L.check(lib.smash_b200_plan_run_forward(plan, C.byref(ms)))   # returns after cudaEventSynchronize
        kernel_ms.append(ms.value)
        L.check(lib.smash_b200_plan_kernel_times(plan, kt))           # events recorded inside the run, read after it
        per_kernel.append([kt[i] for i in range(5)])
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    kms = float(np.mean(kernel_ms))
    value = units * args.steps / wall
    chk = C.c_double(0.0)
    L.check(lib.smash_b200_plan_checksum(plan, C.byref(chk)))
    lib.smash_b200_plan_info(plan, info)
    launches_per_step = int(info[7])

    # ---- roofline: the whole step against 12 B per cell-step; kernels with their time, share and measured DRAM traffic
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass
    step_bytes = 12.0 * units
    pk_ms = np.mean(np.array(per_kernel), axis=0)
    window_pass = lib.smash_b200_plan_stat(plan, b"tick_pass") == 1.0
    sub_engine = lib.smash_b200_plan_stat(plan, b"sub_engine") == 1.0
    names = ["tick_forward_kernel" if window_pass else "vertical_forward_kernel", "route_forward_kernel", "rows_to_domain_kernel"]
    if sub_engine:
        names = ["sub_forward_kernel", "route_pairs_kernel", "rows_to_domain_kernel"]
    kernels = {names[i]: {"ms": float(pk_ms[i]), "share": float(pk_ms[i] / kms), "dram_traffic_bytes": traffic.get(names[i])}
               for i in range(3) if pk_ms[i] > 0.02}
    tr = [k["dram_traffic_bytes"] for k in kernels.values()]
    roofline = {"bound": "hbm", "achieved": step_bytes / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": step_bytes / (kms * 1e-3) / 1e9 / peak, "traffic": float(sum(tr)) if tr and all(t is not None for t in tr) else None,
                "peak_source": peak_src, "scope": "whole forward step (all kernels), SURVEY 8(d): 12 B per active cell-step",
                "algorithmic_bytes_per_step": step_bytes, "device_ms": kms, "kernels": kernels}

    # ---- gradient (first class): forward sweep with the tape + cost + reverse sweeps
    gradient = None
    if not args.no_extra:
        f_ms, r_ms = C.c_float(0), C.c_float(0)
        L.check(lib.smash_b200_plan_run_gradient(plan, C.byref(f_ms), C.byref(r_ms)))
        gms, gk = [], []
        for _ in range(max(3, args.steps // 5)):
            L.check(lib.smash_b200_plan_run_gradient(plan, C.byref(f_ms), C.byref(r_ms)))
            gms.append((f_ms.value, r_ms.value))
            L.check(lib.smash_b200_plan_kernel_times(plan, kt))
            gk.append([kt[i] for i in range(5)])
        gf, gr = float(np.mean([g[0] for g in gms])), float(np.mean([g[1] for g in gms]))
        gk = np.mean(np.array(gk), axis=0)
        ck = lib.smash_b200_plan_stat(plan, b"checkpoint") == 1.0
        gradient = {"evals_per_s": 1e3 / (gf + gr), "cell_timesteps_per_s": units / ((gf + gr) * 1e-3), "ms_per_eval": gf + gr,
                    "forward_tape_ms": gf, "reverse_ms": gr, "frac_of_40B_roofline": 40.0 * units / ((gf + gr) * 1e-3) / 1e9 / peak,
                    "tape": "checkpointed, 256-step windows (forward sweeps: 2)" if ck else "store-all in HBM (forward sweeps: 1)",
                    "tape_gb": lib.smash_b200_plan_stat(plan, b"tape_bytes") / 1e9,
                    "kernels_ms": {"vertical_forward_kernel(tape)": float(gk[0]), "route_forward_kernel(tape)": float(gk[1]),
                                   "route_adjoint_kernel": float(gk[3]), "vertical_adjoint_kernel": float(gk[4])}}
    lib.smash_b200_plan_destroy(plan)

    # ---- the other engines on the same workload (plan API, results resident): the row-based passes and the tick pass
    if not args.no_extra:
        roofline["alternatives"] = {}
        alts = {"row_passes": ({"sub_engine": 0, "tick_pass": 0}, ["vertical_forward_kernel", "route_forward_kernel"],
                               "reservoir pass per cell + routing scan per heavy-path chain over rows [cell][time] (the engine of "
                               "the drop-in calls and of the gradient): 2.2 x the algorithmic DRAM traffic, DESIGN.md section 3"),
                "tick_pass": ({"sub_engine": 0, "tick_pass": 1}, ["tick_forward_kernel"],
                              "every cell advances 8 steps per ticket, discharge blocks handed from producer to consumer; bound by the "
                              "latency of a ticket, DESIGN.md section 3b")}
        for name, (opts, kerns, note) in alts.items():
            for k_, v_ in opts.items():
                lib.smash_b200_set_option(k_.encode(), v_)
            try:
                plan2 = C.c_void_p()
                L.check(lib.smash_b200_plan_create(C.byref(s_), C.byref(m_), 1, C.byref(plan2)))
                L.check(lib.smash_b200_plan_set_forcing(plan2, C.byref(s_), C.byref(i_)))
                L.check(lib.smash_b200_plan_set_fields(plan2, C.byref(p_), C.byref(st_), None, None, 0))
                tms = []
                for k in range(args.warmup + 5):
                    L.check(lib.smash_b200_plan_run_forward(plan2, C.byref(ms)))
                    if k >= args.warmup:
                        tms.append(ms.value)
                chk2 = C.c_double(0.0)
                L.check(lib.smash_b200_plan_checksum(plan2, C.byref(chk2)))
                lib.smash_b200_plan_destroy(plan2)
                tr2 = [traffic.get(kn) for kn in kerns]
                tr2 = float(sum(tr2)) if all(x is not None for x in tr2) else None
                roofline["alternatives"][name] = {
                    "ms_per_step": float(np.mean(tms)), "frac": step_bytes / (float(np.mean(tms)) * 1e-3) / 1e9 / peak,
                    "dram_traffic_bytes": tr2, "traffic_over_algorithmic": (tr2 / step_bytes) if tr2 else None,
                    "checksum_q": chk2.value, "note": note}
            except Exception as exc:                      # a secondary block never costs the headline line
                roofline["alternatives"][name] = {"error": str(exc), "note": note}
            finally:
                lib.smash_b200_set_option(b"sub_engine", -1)
                lib.smash_b200_set_option(b"tick_pass", 0)

    # ---- the other structures on the same mesh and forcing (forward only; reservoir pass of struct_kernels.cu + the routing passes)
    structures = None
    if not args.no_extra:
        structures = {}
        for name in ("gr-b", "gr-c", "gr-d", "vic-a"):
            model.setup.structure = name
            model.parameters.ci[...] = 2.0 if name in ("gr-b", "gr-c") else 1e-6
            try:
                pk2 = L.Packed()
                s2 = L.pack_setup(model.setup, model.mesh, pk2)
                p2, st2 = L.pack_parameters(model.parameters, pk2), L.pack_states(model.states, pk2)
                plan2 = C.c_void_p()
                L.check(lib.smash_b200_plan_create(C.byref(s2), C.byref(m_), 1, C.byref(plan2)))
                L.check(lib.smash_b200_plan_set_forcing(plan2, C.byref(s2), C.byref(i_)))
                L.check(lib.smash_b200_plan_set_fields(plan2, C.byref(p2), C.byref(st2), None, None, 0))
                tms = []
                for k in range(args.warmup + 3):
                    L.check(lib.smash_b200_plan_run_forward(plan2, C.byref(ms)))
                    if k >= args.warmup:
                        tms.append(ms.value)
                kt = (C.c_float * 5)()
                lib.smash_b200_plan_kernel_times(plan2, kt)
                chk2 = C.c_double(0.0)
                L.check(lib.smash_b200_plan_checksum(plan2, C.byref(chk2)))
                lib.smash_b200_plan_destroy(plan2)
                structures[name] = {"ms_per_step": float(np.mean(tms)), "cell_timesteps_per_s": units / (float(np.mean(tms)) * 1e-3),
                                    "frac_of_12B_roofline": step_bytes / (float(np.mean(tms)) * 1e-3) / 1e9 / peak,
                                    "kernels_ms": {("vertical_forward_kernel (gr-d mode)" if name == "gr-d" else "vertical_struct_kernel"): float(kt[0]),
                                                   "route_forward_kernel": float(kt[1])},
                                    "checksum_q": chk2.value}
            except Exception as exc:                      # a secondary block never costs the headline line
                structures[name] = {"error": str(exc)}
            finally:
                model.setup.structure = "gr-a"
                model.parameters.ci[...] = 1e-6
        structures["note"] = ("forward runs of md_forward_structure.f90:216-931 on the row passes; same 12 B per cell-step of algorithmic "
                              "traffic as gr-a; the reservoir pass is instruction-bound (divisions, powf of vic_infiltration); gr-d runs on "
                              "gr-a's kernels with the shares 1 / 0 and no exchange (DESIGN.md section 3f)")

    # ---- the ANN mapping's forward pass at France scale: the graph of _ann_optimize.py:143-168 for nd = 6 descriptors
    ann = None
    if not args.no_extra:
        try:
            from smash_b200.net import Net
            nd_, n1 = 6, int(round(np.sqrt(nac * 6) * 2 / 3))
            net = Net()
            net.add("dense", {"input_shape": (nd_,), "neurons": n1, "kernel_initializer": "glorot_uniform"})
            net.add("activation", {"name": "relu"})
            net.add("dense", {"neurons": round(n1 / 2), "kernel_initializer": "glorot_uniform"})
            net.add("activation", {"name": "relu"})
            net.add("dense", {"neurons": 4, "kernel_initializer": "glorot_uniform"})
            net.add("activation", {"name": "sigmoid"})
            net.compile("adam", {"learning_rate": 0.003}, random_state=11)
            xd = np.random.default_rng(3).uniform(0.0, 1.0, (nac, nd_)).astype(np.float32)
            tm, best = {}, None
            for _ in range(3):
                yd = net._predict_device(xd, timing=tm)
                best = dict(tm) if best is None or tm["ms"] < best["ms"] else best
            from smash_b200.net import DeviceChain
            dev = DeviceChain(net, xd)                       # one training step with the chain resident on the device
            gyd = np.random.default_rng(4).uniform(0.5, 1.5, (nac, 4)).astype(np.float32) * np.float32(1e-6)
            for _ in range(2):
                dev.forward()
                dev.backward(gyd)
            train_ms = (dev.ms_forward, dev.ms_backward)
            dev.close()
            ref = net._predict(xd[:2000].astype(np.float64))
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                    bf16 = float(json.load(f)["bf16_tflops"])
            except Exception:
                bf16 = 2250.0
            ann = {"rows": nac, "graph": [nd_, n1, round(n1 / 2), 4], "device_ms": best["ms"], "tflops": best["tflops"], "dtype": "tf32 (f32 accumulate)",
                   "frac_of_tf32_peak": best["tflops"] / (bf16 / 2.0), "tf32_peak_tflops": bf16 / 2.0,
                   "peak_source": "half the measured dense bf16 rate of MEASURED_PEAKS.json (TF32 runs at half the bf16 rate)",
                   "max_abs_err_vs_numpy_f64": float(np.abs(yd[:2000] - ref).max()),
                   "training_step_ms": {"forward": train_ms[0], "backward": train_ms[1],
                                        "note": "backward = activation derivatives, column sums, grad_weight = a^T g (stream-K over the rows) and "
                                                "g W^T per layer, all on the device; the optimiser update stays on the host"},
                   "kernel": "CUTLASS sm100 collective (TMA + tcgen05.mma kind::tf32, accumulators in TMEM), bias + activation fused"}
        except Exception as exc:                          # a secondary block never costs the headline line
            ann = {"error": str(exc)}

    # ---- e2e through the drop-in call with host buffers
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    model.input_data._forcing_version = 0
    lib.smash_b200_set_option(b"pin_host", 1)      # the forcing / output arrays live until clear_cache() below
    par_bgd, sta_bgd = model.parameters.copy(), model.states.copy()

    def e2e_step():
        smash_b200.forward(model.setup, model.mesh, model.input_data, model.parameters, par_bgd, model.states, sta_bgd, model.output)

    t0 = time.perf_counter()
    e2e_step()
    first_call_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_wall = time.perf_counter() - t0
    h2d = 2 * nac * T * 4 + 7 * model.mesh.nrow * model.mesh.ncol * 4
    d2h = nac * T * 4 + 3 * int(info[1]) * int(info[2]) * 4 + 4
    e2e = {"value": units * e2e_steps / e2e_wall, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": e2e_steps, "ms_per_step": e2e_wall / e2e_steps * 1e3, "first_call_ms_incl_page_locking": first_call_s * 1e3,
           "host_buffers": "caller's NumPy arrays, page-locked in place by the library (option pin_host=1) during the untimed first call",
           "streamed": "256-step windows: upload, kernels and download overlap on three streams"}
    lib.smash_b200_clear_cache()
    lib.smash_b200_set_option(b"pin_host", 0)

    # ---- CPU baseline: the oracle on one host core, same configuration, bounded sample
    import oracle
    Ts = 64
    mc = cases.france(T=Ts, seed=0)
    t0 = time.perf_counter()
    oracle.forward(mc.setup, mc.mesh, mc.input_data, mc.parameters, mc.parameters.copy(), mc.states, mc.states.copy(), mc.output)
    dtc = time.perf_counter() - t0
    cpu = {"value": nac * Ts / dtc, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"France mesh, first {Ts} of {T} steps ({nac * Ts} cell-steps, {dtc:.1f} s), save_qsim_domain, C oracle -O3 "
                     "-march=x86-64-v3, 1 thread (a single forward of the reference is single-threaded)"}

    out = {
        "metric": METRIC_FRANCE, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"France 1km mesh forward gr-a, nac={nac}, T={T}, save_qsim_domain (setup_France.yaml)",
                   "engine": "subtree engine: one pass over the forcing, subtrees of the drainage forest routed inside a warp, results in "
                             "the engine's column order (DESIGN.md section 3e)" if sub_engine else
                             "split: " + ("tick pass (reservoirs + routing in one kernel)" if window_pass
                                          else "reservoir pass per cell + routing scan per heavy-path chain"),
                   "pit_pairs": int(info[6]), "l2": "inputs larger than L2 (5.2 GB forcing streamed once per step)",
                   "parallelism": "1 GPU", "checksum_q": chk.value, "model_build_s": t_build},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": args.steps * launches_per_step, "clocks": clocks,
        "gradient": gradient, "ann_dense": ann, "structures": structures,
    }
    if not args.no_extra:
        ens = bench_ensemble(lib, L, smash_b200, cases, None, 0, 1, max(2, min(5, args.steps)), 1)
        ens.pop("clocks")
        ens["cell_timesteps_per_s"] = ens.pop("value")
        out["ensemble"] = ens
        out["parity"] = parity_numbers(smash_b200, oracle, cases)
        out["extra"] = cance_extras(lib, L, smash_b200, oracle, cases)
    print(json.dumps(out))


def parity_numbers(smash_b200, oracle, cases):
    """Measured differences of the device path to the float32 oracle (and of the float32 oracle to the float64 one, the
    noise floor of the model in the reference's real kind) -- numbers, not pass / fail."""
    from smash_b200.solver._derived_types import ParametersDT, StatesDT
    out = {}

    def rel(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        m = np.abs(b) > 1e-2
        return float((np.abs(a - b)[m] / np.abs(b)[m]).max()) if m.any() else 0.0

    a, b, c = cases.cance(), cases.cance(), cases.cance()
    for m in (a, b, c):
        rng = np.random.default_rng(1)
        act = m.mesh.active_cell == 1
        for name, lo, hi in (("cp", 50, 600), ("cft", 50, 800), ("exc", -5, 5), ("lr", 1, 30)):
            getattr(m.parameters, name)[act] = rng.uniform(lo, hi, int(act.sum())).astype(np.float32)
    pa, sa, pb, sb = ParametersDT(a.mesh), StatesDT(a.mesh), ParametersDT(b.mesh), StatesDT(b.mesh)
    smash_b200.forward_b(a.setup, a.mesh, a.input_data, a.parameters, pa, a.parameters.copy(), None, a.states, sa, a.states.copy(),
                         None, a.output, None)
    oracle.forward_b(b.setup, b.mesh, b.input_data, b.parameters, pb, b.parameters.copy(), b.states, sb, b.states.copy(), b.output)
    oracle.forward(c.setup, c.mesh, c.input_data, c.parameters, c.parameters.copy(), c.states, c.states.copy(), c.output, precision="f64")
    out["cance_T1440_qsim_max_rel_gpu_vs_f32_oracle"] = rel(a.output.qsim, b.output.qsim)
    out["cance_T1440_qsim_max_rel_gpu_vs_f64_oracle"] = rel(a.output.qsim, c.output.qsim)
    out["cance_T1440_qsim_max_rel_f32_oracle_vs_f64_oracle"] = rel(b.output.qsim, c.output.qsim)
    out["cance_T1440_cost_abs_diff"] = abs(float(a.output.cost) - float(b.output.cost))
    for n in ("cp", "cft", "exc", "lr"):
        x, y = np.asarray(getattr(pa, n), np.float64), np.asarray(getattr(pb, n), np.float64)
        out[f"cance_gradient_{n}_max_err_over_inf_norm"] = float(np.abs(x - y).max() / np.abs(y).max())
        out[f"cance_gradient_{n}_cosine"] = float((x * y).sum() / np.sqrt((x * x).sum() * (y * y).sum()))
    return out


def cance_extras(lib, L, smash_b200, oracle, cases):
    """Cance (383 cells, T = 1440): single gradient latency and the variational calibration loop (configs[1])."""
    from smash_b200.solver._derived_types import ParametersDT, StatesDT
    out = {}
    m = cases.cance()
    m.input_data._forcing_version = 1
    pb, sb = ParametersDT(m.mesh), StatesDT(m.mesh)

    def grad():
        smash_b200.forward_b(m.setup, m.mesh, m.input_data, m.parameters, pb, m.parameters.copy(), None, m.states, sb,
                             m.states.copy(), None, m.output, None)

    for _ in range(3):
        grad()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        grad()
    dt = (time.perf_counter() - t0) / n
    mo = cases.cance()
    po, so = ParametersDT(mo.mesh), StatesDT(mo.mesh)
    t0 = time.perf_counter()
    oracle.forward_b(mo.setup, mo.mesh, mo.input_data, mo.parameters, po, mo.parameters.copy(), mo.states, so, mo.states.copy(), mo.output)
    dto = time.perf_counter() - t0
    out["cance_gradient"] = {"evals_per_s": 1.0 / dt, "ms_per_eval_e2e": dt * 1e3, "cpu_port_ms_per_eval": dto * 1e3, "cpu_cores": 1}
    # configs[1]: distributed-mapping variational calibration, L-BFGS-B driven by the adjoint gradient (mw_optimize.f90:484-676)
    from smash_b200 import simulation
    import oracle_solver                                  # the CPU port behind the solver signatures (checker / baseline only)
    iters = 20
    simulation.optimize(cases.cance(), mapping="distributed", options={"maxiter": 2})            # warm-up (plan, forcing)
    t0 = time.perf_counter()
    g = simulation.optimize(cases.cance(), mapping="distributed", options={"maxiter": iters})
    tg = time.perf_counter() - t0
    it_cpu = 3
    t0 = time.perf_counter()
    c = simulation.optimize(cases.cance(), mapping="distributed", options={"maxiter": it_cpu}, solver=oracle_solver)
    tc = time.perf_counter() - t0
    c_same = simulation.optimize(cases.cance(), mapping="distributed", options={"maxiter": it_cpu})
    out["cance_vda_lbfgsb"] = {"iterations": iters, "wall_s": tg, "s_per_iteration": tg / iters, "final_cost": float(g.output.cost),
                               "cpu_port_s_per_iteration": tc / it_cpu, "cpu_port_iterations": it_cpu, "cpu_cores": 1,
                               "cost_after_cpu_port_iterations": {"b200": float(c_same.output.cost), "cpu_port": float(c.output.cost)}}
    return out


if __name__ == "__main__":
    main()
