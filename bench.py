#!/usr/bin/env python
"""bench.py -- forward / adjoint solver throughput on B200 (contract: see the task statement, section 4).

Headline workload (BASELINE.json configs[3], the configuration the HBM-roofline metric is quoted on):
France 1 km flow-direction mesh (906 044 active cells, mesh_France.hdf5 -> tests/golden/france_mesh.npz),
T = 720 synthetic hourly steps (SURVEY.md 8d recipe), gr-a, default parameters, save_qsim_domain as in
setup_France.yaml.  One "step" = one forward run over the whole mesh and all T time steps.

  value   active-cell-timesteps/s, forcing already resident in HBM (plan API), kernel launches only; the result
          (domain discharge in the reference's sparse layout [t][k]) stays in HBM
  e2e     same metric through the drop-in call smash_b200.forward(...) with HOST arrays: per step the forcing
          (5.2 GB) goes host->device and the domain discharge (2.6 GB) comes back
  roofline  the longest kernel of the step (split engine: vertical_forward / route_forward / rows_to_domain), its
          algorithmic bytes / its CUDA-event time; "step" = the whole forward step against the
          12 B per active cell-step of SURVEY.md 8(d) (prcp + pet read, q written)
  cpu_baseline  the C oracle (restatement of the Fortran solver, oracle/) on one host core, bounded sample
  extra   fwd+adjoint gradient on the same mesh, Cance gradient latency, 4096-member Cance ensemble

N > 1 (torchrun): every rank runs the same-size France domain with its own forcing seed (independent regions,
no data-path collective) -> weak scaling; time = max over ranks.
--impl reference: the oracle port timed on the host (the reference's forward is single-threaded for one run).
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

METRIC = "forward active-cell-timesteps/s (France 1km mesh, gr-a); extra: fwd+adjoint gradient evals/s; % HBM roofline"
UNIT = "cell-timesteps/s"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def build_france(T, seed):
    import cases
    t0 = time.time()
    m = cases.france(T=T, seed=seed)
    return m, time.time() - t0


def run_reference(args, rank, world):
    """--impl reference: the oracle port of the Fortran solver on the host, bounded sample per step."""
    if rank != 0:
        return
    import cases
    import oracle
    Ts = args.ref_steps
    m = cases.france(T=Ts, seed=0)
    m.setup.save_qsim_domain = False
    units = m.mesh.nac * Ts

    def step():
        oracle.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(), m.output)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = units / dt
    sample = f"France mesh, {Ts} of {args.T} time steps per step ({units} cell-steps), 1 thread (a single forward run of the reference is single-threaded)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"France 1km mesh forward gr-a, nac=906044, T={args.T} (reference arm: bounded sample)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)   # 200 forward runs ~ 1 s: long enough for clock samples
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--T", type=int, default=720, help="time steps of the France workload")
    ap.add_argument("--ref-steps", type=int, default=8, help="time steps per reference-arm step")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import smash_b200
    from smash_b200 import _lib as L
    lib = L.lib()
    if lib.smash_b200_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: smash_b200 has no CPU fallback")
    L.check(lib.smash_b200_set_device(local_rank))

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    T = args.T
    model, t_build = build_france(T, seed=rank)
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
    barrier()
    kernel_ms, per_kernel = [], []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        L.check(lib.smash_b200_plan_run_forward(plan, C.byref(ms)))   # returns after cudaEventSynchronize
        kernel_ms.append(ms.value)
        L.check(lib.smash_b200_plan_kernel_times(plan, kt))           # events recorded inside the run, read after it
        per_kernel.append([kt[i] for i in range(5)])
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    wall = max_over_ranks(wall)
    kms = max_over_ranks(float(np.mean(kernel_ms)))
    value = world * units * args.steps / wall
    chk = C.c_double(0.0)
    L.check(lib.smash_b200_plan_checksum(plan, C.byref(chk)))

    # ---- roofline (device events on the launching stream, average over the timed launches)
    peak, peak_src = measured_peaks()
    split = int(info[11]) == -1
    step_bytes = 12.0 * units
    roofline_step = {"achieved": step_bytes / (kms * 1e-3) / 1e9, "frac": step_bytes / (kms * 1e-3) / 1e9 / peak,
                     "algorithmic_bytes_per_step": step_bytes, "device_ms": kms}
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass
    if split:
        pk_ms = np.mean(np.array(per_kernel), axis=0)
        nrt, nedge = float(lib.smash_b200_plan_stat(plan, b"routed_cells")), float(lib.smash_b200_plan_stat(plan, b"inflow_edges"))
        names = ["vertical_forward_kernel", "route_forward_kernel", "rows_to_domain_kernel"]
        # algorithmic bytes per launch: reservoir pass 8 B forcing + 4 B series per cell-step;
        # routing 4 B per routed cell-step in and out + 4 B per inflow edge-step; export 8 B per routed cell-step
        kbytes = [12.0 * units, 4.0 * T * (2.0 * nrt + nedge), 8.0 * T * nrt]
        fused_export = pk_ms[2] < 0.02                # option fuse_export: the routing warps write qsim_domain themselves
        ktraffic = [traffic.get(n) for n in names]
        if fused_export:
            kbytes[1] += kbytes[2]                    # (profiles/ncu_traffic.json holds the traffic of the fused kernel)
        kernels = {names[i]: {"ms": float(pk_ms[i]), "share": float(pk_ms[i] / kms), "algorithmic_bytes": kbytes[i],
                              "achieved_gbs": kbytes[i] / (pk_ms[i] * 1e-3) / 1e9, "frac": kbytes[i] / (pk_ms[i] * 1e-3) / 1e9 / peak,
                              "traffic": ktraffic[i]} for i in range(3) if pk_ms[i] > 0 and not (i == 2 and fused_export)}
        if fused_export:
            kernels[names[1]]["includes"] = "export of the routed cells' series to [t][cell] (8 B per routed cell-step)"
        top = max(kernels, key=lambda k: kernels[k]["ms"])
        roofline = {"bound": "hbm", "achieved": kernels[top]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[top]["frac"], "traffic": kernels[top]["traffic"], "peak_source": peak_src, "kernel": top,
                    "kernel_ms": kernels[top]["ms"], "algorithmic_bytes_per_launch": kernels[top]["algorithmic_bytes"],
                    "kernels": kernels, "step": roofline_step}
        lib.smash_b200_plan_info(plan, info)
        launches_per_step = int(info[7])             # kernels launched by the last plan_run_forward (counted by the library)
    else:
        roofline = {"bound": "hbm", "achieved": roofline_step["achieved"], "peak": peak, "unit": "GB/s",
                    "frac": roofline_step["frac"], "traffic": traffic.get("forward_kernel"), "peak_source": peak_src,
                    "kernel": "forward_kernel", "kernel_ms": kms, "algorithmic_bytes_per_launch": step_bytes,
                    "step": roofline_step}
        launches_per_step = 1

    extra = {}
    if not args.no_extra:
        # fwd + adjoint gradient on the same mesh (store-all tape in HBM): 48 B per cell-step moved by design
        f_ms, r_ms = C.c_float(0), C.c_float(0)
        L.check(lib.smash_b200_plan_run_gradient(plan, C.byref(f_ms), C.byref(r_ms)))
        gms = []
        for _ in range(max(2, args.steps // 3)):
            L.check(lib.smash_b200_plan_run_gradient(plan, C.byref(f_ms), C.byref(r_ms)))
            gms.append((f_ms.value, r_ms.value))
        gf, gr = float(np.mean([g[0] for g in gms])), float(np.mean([g[1] for g in gms]))
        extra["france_gradient"] = {"evals_per_s": 1e3 / (gf + gr), "cell_timesteps_per_s": units / ((gf + gr) * 1e-3),
                                    "forward_tape_ms": gf, "reverse_ms": gr,
                                    "hbm_frac_vs_40B": 40.0 * units / ((gf + gr) * 1e-3) / 1e9 / peak}
    lib.smash_b200_plan_destroy(plan)

    # ---- e2e through the drop-in call with host buffers
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    model.input_data._forcing_version = 0
    # the forcing / output arrays live until clear_cache() below: let the library page-lock them in place (opt-in option)
    lib.smash_b200_set_option(b"pin_host", 1)

    par_bgd, sta_bgd = model.parameters.copy(), model.states.copy()     # the background never changes between calls

    def e2e_step():
        smash_b200.forward(model.setup, model.mesh, model.input_data, model.parameters, par_bgd, model.states, sta_bgd,
                           model.output)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_wall = max_over_ranks(time.perf_counter() - t0)
    h2d = 2 * nac * T * 4 + 7 * model.mesh.nrow * model.mesh.ncol * 4
    d2h = nac * T * 4 + 3 * int(info[1]) * int(info[2]) * 4 + 4
    e2e = {"value": world * units * e2e_steps / e2e_wall, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": e2e_steps, "ms_per_step": e2e_wall / e2e_steps * 1e3,
           "host_buffers": "caller's NumPy arrays, page-locked in place by the library (option pin_host=1) during the untimed first call",
           "streamed": "256-step windows: upload, kernels and download overlap on three streams"}
    lib.smash_b200_clear_cache()
    lib.smash_b200_set_option(b"pin_host", 0)

    cpu = None
    if rank == 0 and world == 1:
        import cases
        import oracle
        Ts = 64                                  # ~10 s of single-thread CPU work
        mc = cases.france(T=Ts, seed=0)
        mc.setup.save_qsim_domain = False
        t0 = time.perf_counter()
        oracle.forward(mc.setup, mc.mesh, mc.input_data, mc.parameters, mc.parameters.copy(), mc.states, mc.states.copy(), mc.output)
        dtc = time.perf_counter() - t0
        cpu = {"value": nac * Ts / dtc, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"France mesh, first {Ts} of {T} steps ({nac * Ts} cell-steps, {dtc:.1f} s), C oracle -O3, 1 thread "
                         "(a single forward of the reference is single-threaded)"}
        if not args.no_extra:
            extra.update(cance_extras(lib, L, smash_b200, oracle, cases))
    if world > 1 and not args.no_extra:
        # configs[2]: the 4096-member Cance ensemble split over the ranks (contiguous member blocks, no data-path collective;
        # one all-gather of the costs afterwards) -- total work fixed, i.e. strong scaling of this extra
        import cases
        from smash_b200 import distributed as sdist
        mc = cases.cance()
        mc.input_data._forcing_version = 1
        ns = 4096
        rng = np.random.RandomState(99)
        bounds = [(1e-6, 1e3), (1e-6, 1e3), (-50.0, 50.0), (1e-6, 1e3)]
        smp = np.asfortranarray(np.stack([rng.uniform(lo, hi, ns) for lo, hi in bounds]).astype(np.float32))
        cost, q0 = np.zeros(ns, np.float32), np.zeros((0,), np.float32)

        def ens():
            sdist.multiple_run_sharded(mc.setup, mc.mesh, mc.input_data, mc.parameters, mc.states, mc.output, smp,
                                       cases.IND_CP_CFT_EXC_LR, cost, q0)

        ens()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            ens()
        barrier()
        dte = max_over_ranks(time.perf_counter() - t0) / 3
        # configs[4]: regionalised calibration, one Cance-sized catchment per rank, shared hyper-polynomial mapping,
        # one all-reduce of (cost, gradient) per evaluation
        def catchment():
            mk = cases.cance()
            mk.input_data.qobs = np.asfortranarray(mk.input_data.qobs * np.float32(1.0 + 0.05 * rank))
            mk.input_data._forcing_version = 100 + rank
            cases.set_optimize(mk.setup, mk.mesh, jobs_fun=("nse",), mapping="hyper-polynomial")
            mk.setup._optimize.optim_parameters[[1, 3, 6, 15]] = 1
            mk.setup._optimize.maxiter = 10
            mk.setup._optimize.verbose = False
            return mk
        mk = catchment()
        sdist.optimize_hyper_lbfgsb_sharded([(mk.setup, mk.mesh, mk.input_data, mk.parameters, mk.states, mk.output)])
        mk = catchment()
        barrier()
        t0 = time.perf_counter()
        sdist.optimize_hyper_lbfgsb_sharded([(mk.setup, mk.mesh, mk.input_data, mk.parameters, mk.states, mk.output)])
        barrier()
        dtr = max_over_ranks(time.perf_counter() - t0)
        extra["regional_hyper_polynomial_lbfgsb"] = {"catchments": world, "iterations": 10, "wall_s": dtr,
                                                     "s_per_iteration": dtr / 10, "collective": "all-reduce of 21 float64 per evaluation",
                                                     "rank0_cost": float(mk.output.cost)}
        # ONE France run with its 3 434 drainage basins spread over the ranks (whole basins per rank, no exchange):
        # strong scaling of a single domain, bounded by the largest basin (the Loire, 15 % of the cells)
        masks, load = sdist.basin_masks(model.mesh, world, model.setup)
        keep_mask = model.mesh._local_active_cell
        model.mesh._local_active_cell = masks[rank]
        if hasattr(model.mesh, "_b200_cache"):
            del model.mesh._b200_cache
        pkb = L.Packed()
        sb_, mb_, ib_ = L.pack_setup(model.setup, model.mesh, pkb), L.pack_mesh(model.mesh, model.setup, pkb), L.pack_input(model.input_data, model.setup, model.mesh, pkb)
        planb = C.c_void_p()
        L.check(lib.smash_b200_plan_create(C.byref(sb_), C.byref(mb_), 1, C.byref(planb)))
        L.check(lib.smash_b200_plan_set_forcing(planb, C.byref(sb_), C.byref(ib_)))
        L.check(lib.smash_b200_plan_set_fields(planb, C.byref(p_), C.byref(st_), None, None, 0))
        msb = C.c_float(0.0)
        for _ in range(3):
            L.check(lib.smash_b200_plan_run_forward(planb, C.byref(msb)))
        barrier()
        tb0 = time.perf_counter()
        nrep = 20
        for _ in range(nrep):
            L.check(lib.smash_b200_plan_run_forward(planb, C.byref(msb)))
        barrier()
        dtb = max_over_ranks(time.perf_counter() - tb0) / nrep
        lib.smash_b200_plan_destroy(planb)
        model.mesh._local_active_cell = keep_mask
        if hasattr(model.mesh, "_b200_cache"):
            del model.mesh._b200_cache
        extra["france_one_domain_split_by_basin"] = {"cell_timesteps_per_s": units / dtb, "ms_per_run": dtb * 1e3, "scaling": "strong",
                                                     "cells_per_rank": [int(x) for x in load], "collective": "none"}
        extra["cance_ensemble_4096_sharded"] = {"cell_timesteps_per_s_e2e": ns * 383 * 1440 / dte, "ms_per_call": dte * 1e3,
                                                "members_per_rank": ns // world, "scaling": "strong",
                                                "cost_checksum": float(np.sum(cost[np.isfinite(cost)], dtype=np.float64))}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"France 1km mesh forward gr-a, nac={nac}, T={T}, save_qsim_domain (setup_France.yaml)",
                       "engine": "split (reservoir pass per cell + routing scan per heavy-path chain)" if split else "fused tick wavefront",
                       "ctas": int(info[1]), "cta_size": int(info[2]),
                       **({"chains": int(info[5]), "routing_tasks": int(info[10]), "chain_dependency_height": int(info[3]),
                           "longest_dependency_path_cells": int(info[8])} if split else
                          {"max_skew": int(info[3]), "cross_block_edges": int(info[5])}),
                       "pit_pairs": int(info[6]),
                       "l2": "inputs larger than L2 (5.2 GB forcing streamed once per step)",
                       "parallelism": f"{world} independent domain replica(s), no collective", "checksum_q": chk.value,
                       "model_build_s": t_build},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": args.steps * launches_per_step, "clocks": clocks, "extra": extra,
        }))
    if dist is not None:
        dist.destroy_process_group()


def cance_extras(lib, L, smash_b200, oracle, cases):
    """Cance (383 cells, T=1440): single gradient latency and the 4096-member ensemble (configs[1], configs[2])."""
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
    # ensemble
    ns = 4096
    rng = np.random.RandomState(99)
    bounds = [(1e-6, 1e3), (1e-6, 1e3), (-50.0, 50.0), (1e-6, 1e3)]
    smp = np.asfortranarray(np.stack([rng.uniform(lo, hi, ns) for lo, hi in bounds]).astype(np.float32))
    cost = np.zeros(ns, np.float32)
    q0 = np.zeros((0,), np.float32)

    def ens():
        smash_b200.compute_multiple_run(m.setup, m.mesh, m.input_data, m.parameters, m.states, m.output, smp,
                                        cases.IND_CP_CFT_EXC_LR, cost, q0)

    ens()
    t0 = time.perf_counter()
    for _ in range(3):
        ens()
    dte = (time.perf_counter() - t0) / 3
    nthreads = os.cpu_count() or 1
    nso = 64
    co = np.zeros(nso, np.float32)
    t0 = time.perf_counter()
    oracle.compute_multiple_run(mo.setup, mo.mesh, mo.input_data, mo.parameters, mo.states, mo.output, smp[:, :nso],
                                cases.IND_CP_CFT_EXC_LR, co, q0, nthreads=nthreads)
    dtoe = time.perf_counter() - t0
    cs = 383 * 1440
    out["cance_ensemble_4096"] = {"cell_timesteps_per_s_e2e": ns * cs / dte, "ms_per_call": dte * 1e3,
                                  "cpu_port_cell_timesteps_per_s": nso * cs / dtoe, "cpu_cores": nthreads,
                                  "cpu_sample": f"{nso} members, OpenMP over members"}
    # configs[1]: distributed-mapping variational calibration, L-BFGS-B driven by the adjoint gradient (mw_optimize.f90:484-676)
    from smash_b200 import simulation
    sys.path.insert(0, os.path.join(ROOT, "tests"))
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
