"""France 1 km, long gradient runs: checkpointed adjoint (256-step windows) against the store-all tape.
usage: python tools/ckpt_france.py [--T 2160] [--reps 2] [--no-store-all]   -> one JSON line"""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
from smash_b200 import _lib as L
from smash_b200.solver._derived_types import ParametersDT, StatesDT

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=2160); ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--no-store-all", action="store_true")
a = ap.parse_args()
lib = L.lib()
t0 = time.time()
m = cases.france(T=a.T, ngauge=4, qobs_from_oracle=False)
rng = np.random.default_rng(1)
act = m.mesh.active_cell == 1
for name, lo, hi in (("cp", 50, 600), ("cft", 50, 800), ("exc", -5, 5), ("lr", 1, 30)):
    getattr(m.parameters, name)[act] = rng.uniform(lo, hi, int(act.sum())).astype(np.float32)
build_s = time.time() - t0
out = {"workload": "France 1 km gradient", "T": a.T, "nac": int(m.mesh.nac), "host_build_s": round(build_s, 1)}
grads = {}
for mode in ([1] if a.no_store_all else [1, 0]):
    lib.smash_b200_set_option(b"adjoint_checkpoint", mode)
    lib.smash_b200_set_option(b"plan_small_windows", 1)     # store-all run on the same 256-step routing windows
    pk = L.Packed()
    s_, m_, i_ = L.pack_setup(m.setup, m.mesh, pk), L.pack_mesh(m.mesh, m.setup, pk), L.pack_input(m.input_data, m.setup, m.mesh, pk)
    p_, st_ = L.pack_parameters(m.parameters, pk), L.pack_states(m.states, pk)
    plan = C.c_void_p()
    L.check(lib.smash_b200_plan_create(C.byref(s_), C.byref(m_), 1, C.byref(plan)))
    L.check(lib.smash_b200_plan_set_forcing(plan, C.byref(s_), C.byref(i_)))
    L.check(lib.smash_b200_plan_set_fields(plan, C.byref(p_), C.byref(st_), None, None, 0))
    f, r = C.c_float(0), C.c_float(0)
    ts = []
    for _ in range(a.reps + 1):
        L.check(lib.smash_b200_plan_run_gradient(plan, C.byref(f), C.byref(r))); ts.append((f.value, r.value))
    ts = np.array(ts[1:])
    gp, gs = ParametersDT(m.mesh), StatesDT(m.mesh)
    for n in ("cp", "cft", "exc", "lr"): getattr(gp, n)[...] = 0
    for n in ("hp", "hft", "hlr"): getattr(gs, n)[...] = 0
    gp_, gs_ = L.pack_parameters(gp, pk), L.pack_states(gs, pk)
    L.check(lib.smash_b200_plan_get_gradient(plan, C.byref(gp_), C.byref(gs_)))
    grads[mode] = {n: np.array(getattr(gp, n)) for n in ("cp", "cft", "exc", "lr")} | {n: np.array(getattr(gs, n)) for n in ("hp", "hft", "hlr")}
    units = m.mesh.nac * a.T
    key = "checkpointed" if mode else "store_all"
    out[key] = {"forward_ms": round(float(ts[:, 0].mean()), 3), "reverse_ms": round(float(ts[:, 1].mean()), 3),
                "gradients_per_s": round(1e3 / float(ts.sum(1).mean()), 3),
                "cell_steps_per_s": units / float(ts.sum(1).mean()) * 1e3,
                "frac_of_40B_roofline": units * 40 / float(ts.sum(1).mean()) * 1e3 / 6534.8e9,
                "tape_gb": round(lib.smash_b200_plan_stat(plan, b"tape_bytes") / 1e9, 3),
                "windows": int(lib.smash_b200_plan_stat(plan, b"route_windows")),
                "forward_sweeps": 2 if mode else 1}
    lib.smash_b200_plan_destroy(plan)
if 0 in grads:
    out["max_rel_diff_vs_store_all"] = {n: float(np.abs(grads[1][n].astype(np.float64) - grads[0][n]).max() / max(np.abs(grads[0][n]).max(), 1e-300))
                                        for n in grads[0]}
print(json.dumps(out), flush=True)
