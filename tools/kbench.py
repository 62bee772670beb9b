"""Kernel micro-benchmark through the plan API (device-resident inputs, CUDA-event times).
usage: python tools/kbench.py <cance|france> [--T N] [--block B] [--math M] [--members N] [--grad] [--reps R]"""
import argparse, ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
from smash_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("mesh"); ap.add_argument("--T", type=int, default=None); ap.add_argument("--block", type=int, default=0)
ap.add_argument("--math", type=int, default=1); ap.add_argument("--members", type=int, default=1)
ap.add_argument("--grad", action="store_true"); ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--opt", action="append", default=[], help="name=value library option")
a = ap.parse_args()
lib = L.lib()
lib.smash_b200_set_option(b"block", a.block); lib.smash_b200_set_option(b"math", a.math)
for kv in a.opt:
    k, v = kv.split("="); lib.smash_b200_set_option(k.encode(), int(v))
m = cases.cance(sparse=True, T=a.T) if a.mesh == "cance" else cases.france(T=a.T or 720)
pk = L.Packed()
s_, m_, i_ = L.pack_setup(m.setup, m.mesh, pk), L.pack_mesh(m.mesh, m.setup, pk), L.pack_input(m.input_data, m.setup, m.mesh, pk)
p_, st_ = L.pack_parameters(m.parameters, pk), L.pack_states(m.states, pk)
plan = C.c_void_p()
L.check(lib.smash_b200_plan_create(C.byref(s_), C.byref(m_), a.members, C.byref(plan)))
L.check(lib.smash_b200_plan_set_forcing(plan, C.byref(s_), C.byref(i_)))
smp = ind = None; nvar = 0
if a.members > 1:
    rng = np.random.RandomState(99)
    smp = np.asfortranarray(np.stack([rng.uniform(lo, hi, a.members) for lo, hi in [(1e-6, 1e3), (1e-6, 1e3), (-50, 50), (1e-6, 1e3)]]).astype(np.float32))
    ind = cases.IND_CP_CFT_EXC_LR; nvar = 4
L.check(lib.smash_b200_plan_set_fields(plan, C.byref(p_), C.byref(st_), L._fp(smp) if smp is not None else None, L._ip(ind) if ind is not None else None, nvar))
info = (C.c_int64 * 12)(); lib.smash_b200_plan_info(plan, info)
units = int(info[0]) * m.setup._ntime_step * a.members
ms = C.c_float(0); f = C.c_float(0); r = C.c_float(0)
fw = []
for i in range(a.reps + 2):
    L.check(lib.smash_b200_plan_run_forward(plan, C.byref(ms))); fw.append(ms.value)
fw = np.array(fw[2:])
kt = (C.c_float * 5)(); lib.smash_b200_plan_kernel_times(plan, kt)
kf = [kt[i] for i in range(5)]
eng = "split" if int(info[11]) == -1 else "fused"
line = f"{a.mesh} T={m.setup._ntime_step} engine={eng} B={int(info[2])} blocks={int(info[1])} members={a.members} math={a.math} crit={int(info[8])} | fwd {fw.mean():.3f} ms (min {fw.min():.3f}) {units/fw.mean()/1e-3:.3e} cs/s"
if kf[0] >= 0: line += f" [vert {kf[0]:.3f} route {kf[1]:.3f} export {kf[2]:.3f}]"
if a.grad:
    g = []
    for i in range(a.reps + 1):
        L.check(lib.smash_b200_plan_run_gradient(plan, C.byref(f), C.byref(r))); g.append((f.value, r.value))
    g = np.array(g[1:])
    lib.smash_b200_plan_kernel_times(plan, kt)
    if kt[0] >= 0: line += f" | [vert {kt[0]:.3f} route {kt[1]:.3f} route_b {kt[3]:.3f} vert_b {kt[4]:.3f}]"
    line += f" | grad fwd {g[:,0].mean():.3f} rev {g[:,1].mean():.3f} ms {units/g.sum(1).mean()/1e-3:.3e} cs/s"
print(line, flush=True)
lib.smash_b200_plan_destroy(plan)
