"""Tick-pass experiments on the France mesh through the plan API: one model, several option sets.
usage: python tools/wbench.py [--T N] [--reps R] set1 set2 ...   with set = name=value[,name=value...] ('base' = defaults)"""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
from smash_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=720); ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--structure", default="gr-a")
ap.add_argument("sets", nargs="*", default=["base"])
a = ap.parse_args()
lib = L.lib()
m = cases.france(T=a.T)
m.setup.structure = a.structure
if a.structure in ("gr-b", "gr-c"):
    m.parameters.ci[...] = 2.0
DEFAULTS = {"tick_pass": 0, "sub_engine": -1, "sub_scatter": 0, "tick_nb": 2, "tick_slack": 1, "tick_dbg": 0, "shallow_acc": 32, "tick_variant": 8, "tick_ctas_per_sm": 0, "fuse_export": 4}
for st in a.sets:
    opts = dict(DEFAULTS)
    if st != "base":
        for kv in st.split(","):
            k, v = kv.split("="); opts[k] = int(v)
    for k, v in opts.items():
        lib.smash_b200_set_option(k.encode(), v)
    pk = L.Packed()
    s_, m_, i_ = L.pack_setup(m.setup, m.mesh, pk), L.pack_mesh(m.mesh, m.setup, pk), L.pack_input(m.input_data, m.setup, m.mesh, pk)
    p_, st_ = L.pack_parameters(m.parameters, pk), L.pack_states(m.states, pk)
    plan = C.c_void_p()
    L.check(lib.smash_b200_plan_create(C.byref(s_), C.byref(m_), 1, C.byref(plan)))
    L.check(lib.smash_b200_plan_set_forcing(plan, C.byref(s_), C.byref(i_)))
    L.check(lib.smash_b200_plan_set_fields(plan, C.byref(p_), C.byref(st_), None, None, 0))
    ms = C.c_float(0)
    fw, kts = [], []
    for i in range(a.reps + 2):
        L.check(lib.smash_b200_plan_run_forward(plan, C.byref(ms))); fw.append(ms.value)
        kt = (C.c_float * 5)(); lib.smash_b200_plan_kernel_times(plan, kt); kts.append([kt[j] for j in range(3)])
    fw = np.array(fw[2:]); kts = np.array(kts[2:])
    chk = C.c_double(0)
    L.check(lib.smash_b200_plan_checksum(plan, C.byref(chk)))
    units = m.mesh.nac * a.T
    print(f"{st:50s} fwd {fw.mean():.3f} ms (min {fw.min():.3f}) [pass1 {kts[:,0].mean():.3f} route {kts[:,1].mean():.3f} export {kts[:,2].mean():.3f}]"
          f" {units/fw.mean()*1e3:.3e} cs/s frac12B {units*12/fw.mean()*1e3/6534.8e9:.3f} checksum {chk.value:.9e}", flush=True)
    lib.smash_b200_plan_destroy(plan)
