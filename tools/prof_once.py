"""One run of a configuration for ncu captures (no warm-up: ncu replays every kernel).
usage: python tools/prof_once.py france_fwd | france_grad | ensemble | hyper | struct:<gr-b|gr-c|gr-d|vic-a>"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
from smash_b200 import _lib as L
what = sys.argv[1]
structure = "gr-a"
if what.startswith("struct:"):
    structure, what = what.split(":")[1], "france_fwd"
lib = L.lib()
if what == "ensemble":
    m = cases.cance(sparse=True)
    members = 592
else:
    m = cases.france(T=720, ngauge=4 if what in ("france_grad", "hyper") else 0, nd=6 if what == "hyper" else 0, qobs_from_oracle=False)
    members = 1
m.setup.structure = structure
if structure in ("gr-b", "gr-c"):
    m.parameters.ci[...] = 2.0
if what == "hyper":
    from test_gpu_parity2 import _hyper_objects
    cases.set_optimize(m.setup, m.mesh, jobs_fun=("nse",), mapping="hyper-polynomial", gauge="all")
    hp, hs = _hyper_objects(m, "hyper-polynomial")
pk = L.Packed()
s_, m_, i_ = L.pack_setup(m.setup, m.mesh, pk), L.pack_mesh(m.mesh, m.setup, pk), L.pack_input(m.input_data, m.setup, m.mesh, pk)
p_, st_ = L.pack_parameters(m.parameters, pk), L.pack_states(m.states, pk)
plan = C.c_void_p()
L.check(lib.smash_b200_plan_create(C.byref(s_), C.byref(m_), members, C.byref(plan)))
L.check(lib.smash_b200_plan_set_forcing(plan, C.byref(s_), C.byref(i_)))
smp = ind = None
if members > 1:
    rng = np.random.RandomState(99)
    smp = np.asfortranarray(np.stack([rng.uniform(lo, hi, members) for lo, hi in [(1e-6, 1e3), (1e-6, 1e3), (-50, 50), (1e-6, 1e3)]]).astype(np.float32))
    ind = cases.IND_CP_CFT_EXC_LR
L.check(lib.smash_b200_plan_set_fields(plan, C.byref(p_), C.byref(st_), L._fp(smp) if smp is not None else None,
                                       L._ip(ind) if ind is not None else None, 4 if smp is not None else 0))
ms, f, r = C.c_float(0), C.c_float(0), C.c_float(0)
if what in ("france_fwd", "ensemble"):
    L.check(lib.smash_b200_plan_run_forward(plan, C.byref(ms))); print(what, ms.value, "ms")
elif what == "france_grad":
    L.check(lib.smash_b200_plan_run_gradient(plan, C.byref(f), C.byref(r))); print(what, f.value, r.value, "ms")
else:
    hp_, hs_ = L.pack_parameters(hp, pk), L.pack_states(hs, pk)
    hb = np.zeros((7, m.setup._optimize.nhyper), np.float32); t = (C.c_float * 4)()
    L.check(lib.smash_b200_plan_run_hyper_gradient(plan, C.byref(s_), C.byref(i_), C.byref(hp_), C.byref(hs_), L._fp(hb), t)); print(what, list(t))
lib.smash_b200_plan_destroy(plan)
