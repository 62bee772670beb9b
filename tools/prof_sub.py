import ctypes as C, os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, cases
from smash_b200 import _lib as L
lib = L.lib()
m = cases.france(T=720)
pk = L.Packed()
s_, m_, i_ = L.pack_setup(m.setup, m.mesh, pk), L.pack_mesh(m.mesh, m.setup, pk), L.pack_input(m.input_data, m.setup, m.mesh, pk)
p_, st_ = L.pack_parameters(m.parameters, pk), L.pack_states(m.states, pk)
plan = C.c_void_p()
L.check(lib.smash_b200_plan_create(C.byref(s_), C.byref(m_), 1, C.byref(plan)))
L.check(lib.smash_b200_plan_set_forcing(plan, C.byref(s_), C.byref(i_)))
L.check(lib.smash_b200_plan_set_fields(plan, C.byref(p_), C.byref(st_), None, None, 0))
ms = C.c_float(0)
L.check(lib.smash_b200_plan_run_forward(plan, C.byref(ms))); print("sub", ms.value, "ms")
