"""Diagnostic: GPU gradient vs oracle f32 and f64 on a France window (three-way errors per field)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases, oracle, smash_b200
from smash_b200 import _lib as L
from smash_b200.solver._derived_types import ParametersDT, StatesDT
from test_gpu_parity import random_fields

which = sys.argv[1] if len(sys.argv) > 1 else "window"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 96
def make():
    if which == "window":
        m = cases.france(T=T, sub=(400, 700, 400, 700), ngauge=4)
    elif which == "cance":
        m = cases.cance(T=T)
    else:
        m = cases.france(T=T, ngauge=3)
    m.setup.save_qsim_domain = False
    random_fields(m, seed=2)
    return m
res = {}
for tag in ("f32", "f64", "gpu0", "gpu1"):
    m = make()
    pb, sb = ParametersDT(m.mesh), StatesDT(m.mesh)
    if tag.startswith("gpu"):
        L.lib().smash_b200_set_option(b"math", int(tag[3]))
        smash_b200.forward_b(m.setup, m.mesh, m.input_data, m.parameters, pb, m.parameters.copy(), None, m.states, sb, m.states.copy(), None, m.output, None)
    else:
        oracle.forward_b(m.setup, m.mesh, m.input_data, m.parameters, pb, m.parameters.copy(), m.states, sb, m.states.copy(), m.output, precision=tag)
    res[tag] = (float(m.output.cost), {n: np.asarray(getattr(pb if n in ("cp","cft","exc","lr") else sb, n), np.float64) for n in ("cp","cft","exc","lr","hp","hft","hlr")})
    print(tag, "cost", res[tag][0], flush=True)
def cmp(a, b):
    out = []
    for n in ("cp","cft","exc","lr","hp","hft","hlr"):
        x, y = res[a][1][n], res[b][1][n]
        sc = np.abs(y).max()
        cos = (x*y).sum()/np.sqrt((x*x).sum()*(y*y).sum()) if sc > 0 else 1.0
        out.append(f"{n}: relinf {np.abs(x-y).max()/max(sc,1e-300):.2e} cos {cos:.6f} scale {sc:.1e}")
    print(f"{a} vs {b}: " + " | ".join(out))
for a, b in (("f32","f64"),("gpu0","f64"),("gpu1","f64"),("gpu0","f32"),("gpu1","f32")):
    cmp(a, b)
