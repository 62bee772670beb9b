"""France-scale hyper gradient: per-cell gradient planes (GPU forward_b vs oracle forward_b) on the mapped fields."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases, oracle, smash_b200
from smash_b200.solver._derived_types import Hyper_ParametersDT, Hyper_StatesDT, ParametersDT, StatesDT
from test_gpu_parity2 import _hyper_objects

m = cases.france(T=24, ngauge=4, nd=6)
cases.set_optimize(m.setup, m.mesh, jobs_fun=("nse",), mapping="hyper-polynomial", gauge="all")
hp, hs = _hyper_objects(m, "hyper-polynomial")
a, b = m.copy(), m.copy()
smash_b200.hyper_forward(a.setup, a.mesh, a.input_data, a.parameters, hp, hp.copy(), a.states, hs, hs.copy(), a.output)
oracle.hyper_forward(b.setup, b.mesh, b.input_data, b.parameters, hp, b.states, hs, b.output)
print("cost", float(a.output.cost), float(b.output.cost))
for n in ("cp", "cft", "exc", "lr"):
    x, y = getattr(a.parameters, n), getattr(b.parameters, n)
    print("mapped", n, float(np.abs(x - y).max()), float(y.min()), float(y.max()))
# per-cell gradients on the oracle's mapped fields, initial states as mapped
a2, b2 = m.copy(), m.copy()
for n in ("cp", "cft", "exc", "lr"):
    setattr(a2.parameters, n, getattr(b.parameters, n).copy()); setattr(b2.parameters, n, getattr(b.parameters, n).copy())
cases.set_optimize(a2.setup, a2.mesh, jobs_fun=("nse",), gauge="all"); cases.set_optimize(b2.setup, b2.mesh, jobs_fun=("nse",), gauge="all")
pa, sa, pb, sb = ParametersDT(m.mesh), StatesDT(m.mesh), ParametersDT(m.mesh), StatesDT(m.mesh)
smash_b200.forward_b(a2.setup, a2.mesh, a2.input_data, a2.parameters, pa, a2.parameters.copy(), None, a2.states, sa, a2.states.copy(), None, a2.output, None)
oracle.forward_b(b2.setup, b2.mesh, b2.input_data, b2.parameters, pb, b2.parameters.copy(), b2.states, sb, b2.states.copy(), b2.output)
print("cost2", float(a2.output.cost), float(b2.output.cost))
for n in ("cp", "cft", "exc", "lr"):
    x, y = np.asarray(getattr(pa, n), np.float64), np.asarray(getattr(pb, n), np.float64)
    k = np.unravel_index(np.abs(x - y).argmax(), x.shape)
    print(n, "max|d|", np.abs(x - y).max(), "scale", np.abs(y).max(), "sum", x.sum(), y.sum(), "at", k, x[k], y[k], "nonzero", (x != 0).sum(), (y != 0).sum())
