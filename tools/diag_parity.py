"""Diagnostic (not a test): error of the CUDA path against the f32 and f64 oracles on Cance."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases, oracle, smash_b200
from smash_b200 import _lib as L

def run(fn, **kw):
    m = cases.cance()
    fn(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(), m.output, **kw)
    return np.asarray(m.output.qsim, np.float64), float(m.output.cost)

q32, c32 = run(oracle.forward, precision="f32")
q64, c64 = run(oracle.forward, precision="f64")
def report(name, q, c):
    for ref, qr, cr in (("f32", q32, c32), ("f64", q64, c64)):
        d = np.abs(q - qr)
        big = np.abs(qr) > 1e-2
        print(f"{name:12s} vs {ref}: max abs {d.max():.3e}  max rel(q>1e-2) {(d[big]/np.abs(qr[big])).max():.3e}  "
              f"viol(1e-6+1e-4rel) {(d > 1e-6 + 1e-4*np.abs(qr)).sum():5d}  viol(ref allclose 1e-4,1e-5) {(d > 1e-4 + 1e-5*np.abs(qr)).sum():4d}  dcost {c-cr:+.2e}")
report("oracle f32", q32, c32)
for mode in (0, 1):
    L.lib().smash_b200_set_option(b"math", mode)
    q, c = run(smash_b200.forward)
    report(f"gpu math={mode}", q, c)

# France window (300 x 300, 96 steps): domain discharge of the default engine against the f32 oracle
a, b = cases.france(T=96, sub=(400, 700, 400, 700)), cases.france(T=96, sub=(400, 700, 400, 700))
L.lib().smash_b200_set_option(b"math", 1)
smash_b200.forward(a.setup, a.mesh, a.input_data, a.parameters, a.parameters.copy(), a.states, a.states.copy(), a.output)
oracle.forward(b.setup, b.mesh, b.input_data, b.parameters, b.parameters.copy(), b.states, b.states.copy(), b.output)
qa, qb = np.asarray(a.output.sparse_qsim_domain, np.float64), np.asarray(b.output.sparse_qsim_domain, np.float64)
d = np.abs(qa - qb)
big = np.abs(qb) > 1e-3
print(f"france window: max abs {d.max():.3e}  max rel(q>1e-3) {(d[big]/np.abs(qb[big])).max():.3e}  "
      f"median rel {np.median(d[big]/np.abs(qb[big])):.3e}  viol(1e-6+1e-4rel) {(d > 1e-6 + 1e-4*np.abs(qb)).sum()} of {d.size}")
