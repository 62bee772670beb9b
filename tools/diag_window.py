"""Window pass vs row passes on a France crop: which cells differ?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases, smash_b200
from smash_b200 import _lib as L
from test_gpu_parity import random_fields
lib = L.lib()
A = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 40
def run(win):
    lib.smash_b200_set_option(b"window_pass", win); lib.smash_b200_set_option(b"window_min_cells", 1000)
    lib.smash_b200_set_option(b"shallow_acc", A); lib.smash_b200_clear_cache()
    m = cases.france(T=T, sub=(250, 900, 250, 900), ngauge=0)
    random_fields(m, seed=11)
    smash_b200.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(), m.output)
    return m
a, b = run(1), run(0)
qa, qb = np.asarray(a.output.sparse_qsim_domain, np.float64), np.asarray(b.output.sparse_qsim_domain, np.float64)
bad = np.abs(qa - qb) > 1e-7 + 1e-5 * np.abs(qb)
cells = np.where(bad.any(axis=1))[0]
print("cells", qa.shape[0], "bad cells", len(cells), "first bad steps", np.where(bad.any(axis=0))[0][:10])
# flwacc per sparse index
k = a.mesh._rowcol_to_ind_sparse
fa = np.zeros(a.mesh.nac, np.int64); act = a.mesh.active_cell == 1
fa[k[act] - 1] = a.mesh.flwacc[act]
print("flwacc of bad cells: min", fa[cells].min() if len(cells) else None, "hist", np.bincount(np.minimum(fa[cells], 40))[:41] if len(cells) else None)
print("bad cell idx sample", cells[:20], "mod 32", (cells[:20] % 32))
for c in cells[:5]:
    t = np.where(bad[c])[0][:6]
    print(c, fa[c], t, qa[c, t], qb[c, t])
