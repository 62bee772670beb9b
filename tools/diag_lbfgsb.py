"""Cance distributed-mapping L-BFGS-B: cost after k iterations for the B200 (math = 1 and 0) and the CPU oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases, oracle_solver
from smash_b200 import simulation, _lib as L
lib = L.lib()
rows = {}
for name, solver, math in (("b200 math=1", None, 1), ("b200 math=0", None, 0), ("cpu oracle f32", oracle_solver, None)):
    if math is not None:
        lib.smash_b200_set_option(b"math", math); lib.smash_b200_clear_cache()
    rows[name] = [float(simulation.optimize(cases.cance(), mapping="distributed", options={"maxiter": k}, solver=solver).output.cost)
                  for k in range(0, 9)]
lib.smash_b200_set_option(b"math", 1)
print("iterations      " + " ".join("%9d" % k for k in range(0, 9)))
for n, r in rows.items():
    print("%-15s " % n + " ".join("%9.6f" % x for x in r))
