import sys, time, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, cases, smash_b200
from smash_b200 import _lib as L
lib = L.lib()
m = cases.cance()
ns = 4096
rng = np.random.RandomState(99)
smp = np.asfortranarray(np.stack([rng.uniform(lo, hi, ns) for lo, hi in [(1e-6, 1e3), (1e-6, 1e3), (-50, 50), (1e-6, 1e3)]]).astype(np.float32))
m.input_data._forcing_version = 1
for eng in (0, 1, 0, 1):
    lib.smash_b200_set_option(b"ensemble_engine", eng)
    cost = np.zeros(ns, np.float32); q0 = np.zeros((0,), np.float32)
    f = lambda: smash_b200.compute_multiple_run(m.setup, m.mesh, m.input_data, m.parameters, m.states, m.output, smp, cases.IND_CP_CFT_EXC_LR, cost, q0)
    f(); t0 = time.perf_counter()
    for _ in range(5): f()
    dt = (time.perf_counter() - t0) / 5
    print(f"ensemble_engine={eng}: {dt*1e3:.2f} ms per call, {ns*383*1440/dt:.3e} cs/s", flush=True)
