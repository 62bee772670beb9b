"""adjust_interception_store on the France mesh (device search) next to the oracle on a sample of the same cells.
usage: python tools/ci_bench.py [--T 720]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases, oracle
from smash_b200.solver import _mw_interception_store as dev

ap = argparse.ArgumentParser(); ap.add_argument("--T", type=int, default=720)
a = ap.parse_args()
m = cases.france(T=a.T)
di = (np.arange(a.T) // 24 + 1).astype(np.int32)
nday = int(di.max())
t0 = time.perf_counter()
dev.adjust_interception_store(m.setup, m.mesh, m.input_data, m.parameters, nday, di)
wall = time.perf_counter() - t0
units = m.mesh.nac * a.T * 49
print(f"device: kernel {dev.last_kernel_ms:.2f} ms, call {wall*1e3:.0f} ms (upload of {m.mesh.nac*a.T*8/1e9:.2f} GB of forcing included), "
      f"{units/dev.last_kernel_ms*1e-6:.1f} G candidate-cell-steps/s, forcing read {7*m.mesh.nac*a.T*8/dev.last_kernel_ms*1e-6:.0f} GB/s")
# the oracle on a crop of the same model (one host core)
c = cases.france(T=a.T, sub=(300, 500, 300, 500))
t0 = time.perf_counter()
oracle.adjust_interception_store(c.setup, c.mesh, c.input_data, c.parameters, nday, di)
dt = time.perf_counter() - t0
print(f"oracle: {c.mesh.nac} cells in {dt:.2f} s -> {c.mesh.nac*a.T*49/dt*1e-9:.3f} G candidate-cell-steps/s on one core; "
      f"France at that rate: {m.mesh.nac/c.mesh.nac*dt:.0f} s")
vals, cnt = np.unique(np.asarray(m.parameters.ci)[m.mesh.active_cell == 1], return_counts=True)
print("chosen capacities:", dict(zip([round(float(v), 1) for v in vals[:8]], cnt[:8].tolist())), "...")
