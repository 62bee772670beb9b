"""PCIe floor of the end-to-end forward call: pinned host <-> device copy rates, one direction at a time and both at once
(torch is only the plumbing here).  usage: python tools/pcie_probe.py [GB]"""
import sys, time
import torch

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 2.6
n = int(gb * 1e9 / 4)
h_in, h_out = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.float32, device="cuda"), torch.zeros(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


t = timed(h2d); print(f"H2D  {gb:.2f} GB: {t*1e3:7.1f} ms  {gb/t:6.1f} GB/s")
t = timed(d2h); print(f"D2H  {gb:.2f} GB: {t*1e3:7.1f} ms  {gb/t:6.1f} GB/s")
t = timed(both); print(f"both {2*gb:.2f} GB: {t*1e3:7.1f} ms  {gb/t:6.1f} GB/s per direction")
