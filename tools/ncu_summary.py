"""Summarise an .ncu-rep (read on the CPU box): key raw metrics, stall reasons, hot SASS by opcode.
usage: python tools/ncu_summary.py file.ncu-rep [warp_ticks]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
wt = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:65s} {vals[i]:>22s} {units[i]}")
st = [(h, float(vals[i])) for i, h in enumerate(hdr) if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
tot = sum(v for _, v in st) or 1
print("stall reasons (pc sampling, all samples):")
for h, v in sorted(st, key=lambda x: -x[1])[:10]:
    print(f"   {h.split('stalled_')[1]:25s} {100*v/tot:5.1f}%")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2, data = rows[1], rows[2:]
iS, iE, iW = h2.index("Source"), h2.index("Instructions Executed"), h2.index("Warp Stall Sampling (All Samples)")
tot_e = sum(int(r[iE]) for r in data)
print("instructions executed (warp-level):", tot_e, (f"= {tot_e/wt:.1f} per warp-tick" if wt else ""))
c = Counter()
for r in data:
    t = r[iS].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += int(r[iE])
print("by opcode:", ", ".join(f"{k} {100*v/tot_e:.1f}%" for k, v in c.most_common(18)))
print("top stalled instructions:")
for r in sorted(data, key=lambda r: -int(r[iW]))[:12]:
    print(f"   {r[iW]:>6s} samples  exec {r[iE]:>9s}  {r[iS][:80]}")
