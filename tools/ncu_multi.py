"""Summarise every kernel of an .ncu-rep exported as CSV (raw + source pages).
usage: python tools/ncu_multi.py raw.csv source.csv [units_per_kernel...]"""
import csv, sys
from collections import Counter
raw = list(csv.reader(open(sys.argv[1])))
hdr, units = raw[0], raw[1]
want = ["Kernel Name","gpu__time_duration.sum","launch__grid_size","launch__registers_per_thread","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","lts__t_bytes.sum","l1tex__t_bytes.sum","smsp__warps_eligible.avg.per_cycle_active","smsp__warps_active.avg.per_cycle_active","lts__t_sector_hit_rate.pct","sm__cycles_active.avg","sm__cycles_active.max","sm__cycles_elapsed.max"]
src = list(csv.reader(open(sys.argv[2])))
blocks=[]; cur=None
for r in src:
    if r and r[0]=='Kernel Name': cur={'name':r[1],'rows':[]}; blocks.append(cur); continue
    if r and r[0]=='Address': continue
    if cur is not None and len(r)>6: cur['rows'].append(r)
for bi, v in enumerate(raw[2:]):
    print('=' * 100)
    for i,h in enumerate(hdr):
        if h in want: print(f"{h:62s} {v[i]:>22s} {units[i]}")
    st = [(h, float(v[i])) for i,h in enumerate(hdr) if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
    tot = sum(x for _,x in st) or 1
    print("stalls:", ", ".join(f"{h.split('stalled_')[1]} {100*x/tot:.1f}%" for h,x in sorted(st, key=lambda z:-z[1])[:9]))
    if bi < len(blocks):
        data = blocks[bi]['rows']
        tot_e = sum(int(r[5]) for r in data); tot_s = sum(int(r[2]) for r in data) or 1
        c = Counter()
        for r in data:
            t = r[1].split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]; c[op]+=int(r[5])
        print('instr', tot_e, ':', ', '.join(f"{k} {100*x/tot_e:.1f}%" for k,x in c.most_common(16)))
        for r in sorted(data, key=lambda r:-int(r[2]))[:int(sys.argv[3]) if len(sys.argv)>3 else 14]:
            print(f"   {100*int(r[2])/tot_s:5.1f}%  exec {int(r[5]):>10d}  {r[1][:100]}")
