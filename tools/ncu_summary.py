#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics (raw page) + hottest source lines (source page)."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
for i, h in enumerate(hdr):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and float(vals[i] or 0) > 0.15):
        print(f"{h:90s} {units[i]:14s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# find header row
hi = next((i for i, r in enumerate(rows) if "Source" in r and any("Instructions Executed" in c for c in r)), None)
if hi is None:
    print("no source page"); sys.exit(0)
h = rows[hi]
ci = {name: h.index(name) for name in h}
def col(name):
    return next((k for k in h if k.startswith(name)), None)
c_src, c_inst, c_samp = col("Source"), col("Instructions Executed"), col("Warp Stall Sampling (All")
agg = collections.OrderedDict()
tot_i = tot_s = 0
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    try:
        ni = float(r[ci[c_inst]] or 0); ns = float(r[ci[c_samp]] or 0)
    except ValueError:
        continue
    key = r[ci[c_src]].strip()[:110]
    a = agg.setdefault(key, [0, 0]); a[0] += ni; a[1] += ns
    tot_i += ni; tot_s += ns
print(f"\n total warp-instructions {tot_i:.3e}   stall samples {tot_s:.0f}")
print(" %inst  %stall  source")
for k, (ni, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f" {100*ni/max(tot_i,1):5.1f}  {100*ns/max(tot_s,1):5.1f}   {k}")
