#!/usr/bin/env python
"""Opcode histogram of an .ncu-rep weighted by executed warp-instructions (source page, SASS view)."""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]
px = float(sys.argv[2]) if len(sys.argv) > 2 else None   # pixel*evaluations in the profiled launch
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Instructions Executed" in c for c in r))
h = rows[hi]
c_src = h.index("Source"); c_inst = next(i for i, k in enumerate(h) if k.startswith("Instructions Executed"))
hist = collections.Counter(); tot = 0
for r in rows[hi + 1:]:
    if len(r) <= c_inst: continue
    try: n = float(r[c_inst] or 0)
    except ValueError: continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[c_src])
    if not m: continue
    op = m.group(2)
    hist[op] += n; tot += n
print(f"total warp-instructions {tot:.4e}" + (f"  = {tot*32/px:.1f} thread-instr per pixel*evaluation" if px else ""))
for op, n in hist.most_common(40):
    print(f"  {op:12s} {100*n/tot:6.2f}%" + (f"  {n*32/px:7.2f} /px" if px else ""))
