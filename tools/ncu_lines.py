#!/usr/bin/env python
"""Hottest CUDA source lines of an .ncu-rep by executed warp-instructions and stall samples."""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
key = sys.argv[3] if len(sys.argv) > 3 else "inst"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
out = []; fname = ""
hi = None
for i, r in enumerate(rows):
    if r and r[0] == "File Name": fname = r[1].split("/")[-1]
    if "Source" in r and any(k.startswith("Instructions Executed") for k in r):
        h = r; ii = next(j for j, k in enumerate(h) if k.startswith("Instructions Executed")); si = h.index("Warp Stall Sampling (All Samples)")
        continue
    if len(r) > 8 and r[0].isdigit():
        try: out.append((float(r[ii] or 0), float(r[si] or 0), fname, int(r[0]), r[1].strip()[:105]))
        except (ValueError, NameError): pass
ti = sum(o[0] for o in out); ts = sum(o[1] for o in out)
out.sort(key=lambda t: -(t[0] if key == "inst" else t[1]))
print(f"total warp-instructions {ti:.4e}, stall samples {ts:.0f}")
print(" %inst %stall  file:line  source")
for n, s, f, ln, src in out[:topn]:
    print(f" {100*n/ti:5.2f} {100*s/max(ts,1):5.2f}  {f}:{ln}  {src}")
