#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total us, share."""
import csv, sys, collections
path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if r]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    if r[mu] == "ns": v /= 1e3
    elif r[mu] == "ms": v *= 1e3
    a = agg.setdefault(r[kn], [0, 0.0]); a[0] += 1; a[1] += v
OURS = ("dic::", "gn_solve", "pyramid_level", "rect_grid", "rect_fill", "rect_tiles", "compact_", "tiles_scatter", "expand_spans", "bbox_kernel", "copy_rows")
ours = {k: v for k, v in agg.items() if k.startswith("void dic") or any(t in k for t in OURS)}
tot = sum(v[1] for v in ours.values()); all_t = sum(v[1] for v in agg.values())
print(f"# our kernels: {tot/1e3:.3f} ms of {all_t/1e3:.3f} ms profiled")
print("# kernel | launches | total us | share of our kernels")
for k, (n, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:112]:112s} {n:6d} {t:12.1f} {100*t/tot:6.2f}%")
print("# --- not ours (torch: synthetic speckle generation, L2 flush fill, copies) ---")
for k, (n, t) in sorted(((k, v) for k, v in agg.items() if k not in ours), key=lambda kv: -kv[1][1])[:8]:
    print(f"{k[:112]:112s} {n:6d} {t:12.1f}")
