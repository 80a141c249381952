#!/usr/bin/env python
"""GPU probe: master-CTA timeline per evaluation of one workload (tile kernel)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from correlation_b200 import engine
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
w = bench.workload(name)
und, dfm = bench.make_images(w, torch.device("cuda", 0))
npar = 12 if w["model"] == "quad" else 6
eng = engine.CudaEngine(0, fitting_model=engine.FM_QUADRATIC if npar == 12 else engine.FM_UVUxUyVxVy, arith_mode=mode)
eng.resetImagePyramidsDevice(und.data_ptr(), dfm.data_ptr(), None, w["rows"], w["cols"], w["cols"], pyramid=w["pyramid"])
eng.resetPolygon(0, *w["domain"][1:])
for _ in range(4):
    r = eng.correlate(0, np.zeros(npar, np.float32))
t = eng.timeline()
print(f"{name} mode={mode} total {eng.last_correlate_ms():.3f} ms, evals {r['evaluations'][:5]}, points {r['points_per_level'][:5]}")
print("  eval:  own-pass  wait-others  sum+LM   (us)   since start")
for i, m in enumerate(t):
    print(f"  {i:3d}  {(m[1]-m[0])/1e3:8.1f} {(m[2]-m[1])/1e3:8.1f} {(m[3]-m[2])/1e3:8.1f}   {(m[3]-t[0][0])/1e3:8.1f}")
ct = eng.cta_times(296)
ct = (ct - t[-1][0]) / 1e3  # us since the start of the last pass
ct = ct[ct > 0]
print("last pass: per-CTA end times (us): n", len(ct), "min %.1f p10 %.1f median %.1f p90 %.1f max %.1f" % (ct.min(), np.percentile(ct, 10), np.median(ct), np.percentile(ct, 90), ct.max()))
print("  first 8 CTAs", np.round(ct[:8], 1), " last 8", np.round(ct[-8:], 1))
order = np.argsort(ct)
print("  slowest CTAs", order[-8:], np.round(ct[order[-8:]], 1))
sm = eng.cta_smids(296)[: len(ct)]
print("  CTA -> SM:", " ".join(f"{i}:{s}" for i, s in list(enumerate(sm))[:16]), "...")
by_sm = {}
for i, s in enumerate(sm):
    by_sm.setdefault(int(s), []).append(i)
print("  CTAs per SM histogram:", np.bincount([len(v) for v in by_sm.values()]))
fast = [i for i in range(len(ct)) if ct[i] < np.median(ct) - 0.5 * (np.median(ct) - ct.min())]
print("  fast CTAs:", fast[:40], "their SM mates:", [[j for j in by_sm[int(sm[i])] if j != i] for i in fast[:12]])
print("  end time by SM (max over its CTAs) percentiles:", np.percentile([max(ct[j] for j in v) for v in by_sm.values()], [0, 10, 50, 90, 100]).round(1))
