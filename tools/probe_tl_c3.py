#!/usr/bin/env python
"""GPU probe: CTA 0's per-evaluation timeline of the c3 blob domain (one frame pair, grid form)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from correlation_b200 import engine, synth
w = bench.workload("c3")
dev = torch.device("cuda", 0)
und = synth.make_image(w["rows"], w["cols"], w["seed"], None, w["center"], device=dev)
dfm = synth.make_image(w["rows"], w["cols"], w["seed"], tuple(np.array(w["rate"])), w["center"], device=dev)
d = w["domain"]
contour = synth.star_polygon(d[1], d[2], d[3], n_vertices=d[4], seed=w["seed"])
eng = engine.CudaEngine(0, arith_mode=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
eng.resetImagePyramidsDevice(und.data_ptr(), dfm.data_ptr(), None, w["rows"], w["cols"], w["cols"], pyramid=w["pyramid"])
eng.resetPolygonBlob(0, contour) if hasattr(eng, "resetPolygonBlob") else eng.resetPolygon(0, contour)
for _ in range(4):
    r = eng.correlate(0, np.zeros(6, np.float32))
t = eng.timeline()
print(f"c3 total {eng.last_correlate_ms():.3f} ms, evals {r['evaluations'][:3]}, points {r['points_per_level'][:3]}")
print("  eval:  own-pass  wait-others  sum+LM   (us)   since start")
for i, m in enumerate(t):
    print(f"  {i:3d}  {(m[1]-m[0])/1e3:8.1f} {(m[2]-m[1])/1e3:8.1f} {(m[3]-m[2])/1e3:8.1f}   {(m[3]-t[0][0])/1e3:8.1f}")
