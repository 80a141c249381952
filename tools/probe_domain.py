#!/usr/bin/env python
"""GPU probe: domain-build times (mask compaction + tiles) of the c2 annulus and the c4 grid of 4096 rectangles."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from correlation_b200 import engine
for name in ("c2", "c4"):
    w = bench.workload(name)
    und, dfm = bench.make_images(w, torch.device("cuda", 0))
    eng = engine.CudaEngine(0, fitting_model=engine.FM_QUADRATIC if w["model"] == "quad" else engine.FM_UVUxUyVxVy)
    eng.resetImagePyramidsDevice(und.data_ptr(), dfm.data_ptr(), None, w["rows"], w["cols"], w["cols"], pyramid=w["pyramid"])
    ts = []
    for rep in range(4):
        eng.synchronize()
        t0 = time.perf_counter()
        if name == "c2":
            eng.resetPolygon(0, *w["domain"][1:])
        else:
            eng.resetPolygonRectGrid(0, np.array(bench.subset_boxes(*w["domain"][1:]), np.int32))
        eng.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    n = eng.level_points(0, 0).shape[0] if name == "c2" else 4096 * 125 * 125
    print(f"{name}: domain build {min(ts):.2f} ms (runs: {[round(t, 2) for t in ts]}), {n} level-0 pixels emitted -> {8.0 * n / (min(ts) * 1e-3) / 1e9:.1f} GB/s of the 8 B / emitted pixel figure end to end")
    eng.close()
