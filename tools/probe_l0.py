#!/usr/bin/env python
"""GPU probe for ncu: level-0-only solve of the c2 annulus (steady-state pass, little sync)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from correlation_b200 import engine
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
w = bench.workload("c2")
und, dfm = bench.make_images(w, torch.device("cuda", 0))
eng = engine.CudaEngine(0, fitting_model=engine.FM_QUADRATIC, arith_mode=mode)
eng.resetImagePyramidsDevice(und.data_ptr(), dfm.data_ptr(), None, 4096, 4096, 4096, pyramid=(0, 1, 0))
eng.resetPolygon(0, *w["domain"][1:])
g = np.zeros(12, np.float32); g[:12] = np.array(w["truth"], np.float32) * 0.98
for _ in range(4):
    r = eng.correlate(0, g)
print(f"l0-only mode={mode}: {eng.last_correlate_ms():.3f} ms evals {r['evaluations'][:1]} {r['pixel_evaluations']/eng.last_correlate_ms()/1e6:.1f} Gpx*ev/s")
