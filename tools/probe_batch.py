#!/usr/bin/env python
"""GPU probe for ncu: N subsets of C4 through correlate_batch with a given mode / kernel variant."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from correlation_b200 import engine
n, mode, variant = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
w = bench.workload("c4")
und, dfm = bench.make_images(w, torch.device("cuda", 0))
boxes = bench.subset_boxes(*w["domain"][1:])[:n]
eng = engine.CudaEngine(0, arith_mode=mode)
eng.set_kernel_variant(variant)
eng.resetImagePyramidsDevice(und.data_ptr(), dfm.data_ptr(), None, w["rows"], w["cols"], w["cols"], pyramid=w["pyramid"])
for k, bx in enumerate(boxes):
    eng.resetPolygon(k, *bx)
zero = np.zeros((len(boxes), 6), np.float32)
for _ in range(3):
    rs = eng.correlate_batch(0, zero)
pe = sum(r["pixel_evaluations"] for r in rs)
print(f"batch n={n} mode={mode} variant={variant}: {eng.last_correlate_ms():.3f} ms, {pe/eng.last_correlate_ms()/1e6:.2f} Gpx*ev/s")
