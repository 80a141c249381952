#!/usr/bin/env python
"""GPU probe: host-side overhead of correlate_batch / correlate around the kernel time."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from correlation_b200 import engine
w = bench.workload("c4")
und, dfm = bench.make_images(w, torch.device("cuda", 0))
boxes = bench.subset_boxes(*w["domain"][1:])
eng = engine.CudaEngine(0, arith_mode=engine.MODE_FAST)
eng.resetImagePyramidsDevice(und.data_ptr(), dfm.data_ptr(), None, w["rows"], w["cols"], w["cols"], pyramid=w["pyramid"])
for n in (512, 4096):
    for k, bx in enumerate(boxes[:n]):
        eng.resetPolygon(k, *bx)
    g = np.zeros((n, 6), np.float32); res = np.zeros(n, engine.RESULT_DTYPE)
    for _ in range(3):
        eng.lib.dic_correlate_batch(eng.h, 0, n, g.ctypes.data, res.ctypes.data)
    ts, ks = [], []
    for _ in range(20):
        g[:] = 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.lib.dic_correlate_batch(eng.h, 0, n, g.ctypes.data, res.ctypes.data)
        ts.append(time.perf_counter() - t0); ks.append(eng.last_correlate_ms())
    print(f"batch n={n}: wall {np.median(ts)*1e3:.3f} ms kernel {np.median(ks):.3f} ms overhead {np.median(ts)*1e3-np.median(ks):.3f} ms")
