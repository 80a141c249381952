#!/usr/bin/env python
"""GPU probe: time correlate() on a workload for several pyramid ranges / modes / kernel variants."""
import sys, os, time, itertools
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from correlation_b200 import engine

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = bench.workload(name)
dev = torch.device("cuda", 0)
und, dfm = bench.make_images(w, dev)
npar = 12 if w["model"] == "quad" else 6
pyrs = [w["pyramid"], (0, 1, 0), (w["pyramid"][2], 1, w["pyramid"][2])]
for pyr, mode, variant in itertools.product(pyrs, (engine.MODE_PARITY, engine.MODE_FAST), (0, 1)):
    eng = engine.CudaEngine(0, fitting_model=engine.FM_QUADRATIC if npar == 12 else engine.FM_UVUxUyVxVy, arith_mode=mode)
    eng.set_kernel_variant(variant)
    eng.resetImagePyramidsDevice(und.data_ptr(), dfm.data_ptr(), None, w["rows"], w["cols"], w["cols"], pyramid=pyr)
    d = w["domain"]
    eng.resetPolygon(0, *d[1:])
    guess = np.zeros(npar, np.float32)
    if pyr[2] == 0:
        guess[:len(w["truth"])] = np.array(w["truth"], np.float32) * 0.98
    for _ in range(3):
        r = eng.correlate(0, guess)
    ms = []
    for _ in range(5):
        r = eng.correlate(0, guess)
        ms.append(eng.last_correlate_ms())
    ev = r["evaluations"][:pyr[2] + 1]
    pe = r["pixel_evaluations"]
    print(f"{name} pyr={pyr} mode={'fast' if mode else 'parity'} kernel={'list' if variant else 'tiles'}: "
          f"{np.median(ms):8.3f} ms  evals={ev} px*ev={pe:.3e}  {pe/np.median(ms)/1e6:8.2f} Gpx*ev/s  err={r['error_code']} it={r['iterations']}")
    eng.close()
