#!/usr/bin/env python
"""GPU probe: BASELINE config 4 (4096 subsets, one CTA each) build time, batch solve time, parity sample."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, oracle
from correlation_b200 import engine
w = bench.workload("c4")
und, dfm = bench.make_images(w, torch.device("cuda", 0))
boxes = bench.subset_boxes(*w["domain"][1:])
for mode in (engine.MODE_PARITY, engine.MODE_FAST):
    for variant in (0, 1):
        eng = engine.CudaEngine(0, arith_mode=mode)
        eng.set_kernel_variant(variant)
        eng.resetImagePyramidsDevice(und.data_ptr(), dfm.data_ptr(), None, w["rows"], w["cols"], w["cols"], pyramid=w["pyramid"])
        t0 = time.perf_counter()
        for k, bx in enumerate(boxes):
            eng.resetPolygon(k, *bx)
        eng.synchronize()
        t_build = time.perf_counter() - t0
        zero = np.zeros((len(boxes), 6), np.float32)
        for _ in range(2):
            rs = eng.correlate_batch(0, zero)
        t0 = time.perf_counter()
        rs = eng.correlate_batch(0, zero)
        wall = time.perf_counter() - t0
        pe = sum(r["pixel_evaluations"] for r in rs)
        errs = sum(r["error_code"] != 0 for r in rs)
        print(f"c4 mode={'fast' if mode else 'parity'} kernel={'list' if variant else 'tiles'}: build {t_build:.2f}s  kernel {eng.last_correlate_ms():.3f} ms wall {wall*1e3:.3f} ms  "
              f"px*ev={pe:.3e} -> {pe/eng.last_correlate_ms()/1e6:.2f} Gpx*ev/s errs={errs} evals0={rs[0]['evaluations'][:3]} n0={rs[0]['points_per_level'][:3]}")
        if mode == engine.MODE_PARITY and variant == 0:
            uh, dh = und.cpu().numpy(), dfm.cpu().numpy()
            o = oracle.OracleEngine(n_threads=1, pyramid=w["pyramid"], accum_double=True)
            o.set_image("und", uh); o.set_image("def", dh)
            worst = [0, 0, 0, 0]
            for k in range(0, len(boxes), 257):
                bx = boxes[k]
                want = o.correlate(np.zeros(6), oracle.rect_points(*bx), center=((bx[0]+bx[2])/2, (bx[1]+bx[3])/2))
                d = np.abs(rs[k]["params"] - want["params"])
                worst = [max(worst[0], d[:2].max()), max(worst[1], d[2:].max()), max(worst[2], abs(rs[k]["chi"]-want["chi"])/want["chi"]), max(worst[3], abs(rs[k]["iterations"]-want["iterations"]))]
            print("   parity sample (16 subsets) worst duv %.2e dgrad %.2e relchi %.2e diter %d" % tuple(worst))
        eng.close()
