#!/usr/bin/env python
"""GPU probe: does one evaluation through the tile kernel equal the list kernel / oracle? (parity mode)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from correlation_b200 import engine, synth
truth = (1.1, 0.6, 0.002, -0.001, 0.001, 0.002)
u2, d2 = synth.make_pair(384, 384, 41, truth, center=(192, 192))
eng = engine.CudaEngine(0)
eng.set_max_iters(1)
eng.resetImagePyramids(u2, d2, pyramid=(0, 1, 0))
boxes = [(64 + 64 * i + 1, 64 + 64 * j + 1, 64 + 64 * i + 63, 64 + 64 * j + 63) for i in range(4) for j in range(4)]
o = oracle.OracleEngine(n_threads=1, pyramid=(0, 1, 0), accum_double=True)
o.set_image("und", u2); o.set_image("def", d2)
for k, bx in enumerate(boxes):
    eng.resetPolygon(k, *bx)
    cx, cy = (bx[0] + bx[2]) / 2, (bx[1] + bx[3]) / 2
    p = np.array([1.1 + 0.002 * (cx - 192) - 0.001 * (cy - 192), 0.6 + 0.001 * (cx - 192) + 0.002 * (cy - 192), 0.002, -0.001, 0.001, 0.002], np.float32)
    A, b, chi_eval, oob = eng.evaluate(k, 0, p)
    n = (bx[2] - bx[0] + 1) * (bx[3] - bx[1] + 1)
    o.set_points(oracle.rect_points(*bx), center=(cx, cy))
    Ao, bo, chio, _ = o.evaluate(0, p)
    res = {}
    for variant in (1, 0):
        eng.set_kernel_variant(variant)
        r = eng.correlate(k, p)
        res[variant] = r
    print(f"subset {k}: eval chi/N {chi_eval/n:.8f} oracle {chio/n:.8f} | correlate list chi {res[1]['chi']:.8f} tiles chi {res[0]['chi']:.8f} "
          f"rel(tiles-list) {abs(res[0]['chi']-res[1]['chi'])/res[1]['chi']:.2e}  dparams {np.abs(res[0]['params']-res[1]['params']).max():.2e} evals {res[0]['evaluations'][:1]} {res[1]['evaluations'][:1]}")
