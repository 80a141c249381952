#!/usr/bin/env python
"""LM decision trace of single c4 subsets with the device's / the oracle's evaluation and solver mixed (tests/lm_harness.py)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench, oracle, lm_harness
from correlation_b200 import engine

ids = [int(a) for a in sys.argv[1:]] or [3250, 980, 455]
w = bench.workload("c4")
und_t, dfm_t = bench.make_images(w, torch.device("cuda", 0))
und, dfm = und_t.cpu().numpy(), dfm_t.cpu().numpy()
eng = engine.CudaEngine(0)
eng.resetImagePyramidsDevice(und_t.data_ptr(), dfm_t.data_ptr(), None, 8192, 8192, 8192, pyramid=w["pyramid"])
boxes = bench.subset_boxes(*w["domain"][1:])
o64 = oracle.OracleEngine(n_threads=1, pyramid=w["pyramid"], accum_double=True)
o32 = oracle.OracleEngine(n_threads=1, pyramid=w["pyramid"], accum_double=False)
for o in (o64, o32):
    o.set_image("und", und); o.set_image("def", dfm)
for i in ids:
    bx = boxes[i]
    c = ((bx[0] + bx[2]) / 2, (bx[1] + bx[3]) / 2)
    xy = oracle.rect_points(*bx)
    eng.resetPolygon(0, *bx)
    for o in (o64, o32):
        o.set_points(xy, center=c)
    npts = {lv: len(o64.level_points(lv)) for lv in (0, 1, 2)}
    gpu_eval = lambda lv, p: eng.evaluate(0, lv, p)[:3] + (eng.evaluate(0, lv, p)[3] > 0,)
    def orc_eval(o):
        def f(lv, p):
            A, b, chi, err = o.evaluate(lv, p)
            return A, b, chi, err != 0
        return f
    gpu_solve = lambda A, b, lam, sc: eng.solve_step(A, b, float(lam), float(sc))
    orc_solve = lambda A, b, lam, sc: o64.solve_step(A, b, float(lam), float(sc))
    def np64_solve(A, b, lam, sc):  # exact (fp64) damped solve: the noise-free reference point
        n = len(b)
        M = np.triu(np.asarray(A, np.float64)); M = M + M.T - np.diag(np.diag(M))
        M = M * float(sc); M[np.diag_indices(n)] *= (1.0 + float(lam))
        return np.linalg.solve(M, np.asarray(b, np.float64) * float(sc)).astype(np.float32)
    direct = eng.correlate(0, np.zeros(6, np.float32))
    want = o64.correlate(np.zeros(6, np.float32), xy, center=c)
    print(f"\n=== subset {i}: kernel chi {direct['chi']:.9g} evals {direct['evaluations'][:3]}   oracle(fp64 acc) chi {want['chi']:.9g} evals {want['evaluations'][:3]}"
          f"  rel {abs(direct['chi'] - want['chi']) / want['chi']:.2e}")
    combos = [("gpu eval + gpu cholesky", gpu_eval, gpu_solve), ("gpu eval + oracle QR", gpu_eval, orc_solve), ("gpu eval + exact solve", gpu_eval, np64_solve),
              ("orc64 eval + oracle QR", orc_eval(o64), orc_solve), ("orc64 eval + gpu cholesky", orc_eval(o64), gpu_solve), ("orc64 eval + exact solve", orc_eval(o64), np64_solve),
              ("orc32 eval + oracle QR", orc_eval(o32), orc_solve)]
    base = None
    for name, ev, sv in combos:
        tr = []
        r = lm_harness.newton_raphson(ev, sv, npts, np.zeros(6, np.float32), w["pyramid"], trace=tr)
        if base is None:
            base = r
        path = "".join({"init": "I", "redo": "R", "tent": "T", "accept": "+", "reject": "-"}[t[1]] + ("|" if t[1] == "init" and t[0] != 2 else "") for t in tr)
        lvl0 = [f"{t[2]:.8g}" for t in tr if t[0] == 0 and t[1] in ("init", "redo", "tent")]
        print(f"{name:28s} chi {r['chi']:.9g} rel-to-oracle {abs(r['chi'] - want['chi']) / want['chi']:.2e} u {r['params'][0]:.7f} v {r['params'][1]:.7f} path {path}  level-0 chi: {' '.join(lvl0)}")
