#!/usr/bin/env python
"""GPU probe: BASELINE tolerances of parity / fast mode against the double-accumulator oracle."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, oracle
from correlation_b200 import engine
for name in sys.argv[1:] or ["c1", "c2"]:
    w = bench.workload(name)
    und, dfm = bench.make_images(w, torch.device("cuda", 0))
    uh, dh = und.cpu().numpy(), dfm.cpu().numpy()
    npar = 12 if w["model"] == "quad" else 6
    d = w["domain"]
    if d[0] == "rect":
        xy, center = oracle.rect_points(*d[1:]), ((d[1] + d[3]) / 2.0, (d[2] + d[4]) / 2.0)
    else:
        xy, center = oracle.annulus_points(*d[1:]), None
    model = oracle.FM_QUAD if npar == 12 else oracle.FM_AFFINE
    t0 = time.time()
    refs = {}
    for tag, kw in (("double", dict(accum_double=True)), ("float20", dict())):
        O = oracle.OracleEngine(model=model, n_threads=20, pyramid=w["pyramid"], real_threads=True, **kw)
        O.set_image("und", uh); O.set_image("def", dh)
        refs[tag] = O.correlate(np.zeros(npar, np.float32), xy, center=center)
    print(f"{name}: oracle runs {time.time()-t0:.1f}s; chi double {refs['double']['chi']:.8f} float20 {refs['float20']['chi']:.8f} "
          f"(rel {abs(refs['double']['chi']-refs['float20']['chi'])/refs['double']['chi']:.2e}) evals {refs['double']['evaluations'][:5]}")
    for mode in (engine.MODE_PARITY, engine.MODE_FAST):
        eng = engine.CudaEngine(0, fitting_model=engine.FM_QUADRATIC if npar == 12 else engine.FM_UVUxUyVxVy, arith_mode=mode)
        eng.resetImagePyramidsDevice(und.data_ptr(), dfm.data_ptr(), None, w["rows"], w["cols"], w["cols"], pyramid=w["pyramid"])
        eng.resetPolygon(0, *d[1:])
        r = eng.correlate(0, np.zeros(npar, np.float32))
        for tag, want in refs.items():
            dd = np.abs(r["params"].astype(np.float64) - want["params"])
            print(f"  {'fast  ' if mode else 'parity'} vs {tag:8s}: duv {dd[:2].max():.2e} dgrad {dd[2:6].max():.2e}" +
                  (f" d2nd {dd[6:].max():.2e}" if npar == 12 else "") +
                  f" rel chi {abs(r['chi']-want['chi'])/want['chi']:.2e} iters {r['iterations']}/{want['iterations']} evals {r['evaluations'][:5]} center {r['und_center']} vs {want['und_center']}")
        eng.close()
