#!/usr/bin/env python
"""GPU probe: master-CTA timeline per evaluation + batch-vs-oracle diagnostics."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, oracle
from correlation_b200 import engine, synth
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = bench.workload(name)
und, dfm = bench.make_images(w, torch.device("cuda", 0))
npar = 12 if w["model"] == "quad" else 6
for mode, variant in ((engine.MODE_FAST, 0), (engine.MODE_FAST, 1)):
    eng = engine.CudaEngine(0, fitting_model=engine.FM_QUADRATIC if npar == 12 else engine.FM_UVUxUyVxVy, arith_mode=mode)
    eng.set_kernel_variant(variant)
    eng.resetImagePyramidsDevice(und.data_ptr(), dfm.data_ptr(), None, w["rows"], w["cols"], w["cols"], pyramid=w["pyramid"])
    eng.resetPolygon(0, *w["domain"][1:])
    for _ in range(3):
        r = eng.correlate(0, np.zeros(npar, np.float32))
    t = eng.timeline()
    print(f"{name} kernel={'list' if variant else 'tiles'} total {eng.last_correlate_ms():.3f} ms, evals {r['evaluations'][:4]}, points {r['points_per_level'][:4]}")
    print("  units on the per-pixel (non-staged) path over the whole launch:", getattr(eng, "slow_units", None))
    print("  eval:  own-pass  wait-others  sum+LM   (us)")
    for i, m in enumerate(t):
        print(f"  {i:3d}  {(m[1]-m[0])/1e3:8.1f} {(m[2]-m[1])/1e3:8.1f} {(m[3]-m[2])/1e3:8.1f}")
    eng.close()
# batch diagnostics
truth = (1.1, 0.6, 0.002, -0.001, 0.001, 0.002)
u2, d2 = synth.make_pair(384, 384, 41, truth, center=(192, 192))
eng = engine.CudaEngine(0)
eng.resetImagePyramids(u2, d2, pyramid=(0, 1, 2))
boxes = [(64 + 64 * i + 1, 64 + 64 * j + 1, 64 + 64 * i + 63, 64 + 64 * j + 63) for i in range(4) for j in range(4)]
for k, bx in enumerate(boxes):
    eng.resetPolygon(k, *bx)
o = oracle.OracleEngine(n_threads=1, pyramid=(0, 1, 2), accum_double=True)
o.set_image("und", u2); o.set_image("def", d2)
for variant in (0, 1):
    eng.set_kernel_variant(variant)
    batch = eng.correlate_batch(0, np.zeros((16, 6), np.float32))
    worst = 0
    for k, bx in enumerate(boxes):
        want = o.correlate(np.zeros(6), oracle.rect_points(*bx), center=((bx[0] + bx[2]) / 2, (bx[1] + bx[3]) / 2))
        d = np.abs(batch[k]["params"] - want["params"])
        rc = abs(batch[k]["chi"] - want["chi"]) / want["chi"]
        worst = max(worst, rc)
        if k < 4 or rc > 5e-6:
            print(f"variant {variant} subset {k}: duv {d[:2].max():.2e} dgrad {d[2:].max():.2e} rel chi {rc:.2e} it {batch[k]['iterations']} {want['iterations']} evals {batch[k]['evaluations'][:3]} {want['evaluations'][:3]}")
    print("worst rel chi", worst)
# solver accuracy vs fp64
rng = np.random.default_rng(1)
J = rng.normal(size=(500, 6)) * np.array([1, 1, 40, 40, 40, 40])
A = (J.T @ J).astype(np.float32); b = (J.T @ rng.normal(size=500)).astype(np.float32)
Ad = A.astype(np.float64) / 500; Ad[np.diag_indices(6)] *= 1 + 1e-4
want = np.linalg.solve(Ad, b.astype(np.float64) / 500)
got = eng.solve_step(np.triu(A), b, 1e-4, 1 / 500)
oq = oracle.OracleEngine(n_threads=1).solve_step(np.triu(A), b, 1e-4, 1 / 500)
print("solver rel err: gpu", np.abs(got - want).max() / np.abs(want).max(), " oracle QR", np.abs(oq - want).max() / np.abs(want).max())
