#!/usr/bin/env python
"""Where does the chi spread of small subsets come from? For a stratified sample of the c4 subsets: GPU batch result
vs the oracle (fp64 accumulators), with the LM path (evaluations per level) of both."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, oracle
from correlation_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = bench.workload("c4")
und_t, dfm_t = bench.make_images(w, torch.device("cuda", 0))
eng = engine.CudaEngine(0)
eng.resetImagePyramidsDevice(und_t.data_ptr(), dfm_t.data_ptr(), None, 8192, 8192, 8192, pyramid=w["pyramid"])
boxes = bench.subset_boxes(*w["domain"][1:])
ids = bench.stratified_sample(len(boxes), n)
eng.resetPolygonRectGrid(0, np.array([boxes[i] for i in ids], np.int32))
_, res = eng.correlate_batch_raw(0, np.zeros((len(ids), 6), np.float32))
und, dfm = und_t.cpu().numpy(), dfm_t.cpu().numpy()
o = oracle.OracleEngine(n_threads=8, pyramid=w["pyramid"], accum_double=True, real_threads=True)
o.set_image("und", und); o.set_image("def", dfm)
of = oracle.OracleEngine(n_threads=20, pyramid=w["pyramid"], accum_double=False)
of.set_image("und", und); of.set_image("def", dfm)
ox = oracle.OracleEngine(n_threads=8, pyramid=w["pyramid"], accum_double=True, real_threads=True, solve_double=True)
ox.set_image("und", und); ox.set_image("def", dfm)
rels = {"gpu_vs_exact": [], "gpu_vs_qr64": [], "ref32_vs_exact": [], "qr64_vs_exact": []}
for k, i in enumerate(ids):
    bx = boxes[i]
    c = ((bx[0] + bx[2]) / 2, (bx[1] + bx[3]) / 2)
    want = o.correlate(np.zeros(6), oracle.rect_points(*bx), center=c)
    wf = of.correlate(np.zeros(6), oracle.rect_points(*bx), center=c)
    wx = ox.correlate(np.zeros(6), oracle.rect_points(*bx), center=c)
    rels["gpu_vs_exact"].append(abs(res["chi"][k] - wx["chi"]) / wx["chi"])
    rels["gpu_vs_qr64"].append(abs(res["chi"][k] - want["chi"]) / want["chi"])
    rels["ref32_vs_exact"].append(abs(wf["chi"] - wx["chi"]) / wx["chi"])
    rels["qr64_vs_exact"].append(abs(want["chi"] - wx["chi"]) / wx["chi"])
    rel = abs(res["chi"][k] - want["chi"]) / want["chi"]
    relf = abs(wf["chi"] - want["chi"]) / want["chi"]
    d = np.abs(res["resultingParameters"][k, :6] - want["params"])
    ge = res["evaluationsPerLevel"][k, :3].tolist()
    flag = "" if ge == want["evaluations"][:3] else "  <-- LM path differs"
    if rels["gpu_vs_exact"][-1] > 1e-5 or flag:
        print(f"subset {i:4d} vs exact-solve oracle {rels['gpu_vs_exact'][-1]:.2e} evals {wx['evaluations'][:3]} | vs fp64-acc QR oracle rel dchi {rel:.2e} (fp32-1thread oracle vs fp64 oracle {relf:.2e}) duv {d[:2].max():.1e} dgrad {d[2:].max():.1e} "
              f"evals gpu {ge} oracle {want['evaluations'][:3]} fp32-oracle {wf['evaluations'][:3]} iters {res['iterations'][k]} {want['iterations']}{flag}")

print("rel dchi over the sample: max / median / count > 1e-5")
for k, v in rels.items():
    v = np.array(v)
    print(f"  {k:16s} {v.max():.2e} {np.median(v):.2e} {(v > 1e-5).sum()} of {len(v)}")
