#!/usr/bin/env python
"""GPU probe: tile kernel vs pixel-list kernel (both parity mode) on c4 subsets: chi and parameters."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from correlation_b200 import engine
w = bench.workload("c4")
und, dfm = bench.make_images(w, torch.device("cuda", 0))
boxes = bench.subset_boxes(*w["domain"][1:])
ids = bench.stratified_sample(len(boxes), 256)
eng = engine.CudaEngine(0, arith_mode=0)
eng.resetImagePyramidsDevice(und.data_ptr(), dfm.data_ptr(), None, w["rows"], w["cols"], w["cols"], pyramid=w["pyramid"])
eng.resetPolygonRectGrid(0, np.array([boxes[i] for i in ids], np.int32))
zero = np.zeros((len(ids), 6), np.float32)
out = {}
for variant in (0, 1):
    eng.set_kernel_variant(variant)
    _, res = eng.correlate_batch_raw(0, zero)
    out[variant] = res.copy()
a, b = out[0], out[1]
rel = np.abs(a["chi"] - b["chi"]) / b["chi"]
dp = np.abs(a["resultingParameters"][:, :6] - b["resultingParameters"][:, :6])
same = (a["evaluationsPerLevel"][:, :3] == b["evaluationsPerLevel"][:, :3]).all(1)
print("tile vs list: same LM path", same.sum(), "of", len(ids), "| chi rel median %.2e max %.2e | d uv max %.2e median %.2e | d grad max %.2e" %
      (np.median(rel[same]), rel[same].max(), dp[same][:, :2].max(), np.median(dp[same][:, :2].max(1)), dp[same][:, 2:].max()))
print("numberOfPoints equal:", (a["numberOfPoints"] == b["numberOfPoints"]).all(), a["numberOfPoints"][:4], "pointsPerLevel", a["pointsPerLevel"][0, :3], b["pointsPerLevel"][0, :3])
