#!/usr/bin/env python
"""GPU probe (diagnostics build, -DDIC_BATCH_TIMELINE=1): phases of CTA 0's last subset in a c4 batch."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from correlation_b200 import engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w = bench.workload("c4")
und, dfm = bench.make_images(w, torch.device("cuda", 0))
boxes = bench.subset_boxes(*w["domain"][1:])[:n]
eng = engine.CudaEngine(0, arith_mode=0)
eng.resetImagePyramidsDevice(und.data_ptr(), dfm.data_ptr(), None, w["rows"], w["cols"], w["cols"], pyramid=w["pyramid"])
eng.resetPolygonRectGrid(0, np.array(boxes, np.int32))
zero = np.zeros((len(boxes), 6), np.float32)
for _ in range(3):
    rs = eng.correlate_batch(0, zero)
t = eng.timeline()
print(f"batch n={n}: {eng.last_correlate_ms():.3f} ms")
print("  eval:  pass(all warps)  sum->LM start  LM step  LM end->next pass   (us)")
for i, m in enumerate(t):
    nxt = t[i + 1][0] if i + 1 < len(t) else m[3]
    print(f"  {i:3d}  {(m[1]-m[0])/1e3:8.2f} {(m[2]-m[1])/1e3:8.2f} {(m[3]-m[2])/1e3:8.2f} {(nxt-m[3])/1e3:8.2f}")
tot = (t[-1][3] - t[0][0]) / 1e3
print(f"  subset total {tot:.1f} us: pass {sum(m[1]-m[0] for m in t)/1e3:.1f}, sum->LM {sum(m[2]-m[1] for m in t)/1e3:.1f}, LM step {sum(m[3]-m[2] for m in t)/1e3:.1f}, LM end->next pass {sum(t[i+1][0]-t[i][3] for i in range(len(t)-1))/1e3:.1f}")
