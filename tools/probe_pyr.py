#!/usr/bin/env python
"""GPU probe: device time of one pyramid build (level 0 resident -> levels 1..3) of a 4096^2 image."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from correlation_b200 import engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
img = torch.randint(0, 256, (n, n), dtype=torch.uint8, device="cuda")
eng = engine.CudaEngine(0)
eng.resetImagePyramidsDevice(img.data_ptr(), img.data_ptr(), None, n, n, n, pyramid=(0, 1, 3))
def run(reps):
    for _ in range(reps):
        eng.lib.dic_reset_def_pyramid_device(eng.h, img.data_ptr(), n, n, n)
    eng.synchronize()
run(5)
t0 = time.perf_counter(); run(50); dt = (time.perf_counter() - t0) / 50
src_px = n * n * (1 + 0.25 + 0.0625)
print(f"{n}^2: D2D copy + 3 levels {dt*1e6:.1f} us per image; 1.25 B x {src_px/1e6:.1f} M source px = {1.25*src_px/dt/1e9:.0f} GB/s algorithmic (copy included)")
