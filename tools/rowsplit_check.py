#!/usr/bin/env python
"""torchrun --nproc-per-node G tools/rowsplit_check.py [rows]: one rectangle row-split over G GPUs
vs the same rectangle on one GPU. Prints parity and timing on rank 0."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from correlation_b200 import engine, synth, rowsplit

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = engine.MODE_FAST if (len(sys.argv) > 2 and sys.argv[2] == "fast") else engine.MODE_PARITY
pyr = (0, 1, 4)
truth = (10.0, -7.5, 0.004, -0.003, 0.002, 0.005)
dev = torch.device("cuda", lr)
c = size / 2.0
kw = dict(spectrum=(5.0, 600.0), n_waves=64)
und = synth.make_image(size, size, 5, None, (c, c), device=dev, **kw)
dfm = synth.make_image(size, size, 5, truth, (c, c), device=dev, **kw)
m = size // 32
x0, y0, x1, y1 = m, m, size - m, size - m
eng = engine.CudaEngine(lr, arith_mode=mode)
eng.resetImagePyramidsDevice(und.data_ptr(), dfm.data_ptr(), None, size, size, size, pyramid=pyr)
rowsplit.connect(eng, dist if world > 1 else None)
b0, b1 = rowsplit.equal_row_bands(y0, y1, world)[rank]
eng.resetPolygonRectBand(0, x0, y0, x1, y1, b0, b1)
for _ in range(2):
    if world > 1: dist.barrier()
    r = eng.correlate(0, np.zeros(6, np.float32))
if world > 1: dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
r = eng.correlate(0, np.zeros(6, np.float32))
wall = time.perf_counter() - t0
ms = eng.last_correlate_ms()
# every rank must hold bitwise the same answer
mine = torch.tensor(np.concatenate([r["params"], [r["chi"]]]).astype(np.float32), device=dev)
same = True
if world > 1:
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    same = all(torch.equal(o.view(torch.int32), out[0].view(torch.int32)) for o in out)
if rank == 0:
    eng.rowsplit_disconnect()
    eng.resetPolygon(1, x0, y0, x1, y1)
    for _ in range(2):
        full = eng.correlate(1, np.zeros(6, np.float32))
    ms1 = eng.last_correlate_ms()
    d = np.abs(r["params"] - full["params"])
    pe = full["pixel_evaluations"]
    print(f"rowsplit {size}^2 world={world} mode={'fast' if mode else 'parity'}: err={r['error_code']} ranks bitwise identical={same} "
          f"evals {r['evaluations'][:5]} vs single {full['evaluations'][:5]}; duv {d[:2].max():.2e} dgrad {d[2:].max():.2e} "
          f"rel chi {abs(r['chi']-full['chi'])/full['chi']:.2e}; params {r['params']}")
    print(f"  kernel {ms:.3f} ms (wall {wall*1e3:.3f}) on {world} GPU(s) vs {ms1:.3f} ms on one: speed-up {ms1/ms:.2f}x; "
          f"{pe/ms/1e6:.1f} vs {pe/ms1/1e6:.1f} Gpx*ev/s")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
