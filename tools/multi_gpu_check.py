#!/usr/bin/env python
"""torchrun --nproc-per-node G tools/multi_gpu_check.py : the two multi-GPU forms of the hot path on G GPUs of one box,
asserted against the single-GPU result and the oracle (tests/test_gpu_multi.py drives it when >= 2 GPUs are visible).

  1. independent subsets sharded over the ranks in whole rows of subsets (no collective on the data path): the gathered
     records must equal rank 0's own single-GPU batch of all subsets BIT FOR BIT, and a sample must match the oracle;
  2. one domain row-split over the ranks with the in-kernel NVLink all-reduce of the normal equations: every rank must
     hold bitwise the same answer, with the LM path of the single-GPU solve, within the BASELINE tolerances of it and
     of the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench, oracle
from correlation_b200 import engine, synth, rowsplit, sharding

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
TOL = bench.TOLERANCES

# ---------------------------------------------------------------- 1. sharded subsets
size, n_side = 1024, 8
truth = (1.25, -0.75, .002, -.0015, .001, .0025)
c = size / 2.0
und_t = synth.make_image(size, size, 4, None, (c, c), device=dev)
dfm_t = synth.make_image(size, size, 4, truth, (c, c), device=dev)
boxes = bench.subset_boxes(32, size - 32, n_side)
eng = engine.CudaEngine(lr)
eng.resetImagePyramidsDevice(und_t.data_ptr(), dfm_t.data_ptr(), None, size, size, size, pyramid=(0, 1, 2))
for cluster in (1, 2):
    eng.set_cluster_mode(cluster)
    ids = sharding.shard_grid_rows(n_side, n_side, world, rank)
    assert eng.resetPolygonRectGrid(0, np.array([boxes[i] for i in ids], np.int32)) == 0
    _, mine = eng.correlate_batch_raw(0, np.zeros((len(ids), 6), np.float32))
    item = mine.dtype.itemsize
    buf = torch.zeros((len(boxes) // world + 8) * item, dtype=torch.uint8, device=dev)
    raw = torch.from_numpy(np.frombuffer(mine.tobytes(), np.uint8).copy()).to(dev)
    buf[: raw.numel()] = raw
    out = [torch.zeros_like(buf) for _ in range(world)]
    if world > 1:
        dist.all_gather(out, buf)
    else:
        out = [buf]
    if rank == 0:
        full = np.zeros(len(boxes), mine.dtype)
        for r in range(world):
            rid = sharding.shard_grid_rows(n_side, n_side, world, r)
            full[rid] = np.frombuffer(out[r].cpu().numpy().tobytes()[: len(rid) * item], mine.dtype)
        assert eng.resetPolygonRectGrid(100, np.array(boxes, np.int32)) == 0
        _, one = eng.correlate_batch_raw(100, np.zeros((len(boxes), 6), np.float32))
        assert (one["errorCode"] == 0).all()
        assert one.tobytes() == full.tobytes(), f"sharded records differ from the single-GPU batch (cluster mode {cluster})"
        o = oracle.OracleEngine(n_threads=4, pyramid=(0, 1, 2), accum_double=True, real_threads=True)
        o.set_image("und", und_t.cpu().numpy()); o.set_image("def", dfm_t.cpu().numpy())
        for k in (0, 27, 63):
            bx = boxes[k]
            want = o.correlate(np.zeros(6), oracle.rect_points(*bx), center=((bx[0] + bx[2]) / 2, (bx[1] + bx[3]) / 2))
            d = np.abs(full["resultingParameters"][k, :6] - want["params"])
            assert d[:2].max() < TOL["duv"] and d[2:].max() < TOL["dgrad"], (k, d)
            assert abs(int(full["iterations"][k]) - want["iterations"]) <= 1
        print(f"sharded subsets, {world} rank(s), {cluster} CTA(s) per subset: {len(boxes)} records bitwise equal to the single-GPU batch; sample within tolerance of the oracle", flush=True)
eng.set_cluster_mode(0)
eng.close()

# ---------------------------------------------------------------- 2. row-split
size = 2048
truth = (10.0, -7.5, 0.004, -0.003, 0.002, 0.005)
kw = dict(spectrum=(5.0, 600.0), n_waves=64)
c = size / 2.0
und_t = synth.make_image(size, size, 5, None, (c, c), device=dev, **kw)
dfm_t = synth.make_image(size, size, 5, truth, (c, c), device=dev, **kw)
m = size // 32
x0, y0, x1, y1 = m, m, size - m, size - m
pyr = (0, 1, 3)
eng = engine.CudaEngine(lr)
eng.resetImagePyramidsDevice(und_t.data_ptr(), dfm_t.data_ptr(), None, size, size, size, pyramid=pyr)
rowsplit.connect(eng, dist if world > 1 else None)
b0, b1 = rowsplit.equal_row_bands(y0, y1, world)[rank]
eng.resetPolygonRectBand(0, x0, y0, x1, y1, b0, b1)
for _ in range(3):  # several solves in a row: the exchange's sequence numbers carry over from launch to launch
    if world > 1:
        dist.barrier()
    r = eng.correlate(0, np.zeros(6, np.float32))
assert r["error_code"] == 0, r
mine = torch.tensor(np.concatenate([r["params"], [r["chi"]]]).astype(np.float32), device=dev)
if world > 1:
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    assert all(torch.equal(o_.view(torch.int32), out[0].view(torch.int32)) for o_ in out), "ranks disagree bitwise"
if rank == 0:
    eng.rowsplit_disconnect()
    eng.resetPolygon(1, x0, y0, x1, y1)
    one = eng.correlate(1, np.zeros(6, np.float32))
    assert r["evaluations"] == one["evaluations"], (r["evaluations"], one["evaluations"])
    assert r["number_of_points"] == one["number_of_points"]
    o = oracle.OracleEngine(n_threads=8, pyramid=pyr, accum_double=True, real_threads=True)
    o.set_image("und", und_t.cpu().numpy()); o.set_image("def", dfm_t.cpu().numpy())
    want = o.correlate(np.zeros(6), oracle.rect_points(x0, y0, x1, y1), center=(c, c))
    for name, ref in (("single GPU", one), ("oracle", want)):
        d = np.abs(r["params"] - ref["params"])
        assert d[:2].max() < TOL["duv"] and d[2:].max() < TOL["dgrad"], (name, d)
        assert abs(r["chi"] - ref["chi"]) <= TOL["rel_dchi"] * ref["chi"], (name, r["chi"], ref["chi"])
        assert abs(r["iterations"] - ref["iterations"]) <= 1
    print(f"row-split over {world} rank(s): bitwise identical on all ranks, LM path of the single-GPU solve, within tolerance of it and of the oracle "
          f"(rel dchi vs oracle {abs(r['chi'] - want['chi']) / want['chi']:.1e})", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
if rank == 0:
    print("MULTI_GPU_CHECK OK", flush=True)
