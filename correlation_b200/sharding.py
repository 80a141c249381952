"""Multi-GPU partitioning of the hot path. One process per GPU.

Two ways the path shards:
  * independent units -- subdivision subsets of a rectangle, annular sectors, separate polygon
    domains (the reference loops over them serially, manager_class.cpp:304-547): contiguous
    blocks of sector ids per rank (shard_range) or, for a grid of subsets, whole rows of subsets
    per rank (shard_grid_rows: the rank then needs only a band of image rows), NO collective on
    the data path; one gather of the (<= 176 B) result records at the end;
  * one huge domain -- pixel rows split into bands of equal pixel count; every evaluation ends
    with an all-reduce(sum) of the (n^2 + n)/2 + n + 2 normal-equation floats and every rank runs
    the identical LM state machine (see correlation_b200/rowsplit.py).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_units: int, world: int, rank: int) -> tuple[int, int]:
    """[begin, end) of the units rank owns: unit i belongs to rank floor(i * world / n_units)."""
    if n_units <= 0:
        return 0, 0
    begin = -(-rank * n_units // world)          # ceil(rank * n / world)
    end = -(-(rank + 1) * n_units // world)
    return min(begin, n_units), min(end, n_units)


def owner_of(unit: int, n_units: int, world: int) -> int:
    return unit * world // n_units


def shard_grid_rows(n_h: int, n_v: int, world: int, rank: int) -> list[int]:
    """Sector ids (reference order: id = ih * n_v + iv, horizontal index outer, manager_class.cpp:304-310)
    of the subdivisions whose VERTICAL index lies in rank's block: every rank then owns a band of image
    rows, so it needs only that band of the images (dic_stage_next_pair_rows) -- the PCIe bytes per rank
    fall with the number of GPUs instead of being replicated. Still no collective on the data path."""
    vb, ve = shard_range(n_v, world, rank)
    return [ih * n_v + iv for iv in range(vb, ve) for ih in range(n_h)]


def gather_results(local: np.ndarray, n_units: int, dist=None, dst: int = 0):
    """Gather per-rank structured result arrays (RESULT_DTYPE) into one array of n_units on `dst`.

    Works with any torch.distributed backend (gloo on CPU for tests, nccl on GPUs: the payload
    travels as a uint8 tensor on the backend's device)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    item = local.dtype.itemsize
    per = max(shard_range(n_units, world, r)[1] - shard_range(n_units, world, r)[0] for r in range(world))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = torch.zeros(per * item, dtype=torch.uint8, device=dev)
    raw = torch.from_numpy(np.frombuffer(local.tobytes(), np.uint8).copy())
    buf[: raw.numel()] = raw.to(dev)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    if rank != dst:
        return None
    full = np.zeros(n_units, local.dtype)
    for r in range(world):
        b, e = shard_range(n_units, world, r)
        full[b:e] = np.frombuffer(out[r].cpu().numpy().tobytes()[: (e - b) * item], local.dtype)
    return full


def band_rows(row_counts: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Split rows 0..len-1 into `world` contiguous bands of (nearly) equal pixel count.

    row_counts[r] = number of domain pixels in image row r. Returns [(r_begin, r_end)) per rank."""
    total = float(row_counts.sum())
    cum = np.concatenate([[0.0], np.cumsum(row_counts, dtype=np.float64)])
    cuts = [0]
    for k in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * k / world, side="left")))
    cuts.append(len(row_counts))
    cuts = np.maximum.accumulate(np.array(cuts))
    return [(int(cuts[k]), int(cuts[k + 1])) for k in range(world)]
