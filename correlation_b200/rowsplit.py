"""One huge domain split by pixel rows over the GPUs of a node (BASELINE config 5).

Each rank owns a band of rows of the domain; every evaluation ends with a sum of the
(n^2 + n)/2 + n + 2 normal-equation floats over the ranks. That sum is NOT an NCCL call: the master
CTA of each rank's persistent GN kernel writes its sums into every peer's mailbox (peer-mapped
through CUDA IPC, i.e. NVLink stores), and every rank adds the rows in rank order, so that all
ranks hold bitwise identical totals and run the same LM state machine without a broadcast
(`rowsplit_allreduce` in csrc/dic_kernels.cuh). torch.distributed is used once, to exchange the
64-byte IPC handles.
"""
from __future__ import annotations

import numpy as np


def equal_row_bands(y0: int, y1: int, world: int):
    """Rows y0..y1 (inclusive) of a rectangle in `world` contiguous bands of equal height (+-1)."""
    n = y1 - y0 + 1
    cuts = [y0 + (n * r) // world for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1] - 1) for r in range(world)]


def connect(eng, dist=None):
    """Exchange mailbox handles and wire the engine for row-split operation. Collective."""
    handle = eng.rowsplit_handle()
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        eng.rowsplit_connect(0, 1, handle.reshape(1, 64))
        return 0, 1
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(handle.copy()).to(dev)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    handles = np.stack([t.cpu().numpy() for t in out])
    eng.rowsplit_connect(rank, world, handles)
    dist.barrier()
    return rank, world
