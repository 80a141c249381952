"""CPU suite: the N > 1 host logic (partitioning + result gather) under gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from correlation_b200 import engine, sharding


def test_shard_range_is_a_partition_and_matches_owner_rule():
    for n in (1, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(n, int)
            for r in range(world):
                b, e = sharding.shard_range(n, world, r)
                seen[b:e] += 1
                for u in range(b, e):
                    assert sharding.owner_of(u, n, world) == r
            assert (seen == 1).all()
            sizes = [sharding.shard_range(n, world, r)[1] - sharding.shard_range(n, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_shard_grid_rows_partitions_into_row_bands():
    for n_h, n_v in ((64, 64), (3, 5), (1, 7)):
        for world in (1, 2, 4, 8):
            seen = np.zeros(n_h * n_v, int)
            for r in range(world):
                ids = sharding.shard_grid_rows(n_h, n_v, world, r)
                seen[ids] += 1
                ivs = sorted({i % n_v for i in ids})
                assert ivs == list(range(ivs[0], ivs[-1] + 1)) if ivs else True  # one contiguous band of rows
                assert all(sum(1 for i in ids if i % n_v == iv) == n_h for iv in ivs)  # whole rows of sectors
            assert (seen == 1).all()


def test_band_rows_balances_pixels():
    rows = np.zeros(1000)
    yy = np.arange(1000) - 500.0
    rows[:] = 2 * np.sqrt(np.maximum(0, 450.0**2 - yy**2))  # a disc
    for world in (2, 4, 8):
        bands = sharding.band_rows(rows, world)
        assert bands[0][0] == 0 and bands[-1][1] == 1000
        assert all(bands[k][1] == bands[k + 1][0] for k in range(world - 1))
        px = np.array([rows[b:e].sum() for b, e in bands])
        assert px.max() / px.mean() < 1.02


def _worker(rank, world, port, n_units, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = sharding.shard_range(n_units, world, rank)
    local = np.zeros(e - b, engine.RESULT_DTYPE)
    # a stand-in for correlate_batch on this rank's sectors: result depends only on the sector id
    ids = np.arange(b, e)
    local["resultingParameters"][:, 0] = ids * 0.5
    local["chi"] = ids + 0.25
    local["numberOfPoints"] = ids
    local["evaluationsPerLevel"][:, 0] = 3
    local["pointsPerLevel"][:, 0] = 100 + ids
    full = sharding.gather_results(local, n_units, dist)
    if rank == 0:
        q.put((full["chi"].tolist(), full["numberOfPoints"].tolist(), engine.CudaEngine.pixel_evaluations(full)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_units", [5, 64])
def test_gather_results_world_size_2_gloo(n_units):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_units, q)) for r in range(2)]
    for p in procs:
        p.start()
    chi, npts, work = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert chi == [i + 0.25 for i in range(n_units)]
    assert npts == list(range(n_units))
    assert work == 3.0 * sum(100 + i for i in range(n_units))
