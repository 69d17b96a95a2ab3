"""Multi-GPU host logic on CPU: contiguous (elevation, sector) sharding and the single gather of
the product volume, exercised with world_size 2 and 3 over gloo."""
import os
import socket

import numpy as np
import pytest


def test_shard_bounds_cover_the_volume(wrp):
    from importlib import import_module
    vol = import_module("weather-radar-processing_b200.volume")
    for units, world in [(1287, 8), (1287, 4), (1287, 2), (1287, 1), (5, 8), (0, 3), (64, 8)]:
        spans = [vol.shard_bounds(units, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == units
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1 and max(sizes) == vol.max_shard(units, world)
    assert vol.shard_bounds(1287, 3, 8) == (483, 644)  # ceil(1287*3/8), ceil(1287*4/8)
    assert vol.unit_to_ids(143 * 4 + 17, 143) == (17, 4)
    with pytest.raises(ValueError):
        vol.shard_bounds(10, 2, 2)


def _worker(rank, world, port, units, q):
    import importlib
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vol = importlib.import_module("weather-radar-processing_b200.volume")
    lo, hi = vol.shard_bounds(units, rank, world)
    # stand-in products: gate g of unit k holds (k + g/1000, -k)
    k = torch.arange(lo, hi, dtype=torch.float32)[:, None]
    g = torch.arange(16, dtype=torch.float32)[None, :]
    local = torch.stack([k + g / 1000, -k.expand(-1, 16)], dim=2)
    full = vol.gather_volume(local, units)
    q.put((rank, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,units", [(2, 7), (3, 10), (2, 2)])
def test_gather_volume_gloo(wrp, world, units):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, units, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    k = np.arange(units, dtype=np.float32)[:, None]
    g = np.arange(16, dtype=np.float32)[None, :]
    want = np.stack([k + g / 1000, np.broadcast_to(-k, (units, 16))], axis=2)
    for r in range(world):
        assert got[r].shape == (units, 16, 2) and np.array_equal(got[r], want)
