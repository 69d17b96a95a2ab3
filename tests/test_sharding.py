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


class _FakeChain:
    """Stands in for RadarChain in the CPU test of the volume driver: 'products' of a unit are
    derived from the first bytes of its input record, so misplaced units are detected."""
    M = 8

    def process_host(self, host_iq, n, out):
        rec = np.asarray(host_iq).reshape(n, -1)
        for i in range(n):
            out[i, :, 0] = rec[i, 0] + np.arange(4) / 10
            out[i, :, 1] = -float(rec[i, 1])
        return out

    def process_host_to_device(self, host_iq, n, dev_out_ptr):
        # the "device" of the CPU test is host memory: write the products where the tensor lives
        import ctypes
        out = np.ctypeslib.as_array((ctypes.c_float * (n * 4 * 2)).from_address(dev_out_ptr)).reshape(n, 4, 2)
        self.process_host(host_iq, n, out)


def _volume_worker(rank, world, port, units, q):
    import importlib
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vol = importlib.import_module("weather-radar-processing_b200.volume")
    lo, hi = vol.shard_bounds(units, rank, world)
    shard = np.zeros((hi - lo, 16), np.uint8)
    shard[:, 0] = np.arange(lo, hi)          # unit index
    shard[:, 1] = np.arange(lo, hi) // 5     # "elevation" with 5 sectors per elevation
    full = vol.process_volume(_FakeChain(), shard, units, torch.device("cpu"))
    q.put((rank, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,units", [(2, 15), (3, 15)])
def test_process_volume_gloo(wrp, world, units):
    """The volume driver (shard -> chain -> one gather) with world_size 2 and 3: every rank ends with
    the whole volume in (elevation, sector) order."""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_volume_worker, args=(r, world, port, units, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    k = np.arange(units, dtype=np.float32)
    for r in range(world):
        v = got[r]
        assert v.shape == (units, 4, 2)
        assert np.allclose(v[:, :, 0], k[:, None] + np.arange(4) / 10)
        assert np.array_equal(v[:, 0, 1], -(k // 5))
        flat = importlib_volume().as_sitdim(v, 5, units // 5)
        assert flat.shape == (units * 8,)


def importlib_volume():
    from importlib import import_module
    return import_module("weather-radar-processing_b200.volume")
