"""Sharding of a volume scan across GPUs and assembly of the product volume.

Every (elevation, sector) unit is independent — ``advance()`` only bumps counters
(rpv2.cu:572-579) — so the path is partitioned with no data-path collective: unit
``k = e*S + s`` goes to the rank owning the contiguous block ``[ceil(U*g/G), ceil(U*(g+1)/G))``,
which keeps every rank's slice contiguous in the reference's ``sitdim`` order
``result[x + 2*gate + sector*M + elev*M*S]`` (rpv2.cu:735-736, 607).  The only exchange is the
gather of the finished products, done with ``torch.distributed`` (NCCL on GPUs, gloo in the CPU
tests): padded equal shards through one all_gather.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of units owned by ``rank``: [lo, hi)."""
    if world < 1 or not (0 <= rank < world) or n_units < 0:
        raise ValueError("bad shard arguments")
    lo = -((-n_units * rank) // world)
    hi = -((-n_units * (rank + 1)) // world)
    return lo, hi


def unit_to_ids(k: int, n_sectors: int) -> tuple[int, int]:
    """Flat unit index -> (sector, elevation)."""
    return k % n_sectors, k // n_sectors


def max_shard(n_units: int, world: int) -> int:
    return max(shard_bounds(n_units, r, world)[1] - shard_bounds(n_units, r, world)[0] for r in range(world))


def gather_volume(local, n_units: int, group=None):
    """All-gather each rank's ``[n_local, gates, 2]`` products into the full
    ``[n_units, gates, 2]`` volume (a torch tensor on ``local``'s device).  One collective."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    pad = max_shard(n_units, world)
    gates = local.shape[1]
    buf = torch.zeros((pad, gates, 2), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty((world * pad, gates, 2), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_units, r, world)
        parts.append(out[r * pad: r * pad + (hi - lo)])
    return torch.cat(parts, dim=0)


def process_shard(chain, shard_input, n_local: int, out=None):
    """Run a rank's contiguous block of units (already in ``chain``'s input format, host memory)
    through ``wrp_process_host``; returns float32 ``[n_local, gates, 2]``."""
    if out is None:
        out = np.empty((n_local, chain.M // 2, 2), np.float32)
    if n_local:
        chain.process_host(shard_input, n_local, out)
    return out


def process_volume(chain, shard_input, n_units: int, device, group=None):
    """One volume scan on this rank's shard + the gather: returns the ``[n_units, gates, 2]`` product
    volume as a torch tensor on ``device`` (every rank gets it).  ``shard_input`` holds the units
    ``shard_bounds(n_units, rank, world)`` of this rank."""
    import torch
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_bounds(n_units, rank, world)
    # the products never visit the host: H2D of the input, chain, and the gather all on the device
    local = torch.empty((hi - lo, chain.M // 2, 2), dtype=torch.float32, device=device)
    if hi > lo:
        chain.process_host_to_device(shard_input, hi - lo, local.data_ptr())
    return local if world == 1 else gather_volume(local, n_units, group)


def as_sitdim(volume: np.ndarray, n_sectors: int, n_elevations: int) -> np.ndarray:
    """[E*S, gates, 2] -> the reference's flat ``result`` array (sitdim order, rpv2.cu:736)."""
    v = np.asarray(volume)
    assert v.shape[0] == n_sectors * n_elevations
    return np.ascontiguousarray(v).reshape(-1)
