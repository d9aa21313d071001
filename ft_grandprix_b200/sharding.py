"""Env sharding: worlds are independent, so a fleet is cut into contiguous blocks of worlds, one block
per rank (one process per GPU), with NO collective on the step path (SURVEY §8 e).  Only per-episode
results -- laps, finishing rank, lap times, off-track / wall-contact ticks -- are gathered, once, over
torch.distributed (NCCL on the GPU box, gloo in the CPU tests).

The reference has no counterpart (it steps one MjModel on one thread, ft_grandprix/custom.py:1425);
what is gathered is what its dashboard shows per car (custom.py:335-361) plus `winners` (custom.py:1368).
"""
import numpy as np
import torch
import torch.distributed as dist

from ._lib import LAP_FIELDS, MAX_LAPTIMES

STAT_FIELDS = ("laps", "rank", "finished", "ntimes", "offtrack_ticks", "contact_ticks", "completion")


def shard_range(nworlds, world_size, rank):
    """World w lives on rank w * world_size // nworlds: contiguous blocks, sizes differing by at most one."""
    lo = (nworlds * rank + world_size - 1) // world_size
    hi = (nworlds * (rank + 1) + world_size - 1) // world_size
    return lo, hi


def owner_of(world, nworlds, world_size):
    return world * world_size // nworlds


def track_assignment(nworlds, ntracks):
    """BASELINE config 4: tracks alternate by GLOBAL world index, so a shard's mix does not depend on the cut."""
    return np.arange(nworlds, dtype=np.int32) % ntracks


def pack_stats(lap, times):
    """[ncars, len(STAT_FIELDS) + MAX_LAPTIMES] int32 from a shard's lap state tensors."""
    idx = [LAP_FIELDS.index(f) for f in STAT_FIELDS]
    return torch.cat([lap[:, idx], times[:, :MAX_LAPTIMES]], dim=1).contiguous()


def gather_stats(local, ncars_global, cars_per_world=1, group=None):
    """All ranks call this once per episode; returns the global [ncars_global, F] table on every rank.
    Shards may differ by one world, so they are padded to the largest before the all_gather."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    nworlds = ncars_global // cars_per_world
    sizes = [(shard_range(nworlds, ws, r)[1] - shard_range(nworlds, ws, r)[0]) * cars_per_world for r in range(ws)]
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    pad = torch.zeros(max(sizes), local.shape[1], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)


class ShardedRace:
    """One rank's block of a global fleet.  `make_fleet(ncars_local, track_id_local, first_world)` builds the
    local fleet (a ft_grandprix_b200.Fleet on the GPU box; any object with lap/times tensors in tests)."""

    def __init__(self, nworlds, cars_per_world, ntracks, make_fleet, group=None):
        self.group = group
        self.ws = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nworlds, self.cpw = nworlds, cars_per_world
        self.lo, self.hi = shard_range(nworlds, self.ws, self.rank)
        tid_world = track_assignment(nworlds, ntracks)[self.lo:self.hi]
        self.track_id = np.repeat(tid_world, cars_per_world)
        self.fleet = make_fleet((self.hi - self.lo) * cars_per_world, self.track_id, self.lo)

    def episode_stats(self):
        local = pack_stats(self.fleet.lap, self.fleet.times)
        return gather_stats(local, self.nworlds * self.cpw, self.cpw, self.group)
