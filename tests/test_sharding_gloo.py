"""CPU, world_size 2 over gloo: the sharding layer (partition, track mix, episode-stats gather)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_ranges_partition_every_world():
    from ft_grandprix_b200.sharding import owner_of, shard_range
    for nworlds in (1, 7, 8, 1000, 1048576, 32768):
        for ws in (1, 2, 3, 4, 8):
            cover = []
            for r in range(ws):
                lo, hi = shard_range(nworlds, ws, r)
                cover += [(lo, hi)]
                if nworlds < 5000:
                    for w in range(lo, hi):
                        assert owner_of(w, nworlds, ws) == r
            assert cover[0][0] == 0 and cover[-1][1] == nworlds
            assert all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
            sizes = [hi - lo for lo, hi in cover]
            assert max(sizes) - min(sizes) <= 1


class FakeFleet:
    def __init__(self, ncars, track_id, first_world, cpw):
        from ft_grandprix_b200._lib import LAP_FIELDS, MAX_LAPTIMES
        g = torch.arange(ncars, dtype=torch.int32) + first_world * cpw      # global car index
        self.lap = torch.zeros(ncars, len(LAP_FIELDS), dtype=torch.int32)
        self.lap[:, LAP_FIELDS.index("laps")] = g % 7
        self.lap[:, LAP_FIELDS.index("rank")] = g % 3
        self.lap[:, LAP_FIELDS.index("contact_ticks")] = torch.as_tensor(track_id, dtype=torch.int32) * 100 + 1
        self.times = (g[:, None] * 10 + torch.arange(MAX_LAPTIMES, dtype=torch.int32)[None, :]).to(torch.int32)


def _worker(rank, ws, port, nworlds, cpw, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from ft_grandprix_b200.sharding import STAT_FIELDS, ShardedRace
    race = ShardedRace(nworlds, cpw, 2, lambda n, tid, first: FakeFleet(n, tid, first, cpw))
    stats = race.episode_stats()
    q.put((rank, race.lo, race.hi, stats.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nworlds,cpw", [(11, 1), (6, 8)])
def test_episode_stats_gather_world_size_2(nworlds, cpw):
    from ft_grandprix_b200._lib import MAX_LAPTIMES
    from ft_grandprix_b200.sharding import STAT_FIELDS
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nworlds, cpw, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == nworlds
    n = nworlds * cpw
    g = np.arange(n)
    for _, _, _, stats in res:                      # every rank holds the same global table
        assert stats.shape == (n, len(STAT_FIELDS) + MAX_LAPTIMES)
        assert (stats[:, STAT_FIELDS.index("laps")] == g % 7).all()
        assert (stats[:, STAT_FIELDS.index("rank")] == g % 3).all()
        # tracks alternate by GLOBAL world index regardless of the cut
        assert (stats[:, STAT_FIELDS.index("contact_ticks")] == ((g // cpw) % 2) * 100 + 1).all()
        assert (stats[:, len(STAT_FIELDS):] == g[:, None] * 10 + np.arange(MAX_LAPTIMES)[None, :]).all()
