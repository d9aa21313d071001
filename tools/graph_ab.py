#!/usr/bin/env python
"""GPU box: tick latency of small fleets with and without the CUDA-graph replay.
Prints one JSON line per fleet size: ms per tick eager vs graph (CUDA events around 2 000 ticks, after 300 warm ticks)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ft_grandprix_b200 as ft        # noqa: E402
from bench import make_poses          # noqa: E402

track = ft.Track.bundled("track")
for n in (1, 256, 4096, 16384):
    xy, yaw, _ = make_poses(track.path, n, seed=1, level=True)
    res = {"cars": n}
    for mode in ("eager", "graph"):
        fleet = ft.Fleet(track, n, driver="nidc")
        fleet.lib.ftgp_tick_use_graphs(1 if mode == "graph" else 0)
        if n == 1:
            fleet.reset_grid()
        else:
            fleet.reset(xy, yaw)
        fleet.tick(300); fleet.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(fleet.stream):
            e0.record(fleet.stream)
            for _ in range(20):
                fleet.tick(100)
            e1.record(fleet.stream)
        fleet.sync()
        res[mode + "_ms_per_tick"] = e0.elapsed_time(e1) / 2000
        res[mode + "_state_sum"] = float(fleet.qpos.sum())
        fleet.close()
    fleet.lib.ftgp_tick_use_graphs(1)
    res["bit_identical"] = res["eager_state_sum"] == res["graph_state_sum"]
    res["speedup"] = res["eager_ms_per_tick"] / res["graph_ms_per_tick"]
    print(json.dumps(res), flush=True)
