#!/usr/bin/env python
"""GPU box: how the tick time depends on how long the fleet has been driving.

BASELINE config 3 fleet (65,536 cars on track.png, seed 1, batched nidc): at 200 / 1 000 / 5 000 / 25 000 ticks after the
reset, time 50 ticks (CUDA events, L2 flushed between ticks) and read the regime: mean Newton iterations, fraction of
cars in wall contact, off track, reset by the bad-state check, mean speed.  One JSON line per checkpoint.

    python tools/regime_sweep.py [--cars 65536] [--checkpoints 200,1000,5000,25000] > profiles/bench/regime_sweep.json
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ft_grandprix_b200 as ft        # noqa: E402
from bench import make_poses          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cars", type=int, default=65536)
    ap.add_argument("--checkpoints", default="200,1000,5000,25000")
    ap.add_argument("--timed", type=int, default=50)
    a = ap.parse_args()
    track = ft.Track.bundled("track")
    fleet = ft.Fleet(track, a.cars, driver="nidc")
    xy, yaw, _ = make_poses(track.path, a.cars, seed=1, level=True)
    fleet.reset(xy, yaw)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=fleet.device)
    L = ft.fleet.LAP
    done = 0
    for cp in [int(x) for x in a.checkpoints.split(",")]:
        fleet.tick(cp - done); done = cp
        fleet.sync()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.timed)]
        iters = torch.zeros(a.cars, device=fleet.device); wall = torch.zeros(a.cars, device=fleet.device)
        with torch.cuda.stream(fleet.stream):
            for e0, e1 in ev:
                flush.fill_(1)
                e0.record(fleet.stream)
                fleet.tick(1)
                e1.record(fleet.stream)
                iters += (fleet.status & 0xFF).float(); wall += (((fleet.status >> 16) & 0xFF) > 0).float()
        fleet.sync(); done += a.timed
        ms = [e0.elapsed_time(e1) for e0, e1 in ev]
        st = fleet.status.cpu().numpy(); lap = fleet.lap.cpu().numpy()
        speed = fleet.qvel[:, :2].norm(dim=1)
        print(json.dumps({"ticks_since_reset": cp, "cars": a.cars, "tick_ms_mean": float(np.mean(ms)), "tick_ms_min": float(np.min(ms)),
                          "tick_ms_max": float(np.max(ms)), "newton_iterations_mean": float(iters.mean() / a.timed),
                          "newton_iterations_max_last": int((st & 0xFF).max()),
                          "wall_contact_fraction": float(wall.mean() / a.timed), "off_track_fraction": float(lap[:, L["off_track"]].mean()),
                          "reset_fraction_last": float(((st >> 8) & 1).mean()), "speed_mean_mps": float(speed.mean()),
                          "speed_below_0p05_fraction": float((speed < 0.05).float().mean()), "laps_max": int(lap[:, L["laps"]].max()),
                          "finished_fraction": float(lap[:, L["finished"]].mean())}), flush=True)


if __name__ == "__main__":
    main()
