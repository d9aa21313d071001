"""GPU box (debug): why does a 2048-car shard differ from its slice of a 6144-car fleet?  graphs on/off, ticks 1..40"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ft_grandprix_b200 as ft
from conftest import random_poses
t = ft.Track.bundled("track")
n = 6144
poses = random_poses(t.path, n, seed=9, level=True)
xy = poses[:, :2]; yaw = 2 * np.arctan2(poses[:, 6], poses[:, 3])
lib = ft._lib.load()
for graphs in (0, 1):
    lib.ftgp_tick_use_graphs(graphs)
    for ticks in (1, 2, 5, 40):
        whole = ft.Fleet(t, n); whole.reset(xy, yaw); whole.tick(ticks); whole.sync()
        part = ft.Fleet(t, 2048); part.reset(xy[:2048], yaw[:2048]); part.tick(ticks); part.sync()
        eager = ft.Fleet(t, 2048); eager.reset(xy[:2048], yaw[:2048])
        for _ in range(ticks):
            eager.lap_update(); eager.drive(); eager.lidar(); eager.step(1)
        eager.sync()
        out = []
        for name in ("qpos", "qvel", "warm", "ctrl", "ranges", "lap", "status"):
            a = getattr(part, name).double().reshape(2048, -1); b = getattr(whole, name)[:2048].double().reshape(2048, -1); c = getattr(eager, name).double().reshape(2048, -1)
            out.append(f"{name}: part-whole {int((a != b).any(1).sum())} cars max {float((a - b).abs().max()):.2e}; part-eager {int((a != c).any(1).sum())} cars")
        print(f"graphs={graphs} ticks={ticks}: " + " | ".join(out), flush=True)
        bad = (part.qpos != whole.qpos[:2048]).any(1).nonzero().flatten()[:5].tolist()
        if bad:
            print("   first differing cars", bad, "status part", part.status[bad].tolist(), "whole", whole.status[bad].tolist())
lib.ftgp_tick_use_graphs(1)
