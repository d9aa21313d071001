#!/usr/bin/env python
"""GPU box: where a race tick goes -- 8-car worlds from the start grid, lidar and step timed separately (un-fused calls)."""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ft_grandprix_b200 as ft
nworlds = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cpw = 8
track = ft.Track.bundled("track")
fleet = ft.Fleet(track, nworlds * cpw, cars_per_world=cpw, driver="nidc")
n = fleet.ncars
grid = np.array([track.start_pose(c) for c in range(cpw)])
rng = np.random.default_rng(3)
xy = np.tile(grid[:, :2], (nworlds, 1)) + rng.normal(0, 0.01, (n, 2))
yaw = np.tile(grid[:, 2], nworlds) + rng.normal(0, 0.02, n)
fleet.set_driver_kinds(["nidc" if c % 2 == 0 else "fast" for c in range(cpw)] * nworlds)
fleet.reset(xy, yaw)
for settle in (100, 400, 300):
    fleet.tick(settle); fleet.sync()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(20)]
    with torch.cuda.stream(fleet.stream):
        for e in ev:
            e[0].record(fleet.stream); fleet.lap_update(); fleet.drive(); e[1].record(fleet.stream)
            fleet.lidar(); e[2].record(fleet.stream); fleet.step(1); e[3].record(fleet.stream)
    fleet.sync()
    st = fleet.status.cpu().numpy()
    print(json.dumps({"ticks": fleet.steps, "cars": n, "lap_drive_ms": float(np.mean([e[0].elapsed_time(e[1]) for e in ev])),
                      "lidar_ms": float(np.mean([e[1].elapsed_time(e[2]) for e in ev])), "step_ms": float(np.mean([e[2].elapsed_time(e[3]) for e in ev])),
                      "coupled_cars": int(((st >> 9) & 1).sum()), "near_wall": int(((st >> 10) & 1).sum()), "iters_mean": float((st & 0xFF).mean())}), flush=True)
# the same cars as worlds of one car each
f1 = ft.Fleet(track, n, driver="nidc")
f1.qpos.copy_(fleet.qpos); f1.qvel.copy_(fleet.qvel); f1.warm.copy_(fleet.warm); f1.ctrl.copy_(fleet.ctrl); torch.cuda.synchronize()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(20)]
with torch.cuda.stream(f1.stream):
    for e in ev:
        e[0].record(f1.stream); f1.lidar(); e[1].record(f1.stream); f1.step(1); e[2].record(f1.stream)
f1.sync()
print(json.dumps({"single_car_worlds": True, "lidar_ms": float(np.mean([e[0].elapsed_time(e[1]) for e in ev])), "step_ms": float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))}))
