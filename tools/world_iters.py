#!/usr/bin/env python
"""GPU box: Newton iteration counts and contact counts of the cars advanced by the coupled world solver in the race workload
(8-car worlds from the start grid), per tick, plus the time of the step with and without coupled worlds."""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ft_grandprix_b200 as ft
nworlds = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
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
out = []
for t in range(0, 400, 20):
    fleet.tick(19); fleet.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(fleet.stream):
        e0.record(fleet.stream); fleet.tick(1); e1.record(fleet.stream)
    fleet.sync()
    st = fleet.status.cpu().numpy()
    coupled = (st & 0x200) != 0
    it = st & 0xFF
    row = {"tick": t + 20, "tick_ms": e0.elapsed_time(e1), "coupled_cars": int(coupled.sum()),
           "iters_coupled_mean": float(it[coupled].mean()) if coupled.any() else None, "iters_coupled_max": int(it[coupled].max()) if coupled.any() else None,
           "iters_coupled_hist": np.bincount(np.minimum(it[coupled], 120) // 10, minlength=13).tolist() if coupled.any() else None,
           "iters_fast_mean": float(it[~coupled].mean())}
    print(json.dumps(row), flush=True)
