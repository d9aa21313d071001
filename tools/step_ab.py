#!/usr/bin/env python
"""GPU box: time the tick's two heavy kernels with every variant library under variants/ (tools/build_variant.py) and with
the shipped one: 65,536 cars, 300 settle ticks, 30 timed ticks, CUDA events around lidar and step."""
import ctypes as C, glob, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ft_grandprix_b200 as ft
from ft_grandprix_b200 import _lib
from bench import make_poses
track = ft.Track.bundled("track")
n = 65536
xy, yaw, _ = make_poses(track.path, n, seed=1, level=True)
libs = [("shipped", _lib.LIB_PATH)] + [(os.path.basename(p)[8:-3], p) for p in sorted(glob.glob(os.path.join(ROOT, "variants", "libftgp_*.so")))]
flush = None
for name, path in libs:
    lib = C.CDLL(path)
    for fn, (res, args) in _lib.SIGNATURES.items():
        f = getattr(lib, fn); f.restype, f.argtypes = res, args
    _lib._lib = lib                                       # (tools only: the package itself always loads libftgp.so)
    fleet = ft.Fleet(track, n, driver="nidc")
    fleet.lib = lib
    fleet.reset(xy, yaw)
    if flush is None:
        flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=fleet.device)
    fleet.tick(300); fleet.sync()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(30)]
    with torch.cuda.stream(fleet.stream):
        for e in ev:
            flush.fill_(1)
            fleet.lap_update(); fleet.drive()
            e[0].record(fleet.stream); fleet.lidar(); e[1].record(fleet.stream)
            e[2].record(fleet.stream); fleet.step(1); e[3].record(fleet.stream)
    fleet.sync()
    print(json.dumps({"variant": name, "lidar_ms": float(np.mean([e[0].elapsed_time(e[1]) for e in ev])),
                      "step_ms": float(np.mean([e[2].elapsed_time(e[3]) for e in ev])), "state_sum": float(fleet.qpos.sum())}), flush=True)
    fleet.close(); del fleet
