#!/usr/bin/env python
"""GPU box: time the fused tick (ftgp_tick) of every variant library under variants/ and of the shipped one: 65,536 cars,
300 settle ticks, 50 timed ticks one by one with the L2 flushed in between, CUDA events; final state sums must agree."""
import ctypes as C, glob, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ft_grandprix_b200 as ft
from ft_grandprix_b200 import _lib
from bench import make_poses
track = ft.Track.bundled("track")
sizes = [int(x) for x in sys.argv[1:]] or [65536]
libs = [("shipped", _lib.LIB_PATH)] + [(os.path.basename(p)[8:-3], p) for p in sorted(glob.glob(os.path.join(ROOT, "variants", "libftgp_*.so")))]
flush = None
for n, (name, path) in [(n, l) for n in sizes for l in libs]:
    xy, yaw, _ = make_poses(track.path, n, seed=1, level=True)
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
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(50)]
    with torch.cuda.stream(fleet.stream):
        for e in ev:
            flush.fill_(1)
            e[0].record(fleet.stream); fleet.tick(1); e[1].record(fleet.stream)
    fleet.sync()
    print(json.dumps({"variant": name, "cars": n, "tick_ms": float(np.mean([e[0].elapsed_time(e[1]) for e in ev])),
                      "state_sum": float(fleet.qpos.sum()), "ranges_sum": float(fleet.ranges.double().sum())}), flush=True)
    fleet.close(); del fleet
