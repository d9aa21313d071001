#!/usr/bin/env python
"""GPU box: time lidar_kernel of every variant library under variants/ (tools/build_variant.py) and of the shipped one on
the same settled fleet state (65,536 cars, 300 ticks of driving), L2 flushed between launches, and check that every
variant returns the shipped ranges bit for bit (scheduling knobs must not change a single ray)."""
import ctypes as C, glob, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ft_grandprix_b200 as ft
from ft_grandprix_b200 import _lib
from bench import make_poses
track = ft.Track.bundled("track")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
xy, yaw, _ = make_poses(track.path, n, seed=1, level=True)
libs = [("shipped", _lib.LIB_PATH)] + [(os.path.basename(p)[8:-3], p) for p in sorted(glob.glob(os.path.join(ROOT, "variants", "libftgp_*.so")))]
flush = None; state = None; ref = None
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
    if state is None:
        fleet.tick(300); fleet.sync(); state = fleet.qpos.clone()
    fleet.qpos.copy_(state); torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(20)]
    with torch.cuda.stream(fleet.stream):
        for _ in range(3): fleet.lidar()
        for e in ev:
            flush.fill_(1)
            e[0].record(fleet.stream); r = fleet.lidar(); e[1].record(fleet.stream)
    fleet.sync()
    got = r.clone()
    if ref is None: ref = got
    same = bool(torch.equal(got.view(torch.int32), ref.view(torch.int32)))
    print(json.dumps({"variant": name, "lidar_ms": float(np.mean([e[0].elapsed_time(e[1]) for e in ev])), "bit_identical": same}), flush=True)
    fleet.close(); del fleet
