"""Debug helper (GPU box): dump rays where the CUDA lidar and the oracle disagree."""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import ft_grandprix_b200 as ft
from oracle import pyoracle
from conftest import random_poses
out = {}
for name, n in (("track", 4096), ("circle", 1024), ("inkscape", 1024)):
    t = ft.Track.bundled(name)
    z = np.load("ft_grandprix_b200/assets/tracks.npz"); key = name.replace("-", "_")
    shape = tuple(int(v) for v in z[key + "__shape"])
    ot = pyoracle.Track(np.unpackbits(z[key + "__bits"])[: shape[0] * shape[1]].reshape(shape))
    poses = random_poses(t.path, n, seed=0)
    fleet = ft.Fleet(t, n)
    fleet.qpos[:, :7] = torch.from_numpy(poses).to(fleet.device); torch.cuda.synchronize()
    got = fleet.lidar(); fleet.sync(); got = got.cpu().numpy().astype(np.float64)
    want = ot.scan(poses)
    bad = np.argwhere((np.abs(got - want) > 1e-4) | ((got < 0) != (want < 0)))
    print(name, "bad rays", len(bad), "of", got.size)
    out[name] = [dict(car=int(c), beam=int(b), pose=poses[c].tolist(), got=float(got[c, b]), want=float(want[c, b])) for c, b in bad[:200]]
json.dump(out, open("gpurun_out/lidar_mismatch.json", "w"))

# ---- the off-nominal pose set of tests/test_gpu_lidar.py::test_lidar_level_and_tilted_and_far (anywhere on the map,
# inside wall cells, outside the track, large tilts): dump every disagreeing ray for offline analysis
t = ft.Track.bundled("track")
z = np.load("ft_grandprix_b200/assets/tracks.npz")
shape = tuple(int(v) for v in z["track__shape"])
ot = pyoracle.Track(np.unpackbits(z["track__bits"])[: shape[0] * shape[1]].reshape(shape))
allbad = []
for seed in (3, 4, 5, 6):
    rng = np.random.default_rng(seed)
    n = 4096
    xy = rng.uniform(-2, 42, (n, 2)) * [1, -1]
    yaw = rng.uniform(-np.pi, np.pi, n)
    tilt = rng.normal(0, 0.15, (n, 2))
    zz = rng.uniform(-0.05, 0.3, n)
    cy, sy = np.cos(yaw / 2), np.sin(yaw / 2)
    cp, sp, cr, sr = np.cos(tilt[:, 0] / 2), np.sin(tilt[:, 0] / 2), np.cos(tilt[:, 1] / 2), np.sin(tilt[:, 1] / 2)
    q = np.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], 1)
    poses = np.concatenate([xy, zz[:, None], q], 1)
    fleet = ft.Fleet(t, n)
    fleet.qpos[:, :7] = torch.from_numpy(poses).to(fleet.device); torch.cuda.synchronize()
    got = fleet.lidar(); fleet.sync(); got = got.cpu().numpy().astype(np.float64)
    want = ot.scan(poses)
    bad = np.argwhere((np.abs(got - want) > 1e-4) | ((got < 0) != (want < 0)))
    print("off-nominal seed", seed, "bad rays", len(bad), "of", got.size, "flips", int(((got < 0) != (want < 0)).sum()))
    allbad += [dict(seed=seed, car=int(c), beam=int(b), pose=poses[c].tolist(), got=float(got[c, b]), want=float(want[c, b])) for c, b in bad]
json.dump(allbad, open("gpurun_out/lidar_mismatch_offnominal.json", "w"))
