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
