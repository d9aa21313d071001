"""Build container: classify the rays of gpurun_out/lidar_mismatch_offnominal.json (GPU lidar != oracle on off-nominal
poses).  For each ray: brute-force truth over the explicit triangle mesh, ray geometry (origin height, inside a wall?,
slope, distance to the nearest triangle edge at the hit), so that the cause of the mismatch can be named."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle
from test_oracle_ray_bruteforce import chunk_mesh, brute_ray

z = np.load(os.path.join(ROOT, "ft_grandprix_b200", "assets", "tracks.npz"))
shape = tuple(int(v) for v in z["track__shape"])
wall = np.unpackbits(z["track__bits"])[: shape[0] * shape[1]].reshape(shape)
ot = pyoracle.Track(wall)
tris = chunk_mesh(wall)
bad = json.load(open(os.path.join(ROOT, "gpurun_out", "lidar_mismatch_offnominal.json")))
print(len(bad), "mismatching rays")
rx, rz, lr = -0.0525, 0.065, 0.03
rows = []
for b in bad[:400]:
    ps = np.array(b["pose"]); j = b["beam"]
    w, x, y, zq = ps[3:7] / np.linalg.norm(ps[3:7])
    R = np.array([[1 - 2 * (y * y + zq * zq), 2 * (x * y - w * zq), 2 * (x * zq + w * y)],
                  [2 * (x * y + w * zq), 1 - 2 * (x * x + zq * zq), 2 * (y * zq - w * x)],
                  [2 * (x * zq - w * y), 2 * (y * zq + w * x), 1 - 2 * (x * x + y * y)]])
    bb = np.radians(4 * j - 90)
    d = R @ np.array([np.sin(bb), -np.cos(bb), 0.0]); o = ps[:3] + R @ np.array([rx - lr * np.sin(bb), lr * np.cos(bb), rz])
    truth = brute_ray(tris, o, d)
    # local neighbourhood only for speed would be nicer; the full mesh is fine for a few hundred rays
    who = "gpu" if abs(truth - b["got"]) < 1e-4 and (truth < 0) == (b["got"] < 0) else ("oracle" if abs(truth - b["want"]) < 1e-4 and (truth < 0) == (b["want"] < 0) else "neither")
    rows.append((who, b["got"], b["want"], truth, o[2], d[2], ps[0], ps[1]))
import collections
print(collections.Counter(r[0] for r in rows))
for r in rows[:60]:
    print("right: %-7s got %9.5f oracle %9.5f truth %9.5f | origin z %.4f dir z %+.4f at (%.2f, %.2f)" % r)
