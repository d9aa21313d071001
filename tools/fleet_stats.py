"""GPU box: what regime is the bench fleet in after N ticks? (Newton iterations, contacts, speeds)"""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ft_grandprix_b200 as ft
from bench import make_poses
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
t = ft.Track.bundled("track")
fleet = ft.Fleet(t, n)
xy, yaw, _ = make_poses(t.path, n, 1, True)
fleet.reset(xy, yaw)
for k in (1, 20, 100, 200, 500, 1000, 2000):
    fleet.tick(k - fleet.steps); fleet.sync()
    st = fleet.status.cpu().numpy()
    it = st & 0xFF; wall = (st >> 16) & 0xFF; wheel = (st >> 24) & 0xF; rst = (st >> 8) & 1
    v = fleet.qvel[:, :2].norm(dim=1).cpu().numpy()
    lap = fleet.lap.cpu().numpy()
    print(f"tick {k:5d}: iters mean {it.mean():.2f} hist {np.bincount(it, minlength=8)[:8].tolist()} | wall-contact cars {(wall>0).mean()*100:.1f}% | "
          f"wheel contacts mean {wheel.mean():.2f} | reset {rst.sum()} | speed mean {v.mean():.2f} p95 {np.percentile(v,95):.2f} | "
          f"off-track {lap[:, 7].mean()*100:.1f}% | laps max {lap[:, 2].max()} completion mean {lap[:,1].mean():.1f}")
