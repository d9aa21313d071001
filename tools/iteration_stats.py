import numpy as np, torch, sys
sys.path.insert(0,'/root/repo')
import ft_grandprix_b200 as ft
sys.path.insert(0,'/root/repo'); from bench import make_poses
t = ft.Track.bundled("track")
n = 65536
xy, yaw, _ = make_poses(t.path, n, 1, True)
f = ft.Fleet(t, n); f.reset(xy, yaw)
f.tick(200); f.sync()
prev = f.status.cpu().numpy().copy()
f.tick(1); f.sync()
cur = f.status.cpu().numpy()
it_prev, it = prev & 0xFF, cur & 0xFF
wall_prev = ((prev >> 16) & 0xFF) > 0
print("iters hist:", np.bincount(it, minlength=10)[:12], "mean", it.mean())
print("P(it == it_prev)", (it == it_prev).mean(), " |diff|<=1", (np.abs(it.astype(int)-it_prev.astype(int))<=1).mean())
# CTA max with the kernel's grouping: sort by bin = min(it_prev,7) + 8*wall
b = np.minimum(it_prev, 7) + 8 * wall_prev
order = np.argsort(b, kind='stable')
for cta in (54, 8):
    g = it[order][: (n // cta) * cta].reshape(-1, cta)
    print(f"cars/group {cta}: mean of group max (sorted by last count) {g.max(1).mean():.3f}; unsorted {it[: (n // cta) * cta].reshape(-1, cta).max(1).mean():.3f}; ideal (sorted by this tick's count) {np.sort(it)[: (n // cta) * cta].reshape(-1, cta).max(1).mean():.3f}")
