"""CPU: the oracle's mj_ray restatement (oracle/ray.c: per-hfield top-box clipping, cell-range traversal, side faces)
against an INDEPENDENT brute-force ray caster: every height-field triangle and every side-face trapezoid of every
chunk is built explicitly from the wall mask with the pixel -> world mapping of SURVEY C.2 / B.10 (chunk.py:49-64,
mushr.em.xml:19-20,55,92), and each ray is intersected with all ~400,000 triangles (Moeller-Trumbore, two-sided,
numpy).  Same specification, entirely different algorithm and code: it pins the traversal logic of ray.c (and
therefore of the CUDA lidar kernel that is tested against ray.c), not MuJoCo's own reading of the model."""
import numpy as np
import pytest

from conftest import random_poses

HF_RANGE, HF_Z0, PLANE_Z = 0.3, -0.1, 0.01


def chunk_mesh(wall, scale=2.0, px=20):
    """All triangles of all non-empty chunks: [n, 3, 3] world coordinates."""
    H, W = wall.shape
    hc, vc = -(-W // px), -(-H // px)
    size_x, size_y = 20.0 * scale / hc, 20.0 * scale / vc
    tris = []
    for i in range(hc):
        for j in range(vc):
            blk = wall[j * px:(j + 1) * px, i * px:(i + 1) * px]
            if not blk.any():
                continue
            nrow, ncol = blk.shape
            if nrow < 2 or ncol < 2:
                continue
            # hfield rows are flipped: row r = nrow-1-k is PNG row k; elevation normalised to [0, 1] (constant chunk -> 0)
            e = blk[::-1].astype(np.float64)
            e = (e - e.min()) / (e.max() - e.min()) if e.max() > e.min() else np.zeros_like(e)
            z = HF_Z0 + HF_RANGE * e
            xs = size_x * i - size_x / 2 + np.arange(ncol) * size_x / (ncol - 1)
            ys = -size_y * j - size_y / 2 + np.arange(nrow) * size_y / (nrow - 1)
            P = np.stack([np.broadcast_to(xs[None, :], (nrow, ncol)), np.broadcast_to(ys[:, None], (nrow, ncol)), z], -1)
            a, b, c, d = P[:-1, :-1], P[:-1, 1:], P[1:, :-1], P[1:, 1:]          # (c,r) (c+1,r) (c,r+1) (c+1,r+1)
            keep = (e[:-1, :-1] + e[:-1, 1:] + e[1:, :-1] + e[1:, 1:]) > 0     # flat cells at the base are below the ground plane
            tris.append(np.stack([a, d, b], -2)[keep]); tris.append(np.stack([a, d, c], -2)[keep])
            # vertical side faces of the top box: from the base (z = HF_Z0) up to the boundary elevation profile
            for border in (P[:, 0], P[:, -1], P[0, :], P[-1, :]):
                p0, p1 = border[:-1], border[1:]
                use = (p0[:, 2] > HF_Z0) | (p1[:, 2] > HF_Z0)
                b0, b1 = p0.copy(), p1.copy(); b0[:, 2] = HF_Z0; b1[:, 2] = HF_Z0
                tris.append(np.stack([b0, b1, p1], -2)[use]); tris.append(np.stack([b0, p1, p0], -2)[use])
    return np.concatenate(tris, 0)


def brute_ray(tris, o, d):
    """Nearest non-negative hit of ray o + s d with the triangles and the ground plane, or -1."""
    v0, e1, e2 = tris[:, 0], tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]
    p = np.cross(d, e2)
    det = (e1 * p).sum(1)
    ok = np.abs(det) > 1e-300
    inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
    t = o - v0
    u = (t * p).sum(1) * inv
    q = np.cross(t, e1)
    v = (q * d).sum(1) * inv
    s = (e2 * q).sum(1) * inv
    hit = ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (s >= 0)
    best = s[hit].min() if hit.any() else np.inf
    if d[2] < 0:                                                     # ground plane z = 0.01, front side, 300 m half-size
        sp = (PLANE_Z - o[2]) / d[2]
        h = o + sp * d
        if sp >= 0 and abs(h[0]) <= 300 and abs(h[1]) <= 300:
            best = min(best, sp)
    return best if np.isfinite(best) else -1.0


@pytest.mark.parametrize("name,nposes", [("track", 5), ("small-circle", 3)])
def test_oracle_scan_matches_bruteforce_triangle_mesh(name, nposes, otracks, walls):
    wall, svg = walls[name]
    t = otracks[name]
    path = t.centreline(svg)
    tris = chunk_mesh(wall)
    assert len(tris) > 50000
    poses = random_poses(path, nposes, seed=77)
    got = t.scan(poses)                                              # oracle/ray.c, 90 beams per pose
    rx, rz, lr = -0.0525, 0.065, 0.03                                # mushr.em.xml:101-103,114-116
    worst, edge = 0.0, 0
    for k, ps in enumerate(poses):
        w, x, y, z = ps[3:7] / np.linalg.norm(ps[3:7])
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        for j in range(90):
            b = np.radians(4 * j - 90)
            d = R @ np.array([np.sin(b), -np.cos(b), 0.0])
            o = ps[:3] + R @ np.array([rx - lr * np.sin(b), lr * np.cos(b), rz])
            want = brute_ray(tris, o, d)
            if (want < 0) != (got[k, j] < 0) or abs(want - got[k, j]) > 1e-9:
                # a ray through a shared triangle edge / vertex may legitimately pick either facet: tolerate a few
                edge += 1
                assert abs(want - got[k, j]) < 5e-3, (k, j, want, got[k, j])
            else:
                worst = max(worst, abs(want - got[k, j]))
    assert edge <= 2 and worst < 1e-9, (edge, worst)


def cylinder_mesh(pose, nfacets=1440):
    """Triangles of another car's lidar cylinder (mushr.em.xml:108): r = 0.03, half height 0.015, centred at
    (-0.0525, 0, 0.0575) in the car frame; side facets + two caps.  Chord error r (1 - cos(pi / n)) = 7e-8 m."""
    w, x, y, z = pose[3:7] / np.linalg.norm(pose[3:7])
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    c = np.array([-0.0525, 0.0, 0.0575])
    a = np.linspace(0, 2 * np.pi, nfacets + 1)
    ring = np.stack([0.03 * np.cos(a), 0.03 * np.sin(a), np.zeros_like(a)], 1)
    top, bot = ring + [0, 0, 0.015], ring - [0, 0, 0.015]
    ctop, cbot = np.array([0, 0, 0.015]), np.array([0, 0, -0.015])
    tris = []
    for i in range(nfacets):
        tris += [[bot[i], bot[i + 1], top[i + 1]], [bot[i], top[i + 1], top[i]], [ctop, top[i], top[i + 1]], [cbot, bot[i + 1], bot[i]]]
    t = np.array(tris) + c
    return t @ R.T + pose[:3]


def test_oracle_world_scan_matches_bruteforce_with_other_cars(otracks, walls):
    """Multi-car worlds (BASELINE config 5): each car's 90 rays against the walls AND the other cars' lidar cylinders
    (its own is excluded: mj_ray's bodyexclude; a shadowed car is invisible), oracle scan_world vs the brute-force
    caster over the wall mesh plus explicitly triangulated cylinders, for a tight cluster of four cars."""
    wall, svg = walls["track"]
    t = otracks["track"]
    path = t.centreline(svg)
    tris = chunk_mesh(wall)
    n = 4
    rng = np.random.default_rng(5)
    poses = np.zeros((n, 7))
    off = [(0.0, 0.0), (0.35, 0.05), (-0.30, 0.20), (0.10, -0.38)]          # a tight cluster: several beams end on a cylinder
    for i in range(n):
        yaw = rng.uniform(-3, 3)
        poses[i] = [path[10, 0] + off[i][0], path[10, 1] + off[i][1], 0.0156, np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)]
    vis = np.array([1, 1, 0, 1], dtype=np.uint8)                 # car 2 is shadowed: nobody sees it
    got = t.scan_world(poses, vis)
    alone = t.scan(poses)
    cyl = [cylinder_mesh(poses[i]) for i in range(n)]
    rx, rz, lr = -0.0525, 0.065, 0.03
    seen_cars = 0
    for i in range(n):
        others = [cyl[j] for j in range(n) if j != i and vis[j]]
        mesh = np.concatenate([tris] + others, 0)
        w, x, y, z = poses[i, 3:7]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        for j in range(90):
            b = np.radians(4 * j - 90)
            d = R @ np.array([np.sin(b), -np.cos(b), 0.0])
            o = poses[i, :3] + R @ np.array([rx - lr * np.sin(b), lr * np.cos(b), rz])
            want = brute_ray(mesh, o, d)
            assert (want < 0) == (got[i, j] < 0) and abs(want - got[i, j]) < 2e-6, (i, j, want, got[i, j])
            seen_cars += abs(got[i, j] - alone[i, j]) > 1e-3
    assert seen_cars >= 8                                         # several rays really end on another car's cylinder


# (the wall / ground contact set is checked against this explicit mesh in tests/test_contacts_cpu.py)


def _rotm(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _first_root(F, smax=3.0, n=6000):
    """first s in [0, smax] where the implicit function F(s) crosses zero (dense sampling + bisection), or inf"""
    s = np.linspace(0.0, smax, n)
    f = F(s)
    k = np.nonzero(np.sign(f[:-1]) != np.sign(f[1:]))[0]
    if not len(k):
        return np.inf
    lo, hi = s[k[0]], s[k[0] + 1]
    flo = F(np.array([lo]))[0]
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        fm = F(np.array([mid]))[0]
        if (fm < 0) == (flo < 0):
            lo, flo = mid, fm
        else:
            hi = mid
    return 0.5 * (lo + hi)


def test_oracle_world_scan_sees_wheels_and_chassis_of_pitched_cars(otracks, walls):
    """f1 / mj_ray with geomgroup NULL: a ray also ends on another car's chassis mesh (its own 33 triangles,
    mushr.em.xml:119) and wheel ellipsoids (:69), not only on its lidar cylinder.  On level ground those lie below the
    beams, so the observer cars here are pitched / rolled / lifted.  Independent statement: the chassis triangles of
    tests/golden (assets/meshes.npz, scaled and placed in numpy) go through the brute-force triangle caster; each
    ellipsoid is the zero set of |S^-1 R^T (p - c)|^2 - 1, whose first crossing along the ray is bracketed by dense
    sampling and bisected (no closed-form quadratic, no shared code)."""
    import os
    from conftest import ROOT
    wall, svg = walls["track"]
    t = otracks["track"]
    path = t.centreline(svg)
    tris = chunk_mesh(wall)
    base = np.load(os.path.join(ROOT, "ft_grandprix_b200", "assets", "meshes.npz"))["simple_base_nano__tri"].astype(np.float64)
    chassis_local = base * 0.5 + np.array([0, 0, 0.5 * 0.094655])                      # mushr.em.xml:38,119
    wpos = np.array([[0.06925, 0.0575, 0.0244], [0.06925, -0.0575, 0.0244], [-0.079, 0.0575, 0.0244], [-0.079, -0.0575, 0.0244]])
    S = np.array([0.03, 0.01, 0.03])
    rng = np.random.default_rng(8)
    n = 5
    q = np.zeros((n, 34))
    q[:, [11, 18, 24, 30]] = 1.0
    off = [(0.0, 0.0), (0.30, 0.04), (-0.28, 0.16), (0.08, -0.33), (0.27, -0.25)]
    for i in range(n):
        yaw, pitch, roll = rng.uniform(-3, 3), rng.normal(0, 0.22), rng.normal(0, 0.22)
        cy, sy, cp, sp, cr, sr = np.cos(yaw / 2), np.sin(yaw / 2), np.cos(pitch / 2), np.sin(pitch / 2), np.cos(roll / 2), np.sin(roll / 2)
        q[i, :3] = [path[10, 0] + off[i][0], path[10, 1] + off[i][1], 0.03 + rng.uniform(0, 0.05)]
        q[i, 3:7] = [cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy]
        q[i, [8, 15, 22, 28]] = rng.uniform(-0.03, 0, 4)
        q[i, [9, 16]] = rng.uniform(-0.6, 0.6, 2)
        q[i, [10, 17, 23, 29]] = rng.uniform(-3, 3, 4)                                # throttle angle: must not matter
    vis = np.array([1, 1, 1, 0, 1], dtype=np.uint8)
    got = t.scan_world(q, vis)
    hits = {"chassis": 0, "wheel": 0, "cyl": 0}
    rx, rz, lr = -0.0525, 0.065, 0.03
    for i in range(n):
        R = _rotm(q[i, 3:7])
        meshes, chassis, ells = [tris], [], []
        for c in range(n):
            if c == i or not vis[c]:
                continue
            Rc = _rotm(q[c, 3:7])
            meshes.append(cylinder_mesh(q[c, :7])); chassis.append(chassis_local @ Rc.T + q[c, :3])
            for w in range(4):
                st = q[c, [9, 16][w]] if w < 2 else 0.0
                Rz = np.array([[np.cos(st), -np.sin(st), 0], [np.sin(st), np.cos(st), 0], [0, 0, 1]])
                ells.append((q[c, :3] + Rc @ (wpos[w] + [0, 0, q[c, [8, 15, 22, 28][w]]]), Rc @ Rz))
        mesh = np.concatenate(meshes, 0)
        chassis = np.concatenate(chassis, 0)
        alone = t.scan(q[i:i + 1, :7])[0]
        for j in range(90):
            b = np.radians(4 * j - 90)
            d = R @ np.array([np.sin(b), -np.cos(b), 0.0])
            o = q[i, :3] + R @ np.array([rx - lr * np.sin(b), lr * np.cos(b), rz])
            want = brute_ray(mesh, o, d)
            want = np.inf if want < 0 else want
            kind = "cyl" if abs(want - alone[j]) > 1e-3 else "wall"
            o2 = o.copy()
            v0, e1, e2 = chassis[:, 0], chassis[:, 1] - chassis[:, 0], chassis[:, 2] - chassis[:, 0]
            pp = np.cross(d, e2); det = (e1 * pp).sum(1); okd = np.abs(det) > 1e-300
            inv = np.where(okd, 1.0 / np.where(okd, det, 1.0), 0.0)
            tt = o2 - v0; uu = (tt * pp).sum(1) * inv; qq = np.cross(tt, e1); vv = (qq * d).sum(1) * inv; ss = (e2 * qq).sum(1) * inv
            hit = okd & (uu >= 0) & (vv >= 0) & (uu + vv <= 1) & (ss >= 0)
            if hit.any() and ss[hit].min() < want:
                want, kind = ss[hit].min(), "chassis"
            for cc, Re in ells:
                F = lambda s: ((((o[None] + s[:, None] * d[None] - cc) @ Re) / S) ** 2).sum(1) - 1.0
                se = _first_root(F)
                if se < want:
                    want, kind = se, "wheel"
            want = -1.0 if not np.isfinite(want) else want
            assert (want < 0) == (got[i, j] < 0) and abs(want - got[i, j]) < 5e-6, (i, j, want, got[i, j], kind)
            if kind in hits:
                hits[kind] += 1
    assert hits["wheel"] >= 3 and hits["chassis"] >= 3 and hits["cyl"] >= 3, hits
