"""CPU: the oracle's contact set (oracle/step.c car_contacts: wheel / chassis / lidar cylinder against walls and ground)
against an INDEPENDENT numpy statement of the same rules over the EXPLICIT triangle mesh of the walls
(tests/test_oracle_ray_bruteforce.py::chunk_mesh builds every hfield triangle from the wall mask): no chunk index, no
cell arithmetic, no shared code -- every surface triangle near the car is tried.

Rules (this framework's definition): V = a chassis hull vertex below the surface triangle it
projects into / below the ground plane; S = for a wheel ellipsoid or the lidar cylinder, the support point in the
direction opposite to a triangle's normal, if it projects into that triangle and lies below its plane; the deepest
candidate is the geom's one wall contact."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_poses
from test_oracle_ray_bruteforce import chunk_mesh

PLANE_Z = 0.01
WHEEL = np.array([0.03, 0.01, 0.03])
CYL = (0.03, 0.015)
CYL_POS = np.array([-0.0525, 0.0, 0.0575])
WHEEL_POS = np.array([[0.06925, 0.0575, 0.0244], [0.06925, -0.0575, 0.0244], [-0.079, 0.0575, 0.0244], [-0.079, -0.0575, 0.0244]])


def rot(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def surface_triangles(tris):
    """the non-vertical triangles of the mesh (the height-field surface), with upward unit normals"""
    n = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0])
    keep = np.abs(n[:, 2]) > 1e-12
    tris, n = tris[keep], n[keep]
    n = n / np.linalg.norm(n, axis=1, keepdims=True)
    n[n[:, 2] < 0] *= -1
    return tris, n


def inside_xy(tri, p):
    """p's vertical projection inside the (closed) triangles: barycentric coordinates in the xy plane"""
    a, b, c = tri[:, 0, :2], tri[:, 1, :2], tri[:, 2, :2]
    d = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])
    u = ((p[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (p[:, 1] - a[:, 1])) / d
    v = ((b[:, 0] - a[:, 0]) * (p[:, 1] - a[:, 1]) - (p[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])) / d
    eps = 1e-12
    return (u >= -eps) & (v >= -eps) & (u + v <= 1 + eps)


def support(kind, R, c, d):
    """support points of the geom in the world directions d [n, 3]"""
    dl = d @ R                                                        # R^T d per row
    if kind == "ellipsoid":
        s = WHEEL * dl
        s = WHEEL * s / np.linalg.norm(s, axis=1, keepdims=True)
    else:
        h = np.linalg.norm(dl[:, :2], axis=1, keepdims=True)
        s = np.concatenate([np.where(h > 1e-15, CYL[0] * dl[:, :2] / np.maximum(h, 1e-300), 0.0), np.where(dl[:, 2:3] >= 0, CYL[1], -CYL[1])], 1)
    return c + s @ R.T


def rule_S(tris, nrm, kind, R, c, bound):
    near = (np.abs(tris[:, :, 0] - c[0]).min(1) < bound + 0.06) & (np.abs(tris[:, :, 1] - c[1]).min(1) < bound + 0.06)
    T, N = tris[near], nrm[near]
    if not len(T):
        return None
    s = support(kind, R, c, -N)
    dist = ((s - T[:, 0]) * N).sum(1)
    ok = inside_xy(T, s) & (dist < 0)
    if not ok.any():
        return None
    k = np.argmin(np.where(ok, dist, np.inf))
    return dist[k], s[k] - N[k] * dist[k] / 2, N[k]


def brute_contacts(tris, nrm, hull, q):
    """(body, dist, pos, normal) in the oracle's order; bodies: 1 = car, 3/5/7/9 = wheels"""
    R1, p1 = rot(q[3:7]), q[:3]
    out = []
    wheels = []
    for w, (qa, steer) in enumerate(((8, True), (15, True), (22, False), (28, False))):
        pw = p1 + R1 @ (WHEEL_POS[w] + [0, 0, q[qa]])
        ang = q[qa + 1] if steer else 0.0
        thr = q[qa + 2] if steer else q[qa + 1]
        Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
        Ry = np.array([[np.cos(thr), 0, np.sin(thr)], [0, 1, 0], [-np.sin(thr), 0, np.cos(thr)]])
        wheels.append((3 + 2 * w, R1 @ Rz @ Ry, pw))
    for b, Rw, pw in wheels:                                          # wheel-ground
        s = support("ellipsoid", Rw, pw, np.array([[0, 0, -1.0]]))[0]
        if s[2] - PLANE_Z <= 0:
            d = s[2] - PLANE_Z
            out.append((b, d, s - np.array([0, 0, d / 2]), np.array([0, 0, 1.0])))
    for b, Rw, pw in wheels:                                          # wheel-wall
        r = rule_S(tris, nrm, "ellipsoid", Rw, pw, 0.03)
        if r:
            out.append((b, *r))
    nb = 0
    for v in hull:                                                    # chassis hull vertices
        p = p1 + R1 @ v
        near = (np.abs(tris[:, :, 0] - p[0]).min(1) < 0.06) & (np.abs(tris[:, :, 1] - p[1]).min(1) < 0.06)
        T, N = tris[near], nrm[near]
        if len(T) and nb < 8:
            ins = inside_xy(T, np.broadcast_to(p, (len(T), 3)))
            if ins.any():
                # (a vertex exactly on an edge belongs to two triangles: the oracle takes the one of its cell split, v <= u)
                dist = ((p - T[:, 0]) * N).sum(1)
                k = np.nonzero(ins)[0][0]
                if dist[k] < 0:
                    out.append((1, dist[k], p - N[k] * dist[k] / 2, N[k])); nb += 1
        if p[2] - PLANE_Z < 0 and nb < 8:
            d = p[2] - PLANE_Z
            out.append((1, d, p - np.array([0, 0, d / 2]), np.array([0, 0, 1.0]))); nb += 1
    pc = p1 + R1 @ CYL_POS
    r = rule_S(tris, nrm, "cylinder", R1, pc, 0.0336)
    if r and nb < 8:
        out.append((1, *r)); nb += 1
    s = support("cylinder", R1, pc, np.array([[0, 0, -1.0]]))[0]
    if s[2] - PLANE_Z < 0 and nb < 8:
        d = s[2] - PLANE_Z
        out.append((1, d, s - np.array([0, 0, d / 2]), np.array([0, 0, 1.0])))
    return out


def _poses_into_walls(path, wall_xy, n, seed):
    """car poses scattered around wall pixels: on top of, beside and tilted against walls"""
    rng = np.random.default_rng(seed)
    pick = wall_xy[rng.integers(0, len(wall_xy), n)]
    xy = pick + rng.normal(0, 0.07, (n, 2))
    yaw, roll, pitch = rng.uniform(-np.pi, np.pi, n), rng.normal(0, 0.25, n), rng.normal(0, 0.25, n)
    z = rng.uniform(-0.01, 0.12, n)
    cy, sy, cp, sp, cr, sr = np.cos(yaw / 2), np.sin(yaw / 2), np.cos(pitch / 2), np.sin(pitch / 2), np.cos(roll / 2), np.sin(roll / 2)
    q = np.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], 1)
    return np.concatenate([xy, z[:, None], q], 1)


def test_contact_set_matches_bruteforce_over_the_explicit_mesh(oracle, otracks, walls):
    wall, svg = walls["small-circle"]
    t = otracks["small-circle"]
    model = oracle.Model()
    hull = np.array(json.load(open(os.path.join(GOLDEN, "mushr_mesh.json")))["chassis"]["hull"])
    tris, nrm = surface_triangles(chunk_mesh(wall))
    # wall pixel -> world
    H, W = wall.shape
    hc, vc = -(-W // 20), -(-H // 20)
    sx, sy = 40.0 / hc, 40.0 / vc
    py, px = np.nonzero(wall)
    wx = sx * (px // 20 - 0.5 + (px % 20) / 19); wy = -sy * (py // 20 - 0.5 + (py % 20) / 19)
    poses = _poses_into_walls(None, np.stack([wx, wy], 1), 160, seed=5)
    rng = np.random.default_rng(6)
    kinds = {"wheel_wall": 0, "body_wall": 0, "body_ground": 0, "wheel_ground": 0, "cyl": 0}
    compared = 0
    for ps in poses:
        q, _, _ = model.reset(0.0, 0.0, 0.0)
        q[:7] = ps
        q[8], q[15], q[22], q[28] = rng.uniform(-0.03, 0, 4)              # suspension travel
        q[9], q[16] = rng.uniform(-0.5, 0.5, 2); q[10], q[17], q[23], q[29] = rng.uniform(-3, 3, 4)
        got = model.contacts(t, q)
        want = brute_contacts(tris, nrm, hull, q)
        assert len(got) == len(want), (len(got), len(want), ps)
        for g, (b, d, pos, n) in zip(got, want):
            assert int(g[0]) == b
            # the same contact up to rounding; a support point within 1e-9 of a triangle edge may pick the neighbour
            assert abs(g[1] - d) < 1e-9 and np.abs(g[2:5] - pos).max() < 1e-9 and np.abs(g[5:8] - n).max() < 1e-9, (g, d, pos, n)
            compared += 1
            if b != 1:
                kinds["wheel_wall" if n[2] < 0.999999 or pos[2] > 0.02 else "wheel_ground"] += 1
            elif n[2] > 0.999999 and abs(pos[2] + d / 2 - 0.01 - d) < 1e-6:
                kinds["body_ground"] += 1
            else:
                kinds["body_wall"] += 1
    assert compared > 300 and kinds["wheel_wall"] > 20 and kinds["body_wall"] > 50 and kinds["body_ground"] > 5, (compared, kinds)


def test_bubble_wrap_softener_contacts_match_bruteforce(oracle, otracks, walls):
    """option bubble_wrap: rule S for the softener spheres (radius / centre fitted to the wheel mesh, tests/golden/
    mushr_mesh.json) on the softener bodies 4, 6, 8, 10, between the wheel-wall and the car-body contacts"""
    wall, svg = walls["small-circle"]
    t = otracks["small-circle"]
    model = oracle.Model(); model.set_bubble_wrap(True)
    mesh = json.load(open(os.path.join(GOLDEN, "mushr_mesh.json")))
    radius, centre = mesh["softener"]["radius"], np.array(mesh["softener"]["center"])
    tris, nrm = surface_triangles(chunk_mesh(wall))
    H, W = wall.shape
    hc, vc = -(-W // 20), -(-H // 20)
    sx, sy = 40.0 / hc, 40.0 / vc
    py, px = np.nonzero(wall)
    wx = sx * (px // 20 - 0.5 + (px % 20) / 19); wy = -sy * (py // 20 - 0.5 + (py % 20) / 19)
    poses = _poses_into_walls(None, np.stack([wx, wy], 1), 60, seed=9)
    rng = np.random.default_rng(10)
    seen = 0
    for ps in poses:
        q, _, _ = model.reset(0.0, 0.0, 0.0)
        q[:7] = ps
        q[8], q[15], q[22], q[28] = rng.uniform(-0.03, 0, 4); q[9], q[16] = rng.uniform(-0.5, 0.5, 2)
        for a in (11, 18, 24, 30):
            q[a:a + 4] = rng.normal(size=4); q[a:a + 4] /= np.linalg.norm(q[a:a + 4])
        got = [c for c in model.contacts(t, q, maxcon=20) if int(c[0]) in (4, 6, 8, 10)]
        R1, p1 = rot(q[3:7]), q[:3]
        want = []
        for w, (qa, steer) in enumerate(((8, True), (15, True), (22, False), (28, False))):
            ang = q[qa + 1] if steer else 0.0
            thr = q[qa + 2] if steer else q[qa + 1]
            Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
            Ry = np.array([[np.cos(thr), 0, np.sin(thr)], [0, 1, 0], [-np.sin(thr), 0, np.cos(thr)]])
            Rs = R1 @ Rz @ Ry @ rot(q[qa + (3 if steer else 2):qa + (7 if steer else 6)])
            c = p1 + R1 @ (WHEEL_POS[w] + [0, 0, q[qa]]) + Rs @ centre
            near = (np.abs(tris[:, :, 0] - c[0]).min(1) < 0.11) & (np.abs(tris[:, :, 1] - c[1]).min(1) < 0.11)
            T, N = tris[near], nrm[near]
            if not len(T):
                continue
            s = c - radius * N                                         # sphere support point against each plane
            dist = ((s - T[:, 0]) * N).sum(1)
            ok = inside_xy(T, s) & (dist < 0)
            if ok.any():
                k = np.argmin(np.where(ok, dist, np.inf))
                want.append((4 + 2 * w, dist[k], s[k] - N[k] * dist[k] / 2, N[k]))
        assert len(got) == len(want), (len(got), len(want))
        for g, (b, d, pos, n) in zip(got, want):
            assert int(g[0]) == b and abs(g[1] - d) < 1e-9 and np.abs(g[2:5] - pos).max() < 1e-9 and np.abs(g[5:8] - n).max() < 1e-9
            assert g[8] == 1.0 and g[9] == 0.9
            seen += 1
    assert seen > 40


def test_level_driving_car_has_only_wheel_ground_contacts(oracle, otracks, walls):
    """on the open track the new rules add nothing: 4 wheel-ground contacts, as before"""
    wall, svg = walls["track"]
    t = otracks["track"]
    path = t.centreline(svg)
    model = oracle.Model()
    for k in (10, 12, 30, 55):
        d = path[k + 1] - path[k]
        q, v, w = model.reset(path[k, 0], path[k, 1], float(np.arctan2(d[1], d[0])))
        for _ in range(60):
            model.step(t, q, v, w, np.array([1.0, 0.0]))
        c = model.contacts(t, q)
        assert len(c) == 4 and set(c[:, 0].astype(int)) == {3, 5, 7, 9} and (c[:, 8] == 0.5).all()
