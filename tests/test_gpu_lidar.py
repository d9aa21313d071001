"""GPU: lidar / drivers / lap / reset kernels against the CPU oracle, through the C ABI."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_poses

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ft():
    import ft_grandprix_b200 as ft
    return ft


def _scan_gpu(ft, track, poses, **kw):
    fleet = ft.Fleet(track, len(poses), **kw)
    fleet.qpos[:, :7] = torch.from_numpy(poses).to(fleet.device)
    torch.cuda.synchronize()
    r = fleet.lidar()
    fleet.sync()
    return r.cpu().numpy().astype(np.float64), fleet


def _check(got, want, tol=1e-4):
    miss_g, miss_w = got < 0, want < 0
    assert (miss_g == miss_w).all(), f"{int((miss_g != miss_w).sum())} rays disagree on hit/miss"
    err = np.abs(got - want)[~miss_w]
    assert err.max() <= tol, f"max |d range| = {err.max():.3e} m"
    return err.max()


@pytest.mark.parametrize("name", ["track", "circle", "small-circle", "inkscape"])
def test_lidar_config2_parity(ft, otracks, name):
    """BASELINE config 2: 4096 random on-track poses x 90 beams, |d range| <= 1e-4 m, misses agree."""
    t = ft.Track.bundled(name)
    n = 4096 if name == "track" else 1024
    poses = random_poses(t.path, n, seed=0)
    got, _ = _scan_gpu(ft, t, poses)
    want = otracks[name].scan(poses)
    _check(got, want)


def test_lidar_level_and_tilted_and_far(ft, otracks):
    t = ft.Track.bundled("track")
    rng = np.random.default_rng(3)
    n = 1024
    # anywhere on the 40 m map (including inside wall cells and outside the track), large tilts
    xy = rng.uniform(-2, 42, (n, 2)) * [1, -1]
    yaw = rng.uniform(-np.pi, np.pi, n)
    tilt = rng.normal(0, 0.15, (n, 2))
    z = rng.uniform(-0.05, 0.3, n)
    cy, sy = np.cos(yaw / 2), np.sin(yaw / 2)
    cp, sp, cr, sr = np.cos(tilt[:, 0] / 2), np.sin(tilt[:, 0] / 2), np.cos(tilt[:, 1] / 2), np.sin(tilt[:, 1] / 2)
    q = np.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                  cr * cp * sy - sr * sp * cy], 1)
    poses = np.concatenate([xy, z[:, None], q], 1)
    got, _ = _scan_gpu(ft, t, poses)
    want = otracks["track"].scan(poses)
    # Round 1 needed slack here (hit / miss flips on 2e-4 of the rays).  Cause, found by dumping the offending rays
    # (tools/lidar_mismatch.py, tools/lidar_offnominal_analysis.py): every one of them started BELOW the ground plane
    # (origin z < 0.01, possible only for a car sunk 5 cm into the floor) and ended exactly on the hfield's base plane
    # z = -0.1, the lower face of the height slab, where fp32 rounding decided whether the flat floor cells were candidates.
    # With a margin on that test the kernel agrees with the oracle exactly.
    _check(got, want)


def test_lidar_multi_car_world(ft, otracks):
    """config 5 flavour: 8 cars per world see each other's lidar cylinder; shadowed cars are invisible."""
    t = ft.Track.bundled("track")
    nworlds, cpw = 64, 8
    rng = np.random.default_rng(11)
    poses = np.zeros((nworlds * cpw, 7))
    for w in range(nworlds):
        for c in range(cpw):
            x, y, yaw = t.start_pose(c)
            poses[w * cpw + c] = [x + rng.normal(0, 0.05), y + rng.normal(0, 0.05), 0.0156,
                                  np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)]
    vis = (rng.random(nworlds * cpw) > 0.2).astype(np.uint8)
    fleet = ft.Fleet(t, nworlds * cpw, cars_per_world=cpw)
    fleet.qpos[:, :7] = torch.from_numpy(poses).to(fleet.device)
    v = torch.from_numpy(vis).to(fleet.device)
    torch.cuda.synchronize()
    got = fleet.lidar(visible=v); fleet.sync()
    got = got.cpu().numpy().astype(np.float64)
    want = np.concatenate([otracks["track"].scan_world(poses[w * cpw:(w + 1) * cpw], vis[w * cpw:(w + 1) * cpw])
                           for w in range(nworlds)])
    _check(got, want)
    alone = otracks["track"].scan(poses)
    assert (np.abs(alone - want) > 1e-3).sum() > 50       # the other cars really are seen


@pytest.mark.parametrize("seed,spread,tilt,seen", [(12, 0.22, 0.2, 300), (13, 0.08, 0.5, 300), (14, 1.2, 0.35, 40)])
def test_lidar_multi_car_world_pitched_cars_see_chassis_and_wheels(ft, otracks, seed, spread, tilt, seen):
    """f1: mj_ray also ends on the other cars' chassis mesh and wheel ellipsoids (suspension travel and steering angle from
    the state rows); on level ground those lie below the beams, so the cars here are pitched, rolled and lifted.  The kernel
    finds these hits in a pass of its own over the window of beams that can reach the other car's bounding sphere: packs so
    tight that the window is the whole circle (spread 0.08 m, interpenetrating cars), loose ones where it is a few beams, and
    every azimuth (the window wraps around beam 89 -> 0) must all agree with the oracle, which tests every ray against every car."""
    t = ft.Track.bundled("track")
    nworlds, cpw = 48, 8
    n = nworlds * cpw
    rng = np.random.default_rng(seed)
    q = np.zeros((n, 34)); q[:, [11, 18, 24, 30]] = 1.0
    for w in range(nworlds):
        cx, cy = t.path[rng.integers(0, 100)]
        for c in range(cpw):
            yaw, pitch, roll = rng.uniform(-3.2, 3.2), rng.normal(0, tilt), rng.normal(0, tilt)
            a, b, cc_, d, e, f = np.cos(yaw / 2), np.sin(yaw / 2), np.cos(pitch / 2), np.sin(pitch / 2), np.cos(roll / 2), np.sin(roll / 2)
            i = w * cpw + c
            q[i, :3] = [cx + rng.normal(0, spread), cy + rng.normal(0, spread), 0.02 + rng.uniform(0, 0.06)]
            q[i, 3:7] = [e * cc_ * a + f * d * b, f * cc_ * a - e * d * b, e * d * a + f * cc_ * b, e * cc_ * b - f * d * a]
            q[i, [8, 15, 22, 28]] = rng.uniform(-0.03, 0, 4); q[i, [9, 16]] = rng.uniform(-0.6, 0.6, 2)
            q[i, [10, 17, 23, 29]] = rng.uniform(-3, 3, 4)
    vis = (rng.random(n) > 0.15).astype(np.uint8)
    fleet = ft.Fleet(t, n, cars_per_world=cpw)
    fleet.qpos.copy_(torch.from_numpy(q)); v = torch.from_numpy(vis).to(fleet.device)
    torch.cuda.synchronize()
    got = fleet.lidar(visible=v); fleet.sync()
    got = got.cpu().numpy().astype(np.float64)
    ot = otracks["track"]
    want = np.concatenate([ot.scan_world(q[w * cpw:(w + 1) * cpw], vis[w * cpw:(w + 1) * cpw]) for w in range(nworlds)])
    miss = (got < 0) != (want < 0)
    bad = np.abs(got - want) > 1e-4
    assert miss.sum() == 0 and bad.mean() < 2e-4, (miss.sum(), bad.sum(), np.abs(got - want).max())
    # the new targets matter: compare with a scan that only knows the other cars' lidar cylinders' nominal neighbours
    alone = ot.scan(q[:, :7])
    assert (np.abs(alone - want) > 1e-3).sum() > seen


def test_lidar_host_entry_and_ragged(ft, otracks):
    """C ABI with host buffers, ncars = 0 / 1 / non-multiple of the warp count."""
    import ctypes as C
    t = ft.Track.bundled("small-circle")
    g = ft.Geometry(t)
    lib = ft._lib.load()
    for n in (0, 1, 37):
        poses = random_poses(t.path, max(n, 1), seed=n)[:n]
        qpos = np.zeros((n, 34)); qpos[:, :7] = poses
        out = np.full((n, 90), 7.0, dtype=np.float32)
        rc = lib.ftgp_lidar_host(g._ptr, qpos.ctypes.data_as(C.c_void_p), 34, None, n, out.ctypes.data_as(C.c_void_p))
        assert rc == 0
        if n:
            _check(out.astype(np.float64), otracks["small-circle"].scan(poses))
    assert lib.ftgp_lidar_host(g._ptr, None, 34, None, 4, None) == 1       # FTGP_ERR_ARG
    assert b"bad argument" in lib.ftgp_last_error()


def test_lidar_mixed_tracks(ft, otracks):
    """config 4 flavour: circle and small-circle alternate by car index in one fleet."""
    ta, tb = ft.Track.bundled("circle"), ft.Track.bundled("small-circle")
    n = 512
    tid = np.arange(n) % 2
    pa, pb = random_poses(ta.path, n, seed=1), random_poses(tb.path, n, seed=2)
    poses = np.where(tid[:, None] == 0, pa, pb)
    got, _ = _scan_gpu(ft, [ta, tb], poses, track_id=tid)
    want = np.where(tid[:, None] == 0, otracks["circle"].scan(poses), otracks["small-circle"].scan(poses))
    _check(got, want)


def test_drivers_kernel_matches_reference_goldens(ft):
    """a5: the CUDA drivers reproduce ft_grandprix.nidc / fast on the golden scans (fp32-representable)."""
    z = np.load(os.path.join(GOLDEN, "drivers.npz"))
    scans32 = z["scans"].astype(np.float32)
    t = ft.Track.bundled("small-circle")
    fleet = ft.Fleet(t, len(scans32))
    from oracle import pyoracle
    for kind, name in ((0, "nidc"), (1, "fast"), (2, "lobotomy")):
        fleet.default_driver = kind
        fleet.ranges.copy_(torch.from_numpy(scans32))
        fleet.ctrl.fill_(99.0)
        torch.cuda.synchronize()
        fleet.drive(); fleet.sync()
        got = fleet.ctrl.cpu().numpy()
        want = np.array([pyoracle.driver(kind, s.astype(np.float64)) for s in scans32])
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
        # where the golden scan is exactly representable in fp32 the reference's own answer applies
        if name != "lobotomy":
            exact = (scans32.astype(np.float64) == z["scans"]).all(1)
            assert exact.sum() >= 3
            np.testing.assert_allclose(got[exact], z[name][exact], rtol=0, atol=1e-12)
    # per-car kinds + finished cars are lobotomised
    kinds = np.arange(len(scans32)) % 3
    fleet.set_driver_kinds(kinds.tolist())
    fleet.lap[::5, ft.fleet.LAP["finished"]] = 1
    torch.cuda.synchronize()
    fleet.drive(); fleet.sync()
    got = fleet.ctrl.cpu().numpy()
    for i, s in enumerate(scans32):
        k = 2 if i % 5 == 0 else int(kinds[i])
        np.testing.assert_allclose(got[i], pyoracle.driver(k, s.astype(np.float64)), rtol=0, atol=1e-12)


def test_lap_kernel_integer_exact(ft, otracks, walls):
    from oracle import pyoracle
    t = ft.Track.bundled("track")
    nworlds, cpw = 16, 4
    n = nworlds * cpw
    fleet = ft.Fleet(t, n, cars_per_world=cpw, lap_target=2)
    fleet.reset_grid()
    path = t.path
    rng = np.random.default_rng(8)
    laps = [pyoracle.Lap(offset=(i % cpw + 5) * 2, max_times=16) for i in range(n)]
    nwin = [0] * nworlds
    pos = np.array([(i % cpw + 5) * 2 for i in range(n)], dtype=float)
    speed = rng.uniform(-0.4, 1.5, n)
    for step in range(400):
        pos += speed
        k = np.floor(pos).astype(int) % 100
        xy = path[k] + rng.normal(0, 0.05, (n, 2)) + (rng.random((n, 1)) < 0.02) * 3.0
        fleet.qpos[:, :2] = torch.from_numpy(xy).to(fleet.device)
        torch.cuda.synchronize()
        fleet.steps = step
        fleet.lap_update()
        for i in range(n):
            nwin[i // cpw] = laps[i].update(path, xy[i], step, 2, nwin[i // cpw])
    fleet.sync()
    lap = fleet.lap.cpu().numpy(); times = fleet.times.cpu().numpy()
    L = ft.fleet.LAP
    for i in range(n):
        s = laps[i].s
        for f in ("offset", "completion", "laps", "start", "good_start", "finished", "ntimes", "off_track", "rank", "delta"):
            assert lap[i, L[f]] == getattr(s, f), (i, f)
        assert times[i, : min(s.ntimes, 16)].tolist() == laps[i].times[: min(s.ntimes, 16)].tolist()
    assert fleet.winners.cpu().numpy().tolist() == nwin
    assert lap[:, L["finished"]].sum() > 0


def test_reset_matches_reference_spawn(ft):
    t = ft.Track.bundled("track")
    fleet = ft.Fleet(t, 8, cars_per_world=8)
    fleet.reset_grid()
    q = fleet.qpos.cpu().numpy()
    for i in range(8):
        x, y, yaw = t.start_pose(i)
        want = np.zeros(34); want[[0, 1]] = x, y
        want[3], want[6] = np.cos(yaw / 2), np.sin(yaw / 2)
        want[[11, 18, 24, 30]] = 1
        np.testing.assert_allclose(q[i], want, atol=1e-15)
    assert fleet.lap[:, ft.fleet.LAP["offset"]].cpu().tolist() == [(i + 5) * 2 for i in range(8)]
    assert float(fleet.ranges.abs().sum()) == 0.0       # zeros on the first tick


def test_lidar_four_tracks_geometry_read_from_global_memory(ft, otracks):
    """All four bundled tracks in one geometry do not fit in shared memory: the kernel reads them through L2."""
    names = ["track", "circle", "small-circle", "inkscape"]
    tracks = [ft.Track.bundled(nm) for nm in names]
    n = 512
    tid = np.arange(n) % 4
    poses = np.zeros((n, 7))
    for k, t in enumerate(tracks):
        poses[tid == k] = random_poses(t.path, n, seed=20 + k)[tid == k]
    got, fleet = _scan_gpu(ft, tracks, poses, track_id=tid)
    assert fleet.geom.nbytes > 200 * 1024
    want = np.zeros((n, 90))
    for k, nm in enumerate(names):
        want[tid == k] = otracks[nm].scan(poses[tid == k])
    _check(got, want)


def test_snapshot_tensors_match_host_snapshots(ft):
    """v2 driver input on device == the per-car VehicleStateSnapshot objects of the host shim."""
    t = ft.Track.bundled("track")
    n = 64
    fleet = ft.Fleet(t, n)
    fleet.reset_grid() if n <= 40 else None
    rng = np.random.default_rng(5)
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    fleet.qpos[:, 3:7] = torch.from_numpy(q).to(fleet.device)
    fleet.qvel[:, :3] = torch.from_numpy(rng.normal(size=(n, 3))).to(fleet.device)
    from ft_grandprix_b200.fleet import LAP
    fleet.lap[:, LAP["laps"]] = torch.from_numpy(rng.integers(0, 5, n).astype(np.int32)).to(fleet.device)
    fleet.lap[:, LAP["completion"]] = torch.from_numpy(rng.integers(0, 100, n).astype(np.int32)).to(fleet.device)
    fleet.lap[:, LAP["good_start"]] = torch.from_numpy(rng.integers(0, 2, n).astype(np.int32)).to(fleet.device)
    fleet.steps = 123
    torch.cuda.synchronize()
    d = fleet.snapshot_tensors(); fleet.sync()
    host = fleet.snapshots()
    for i, sn in enumerate(host):
        assert int(d["laps"][i]) == sn.laps and int(d["lap_completion"][i]) == sn.lap_completion
        assert int(d["absolute_completion"][i]) == sn.absolute_completion
        for k in ("yaw", "pitch", "roll"):
            assert abs(float(d[k][i]) - getattr(sn, k)) < 1e-12
        assert np.array_equal(d["velocity"][i].cpu().numpy(), sn.velocity)
    assert d["time"] == host[0].time
