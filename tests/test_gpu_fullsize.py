"""GPU, at BASELINE.json's full sizes: properties that do not need the oracle to finish a 65,536- or 1,048,576-car
run -- permutation invariance (the step kernel regroups cars by Newton iteration count and keeps converged cars of a
CTA idling: neither may change a single bit), replication (a fleet made of copies of a block gives copies of the
block's results), determinism, unit quaternions and sane ranges.  The small-size tests compare with the oracle;
these extend that parity to the sizes the benchmark runs at."""
import numpy as np
import pytest

from conftest import random_poses

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ft():
    import ft_grandprix_b200 as ft
    return ft


def _start(t, n, seed):
    poses = random_poses(t.path, n, seed=seed, level=True)
    return poses[:, :2], 2 * np.arctan2(poses[:, 6], poses[:, 3])


def test_full_tick_65536_cars_is_permutation_invariant_and_deterministic(ft):
    """BASELINE config 3 size.  Fleet B holds the cars of fleet A in a random order: after 60 ticks every array of B
    is A's, permuted, bit for bit; a second run of A reproduces A."""
    t = ft.Track.bundled("track")
    n, ticks = 65536, 60
    xy, yaw = _start(t, n, seed=21)
    perm = np.random.default_rng(22).permutation(n)
    a = ft.Fleet(t, n); a.reset(xy, yaw); a.tick(ticks); a.sync()
    b = ft.Fleet(t, n); b.reset(xy[perm], yaw[perm]); b.tick(ticks); b.sync()
    c = ft.Fleet(t, n); c.reset(xy, yaw); c.tick(ticks); c.sync()
    p = torch.from_numpy(perm).to(a.device)
    for name in ("qpos", "qvel", "warm", "ctrl", "ranges", "lap", "times", "status"):
        A, B, Cc = getattr(a, name), getattr(b, name), getattr(c, name)
        assert torch.equal(A[p], B), name
        assert torch.equal(A, Cc), name
    q = a.qpos
    assert torch.isfinite(q).all() and torch.isfinite(a.qvel).all()
    for lo in (3, 11, 18, 24, 30):                                   # free joint + four ball joints stay unit quaternions
        assert (q[:, lo:lo + 4].norm(dim=1) - 1).abs().max() < 1e-9
    r = a.ranges
    assert ((r == -1) | ((r >= 0) & (r < 450))).all()                # -1 = no hit, else a wall or the 300 m ground square
    st = a.status
    assert int((st & 0xFF).max()) < 100 and int(((st >> 8) & 1).sum()) == 0      # Newton converged everywhere, no resets
    assert float(((st >> 24) & 0xF).float().mean()) > 3.0            # the cars are on their wheels (still settling from the spawn pop-up)


def test_lidar_1048576_cars_replicated_blocks(ft):
    """BASELINE config 4 size, mixed tracks: 16 copies of a 65,536-pose block on circle / small-circle give 16 copies
    of the block's ranges (checksum of checksums), and the block itself matches a 65,536-car fleet."""
    tracks = [ft.Track.bundled("circle"), ft.Track.bundled("small-circle")]
    from ft_grandprix_b200.track import Geometry
    nb, copies = 65536, 16
    tid = (np.arange(nb) % 2).astype(np.int32)
    poses = np.zeros((nb, 7))
    for k in range(2):
        sel = np.nonzero(tid == k)[0]
        poses[sel] = random_poses(tracks[k].path, len(sel), seed=30 + k)
    geom = Geometry(tracks, device=0)
    small = ft.Fleet(geom, nb, track_id=tid)
    small.qpos[:, :7] = torch.from_numpy(poses).to(small.device)
    torch.cuda.synchronize()
    small.lidar(); small.sync()
    ref = small.ranges.clone()
    big = ft.Fleet(geom, nb * copies, track_id=np.tile(tid, copies))
    big.qpos[:, :7] = torch.from_numpy(np.tile(poses, (copies, 1))).to(big.device)
    torch.cuda.synchronize()
    got = big.lidar(); big.sync()
    got = got.view(copies, nb, 90)
    sums = got.double().sum(dim=(1, 2))
    assert (sums == sums[0]).all()
    assert torch.equal(got[0], ref) and torch.equal(got[copies - 1], ref)
    assert float((ref >= 0).float().mean()) > 0.99                   # on-track poses see walls


def test_step_1048576_cars_matches_65536_block(ft):
    """The vehicle step at config 4 size: 16 copies of a driven 65,536-car state advance exactly like the block."""
    t = ft.Track.bundled("track")
    nb, copies = 65536, 16
    xy, yaw = _start(t, nb, seed=41)
    small = ft.Fleet(t, nb); small.reset(xy, yaw); small.tick(25); small.sync()
    big = ft.Fleet(t, nb * copies)
    for name in ("qpos", "qvel", "warm", "ctrl"):
        getattr(big, name).copy_(getattr(small, name).repeat(copies, 1))
    torch.cuda.synchronize()
    small.step(3); small.sync()
    big.step(3); big.sync()
    for name in ("qpos", "qvel", "warm"):
        B = getattr(big, name).view(copies, nb, -1)
        assert torch.equal(B[0], getattr(small, name)) and torch.equal(B[copies - 1], getattr(small, name)), name
