"""CPU: pin the oracle (oracle/*.c) against everything the reference lets us generate here
(tests/golden/, made by tests/golden/make_golden.py importing /root/reference)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def chunks_golden():
    return json.load(open(os.path.join(GOLDEN, "chunks.json")))


@pytest.mark.parametrize("name", ["track", "circle", "small-circle", "inkscape"])
def test_track_compiler_matches_chunk_py(name, otracks, walls, chunks_golden):
    """a8: same chunk list, same order, same per-chunk wall pixels as ft_grandprix.chunk.chunk()."""
    g = chunks_golden[name]
    t = otracks[name]
    assert (t.hc, t.vc) == (g["horizontal_chunks"], g["vertical_chunks"])
    assert t.nchunks == len(g["chunks"])
    assert t.chunks().tolist() == g["chunks"]
    wall = walls[name][0]
    assert (wall.shape[1], wall.shape[0]) == (g["width"], g["height"])
    for k in range(0, t.nchunks, 7):
        d = t.chunk_data(k)
        cnt, w, h = g["chunk_counts_w_h"][k]
        assert d.shape == (h, w)
        assert int(d.sum()) == cnt
        i, j = g["chunks"][k]
        # MuJoCo flips the rows of a PNG hfield: hfield row 0 is the image's bottom row
        np.testing.assert_array_equal(d[::-1], wall[j * 20:j * 20 + h, i * 20:i * 20 + w].astype(np.float32))


def test_survey_chunk_counts(otracks):
    assert {k: t.nchunks for k, t in otracks.items()} == {"track": 483, "circle": 506, "small-circle": 330, "inkscape": 746}


def test_drivers_match_reference_classes(oracle):
    """a5: fto_driver == ft_grandprix.nidc.Driver / ft_grandprix.fast.Driver on 403 scans."""
    z = np.load(os.path.join(GOLDEN, "drivers.npz"))
    for kind, key in ((0, "nidc"), (1, "fast")):
        for s, want in zip(z["scans"], z[key]):
            got = oracle.driver(kind, s)
            assert got is not None
            assert got == (want[0], want[1]), (key, got, want)


def test_driver_known_answers(oracle):
    """SURVEY Appendix D known answers."""
    assert oracle.driver(0, np.zeros(90)) == (1.249365981851197, -1.5707963267948966)
    assert oracle.driver(1, np.zeros(90)) == (1.25, -1.5707963267948966)
    a = np.full(90, 2.0); a[50:60] = 6.0
    assert oracle.driver(0, a) == (2.1109138610203724, 0.4886921905584123)
    assert oracle.driver(1, a) == (2.0, 0.4188790204786391)
    assert oracle.driver(2, a) == (0.0, 0.0)
    a = np.full(90, 2.0); a[10] = np.nan; a[40] = np.nan
    # NaN never forms a disparity (nan > 0.6 is False); np.argmax returns the first NaN (proc index 29):
    # value obtained from ft_grandprix.nidc.Driver on the same scan
    assert oracle.driver(0, a) == (2.2220813293002664, -0.3490658503988659)


def _py_lap(state, times, path, xy, steps, lap_target, winners, car_id):
    """Line-by-line replay of ft_grandprix/custom.py:1340-1372 on a dict."""
    distances = ((path - xy) ** 2).sum(1)
    closest = distances.argmin()
    state["off_track"] = bool(distances[closest] > 1)
    if not state["off_track"]:
        completion = (closest - state["offset"]) % 100
        delta = completion - state["completion"]
        state["delta"] = (completion - state["completion"] + 50) % 100 - 50
        if abs(delta) > 90:
            lap_time = steps - state["start"]
            if state["delta"] < 0:
                state["laps"] -= 1
                state["good_start"] = False
                if len(times) != 0:
                    times.pop()
            elif state["delta"] > 0:
                if state["good_start"]:
                    times.append(lap_time)
                    state["start"] = steps
                state["laps"] += 1
                state["good_start"] = True
        if state["laps"] >= lap_target:
            if car_id not in winners:
                winners[car_id] = len(winners) + 1
            state["finished"] = True
        state["completion"] = completion


def test_lap_logic_matches_python_replay(oracle, otracks, walls):
    path = otracks["track"].centreline(walls["track"][1])
    rng = np.random.default_rng(5)
    ncars = 6
    laps = [oracle.Lap(offset=(i + 5) * 2) for i in range(ncars)]
    ref = [dict(offset=(i + 5) * 2, completion=0, laps=0, start=0, good_start=True, finished=False, delta=0,
                off_track=False) for i in range(ncars)]
    rtimes = [[] for _ in range(ncars)]
    winners, nwin = {}, 0
    pos = np.array([(i + 5) * 2 for i in range(ncars)], dtype=float)
    for step in range(4000):
        # cars run forwards at different speeds, sometimes reverse, sometimes leave the track
        vel = np.array([0.9, 0.6, 0.45, -0.3, 0.75, 0.2]) + (rng.random(ncars) < 0.01) * rng.normal(0, 30, ncars)
        pos += vel
        for i in range(ncars):
            k = int(np.floor(pos[i])) % 100
            xy = path[k] + rng.normal(0, 0.05, 2) + (rng.random() < 0.02) * np.array([3.0, 3.0])
            nwin = laps[i].update(path, xy, step, 3, nwin)
            _py_lap(ref[i], rtimes[i], path, xy, step, 3, winners, i)
            s = laps[i].s
            assert (s.completion, s.laps, s.start, bool(s.good_start), bool(s.finished), s.delta, bool(s.off_track)) == \
                   (ref[i]["completion"], ref[i]["laps"], ref[i]["start"], ref[i]["good_start"], ref[i]["finished"],
                    ref[i]["delta"], ref[i]["off_track"]), (step, i)
            assert s.ntimes == len(rtimes[i])
            assert list(laps[i].times[: min(s.ntimes, 32)]) == rtimes[i][:32]
            assert s.rank == winners.get(i, 0)
    assert len(winners) >= 3 and any(r["laps"] < 0 for r in ref)


def test_centreline_properties(otracks, walls):
    """a10: svg.path restated (unpinned: svg.path is absent) -- lap lengths from SURVEY C.3."""
    for name in ("track", "circle", "small-circle", "inkscape"):
        p = otracks[name].centreline(walls[name][1])
        assert p.shape == (100, 2)
        assert (p[:, 0] > 0).all() and (p[:, 0] < 40).all() and (p[:, 1] < 0).all() and (p[:, 1] > -40).all()
        seg = np.linalg.norm(np.diff(np.vstack([p, p[:1]]), axis=0), axis=1)
        assert seg.max() < 2.5 and seg.sum() > 20
        if name == "track":
            assert abs(seg.sum() - 76.7) < 0.6, seg.sum()
            assert 0.2 < seg.min() and seg.max() < 1.1
    p = otracks["track"].centreline(walls["track"][1])
    # first point = the path's `m x,y` scaled by px/W*40 (custom.py:1185-1186)
    np.testing.assert_allclose(p[0], [763.16038 / 1600 * 40, -494.01885 / 1600 * 40], rtol=1e-12)


def test_ray_unit_cases(oracle, otracks):
    t = otracks["track"]
    # straight down onto the ground plane (z = 0.01), away from walls
    assert abs(t.ray([100.0, 100.0, 1.01], [0, 0, -1.0]) - 1.0) < 1e-12
    # upwards: nothing
    assert t.ray([100.0, 100.0, 1.0], [0, 0, 1.0]) == -1
    # plane is only hit inside its 300 m half-size
    assert t.ray([0, 0, 1.0], [1.0, 0, -1e-3]) == -1
    # a level ray from far outside towards the map hits a wall before crossing it
    d = t.ray([-5.0, -20.0, 0.08], [1.0, 0.0, 0.0])
    assert 5 < d < 45


def test_ray_wall_slope_geometry(oracle):
    """a wall vertex column next to free vertices gives a slope of 0.3 m per cell; a level
    ray at height z meets it (z + 0.1) / 0.3 of a cell before the wall vertex."""
    wall = np.zeros((40, 40), dtype=np.uint8)
    wall[:, 30:32] = 1                     # vertical stroke in chunk (1, *), columns 10-11 of the chunk
    t = oracle.Track(wall)
    dx = t.size_x / 19
    xw = t.size_x * (1 - 0.5 + 10 / 19)    # world x of the first wall vertex column
    for z in (0.02, 0.08, 0.15):
        d = t.ray([0.0, -0.3, z], [1.0, 0.0, 0.0])
        want = xw - dx * (1 - (z + 0.1) / 0.3)
        assert abs(d - want) < 1e-9, (z, d, want)
    # above the wall top (0.2): miss
    assert t.ray([0.0, -0.3, 0.25], [1.0, 0.0, 0.0]) == -1


def test_centreline_against_independent_python_restatement(otracks, walls):
    """a10 / svg.path 6.3 semantics restated a second time, independently, in Python: parse `m x,y c ... z` (relative
    cubics + closing line), segment lengths by dense numerical integration (the C code subdivides recursively),
    Path.point(t) = bisect over the cumulative length fractions, then the segment's OWN Bezier parameter, 100 samples at
    t = i/100, scaled by px / W * 40 and -px / H * 40 (custom.py:1185-1186).  svg.path itself is absent, so this pins
    the C restatement against a second reading of the same published behaviour, not against the library."""
    import bisect
    import re
    for name in ("track", "circle", "small-circle", "inkscape"):
        wall, d = walls[name]
        H, W = wall.shape
        tok = re.findall(r"[a-zA-Z]|-?\d*\.?\d+(?:[eE][-+]?\d+)?", d)
        segs, i, cmd = [], 0, None
        cur = start = None
        num = lambda k: complex(float(tok[k]), float(tok[k + 1]))
        while i < len(tok):
            if tok[i].isalpha():
                cmd = tok[i]; i += 1
                if cmd in "zZ":
                    if abs(cur - start) > 0:
                        segs.append(("L", cur, start))
                    cur = start
                continue
            rel = cmd.islower()
            if cmd in "mM":
                pt = num(i); i += 2
                cur = (cur + pt) if (rel and cur is not None) else pt
                start = cur
                cmd = "l" if rel else "L"                           # further pairs after a move are implicit line-tos
            elif cmd in "cC":
                base = cur if rel else 0
                c1, c2, e = base + num(i), base + num(i + 2), base + num(i + 4); i += 6
                segs.append(("C", cur, c1, c2, e)); cur = e
            elif cmd in "lL":
                e = (cur if rel else 0) + num(i); i += 2
                segs.append(("L", cur, e)); cur = e
            else:
                raise AssertionError(f"unexpected path command {cmd}")

        def point(seg, u):
            if seg[0] == "L":
                return seg[1] + (seg[2] - seg[1]) * u
            _, p0, p1, p2, p3 = seg
            return (1 - u) ** 3 * p0 + 3 * (1 - u) ** 2 * u * p1 + 3 * (1 - u) * u * u * p2 + u ** 3 * p3
        us = np.linspace(0, 1, 200001)
        lengths = [0.0] + [float(np.abs(np.diff(point(sg, us))).sum()) for sg in segs]      # the Move has length 0
        frac = np.cumsum(lengths) / sum(lengths)
        all_segs = [("L", start, start)] + segs
        pts = []
        for k in range(100):
            t = k / 100
            if t == 0:
                z = start
            else:
                j = bisect.bisect(list(frac), t)
                u = t / frac[0] if j == 0 else (t - frac[j - 1]) / (frac[j] - frac[j - 1])
                z = point(all_segs[j], u)
            pts.append([z.real / W * 40, -z.imag / H * 40])
        got = otracks[name].centreline(d)
        assert np.abs(got - np.array(pts)).max() < 2e-6, (name, np.abs(got - np.array(pts)).max())
