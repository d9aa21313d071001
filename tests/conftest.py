import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement (oracle/*.c) -- test infrastructure, never imported by the package."""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def walls():
    """Wall masks + SVG path data of the four bundled tracks (frozen by tools/make_assets.py)."""
    import json
    z = np.load(os.path.join(ROOT, "ft_grandprix_b200", "assets", "tracks.npz"))
    paths = json.load(open(os.path.join(ROOT, "ft_grandprix_b200", "assets", "paths.json")))
    out = {}
    for name in ("track", "circle", "small-circle", "inkscape"):
        key = name.replace("-", "_")
        shape = tuple(int(v) for v in z[key + "__shape"])
        out[name] = (np.unpackbits(z[key + "__bits"])[: shape[0] * shape[1]].reshape(shape), paths[name])
    return out


@pytest.fixture(scope="session")
def otracks(oracle, walls):
    return {name: oracle.Track(w) for name, (w, _) in walls.items()}


def random_poses(path, n, seed, level=False):
    """BASELINE config 2 poses (SURVEY §8 d): on-track positions with jitter."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, 100, n)
    nxt = (idx + 1) % 100
    heading = np.arctan2(path[nxt, 1] - path[idx, 1], path[nxt, 0] - path[idx, 0])
    xy = path[idx] + rng.normal(0, 0.10, (n, 2))
    yaw = heading + rng.normal(0, 0.3, n)
    z = np.zeros(n) if level else 0.0156 + rng.uniform(-0.002, 0.002, n)
    roll = np.zeros(n) if level else rng.normal(0, 0.01, n)
    pitch = np.zeros(n) if level else rng.normal(0, 0.01, n)
    cy, sy, cp, sp, cr, sr = np.cos(yaw / 2), np.sin(yaw / 2), np.cos(pitch / 2), np.sin(pitch / 2), np.cos(roll / 2), np.sin(roll / 2)
    q = np.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                  cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], 1)
    return np.concatenate([xy, z[:, None], q], 1)
