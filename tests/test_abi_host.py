"""CPU: the C-ABI library loads, exports every symbol include/ftgp.h declares, and its host-side
entry points (track compiler, centreline) agree with the oracle and the reference goldens.
No device compute is called here."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ftgp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ftgp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ft_grandprix_b200 import _lib
    lib = C.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"libftgp.so does not export {s}"
    # and the ctypes table mirrors the header one to one
    assert sorted(_lib.SIGNATURES) == syms


def test_no_oracle_in_product():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "ft_grandprix_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "ftgp_oracle" not in txt and "fto_" not in txt, f


def test_device_calls_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ft_grandprix_b200 as ft
    t = ft.Track.bundled("small-circle")
    with pytest.raises(ft._lib.FtgpError):
        ft.Geometry(t, device=0)
    with pytest.raises(ft._lib.FtgpError):
        ft.Fleet(t, 4)


@pytest.mark.parametrize("name", ["track", "circle", "small-circle", "inkscape"])
def test_product_track_compiler_matches_chunk_py(name):
    import ft_grandprix_b200 as ft
    g = json.load(open(os.path.join(GOLDEN, "chunks.json")))[name]
    t = ft.Track.bundled(name)
    md = t.metadata
    for key in ("original_width", "original_height", "chunk_width", "chunk_height", "horizontal_chunks",
                "vertical_chunks", "chunks", "width", "height", "scale"):
        assert md[key] == g[key], key
    assert t.chunk_wall_pixels.tolist() == [c[0] for c in g["chunk_counts_w_h"]]


def test_product_centreline_matches_oracle(otracks, walls):
    import ft_grandprix_b200 as ft
    for name in ("track", "circle", "small-circle", "inkscape"):
        t = ft.Track.bundled(name)
        np.testing.assert_array_equal(t.path, otracks[name].centreline(walls[name][1]))


def test_rgb_threshold_rule():
    """chunk.py:39-43: wall iff R+G+B == 765; greys are not wall."""
    import ft_grandprix_b200 as ft
    rgb = np.zeros((40, 40, 3), dtype=np.uint8)
    rgb[5, 5] = 255
    rgb[25, 25] = (255, 255, 254)
    rgb[30, 10] = 200
    t = ft.Track(rgb)
    assert t.chunks.tolist() == [[0, 0]] and t.chunk_wall_pixels.tolist() == [1]
    with pytest.raises(ft._lib.FtgpError):
        ft.Track(np.zeros((10, 10, 2), dtype=np.uint8))


def test_batched_torch_drivers_match_reference_classes():
    import torch
    import ft_grandprix_b200 as ft
    z = np.load(os.path.join(GOLDEN, "drivers.npz"))
    S = torch.from_numpy(z["scans"])
    for cls, key in ((ft.BatchedNidcDriver, "nidc"), (ft.BatchedFastDriver, "fast")):
        sp, st = cls().process_lidar(S)
        np.testing.assert_array_equal(sp.numpy(), z[key][:, 0])
        np.testing.assert_array_equal(st.numpy(), z[key][:, 1])


def test_headless_runner_host_logic():
    """ft_grandprix_b200.run without a GPU: ordinals (custom.py:47-55), driver loading rules (custom.py:1097-1109) and
    the dashboard lines (custom.py:335-361) from a stand-in fleet."""
    import torch
    from ft_grandprix_b200 import run as runner
    from ft_grandprix_b200.fleet import DRIVER_KINDS, LAP
    from ft_grandprix_b200._lib import LAP_FIELDS, MAX_LAPTIMES
    assert [runner.ordinal(n) for n in (1, 2, 3, 4, 11, 12, 13, 21, 22, 23, 101, 111)] == \
        ["1st", "2nd", "3rd", "4th", "11th", "12th", "13th", "21st", "22nd", "23rd", "101st", "111th"]
    kind, obj, path = runner.load_driver("ft_grandprix.nidc")
    assert kind == DRIVER_KINDS["nidc"] and obj is None
    kind, obj, path = runner.load_driver("file://" + os.path.join(ROOT, "tests", "drivers", "slowpoke.py"))
    assert kind is None and obj.process_lidar(np.full(90, 2.0))[0] == 0.8
    kind, obj, path = runner.load_driver("no.such.module")
    assert kind == DRIVER_KINDS["lobotomy"] and obj is None           # import failure -> inert driver
    kind, obj, path = runner.load_driver("http://example.org/x.py")
    assert kind == DRIVER_KINDS["lobotomy"]

    class Stand:
        cars_per_world = 3
        lap = torch.zeros(3, len(LAP_FIELDS), dtype=torch.int32)
        times = torch.zeros(3, MAX_LAPTIMES, dtype=torch.int32)
    f = Stand()
    f.lap[:, LAP["good_start"]] = torch.tensor([1, 1, 0], dtype=torch.int32)
    f.lap[:, LAP["completion"]] = torch.tensor([40, 10, 97], dtype=torch.int32)
    f.lap[:, LAP["laps"]] = torch.tensor([0, 1, 0], dtype=torch.int32)
    f.lap[1, LAP["finished"]] = 1; f.lap[1, LAP["ntimes"]] = 1; f.times[1, 0] = 12500
    lines = runner.dashboard(f, ["red car", "orange car", "green car"])
    assert lines[0].strip().startswith("1st  Car #1 - orange car: Car finished 1 laps in 50.00 seconds!") and "[50.00]" in lines[0]
    assert lines[1].strip().startswith("2nd  Car #0 - red car: Laps: 0  Completion: 40%")
    assert lines[2].strip().startswith("3rd  Car #2 - green car: Laps: 0  Completion: -3%")      # going backwards (custom.py:132-140)
