"""CPU: the MJCF / rendered/ emitter (ft_grandprix_b200/mjcf.py) against goldens expanded from the
reference's own template files and produced by the reference's own chunk() (tests/golden/make_mjcf_golden.py)."""
import gzip
import hashlib
import json
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from conftest import GOLDEN


def canonical(xml_text):
    """Same canonical form as tests/golden/make_mjcf_golden.py: one line per element, sorted attributes."""
    root = ET.fromstring(xml_text)
    lines = []

    def walk(e, d):
        attrs = " ".join(f'{k}="{" ".join(v.split())}"' for k, v in sorted(e.attrib.items()))
        lines.append(f"{'  ' * d}{e.tag} {attrs}".rstrip())
        for c in e:
            walk(c, d + 1)
    walk(root, 0)
    return "\n".join(lines) + "\n"


CARS3 = [{"driver": "ft_grandprix.nidc", "name": "red car", "primary": "red", "secondary": "pink", "icon": "white.png"},
         {"driver": "ft_grandprix.fast", "name": "orange car", "primary": "orange", "secondary": "darkorange", "icon": "white.png"},
         {"driver": "ft_grandprix.nidc", "name": "green car", "primary": "blue", "secondary": "green", "icon": "white.png"}]
CARS7 = [dict(c, driver=f"file://ft_grandprix/{c['driver'].split('.')[-1]}.py") for c in CARS3] + [
    {"driver": "file://ft_grandprix/nidc.py", "name": "trinity car", "primary": "white", "secondary": "lightblue", "icon": "trinity.png"},
    {"driver": "file://ft_grandprix/nidc.py", "name": "maynooth car", "primary": "brown", "secondary": "maroon", "icon": "nuim.png"},
    {"driver": "file://ft_grandprix/nidc.py", "name": "TU car", "primary": "rgb(2, 109, 153)", "secondary": "blue", "icon": "tu.png"},
    {"driver": "file://ft_grandprix/nidc.py", "name": "purple car", "primary": "purple", "secondary": "yellow", "icon": "white.png"}]


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(GOLDEN, "mjcf_golden.json")))


@pytest.mark.parametrize("template,cars_name,track", [("mushr", "cars", "track"), ("car", "cars", "track"),
                                                      ("mushr", "all", "track"), ("mushr", "cars", "circle")])
def test_car_xml_equals_the_expanded_reference_template(tmp_path, golden, template, cars_name, track):
    import ft_grandprix_b200 as ft
    from ft_grandprix_b200 import mjcf
    t = ft.Track.bundled(track)
    cars = CARS3 if cars_name == "cars" else CARS7
    xml_path = mjcf.produce_mjcf(cars, t.metadata, str(tmp_path), rangefinders=90, map_color=[1, 0, 0], tricycle=(template == "car"))
    got = canonical(open(xml_path).read())
    name = f"mjcf_{template}_{cars_name}_{track}.txt.gz"
    want = gzip.open(os.path.join(GOLDEN, name)).read().decode()
    if got != want:                                                  # show the first difference, not 1 800 lines
        for k, (a, b) in enumerate(zip(got.splitlines(), want.splitlines())):
            assert a == b, f"line {k}"
    assert got == want
    assert hashlib.sha256(got.encode()).hexdigest() == golden["mjcf"][name]["sha256"]
    assert json.load(open(tmp_path / "car.json")) == golden["mjcf"][name]["car_json"]      # map.py:67-72
    # a runnable rendered/ directory: the files car.xml names exist
    for rel in {e.attrib["file"] for e in ET.parse(xml_path).getroot().iter() if "file" in e.attrib and not e.attrib["file"].startswith("chunks/")}:
        rel = rel if template == "mushr" else os.path.join("icons", rel)              # car.em.xml: texturedir="icons/"
        assert os.path.exists(tmp_path / rel), rel


@pytest.mark.parametrize("track", ["track", "circle"])
def test_rendered_chunks_are_byte_identical_to_chunk_py(tmp_path, golden, track):
    """chunks/*.png and metadata.json byte for byte as ft_grandprix.chunk.chunk(scale=2.0) writes them."""
    import ft_grandprix_b200 as ft
    from ft_grandprix_b200 import mjcf
    t = ft.Track.bundled(track)
    out = tmp_path / "chunks"
    meta = mjcf.write_chunks(t.wall, str(out), name=t.name, scale=t.scale)
    want = golden["chunks"][track]
    files = sorted(os.listdir(out))
    assert len(files) == want["nfiles"]
    assert hashlib.sha256(open(out / "metadata.json", "rb").read()).hexdigest() == want["metadata_sha256"]
    blob = b"".join(f.encode() + open(out / f, "rb").read() for f in files if f.endswith(".png"))
    assert hashlib.sha256(blob).hexdigest() == want["sha256_of_all_pngs"]
    assert meta == t.metadata                                        # the product track compiler agrees with the emitter


def test_stl_writer_round_trips_the_chassis_mesh(tmp_path):
    from ft_grandprix_b200 import mjcf
    z = np.load(os.path.join(mjcf.ASSETS, "meshes.npz"))
    p = tmp_path / "m.stl"
    mjcf.write_stl(str(p), z["simple_base_nano__tri"], z["simple_base_nano__nrm"])
    b = open(p, "rb").read()
    n = int.from_bytes(b[80:84], "little")
    assert n == 33 and len(b) == 84 + 50 * n
    rec = np.frombuffer(b[84:], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
    assert np.array_equal(rec["v"], z["simple_base_nano__tri"])


def test_render_world_cli_layout(tmp_path):
    """render_world = stage() up to the MuJoCo compile: chunks/, car.xml, car.json, meshes/, icons/."""
    import ft_grandprix_b200 as ft
    from ft_grandprix_b200 import mjcf
    t = ft.Track.bundled("small-circle")
    mjcf.render_world(t, CARS3[:1], str(tmp_path / "rendered"))
    got = set(os.listdir(tmp_path / "rendered"))
    assert {"chunks", "car.xml", "car.json", "meshes", "icons"} <= got
    root = ET.parse(tmp_path / "rendered" / "car.xml").getroot()
    assert len(root.find("asset").findall("hfield")) == t.nchunks == 330
    assert len(root.find("sensor").findall("rangefinder")) == 90
