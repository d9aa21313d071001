#!/usr/bin/env python
"""Derive the bundled track assets from the reference's template/ directory.

Runs only in the build container (needs /root/reference).  The GPU box has no
reference tree, so the inputs of the track compiler (a8) are
frozen here as data:

  ft_grandprix_b200/assets/tracks.npz   per track: RGB wall mask, bit-packed
                                        (pixel is wall iff R+G+B == 765,
                                        ft_grandprix/chunk.py:39-43)
  ft_grandprix_b200/assets/paths.json   per track: the `d` attribute of the first
                                        <g><path> of template/<track>-path.svg
                                        (ft_grandprix/curve.py:11-14)
  ft_grandprix_b200/assets/meshes.npz   triangles + facet normals (float32) of
                                        template/meshes/{simple_base_nano,mushr_wheel}.stl
                                        (mushr.em.xml:38-39), re-serialised as binary STL by
                                        the rendered/ emitter (ft_grandprix_b200/mjcf.py)

No reference source code is copied, only the image/SVG *data* the reference
ships as race inputs.
"""
import json, os, re, sys
import xml.etree.ElementTree as ET
import numpy as np
from PIL import Image

REF = os.environ.get("FTGP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ft_grandprix_b200", "assets")
TRACKS = ["track", "circle", "small-circle", "inkscape"]

def main():
    arrays, paths = {}, {}
    for name in TRACKS:
        img = Image.open(os.path.join(REF, "template", name + ".png")).convert("RGB")
        rgb = np.array(img)
        wall = rgb.sum(2) == 255 * 3                      # chunk.py:39-43
        key = name.replace("-", "_")
        arrays[key + "__bits"] = np.packbits(wall, axis=None)
        arrays[key + "__shape"] = np.array(wall.shape, dtype=np.int64)  # (height, width)
        root = ET.parse(os.path.join(REF, "template", name + "-path.svg")).getroot()
        ns = re.match(r"\{(.+)\}", root.tag)
        ns = "{" + ns.group(1) + "}" if ns else ""
        paths[name] = root.find(f"{ns}g").find(f"{ns}path").attrib["d"]   # curve.py:14
        print(name, wall.shape, int(wall.sum()), "wall px", file=sys.stderr)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "tracks.npz"), **arrays)
    with open(os.path.join(OUT, "paths.json"), "w") as f:
        json.dump(paths, f, indent=1)
    meshes = {}
    for name in ("simple_base_nano", "mushr_wheel"):
        b = open(os.path.join(REF, "template", "meshes", name + ".stl"), "rb").read()
        n = int.from_bytes(b[80:84], "little")
        assert 84 + 50 * n == len(b), "binary STL expected"
        rec = np.frombuffer(b[84:], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
        meshes[name + "__tri"] = rec["v"].copy(); meshes[name + "__nrm"] = rec["n"].copy()
        print(name, n, "triangles", file=sys.stderr)
    np.savez_compressed(os.path.join(OUT, "meshes.npz"), **meshes)

if __name__ == "__main__":
    main()
