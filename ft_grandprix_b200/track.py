"""Track-to-geometry compiler front end (host side of SURVEY.md §8 rows a8-a10).

`Track` mirrors what the reference produces with chunk() + produce_mjcf()
(ft_grandprix/chunk.py:10-80, ft_grandprix/map.py:10-72): the metadata dict has the same
keys as rendered/chunks/metadata.json, `chunks` the same [i, j] list in the same order.
`Geometry` is the device-resident world that replaces MjModel for the static part
(ft_grandprix/custom.py:1178).
"""
import ctypes as C
import json
import os

import numpy as np

from . import _lib

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")
BUNDLED = ("track", "circle", "small-circle", "inkscape")

def _p(a):
    return a.ctypes.data_as(C.c_void_p)

def centreline(svg_d, width, height, chunk_width=20, chunk_height=20, scale=2.0, points=100):
    """extract_path_from_svg (curve.py:6-18) + the metre scaling of custom.py:1184-1186."""
    out = np.zeros((points, 2), dtype=np.float64)
    _lib.check(_lib.load().ftgp_centreline(svg_d.encode(), points, width, height, chunk_width, chunk_height,
                                           float(scale), _p(out)), "ftgp_centreline")
    return out

class Track:
    def __init__(self, pixels, scale=2.0, chunk_px=20, name="track", svg_d=None):
        px = np.asarray(pixels)
        if px.dtype == np.bool_:
            px = px.astype(np.uint8)
        if px.dtype != np.uint8 or px.ndim not in (2, 3):
            raise ValueError("pixels must be uint8/bool [H,W] (wall mask) or uint8 [H,W,3|4] (RGB)")
        px = np.ascontiguousarray(px)
        ch = 1 if px.ndim == 2 else px.shape[2]
        # the wall mask itself (pixel is wall iff R+G+B == 765, chunk.py:39-43), kept for the rendered/ emitter (mjcf.py)
        self.wall = px != 0 if px.ndim == 2 else px[:, :, :3].astype(np.int32).sum(2) == 765
        self.height, self.width = px.shape[:2]
        self.scale, self.chunk_px, self.name = float(scale), int(chunk_px), name
        lib = _lib.load()
        self._ptr = lib.ftgp_track_create(_p(px), self.width, self.height, ch, self.scale, self.chunk_px)
        if not self._ptr:
            raise _lib.FtgpError("ftgp_track_create: " + _lib.last_error())
        meta = np.zeros(6, dtype=np.int32); size = np.zeros(2)
        _lib.check(lib.ftgp_track_meta(self._ptr, _p(meta), _p(size)), "ftgp_track_meta")
        self.horizontal_chunks, self.vertical_chunks, self.nchunks = int(meta[0]), int(meta[1]), int(meta[2])
        self.size_x, self.size_y = float(size[0]), float(size[1])
        ij = np.zeros((self.nchunks, 2), dtype=np.int32); counts = np.zeros(self.nchunks, dtype=np.int32)
        _lib.check(lib.ftgp_track_chunks(self._ptr, _p(ij), _p(counts)), "ftgp_track_chunks")
        self.chunks, self.chunk_wall_pixels = ij, counts
        self.svg_d = svg_d
        self.path = None if svg_d is None else centreline(svg_d, self.width, self.height, chunk_px, chunk_px, scale)

    @classmethod
    def bundled(cls, name="track", scale=2.0):
        """One of the reference's template tracks (template/<name>.png + <name>-path.svg),
        frozen as data by tools/make_assets.py."""
        if name not in BUNDLED:
            raise ValueError(f"unknown bundled track {name!r}; have {BUNDLED}")
        z = np.load(os.path.join(ASSETS, "tracks.npz"))
        key = name.replace("-", "_")
        shape = tuple(int(v) for v in z[key + "__shape"])
        wall = np.unpackbits(z[key + "__bits"])[: shape[0] * shape[1]].reshape(shape)
        with open(os.path.join(ASSETS, "paths.json")) as f:
            d = json.load(f)[name]
        return cls(wall, scale=scale, name=name, svg_d=d)

    @classmethod
    def from_png(cls, image_path, svg_path=None, scale=2.0):
        """chunk(image_path, scale=...) for a user-supplied track image (chunk.py:38-43)."""
        from PIL import Image
        rgb = np.array(Image.open(image_path).convert("RGB"))
        d = None
        if svg_path is not None:
            import re
            import xml.etree.ElementTree as ET
            root = ET.parse(svg_path).getroot()
            ns = re.match(r"\{(.+)\}", root.tag)
            ns = "{" + ns.group(1) + "}" if ns else ""
            d = root.find(f"{ns}g").find(f"{ns}path").attrib["d"]
        name = ".".join(os.path.basename(image_path).split(".")[:-1])
        return cls(rgb, scale=scale, name=name, svg_d=d)

    @property
    def metadata(self):
        """Same keys as rendered/chunks/metadata.json (chunk.py:67-79)."""
        return {"original_width": self.width, "original_height": self.height,
                "chunk_width": self.chunk_px, "chunk_height": self.chunk_px,
                "horizontal_chunks": self.horizontal_chunks, "vertical_chunks": self.vertical_chunks,
                "chunks": self.chunks.tolist(), "width": self.width, "height": self.height,
                "name": self.name, "scale": self.scale}

    def start_pose(self, i):
        """position_vehicles for car i (custom.py:1110-1118,1232-1245): (x, y, yaw)."""
        k = (i + 5) * 2
        delta = self.path[k + 1] - self.path[k]
        return float(self.path[k, 0]), float(self.path[k, 1]), float(np.arctan2(delta[1], delta[0]))

    def __del__(self):
        p = getattr(self, "_ptr", None)
        if p:
            try:
                _lib.load().ftgp_track_destroy(p)
            except Exception:        # interpreter shutdown
                pass
            self._ptr = None

class Geometry:
    """Up to 4 compiled tracks packed into one device blob (replicated per GPU)."""
    def __init__(self, tracks, device=0):
        if isinstance(tracks, Track):
            tracks = [tracks]
        self.tracks = list(tracks)
        n = len(self.tracks)
        arr = (C.c_void_p * n)(*[t._ptr for t in self.tracks])
        self._paths = [None if t.path is None else np.ascontiguousarray(t.path, dtype=np.float64) for t in self.tracks]
        parr = (C.c_void_p * n)(*[None if p is None else p.ctypes.data for p in self._paths])
        self.device = int(device)
        self._ptr = _lib.load().ftgp_geom_create(arr, parr, n, self.device)
        if not self._ptr:
            raise _lib.FtgpError("ftgp_geom_create: " + _lib.last_error())
        self.nbytes = int(_lib.load().ftgp_geom_bytes(self._ptr))

    def __del__(self):
        p = getattr(self, "_ptr", None)
        if p:
            try:
                _lib.load().ftgp_geom_destroy(p)
            except Exception:        # interpreter shutdown
                pass
            self._ptr = None
