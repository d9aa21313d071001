#!/usr/bin/env python
"""Generate golden vectors by IMPORTING the reference (build container only).

    python tests/golden/make_golden.py

Writes, next to this file:
  chunks.json   ft_grandprix.chunk.chunk() on the four bundled tracks at scale=2.0:
                metadata (minus paths) + per-chunk wall-pixel count + sha1 over the
                chunk PNGs' first channel, read back from rendered/chunks/*.png.
  drivers.npz   ft_grandprix.nidc.Driver / ft_grandprix.fast.Driver called on seeded
                scans (including the -1 / 0 / tie corner cases of SURVEY.md App. D).
The reference tree is not present on the GPU box; tests read only these files.
"""
import hashlib, io, json, os, sys, tempfile, contextlib
import numpy as np

REF = os.environ.get("FTGP_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

def gen_chunks():
    from PIL import Image
    from ft_grandprix.chunk import chunk
    out = {}
    for name in ["track", "circle", "small-circle", "inkscape"]:
        with tempfile.TemporaryDirectory() as tmp:
            cwd = os.getcwd()
            os.chdir(tmp)
            try:
                with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                    chunk(os.path.join(REF, "template", name + ".png"), verbose=False, force=True, scale=2.0)
                meta = json.load(open("rendered/chunks/metadata.json"))
                counts, h = [], hashlib.sha1()
                for (i, j) in meta["chunks"]:
                    a = np.array(Image.open(f"rendered/chunks/{i:03}x{j:03}.png").convert("RGB"))
                    assert set(np.unique(a)) <= {0, 255}
                    counts.append([int((a[:, :, 0] == 255).sum()), a.shape[1], a.shape[0]])
                    h.update(np.ascontiguousarray(a[:, :, 0]).tobytes())
            finally:
                os.chdir(cwd)
        meta["chunk_counts_w_h"] = counts
        meta["chunks_sha1"] = h.hexdigest()
        out[name] = meta
    json.dump(out, open(os.path.join(HERE, "chunks.json"), "w"))

def scans(rng):
    s = []
    s.append(np.zeros(90))
    a = np.full(90, 2.0); a[50:60] = 6.0; s.append(a)
    a = np.full(90, 2.0); a[45:] = 5.0; s.append(a)
    for k in range(400):
        kind = k % 8
        base = rng.uniform(0.2, 6.0, 90)
        if kind == 0:      # smooth corridor-like
            th = np.radians(4 * np.arange(90) - 90)
            base = 0.6 / np.maximum(np.abs(np.sin(th + rng.normal(0, .3))), 0.05) + rng.uniform(0, .05, 90)
        elif kind == 1:    # blocks of constant range -> many disparities and ties
            base = np.repeat(rng.uniform(0.1, 8.0, 15), 6)
        elif kind == 2:    # misses
            base[rng.random(90) < 0.15] = -1.0
        elif kind == 3:    # zeros (first tick / touching)
            base[rng.random(90) < 0.1] = 0.0
        elif kind == 4:    # quantised -> exact ties
            base = np.round(base * 2) / 2
        elif kind == 5:    # long ranges
            base = rng.uniform(0.05, 30.0, 90)
        elif kind == 6:    # very close walls: huge cover counts
            base = rng.uniform(0.01, 0.3, 90); base[rng.integers(0, 90, 5)] = 5.0
        s.append(base.astype(np.float64))
    return np.stack(s)

def gen_drivers():
    from ft_grandprix.nidc import Driver as Nidc
    from ft_grandprix.fast import Driver as Fast
    rng = np.random.default_rng(12345)
    S = scans(rng)
    out = {"scans": S}
    for name, cls in (("nidc", Nidc), ("fast", Fast)):
        d = cls()
        res = np.zeros((len(S), 2))
        with np.errstate(all="ignore"):
            for i, r in enumerate(S):
                sp, st = d.process_lidar(r.copy())
                res[i] = (float(sp), float(st))
        out[name] = res
    np.savez_compressed(os.path.join(HERE, "drivers.npz"), **out)

if __name__ == "__main__":
    import warnings; warnings.simplefilter("ignore")
    gen_chunks(); gen_drivers()
    print("golden written to", HERE)
