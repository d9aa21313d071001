"""Run by tests/test_step_cpu.py in a subprocess (killed on time-out): four cars near WALLS stepped as one warp of four
quads, every collective a barrier over all 16 host threads, where some quads step a SHADOWED car (no walls) and the
option bubble_wrap is on.  A collective that only the quads with walls reach (the bug that hung a B200 for 25 minutes in
round 2: `walls.enabled() && qd.any(...)`) dead-locks this script.  The results must equal each car stepped alone."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, sys.argv[1])
from oracle import pyoracle                                    # noqa: E402  (test infrastructure)
import ft_grandprix_b200 as ft                                 # noqa: E402  (host-only calls: track compiler + geometry blob)

P = lambda a: a.ctypes.data_as(C.c_void_p)
model = pyoracle.Model()
hq = C.CDLL(sys.argv[2])
t = ft.Track.bundled("track")
lib = ft._lib.load()
arr = (C.c_void_p * 1)(t._ptr)
nw = lib.ftgp_geom_blob(arr, None, 1, None, 0)
blob = np.zeros(nw, dtype=np.uint32)
lib.ftgp_geom_blob(arr, None, 1, P(blob), nw)
out4 = np.zeros(4, dtype=np.int32); size = np.zeros(2)
lib.ftgp_blob_track_view(P(blob), 0, P(out4), P(size))
hq.hq_set_walls.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
hq.hq_set_walls(C.c_void_p(blob.ctypes.data + 4 * int(out4[0])), C.c_void_p(blob.ctypes.data + 4 * int(out4[1])), int(out4[2]), int(out4[3]),
                float(size[0]), float(size[1]))
hq.hq_set_bubble_wrap(1)
rng = np.random.default_rng(5)
n = 4
Q = np.zeros((n, 34)); V = np.zeros((n, 29)); W = np.zeros((n, 29)); U = np.tile([0.5, 0.0], (n, 1))
yaws = []
for c in range(n):
    k = (10, 22, 37, 64)[c]
    d = t.path[k + 1] - t.path[k]
    yaws.append(float(np.arctan2(d[1], d[0])))
    Q[c], V[c], W[c] = model.reset(float(t.path[k, 0]), float(t.path[k, 1]), yaws[c])
ncon = 0
for k in range(int(sys.argv[3])):
    if k == 40:
        for c in range(n):
            V[c, 0:2] = 2.5 * np.array([-np.sin(yaws[c]), np.cos(yaws[c])]) * (1 if c % 2 else -1)
    for mask in (0b0000, 0b0101, 0b1000):
        hq.hq_set_shadow_mask(mask)
        Qa, Va, Wa = Q.copy(), V.copy(), W.copy(); Ia = np.zeros((n, 4), dtype=np.int32)
        for c in range(n):                                      # each car alone (a quad by itself), with / without walls
            if mask >> c & 1:
                hq.hq_set_walls(None, None, 0, 0, 1.0, 1.0)
            ii = np.zeros(4, dtype=np.int32)
            hq.hq_step_ghost(P(Qa[c]), P(Va[c]), P(Wa[c]), P(U[c]), C.c_long(1), 1, P(ii), 0, 0)
            Ia[c] = ii
            hq.hq_set_walls(C.c_void_p(blob.ctypes.data + 4 * int(out4[0])), C.c_void_p(blob.ctypes.data + 4 * int(out4[1])), int(out4[2]),
                            int(out4[3]), float(size[0]), float(size[1]))
        Qb, Vb, Wb = Q.copy(), V.copy(), W.copy(); Ib = np.zeros((n, 4), dtype=np.int32)
        hq.hq_step_warp4(P(Qb), P(Vb), P(Wb), P(U), P(Ib), 2)
        assert np.array_equal(Qa, Qb) and np.array_equal(Va, Vb) and np.array_equal(Wa, Wb) and np.array_equal(Ia, Ib), (k, mask)
        if mask == 0:
            ncon += int(Ia[:, 2].sum())
            Qn, Vn, Wn = Qa, Va, Wa
    Q, V, W = Qn, Vn, Wn
hq.hq_set_shadow_mask(0)
print("ok", ncon)
