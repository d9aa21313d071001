"""Run by tests/test_step_cpu.py in a subprocess (so that a dead-locked harness can be killed): four cars stepped as ONE
warp of four quads -- every collective is a barrier over all 16 host threads -- with a staged solve (suspend after k1 Newton
rounds, resume to convergence) must give exactly what each car gives alone.  Control flow that is not uniform across the
quads of a warp hangs this script instead of a GPU."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, sys.argv[1])
from oracle import pyoracle                                    # noqa: E402  (test infrastructure)

P = lambda a: a.ctypes.data_as(C.c_void_p)
model = pyoracle.Model()
hq = C.CDLL(sys.argv[2])
rng = np.random.default_rng(21)
n = 4
Q = np.zeros((n, 34)); V = np.zeros((n, 29)); W = np.zeros((n, 29)); U = np.zeros((n, 2))
for c in range(n):
    Q[c], V[c], W[c] = model.reset(rng.normal(), rng.normal(), rng.uniform(-3, 3))
Q[3, 2] = 0.25                                                  # one car dropped from 25 cm: airborne, then impact
nsus = 0
for k in range(int(sys.argv[3])):
    if k % 15 == 0:
        U = np.stack([rng.uniform(0, 4, n), rng.uniform(-0.7, 0.7, n)], 1)
    Qa, Va, Wa = Q.copy(), V.copy(), W.copy(); Ia = np.zeros((n, 4), dtype=np.int32)
    for c in range(n):
        ii = np.zeros(4, dtype=np.int32)
        hq.hq_step_ghost(P(Qa[c]), P(Va[c]), P(Wa[c]), P(U[c]), C.c_long(1), 1, P(ii), 0, 0)
        Ia[c] = ii
    for k1 in (1, 2):
        Qb, Vb, Wb = Q.copy(), V.copy(), W.copy(); Ib = np.zeros((n, 4), dtype=np.int32)
        nsus += hq.hq_step_warp4(P(Qb), P(Vb), P(Wb), P(U), P(Ib), k1)
        assert np.array_equal(Qa, Qb) and np.array_equal(Va, Vb) and np.array_equal(Wa, Wb) and np.array_equal(Ia, Ib), (k, k1)
    Q, V, W = Qa, Va, Wa
print("ok", nsus)
