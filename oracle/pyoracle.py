"""ctypes binding of the CPU oracle (oracle/*.c).  TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs -- never by the ft_grandprix_b200 package.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libftgp_oracle.so")

def build(force=False):
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB

_lib = None
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)

class LapState(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("offset", "completion", "laps", "start", "good_start", "finished", "ntimes",
                 "off_track", "rank", "delta")]

def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.fto_track_create.restype = C.c_void_p
        L.fto_track_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]
        L.fto_track_destroy.argtypes = [C.c_void_p]
        L.fto_track_nchunks.argtypes = [C.c_void_p]
        L.fto_track_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), dp, dp]
        L.fto_track_chunks.argtypes = [C.c_void_p, C.c_void_p]
        L.fto_track_chunk_data.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.fto_centreline.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.fto_lidar_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.fto_lidar_scan_world.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.fto_ray.restype = C.c_double
        L.fto_ray.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.fto_driver.argtypes = [C.c_int, C.c_void_p, C.c_int, dp, dp]
        L.fto_lap_update.argtypes = [C.POINTER(LapState), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_int32, C.c_int32, ip]
        if hasattr(L, "fto_model_create"):
            L.fto_model_create.restype = C.c_void_p
            L.fto_model_destroy.argtypes = [C.c_void_p]
            L.fto_model_constants.argtypes = [C.c_void_p] + [C.c_void_p] * 6
            L.fto_step.argtypes = [C.c_void_p] * 7
            L.fto_step_n.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int64, C.c_int, C.c_void_p]
            L.fto_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double]
            L.fto_mass_matrix.argtypes = [C.c_void_p] * 3
            L.fto_bias.argtypes = [C.c_void_p] * 4
            L.fto_inverse.argtypes = [C.c_void_p] * 5
            L.fto_energy.restype = C.c_double
            L.fto_energy.argtypes = [C.c_void_p] * 3
        _lib = L
    return _lib

def _p(a):
    return a.ctypes.data_as(C.c_void_p)

class Track:
    """fto_track: chunk.py + the hfield placement of mushr.em.xml."""
    def __init__(self, wall, scale=2.0, chunk_px=20):
        wall = np.ascontiguousarray(wall, dtype=np.uint8)
        self.h, self.w = wall.shape
        self.scale, self.chunk_px = scale, chunk_px
        self.ptr = lib().fto_track_create(_p(wall), self.w, self.h, scale, chunk_px)
        if not self.ptr:
            raise ValueError("fto_track_create failed")
        hc, vc, sx, sy = C.c_int(), C.c_int(), C.c_double(), C.c_double()
        lib().fto_track_dims(self.ptr, hc, vc, sx, sy)
        self.hc, self.vc, self.size_x, self.size_y = hc.value, vc.value, sx.value, sy.value
        self.nchunks = lib().fto_track_nchunks(self.ptr)

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().fto_track_destroy(self.ptr); self.ptr = None

    def chunks(self):
        out = np.zeros((self.nchunks, 2), dtype=np.int32)
        lib().fto_track_chunks(self.ptr, _p(out))
        return out

    def chunk_data(self, k):
        out = np.zeros(self.chunk_px * self.chunk_px, dtype=np.float32)
        nr, nc = C.c_int(), C.c_int()
        lib().fto_track_chunk_data(self.ptr, k, _p(out), nr, nc)
        return out[: nr.value * nc.value].reshape(nr.value, nc.value)

    def centreline(self, d, npoints=100):
        out = np.zeros((npoints, 2))
        rc = lib().fto_centreline(d.encode(), npoints, self.w, self.h, self.chunk_px, self.chunk_px, self.scale, _p(out))
        if rc:
            raise ValueError(f"fto_centreline rc={rc}")
        return out

    def scan(self, poses, threads=None):
        """90-beam scan per pose; cars are independent, so they are spread over host threads
        (ctypes releases the GIL around the C call)."""
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 7)
        out = np.zeros((len(poses), 90))
        L = lib()
        def work(ix):
            for i in ix:
                L.fto_lidar_scan(self.ptr, _p(poses[i]), _p(out[i]))
        threads = threads or min(os.cpu_count() or 1, 32)
        if threads <= 1 or len(poses) < 64:
            work(range(len(poses)))
        else:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(work, np.array_split(np.arange(len(poses)), threads * 4)))
        return out

    def scan_world(self, poses, visible=None):
        """Rangefinders of every car of ONE world: poses [n, 7] (joints at qpos0) or full state rows [n, 34]."""
        poses = np.ascontiguousarray(poses, dtype=np.float64)
        poses = poses.reshape(-1, 34) if poses.shape[-1] == 34 else poses.reshape(-1, 7)
        n, stride = poses.shape
        vis = None if visible is None else np.ascontiguousarray(visible, dtype=np.uint8)
        out = np.zeros((n, 90))
        L = lib()
        L.fto_lidar_scan_world_stride.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        for i in range(n):
            L.fto_lidar_scan_world_stride(self.ptr, _p(poses), stride, n, i, _p(vis) if vis is not None else None, _p(out[i]))
        return out

    def ray(self, pnt, vec):
        pnt = np.ascontiguousarray(pnt, dtype=np.float64); vec = np.ascontiguousarray(vec, dtype=np.float64)
        return lib().fto_ray(self.ptr, _p(pnt), _p(vec))

def driver(kind, ranges):
    """kind: 0 nidc, 1 fast, 2 lobotomy.  Returns (speed, steer) or None if the Python driver would raise."""
    r = np.ascontiguousarray(ranges, dtype=np.float64)
    sp, st = C.c_double(), C.c_double()
    rc = lib().fto_driver(kind, _p(r), len(r), sp, st)
    return None if rc else (sp.value, st.value)

class Lap:
    def __init__(self, offset, max_times=32):
        self.s = LapState(offset=offset, good_start=1)
        self.times = np.zeros(max_times, dtype=np.int32)
    def update(self, path, xy, steps, lap_target, nwinners):
        path = np.ascontiguousarray(path, dtype=np.float64); xy = np.ascontiguousarray(xy, dtype=np.float64)
        nw = C.c_int32(nwinners)
        lib().fto_lap_update(C.byref(self.s), _p(self.times), len(self.times), _p(path), _p(xy), steps, lap_target, nw)
        return nw.value

class Model:
    """fto_model: the single-car mushr model (template/mushr.em.xml)."""
    def __init__(self):
        self.ptr = lib().fto_model_create()
    def __del__(self):
        if getattr(self, "ptr", None):
            lib().fto_model_destroy(self.ptr); self.ptr = None
    def set_bubble_wrap(self, on):
        """option bubble_wrap (custom.py:970-972): softener spheres collide with the walls"""
        lib().fto_model_set_bubble_wrap.argtypes = [C.c_void_p, C.c_int]
        lib().fto_model_set_bubble_wrap(self.ptr, int(bool(on)))
    def constants(self):
        dinv = np.zeros(29); binv = np.zeros((11, 2)); mass = np.zeros(11)
        inertia = np.zeros((11, 3, 3)); ipos = np.zeros((11, 3)); mean = np.zeros(1)
        lib().fto_model_constants(self.ptr, _p(dinv), _p(binv), _p(mass), _p(inertia), _p(ipos), _p(mean))
        return dict(dof_invweight0=dinv, body_invweight0=binv, body_mass=mass, body_inertia=inertia,
                    body_ipos=ipos, meaninertia=float(mean[0]))
    def reset(self, x, y, yaw):
        q = np.zeros(34); v = np.zeros(29); w = np.zeros(29)
        lib().fto_reset(self.ptr, _p(q), _p(v), _p(w), x, y, yaw)
        return q, v, w
    def step(self, track, qpos, qvel, warm, ctrl):
        """in-place single-car step; returns (rc, info[8])"""
        info = np.zeros(8, dtype=np.int32)
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
        rc = lib().fto_step(self.ptr, track.ptr if track is not None else None, _p(qpos), _p(qvel), _p(warm), _p(ctrl), _p(info))
        return rc, info
    def step_n(self, track, qpos, qvel, warm, ctrl, nthreads=1):
        """batch of independent cars, [n,34],[n,29],[n,29],[n,2] in place"""
        n = qpos.shape[0]
        info = np.zeros((n, 8), dtype=np.int32)
        lib().fto_step_n(self.ptr, track.ptr if track is not None else None, _p(qpos), _p(qvel), _p(warm), _p(ctrl), n, nthreads, _p(info))
        return info
    def mass_matrix(self, qpos):
        M = np.zeros((29, 29)); lib().fto_mass_matrix(self.ptr, _p(np.ascontiguousarray(qpos)), _p(M)); return M
    def bias(self, qpos, qvel):
        b = np.zeros(29); lib().fto_bias(self.ptr, _p(np.ascontiguousarray(qpos)), _p(np.ascontiguousarray(qvel)), _p(b)); return b
    def inverse(self, qpos, qvel, qacc):
        t = np.zeros(29); lib().fto_inverse(self.ptr, _p(np.ascontiguousarray(qpos)), _p(np.ascontiguousarray(qvel)), _p(np.ascontiguousarray(qacc)), _p(t)); return t
    def constraint_problem(self, track, qpos, qvel, ctrl, maxrows=128):
        """TEST SUPPORT: (M, qfrc_smooth, J, D, R, aref, floss, type) of the convex problem mj_fwdConstraint solves."""
        M = np.zeros((29, 29)); qfs = np.zeros(29); J = np.zeros((maxrows, 29))
        D = np.zeros(maxrows); R = np.zeros(maxrows); aref = np.zeros(maxrows); fl = np.zeros(maxrows)
        ty = np.zeros(maxrows, dtype=np.int32); pos = np.zeros(maxrows)
        L = lib()
        L.fto_constraint_problem.restype = C.c_int
        n = L.fto_constraint_problem(self.ptr, track.ptr if track is not None else None, _p(np.ascontiguousarray(qpos)),
                                     _p(np.ascontiguousarray(qvel)), _p(np.ascontiguousarray(ctrl, dtype=np.float64)), maxrows,
                                     _p(M), _p(qfs), _p(J), _p(D), _p(R), _p(aref), _p(fl), _p(ty), _p(pos))
        self.last_pos = pos[:n]
        return M, qfs, J[:n], D[:n], R[:n], aref[:n], fl[:n], ty[:n]
    def world_step(self, track, qpos, qvel, warm, ctrl, shadowed=None):
        """One mj_step of an N-car world (one Newton problem over all cars), in place; returns (rc, info[3])."""
        n = qpos.shape[0]
        info = np.zeros(4, dtype=np.int32)
        sh = None if shadowed is None else np.ascontiguousarray(shadowed, dtype=np.uint8)
        L = lib(); L.fto_world_step.restype = C.c_int
        rc = L.fto_world_step(self.ptr, track.ptr if track is not None else None, n, _p(qpos), _p(qvel), _p(warm),
                              _p(np.ascontiguousarray(ctrl, dtype=np.float64)), _p(sh) if sh is not None else None, _p(info))
        return rc, info
    def world_problem(self, track, qpos, qvel, ctrl, maxrows=1024):
        n = qpos.shape[0]; nv = 29 * n
        M = np.zeros((nv, nv)); qfs = np.zeros(nv); J = np.zeros((maxrows, nv))
        D = np.zeros(maxrows); R = np.zeros(maxrows); aref = np.zeros(maxrows); fl = np.zeros(maxrows)
        ty = np.zeros(maxrows, dtype=np.int32); ncc = C.c_int(0)
        L = lib(); L.fto_world_problem.restype = C.c_int
        k = L.fto_world_problem(self.ptr, track.ptr if track is not None else None, n, _p(np.ascontiguousarray(qpos)),
                                _p(np.ascontiguousarray(qvel)), _p(np.ascontiguousarray(ctrl, dtype=np.float64)), maxrows,
                                _p(M), _p(qfs), _p(J), _p(D), _p(R), _p(aref), _p(fl), _p(ty), C.byref(ncc))
        assert k >= 0
        return M, qfs, J[:k], D[:k], R[:k], aref[:k], fl[:k], ty[:k], ncc.value
    def contacts(self, track, qpos, maxcon=16):
        """TEST SUPPORT: [(body, dist, pos[3], normal[3], mu, d0)] of one car (oracle/step.c car_contacts)."""
        out = np.zeros((maxcon, 10))
        L = lib(); L.fto_contacts.restype = C.c_int
        n = L.fto_contacts(self.ptr, track.ptr if track is not None else None, _p(np.ascontiguousarray(qpos, dtype=np.float64)), _p(out), maxcon)
        return out[:n]
    def energy(self, qpos, qvel):
        return lib().fto_energy(self.ptr, _p(np.ascontiguousarray(qpos)), _p(np.ascontiguousarray(qvel)))
