"""TEST-ONLY stand-in for the handful of `mujoco` entry points tests/golden/make_mujoco_golden.py uses, backed by the
CPU oracle.  It exists so that the capture script and tests/test_oracle_vs_mujoco.py can be exercised end to end on
machines without MuJoCo (the comparison is then the oracle against itself: a plumbing check, not parity evidence)."""
import types

import numpy as np

from oracle import pyoracle

__version__ = "stand-in (oracle)"
mjtObj = types.SimpleNamespace(mjOBJ_GEOM=5, mjOBJ_SENSOR=18, mjOBJ_ACTUATOR=19)
mjtGeom = types.SimpleNamespace(mjGEOM_HFIELD=1)
_GEOMS = ["plane", "car #0 lidar", "chasis #0", "", "buddy_wheel_fl_throttle #0", "fl softener #0", "buddy_wheel_fr_throttle #0",
          "fr softener #0", "buddy_wheel_bl_throttle #0", "bl softener #0", "buddy_wheel_br_throttle #0", "br softener #0"]
TRACK = None            # the oracle track the stand-in "compiles" (set by the test before from_xml_path)


class MjModel:
    @classmethod
    def from_xml_path(cls, path):
        m = cls()
        m.nq, m.nv, m.nsensordata, m.ngeom = 34, 29, 90 + 15, len(_GEOMS)
        m._model = pyoracle.Model()
        c = m._model.constants()
        m.body_mass, m.body_ipos = c["body_mass"], c["body_ipos"]
        m.body_inertia_full = c["body_inertia"]
        m.body_invweight0, m.dof_invweight0 = c["body_invweight0"], c["dof_invweight0"]
        m.stat = types.SimpleNamespace(meaninertia=c["meaninertia"])
        m.opt = types.SimpleNamespace(timestep=0.004, tolerance=1e-8, ls_tolerance=0.01, iterations=100, ls_iterations=50,
                                      impratio=1.0, cone=0, solver=2, integrator=0, noslip_iterations=0)
        m.geom_type = np.full(m.ngeom, 6); m.geom_type[0] = 0
        return m


class MjData:
    def __init__(self, m):
        self.qpos = np.zeros(34); self.qvel = np.zeros(29); self.qacc_warmstart = np.zeros(29)
        self.ctrl = np.zeros(2); self.sensordata = np.zeros(m.nsensordata); self.ncon = 0; self.contact = []


def mj_name2id(m, kind, name):
    if kind == mjtObj.mjOBJ_SENSOR:
        return int(name.split("#")[-1])
    if kind == mjtObj.mjOBJ_ACTUATOR:
        return {"turn #0": 0, "forward #0": 1}[name]
    return _GEOMS.index(name)


def mj_id2name(m, kind, i):
    return _GEOMS[i]


def mj_resetData(m, d):
    q, v, w = m._model.reset(0.0, 2.0, 0.0)
    d.qpos[:] = q; d.qvel[:] = 0; d.qacc_warmstart[:] = 0; d.ctrl[:] = 0; d.sensordata[:] = 0


def mj_forward(m, d):
    d.sensordata[:90] = TRACK.scan(d.qpos[None, :7], threads=1)[0]


def mj_step(m, d):
    mj_forward(m, d)
    u = np.array([d.ctrl[1], d.ctrl[0]])
    rc, info = m._model.step(TRACK, d.qpos, d.qvel, d.qacc_warmstart, u)
    d.ncon = 0                    # (the stand-in exposes no contact list)
