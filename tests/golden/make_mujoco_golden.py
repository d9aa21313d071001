#!/usr/bin/env python
"""Capture golden vectors from REAL MuJoCo for the two functions whose oracle is still parity-unpinned
(mj_ray -> oracle/ray.c, mj_step -> oracle/step.c).

Run this on any machine where `import mujoco` works (3.2.2 = requirements.txt:4, or 3.3.2 = uv.lock:104-105):

    python tests/golden/make_mujoco_golden.py            # writes tests/golden/mujoco_golden.npz

It needs nothing from the reference checkout: the world is written by this repo's own emitter
(ft_grandprix_b200/mjcf.py -> rendered/car.xml + chunks, checked element by element against the reference's
templates in tests/test_mjcf.py), compiled with mujoco.MjModel.from_xml_path exactly as custom.py:1178 does, and
driven the way custom.py:1337-1426 drives it.  What is captured (all float64):

  const_*      compile-time constants MuJoCo derives from the MJCF: body_mass / body_inertia /
               body_ipos / body_iquat, dof_invweight0, body_invweight0, stat.meaninertia, geom_size (softener sphere fit),
               dof_armature / dof_damping / dof_frictionloss, jnt_range, opt.* -- tools/make_model.py --from-golden
               rewrites oracle/mushr_mesh.h and csrc/mushr_mesh.h from these
  ray_*        BASELINE config 2: 256 seeded poses on track.png (tests/conftest.py random_poses, seed 0) -> the 90
               rangefinder readings of car 0 after mj_forward (the pose is written into qpos; -1 = no hit)
  step_*       256 diverse driving states reached by MuJoCo itself from the reference spawn under seeded random controls:
               (qpos, qvel, qacc_warmstart, ctrl) before and (qpos, qvel, qacc_warmstart) after ONE mj_step, plus ncon
  wall_*       64 cars driven into walls (ctrl = (4, 0) from seeded poses): state before / after one step at the first
               tick with a chassis- or wheel-vs-hfield contact, with the contact list (geom ids, dist, pos, normal)
  traj_*       BASELINE config 1: one car, nidc driver (this repo's bit-exact restatement, oracle/driver.c), start grid
               slot 0 of track.png, 2 500 ticks: qpos, qvel, ctrl, ranges every tick (custom.py:1337-1426 order:
               driver on last step's sensordata, ctrl write, mj_step)

tests/test_oracle_vs_mujoco.py compares the oracle with whatever this file holds and is skipped while it is absent.
`capture(mujoco_module, ...)` takes the module as an argument so that the plumbing can be exercised with a stand-in
(tests/fake_mujoco.py, backed by the oracle) on machines without MuJoCo.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ONE_CAR = [{"driver": "ft_grandprix.nidc", "name": "red car", "primary": "red", "secondary": "pink", "icon": "white.png"}]


def spawn(mj, m, d, x, y, yaw):
    """mj_resetData + position_vehicles for car 0 (custom.py:1092,1232-1245): z stays at qpos0's 0."""
    mj.mj_resetData(m, d)
    d.qpos[0] = x; d.qpos[1] = y
    d.qpos[3] = np.cos(yaw / 2); d.qpos[4] = 0.0; d.qpos[5] = 0.0; d.qpos[6] = np.sin(yaw / 2)   # euler_to_quaternion([yaw,0,0])


def set_ctrl(d, speed, steer):
    """custom.py:1422-1423 with the actuator order of mushr.em.xml:179-180: ctrl[0] = 'turn #0', ctrl[1] = 'forward #0'."""
    d.ctrl[0] = steer
    d.ctrl[1] = speed


def capture(mj, rendered_dir, track, nidc, n_ray=256, n_step=256, n_wall=64, traj_ticks=2500, log=print):
    """mj: the mujoco module (or a stand-in with the same few entry points); nidc: callable ranges[90] -> (speed, steer) or None."""
    from conftest import random_poses
    m = mj.MjModel.from_xml_path(os.path.join(rendered_dir, "car.xml"))
    d = mj.MjData(m)
    out = {"meta": json.dumps({"mujoco_version": getattr(mj, "__version__", "?"), "track": track.name, "cars": 1,
                               "nq": int(m.nq), "nv": int(m.nv), "nsensordata": int(m.nsensordata)})}
    assert m.nq == 34 and m.nv == 29, (m.nq, m.nv)
    # ---- compile-time constants
    for name in ("body_mass", "body_inertia", "body_ipos", "body_iquat", "body_invweight0", "dof_invweight0", "dof_armature",
                 "dof_damping", "dof_frictionloss", "jnt_range", "geom_size", "geom_pos", "geom_rbound", "qpos0"):
        if hasattr(m, name):
            out["const_" + name] = np.array(getattr(m, name), dtype=np.float64)
    out["const_meaninertia"] = np.array([m.stat.meaninertia])
    out["const_opt"] = np.array([m.opt.timestep, m.opt.tolerance, m.opt.ls_tolerance, m.opt.iterations, m.opt.ls_iterations,
                                 m.opt.impratio, m.opt.cone, m.opt.solver, m.opt.integrator, m.opt.noslip_iterations], dtype=np.float64)
    softeners = [mj.mj_name2id(m, mj.mjtObj.mjOBJ_GEOM, f"{t} softener #0") for t in ("fl", "fr", "bl", "br")]
    out["const_softener_geom_ids"] = np.array(softeners, dtype=np.int64)
    # ---- rangefinders of car 0: sensor ids as custom.py:130 gathers them (all earlier sensors are 1-D, so id == address)
    sens = np.array([mj.mj_name2id(m, mj.mjtObj.mjOBJ_SENSOR, f"rangefinder #0.#{j}") for j in range(90)])

    # ---- config 2: rays from seeded poses
    poses = random_poses(track.path, n_ray, seed=0, level=False)
    ranges = np.zeros((n_ray, 90))
    for k in range(n_ray):
        mj.mj_resetData(m, d)
        d.qpos[0:7] = poses[k]
        mj.mj_forward(m, d)                       # the real mj_forward (the reference monkey-patches only its own name)
        ranges[k] = d.sensordata[sens]
    out["ray_poses"], out["ray_ranges"] = poses, ranges
    log(f"rays: {n_ray} poses, {float((ranges < 0).mean()):.4f} misses")

    # ---- single steps from diverse driving states
    rng = np.random.default_rng(0)
    q0 = np.zeros((n_step, 34)); v0 = np.zeros((n_step, 29)); w0 = np.zeros((n_step, 29)); u0 = np.zeros((n_step, 2))
    q1 = np.zeros((n_step, 34)); v1 = np.zeros((n_step, 29)); w1 = np.zeros((n_step, 29)); nc = np.zeros(n_step, dtype=np.int64)
    for k in range(n_step):
        spawn(mj, m, d, rng.uniform(2, 38), -rng.uniform(2, 38), rng.uniform(-3, 3))       # far from walls or not: both occur
        ctrl = np.array([rng.uniform(0, 5), rng.uniform(-0.8, 0.8)])
        for t in range(int(rng.integers(0, 300))):
            if t % 40 == 0:
                ctrl = np.array([rng.uniform(0, 5), rng.uniform(-0.8, 0.8)])
            set_ctrl(d, *ctrl)
            mj.mj_step(m, d)
        set_ctrl(d, *ctrl)
        q0[k], v0[k], w0[k], u0[k] = d.qpos, d.qvel, d.qacc_warmstart, ctrl        # step_ctrl is (speed, steer) = (forward, turn)
        mj.mj_step(m, d)
        q1[k], v1[k], w1[k], nc[k] = d.qpos, d.qvel, d.qacc_warmstart, d.ncon
    out.update(step_qpos0=q0, step_qvel0=v0, step_warm0=w0, step_ctrl=u0, step_qpos1=q1, step_qvel1=v1, step_warm1=w1, step_ncon=nc)
    log(f"steps: {n_step} states, ncon min/max {nc.min()}/{nc.max()}")

    # ---- wall contacts: drive into walls, keep the first tick where something touches an hfield
    hfield = int(mj.mjtGeom.mjGEOM_HFIELD)
    wposes = random_poses(track.path, n_wall, seed=4, level=True)
    wq0 = np.zeros((n_wall, 34)); wv0 = np.zeros((n_wall, 29)); ww0 = np.zeros((n_wall, 29))
    wq1 = np.zeros((n_wall, 34)); wv1 = np.zeros((n_wall, 29)); wtick = np.full(n_wall, -1, dtype=np.int64)
    wcon = np.zeros((n_wall, 16, 9))                    # per contact: geom id of the car geom, dist, pos[3], normal[3], hfield geom id
    wncon = np.zeros(n_wall, dtype=np.int64)
    for k in range(n_wall):
        yaw = 2 * np.arctan2(wposes[k, 6], wposes[k, 3])
        spawn(mj, m, d, wposes[k, 0], wposes[k, 1], yaw)
        for t in range(900):
            set_ctrl(d, 4.0, 0.0)
            pre = (d.qpos.copy(), d.qvel.copy(), d.qacc_warmstart.copy())
            mj.mj_step(m, d)
            rows = []
            for c in range(int(d.ncon)):
                con = d.contact[c]
                g1, g2 = (int(con.geom1), int(con.geom2)) if hasattr(con, "geom1") else (int(con.geom[0]), int(con.geom[1]))
                t1, t2 = int(m.geom_type[g1]), int(m.geom_type[g2])
                if t1 == hfield or t2 == hfield:
                    car_g, hf_g = (g2, g1) if t1 == hfield else (g1, g2)
                    rows.append([car_g, float(con.dist), *np.array(con.pos, dtype=float), *np.array(con.frame[:3], dtype=float), hf_g])
            if rows:
                wq0[k], wv0[k], ww0[k] = pre
                wq1[k], wv1[k] = d.qpos, d.qvel
                wtick[k] = t; wncon[k] = len(rows)
                wcon[k, :min(16, len(rows))] = np.array(rows[:16])
                break
    out.update(wall_qpos0=wq0, wall_qvel0=wv0, wall_warm0=ww0, wall_qpos1=wq1, wall_qvel1=wv1, wall_tick=wtick, wall_ncon=wncon,
               wall_contacts=wcon, wall_geom_names=np.array([mj.mj_id2name(m, mj.mjtObj.mjOBJ_GEOM, g) or "" for g in range(int(m.ngeom))
                                                              if int(m.geom_type[g]) != hfield]))
    log(f"walls: {(wtick >= 0).sum()} of {n_wall} cars touched a wall")

    # ---- config 1 trajectory
    x, y, yaw = track.start_pose(0)
    spawn(mj, m, d, x, y, yaw)
    tq = np.zeros((traj_ticks + 1, 34)); tv = np.zeros((traj_ticks + 1, 29)); tu = np.zeros((traj_ticks, 2)); tr = np.zeros((traj_ticks, 90))
    tq[0], tv[0] = d.qpos, d.qvel
    for t in range(traj_ticks):
        r = np.array(d.sensordata[sens], dtype=np.float64)         # custom.py:1395: last step's readings, zeros at t = 0
        tr[t] = r
        res = nidc(r.copy())
        if res is not None:                                          # an exception keeps the previous ctrl (custom.py:1409-1411)
            set_ctrl(d, res[0], res[1])
        tu[t] = (d.ctrl[1], d.ctrl[0])                               # stored as (forward, turn)
        mj.mj_step(m, d)
        tq[t + 1], tv[t + 1] = d.qpos, d.qvel
    out.update(traj_qpos=tq, traj_qvel=tv, traj_ctrl_forward_turn=tu, traj_ranges=tr)
    log(f"trajectory: {traj_ticks} ticks, travelled {float(np.hypot(*(tq[-1, :2] - tq[0, :2]))):.2f} m from the grid slot")
    return out


def actuator_order_check(mj, m):
    """ctrl[0] = 'turn #0', ctrl[1] = 'forward #0' (mushr.em.xml:179-180 declares turn first)."""
    turn = mj.mj_name2id(m, mj.mjtObj.mjOBJ_ACTUATOR, "turn #0")
    fwd = mj.mj_name2id(m, mj.mjtObj.mjOBJ_ACTUATOR, "forward #0")
    assert (turn, fwd) == (0, 1), (turn, fwd)


def main():
    import mujoco                                                   # fails loudly where MuJoCo is absent
    import ft_grandprix_b200 as ft
    from ft_grandprix_b200 import mjcf
    from oracle import pyoracle
    track = ft.Track.bundled("track")
    with tempfile.TemporaryDirectory() as tmp:
        rendered = os.path.join(tmp, "rendered")
        mjcf.render_world(track, ONE_CAR, rendered)
        m = mujoco.MjModel.from_xml_path(os.path.join(rendered, "car.xml"))
        actuator_order_check(mujoco, m)
        out = capture(mujoco, rendered, track, lambda r: pyoracle.driver(0, r))
    path = os.path.join(HERE, "mujoco_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"({os.path.getsize(path) / 1e6:.1f} MB) from mujoco", mujoco.__version__)


if __name__ == "__main__":
    main()
