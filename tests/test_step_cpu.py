"""CPU: the mj_step restatement (oracle/step.c) checked against physical invariants, and the product's
kernel source (csrc/mushr_step.cuh) compiled for the host and compared with it.

The host build of the kernel source exists ONLY here (tests/host_harness): the GPU box runs the same
source as a CUDA kernel, the product never links a CPU build of it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

P = lambda a: a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def model(oracle):
    return oracle.Model()


@pytest.fixture(scope="module")
def host_kernel():
    src = os.path.join(ROOT, "tests", "host_harness", "step_host.cpp")
    out = os.path.join(ROOT, "tests", "host_harness", "libstep_host.so")
    deps = [src] + [os.path.join(ROOT, "ft_grandprix_b200", "csrc", f) for f in ("mushr_step.cuh", "mushr_consts.h", "mushr_mesh.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", out, src])
    return C.CDLL(out)


@pytest.fixture(scope="module")
def host_quad_kernel():
    """The quad-per-car kernel source compiled for the host: four OS threads play the four lanes of a quad."""
    src = os.path.join(ROOT, "tests", "host_harness", "step_quad_host.cpp")
    out = os.path.join(ROOT, "tests", "host_harness", "libstep_quad_host.so")
    deps = [src] + [os.path.join(ROOT, "ft_grandprix_b200", "csrc", f) for f in ("mushr_step_quad.cuh", "mushr_step.cuh", "mushr_consts.h", "mushr_mesh.h", "hfield_contact.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-pthread", "-o", out, src])
    lib = C.CDLL(out)
    lib.hq_step_ghost.argtypes = [C.c_void_p] * 4 + [C.c_long, C.c_int, C.c_void_p, C.c_int, C.c_int]
    return lib


def random_state(model, rng):
    q = model.reset(0, 0, 0)[0]
    q[0:3] = rng.normal(size=3)
    q[3:7] = rng.normal(size=4); q[3:7] /= np.linalg.norm(q[3:7])
    for a in (11, 18, 24, 30):
        q[a:a + 4] = rng.normal(size=4); q[a:a + 4] /= np.linalg.norm(q[a:a + 4])
    q[[7, 8, 9, 10, 15, 16, 17, 22, 23, 28, 29]] = rng.normal(size=11) * 0.3
    return q


def test_model_constants(model):
    c = model.constants()
    assert abs(c["body_mass"].sum() - 5.6328) < 1e-3                      # SURVEY A.1 total mass
    assert abs(c["body_mass"][1] - (3.542137 + 1000 * np.pi * 0.03 ** 2 * 0.03)) < 1e-12
    # ellipsoid inertia m/5 (b^2+c^2, a^2+c^2, a^2+b^2)
    np.testing.assert_allclose(np.diag(c["body_inertia"][3]), 0.498952 / 5 * np.array([0.001, 0.0018, 0.001]), rtol=1e-12)
    # translational invweight of a free body of mass m coupled to nothing would be 1/m; the car's is below 1/m_chassis
    assert 1 / 5.64 < c["dof_invweight0"][0] < 0.5    # origin is off the CoM and the wheels slide in z: above 1/m_total
    assert np.allclose(c["dof_invweight0"][0:3], c["dof_invweight0"][0]) and np.allclose(c["dof_invweight0"][10:13], c["dof_invweight0"][10])
    assert 0.5 < c["meaninertia"] < 0.8


def test_mass_matrix_spd_and_inverse_dynamics_consistency(model):
    """CRB mass matrix is SPD; RNE with acceleration equals M qacc + bias (two independent recursions)."""
    rng = np.random.default_rng(0)
    for _ in range(20):
        q = random_state(model, rng)
        v, a = rng.normal(size=29), rng.normal(size=29)
        M = model.mass_matrix(q)
        assert np.abs(M - M.T).max() == 0 and np.linalg.eigvalsh(M).min() > 0
        np.testing.assert_allclose(M @ a + model.bias(q, v), model.inverse(q, v, a), rtol=0, atol=1e-11)


def test_bias_matches_lagrangian_finite_differences(model):
    """qfrc_bias = C(q, v) + dV/dq: checked for the scalar joints (hinge/slide) with central differences of
    the kinetic energy T = v^T M v / 2 and the gravitational potential, other joints at rest."""
    rng = np.random.default_rng(1)
    sj = [(7, 6), (8, 7), (9, 8), (10, 9), (15, 13), (16, 14), (17, 15), (22, 19), (23, 20), (28, 24), (29, 25)]   # (qadr, dadr)
    q = random_state(model, rng)
    v = np.zeros(29)
    for _, d in sj:
        v[d] = rng.normal()
    bias = model.bias(q, v)
    eps = 1e-6
    def T(qq): return 0.5 * v @ model.mass_matrix(qq) @ v
    def V(qq): return model.energy(qq, np.zeros(29)) - 0.5 * 500 * sum((qq[a] + 0.015) ** 2 for a in (8, 15, 22, 28))
    # d/dt (dT/dv_i) - dT/dq_i + dV/dq_i with qacc = 0: sum_k dM_ij/dq_k v_k v_j - dT/dq_i + dV/dq_i
    Mdot = np.zeros((29, 29))
    for qa, d in sj:
        qp, qm = q.copy(), q.copy(); qp[qa] += eps; qm[qa] -= eps
        Mdot += (model.mass_matrix(qp) - model.mass_matrix(qm)) / (2 * eps) * v[d]
    for qa, d in sj:
        qp, qm = q.copy(), q.copy(); qp[qa] += eps; qm[qa] -= eps
        want = (Mdot @ v)[d] - (T(qp) - T(qm)) / (2 * eps) + (V(qp) - V(qm)) / (2 * eps)
        assert abs(bias[d] - want) < 1e-6 * max(1, abs(want)), (d, bias[d], want)


def test_free_flight_conserves_momentum_and_falls_at_g(model):
    q, v, w = model.reset(0, 0, 0)
    q[2] = 5.0
    v[0], v[5] = 1.0, 2.0
    c = model.constants()
    mass = c["body_mass"].sum()
    for k in range(50):
        rc, info = model.step(None, q, v, w, np.zeros(2))
        assert rc == 0 and info[2] == 0            # no contacts
    # vertical velocity of the free joint ~ -g t (origin is ~2 cm from the CoM and the body spins: small wobble)
    assert abs(v[2] + 9.81 * 0.2) < 5e-3


def test_settles_on_ground_and_reaches_servo_speed(model):
    q, v, w = model.reset(1.0, 2.0, 0.3)
    for k in range(2500):
        rc, info = model.step(None, q, v, w, np.array([2.0, 0.0]))
        assert rc == 0
    assert info[2] == 4 and info[0] <= 3                                   # 4 wheel contacts, Newton converged fast
    assert abs(q[2] - 0.0151) < 5e-4                                       # ride height (~0.0156 minus sag)
    assert np.all(q[[8, 15, 22, 28]] > -0.002) and np.all(q[[8, 15, 22, 28]] < 0.002)   # suspensions on their stops
    spin = v[[9, 15, 20, 25]]
    # velocity servo kv=100, gear 0.04 on the mean spin, against 0.01 N m s wheel damping: 100 (u - 0.04 s) 0.01 = 0.01 s
    assert np.allclose(spin, 2.0 / 0.05, rtol=2e-3)
    assert abs(np.hypot(v[0], v[1]) - 0.03 * spin.mean()) < 2e-2          # rolling without slipping, r = 0.03
    assert abs(np.arctan2(v[1], v[0]) - 0.3) < 0.03                       # straight along the spawn heading


def test_ackermann_equality_holds(model):
    q, v, w = model.reset(0, 0, 0)
    for k in range(600):
        model.step(None, q, v, w, np.array([1.0, 0.3]))
    x = q[7]
    assert abs(x - 0.3) < 2e-3
    pl = x + 0.375 * x ** 2 + 0.140625 * x ** 3 - 0.0722656 * x ** 4
    pr = x - 0.375 * x ** 2 + 0.140625 * x ** 3 + 0.0722656 * x ** 4
    assert abs(q[9] - pl) < 2e-3 and abs(q[16] - pr) < 2e-3
    assert v[5] > 0.3                                                       # it turns left


def test_bad_state_resets_like_mujoco(model):
    q, v, w = model.reset(3, 4, 1)
    v[0] = 1e11
    rc, _ = model.step(None, q, v, w, np.zeros(2))
    assert rc == 1 and abs(q[1] - 2.0) < 1e-2 and abs(q[0]) < 1e-2          # back at qpos0 (0, 2, 0)


def test_product_kernel_source_matches_oracle_constants(model, host_kernel):
    d, w4, c2 = np.zeros(31), np.zeros(4), np.zeros(2)
    host_kernel.hh_constants(P(d), P(w4), P(c2))
    c = model.constants()
    pad = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, -1, 20, 21, 22, 23, 24, -1, 25, 26, 27, 28]
    for p, dof in enumerate(pad):
        if dof >= 0:
            assert abs(d[p] / c["dof_invweight0"][dof] - 1) < 1e-12
    np.testing.assert_allclose(w4, c["body_invweight0"][[3, 5, 7, 9], 0], rtol=1e-12)
    np.testing.assert_allclose(c2, [c["body_invweight0"][1, 0], c["meaninertia"]], rtol=1e-12)


def test_product_kernel_source_single_step_parity(model, host_kernel):
    """1e-5 relative bar (north star) -- the two implementations (generic dense vs block-arrow) agree to ~1e-13."""
    rng = np.random.default_rng(3)
    worst = 0
    for car in range(6):
        q, v, w = model.reset(rng.normal(), rng.normal(), rng.uniform(-3, 3))
        for k in range(250):
            ctrl = np.array([rng.uniform(0, 4), rng.uniform(-0.6, 0.6)]) if k % 25 == 0 else ctrl
            qo, vo, wo = q.copy(), v.copy(), w.copy()
            model.step(None, qo, vo, wo, ctrl)
            qh, vh, wh = q.copy(), v.copy(), w.copy()
            info = np.zeros(4, dtype=np.int32)
            host_kernel.hh_step(P(qh), P(vh), P(wh), P(ctrl), 1, P(info))
            np.testing.assert_allclose(qh, qo, rtol=1e-5, atol=1e-10)
            np.testing.assert_allclose(vh, vo, rtol=1e-5, atol=1e-9)
            worst = max(worst, np.abs(vh - vo).max())
            q, v, w = qo, vo, wo
    assert worst < 1e-10


def test_quad_kernel_source_single_step_parity(model, host_quad_kernel):
    """The quad-per-car kernel (four lanes per car, csrc/mushr_step_quad.cuh) against the oracle: 1e-5 relative bar,
    same Newton iteration counts; and a quad that is kept going after its car converged (as in a warp / CTA whose
    other cars need more iterations) must leave bit-identical results."""
    rng = np.random.default_rng(3)
    worst = 0
    for car in range(4):
        q, v, w = model.reset(rng.normal(), rng.normal(), rng.uniform(-3, 3))
        for k in range(200):
            ctrl = np.array([rng.uniform(0, 4), rng.uniform(-0.6, 0.6)]) if k % 25 == 0 else ctrl
            qo, vo, wo = q.copy(), v.copy(), w.copy()
            _, info_o = model.step(None, qo, vo, wo, ctrl)
            ref = None
            for gw, gc in ((0, 0), (3, 2)) if k % 10 == 0 else ((0, 0),):
                qh, vh, wh = q.copy(), v.copy(), w.copy()
                info = np.zeros(4, dtype=np.int32)
                host_quad_kernel.hq_step_ghost(P(qh), P(vh), P(wh), P(ctrl), 1, 1, P(info), gw, gc)
                if ref is None:
                    ref = (qh, vh, wh, info)
                else:
                    assert all(np.array_equal(a, b) for a, b in zip(ref, (qh, vh, wh, info)))
            qh, vh, wh, info = ref
            np.testing.assert_allclose(qh, qo, rtol=1e-5, atol=1e-10)
            np.testing.assert_allclose(vh, vo, rtol=1e-5, atol=1e-9)
            np.testing.assert_allclose(wh, wo, rtol=1e-5, atol=1e-6)
            assert info[0] == info_o[0] and info[1] == info_o[2]
            worst = max(worst, np.abs(vh - vo).max())
            q, v, w = qo, vo, wo
    assert worst < 1e-10


def test_quad_kernel_source_resets_bad_state(model, host_quad_kernel):
    q, v, w = model.reset(3, 4, 1)
    v[0] = 1e11
    qo, vo, wo = q.copy(), v.copy(), w.copy()
    rc, _ = model.step(None, qo, vo, wo, np.zeros(2))
    info = np.zeros(4, dtype=np.int32)
    ctrl = np.zeros(2)
    host_quad_kernel.hq_step_ghost(P(q), P(v), P(w), P(ctrl), 1, 1, P(info), 0, 0)
    assert rc == 1 and info[3] == 1
    np.testing.assert_allclose(q, qo, rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(v, vo, rtol=1e-5, atol=1e-9)


def test_quad_kernel_source_staged_solve_is_bit_identical(model, host_quad_kernel):
    """Staged solve (a car that is not done after k1 Newton rounds is suspended into a record and resumed by a later
    launch): suspend after 1 or 2 rounds, resume for 1 more, then to convergence -- same bits as the one-shot step, with
    the shared slots scribbled over between the stages."""
    host_quad_kernel.hq_step_staged.argtypes = [C.c_void_p] * 4 + [C.c_long, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(9)
    nsus = 0
    for car in range(3):
        q, v, w = model.reset(rng.normal(), rng.normal(), rng.uniform(-3, 3))
        for k in range(150):
            ctrl = np.array([rng.uniform(0, 4), rng.uniform(-0.6, 0.6)]) if k % 20 == 0 else ctrl
            qa, va, wa = q.copy(), v.copy(), w.copy(); ia = np.zeros(4, dtype=np.int32)
            host_quad_kernel.hq_step_ghost(P(qa), P(va), P(wa), P(ctrl), 1, 1, P(ia), 0, 0)
            for k1, k2 in ((1, 1), (2, 1), (1, 50)):
                qb, vb, wb = q.copy(), v.copy(), w.copy(); ib = np.zeros(4, dtype=np.int32); ns = C.c_int(0)
                host_quad_kernel.hq_step_staged(P(qb), P(vb), P(wb), P(ctrl), 1, P(ib), k1, k2, C.byref(ns))
                assert np.array_equal(qa, qb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(ia, ib)
                nsus += ns.value
            q, v, w = qa, va, wa
    assert nsus > 200                                                         # the suspend / resume path really ran


def _track_svg():
    import json
    return json.load(open(os.path.join(ROOT, "ft_grandprix_b200", "assets", "paths.json")))["track"]


def _independent_minimiser(M, qfs, J, D, R, aref, fl, ty):
    """Semi-smooth Newton with a bisection line search on the derivative, written from the cost definition of SURVEY B.8
    alone (equality: D r^2 / 2; limit / contact: D r^2 / 2 for r < 0; friction loss: Huber with knee R eta), in numpy.
    Shares no code and no algorithmic detail (warm start, line search, stopping rule) with oracle/step.c."""
    a_s = np.linalg.solve(M, qfs)

    def force_and_active(a):
        r = J @ a - aref
        f = np.zeros_like(r); act = np.zeros_like(r, dtype=bool)
        eq = ty == 0
        f[eq] = -D[eq] * r[eq]; act[eq] = True
        one = (ty == 2) | (ty == 3)
        on = one & (r < 0)
        f[on] = -D[on] * r[on]; act[on] = True
        fr = ty == 1
        knee = R * fl
        quad = fr & (np.abs(r) < knee)
        f[quad] = -D[quad] * r[quad]; act[quad] = True
        lin = fr & ~quad
        f[lin] = -np.sign(r[lin]) * fl[lin]
        return f, act

    def grad(a):
        f, _ = force_and_active(a)
        return M @ (a - a_s) - J.T @ f

    a = a_s.copy()
    for it in range(200):
        g = grad(a)
        if np.linalg.norm(g) < 1e-11 * (1 + np.linalg.norm(qfs)):
            break
        _, act = force_and_active(a)
        H = M + J[act].T @ (D[act, None] * J[act])
        d = -np.linalg.solve(H, g)
        # exact line search: the directional derivative phi'(t) = grad(a + t d) . d is increasing (convex cost)
        lo, hi = 0.0, 1.0
        while grad(a + hi * d) @ d < 0 and hi < 64:
            lo, hi = hi, 2 * hi
        for _ in range(80):
            mid = 0.5 * (lo + hi)
            if grad(a + mid * d) @ d < 0:
                lo = mid
            else:
                hi = mid
        a = a + 0.5 * (lo + hi) * d
    return a


def test_newton_solution_is_the_minimiser_found_by_an_independent_solver(model, otracks):
    """mj_fwdConstraint pinned as an optimisation problem: for driving, cornering, airborne and wall-contact states the
    acceleration the oracle's Newton solver returns (and therefore the CUDA kernels', which agree with it to 1e-13) equals
    the minimiser an independent numpy solver finds for the exported problem, and satisfies its optimality condition."""
    rng = np.random.default_rng(12)
    t = otracks["track"]
    checked = 0
    for car in range(4):
        q, v, w = model.reset(rng.uniform(2, 38), -rng.uniform(2, 38), rng.uniform(-3, 3))
        if car == 3:
            q[2] = 0.3                                       # dropped from 30 cm: free flight, then impact
        ctrl = np.array([3.0, 0.5])
        for k in range(160):
            if k % 20 == 0:
                ctrl = np.array([rng.uniform(0, 5), rng.uniform(-0.8, 0.8)])
            if k % 8 == 0:
                M, qfs, J, D, R, aref, fl, ty = model.constraint_problem(t, q, v, ctrl)
                q2, v2, w2 = q.copy(), v.copy(), w.copy()
                model.step(t, q2, v2, w2, ctrl)              # w2 = qacc of the oracle's solver
                a = _independent_minimiser(M, qfs, J, D, R, aref, fl, ty)
                scale = np.abs(a).max() + 1
                assert np.abs(a - w2).max() < 1e-8 * scale, (car, k, np.abs(a - w2).max(), scale)   # measured: 2e-11
                checked += 1
            model.step(t, q, v, w, ctrl)
    assert checked >= 70
    # straight into a wall: states with chassis-wall contacts (this framework's contact definition, same solver)
    walls_checked = 0
    q, v, w = model.reset(*[float(x) for x in otracks["track"].centreline(_track_svg())[10]], 0.4)
    ctrl = np.array([3.0, 0.0])
    for k in range(900):
        M, qfs, J, D, R, aref, fl, ty = model.constraint_problem(t, q, v, ctrl)
        q2, v2, w2 = q.copy(), v.copy(), w.copy()
        _, info = model.step(t, q2, v2, w2, ctrl)
        if info[3] > 0 and walls_checked < 12:
            a = _independent_minimiser(M, qfs, J, D, R, aref, fl, ty)
            assert np.abs(a - w2).max() < 1e-8 * (np.abs(a).max() + 1), (k, np.abs(a - w2).max())
            walls_checked += 1
        q, v, w = q2, v2, w2
        if walls_checked >= 12:
            break
    assert walls_checked >= 5


def _integrate_pos(q, v, h):
    """q (+) h v for the mushr joint layout (mj_integratePos): free joint, hinges / slides, four ball joints."""
    def qmul(a, b):
        return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                         a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])

    def qint(quat, w):
        ang = np.linalg.norm(w) * h
        if ang == 0:
            return quat / np.linalg.norm(quat)
        ax = w / np.linalg.norm(w)
        return qmul(quat / np.linalg.norm(quat), np.concatenate([[np.cos(ang / 2)], ax * np.sin(ang / 2)]))
    out = q.copy()
    out[0:3] += h * v[0:3]
    out[3:7] = qint(q[3:7], v[3:6])
    scalar = [(7, 6), (8, 7), (9, 8), (10, 9), (15, 13), (16, 14), (17, 15), (22, 19), (23, 20), (28, 24), (29, 25)]
    for qa, d in scalar:
        out[qa] += h * v[d]
    for qa, d in ((11, 10), (18, 16), (24, 21), (30, 26)):
        out[qa:qa + 4] = qint(q[qa:qa + 4], v[d:d + 3])
    return out


def test_constraint_jacobians_are_the_derivatives_of_the_violations(model, otracks):
    """Row Jacobians against central finite differences of the exported constraint violations: for a random velocity v,
    (pos(q (+) eps v) - pos(q (-) eps v)) / (2 eps) = J v for equality, limit and (pyramid pairs averaged) contact-normal
    rows.  Catches sign, frame and indexing errors in the row assembly independently of the solver."""
    rng = np.random.default_rng(4)
    t = otracks["track"]
    checked = {0: 0, 2: 0, 3: 0}
    for trial in range(12):
        q, v, w = model.reset(rng.uniform(5, 35), -rng.uniform(5, 35), rng.uniform(-3, 3))
        ctrl = np.array([rng.uniform(0, 4), rng.uniform(-0.9, 0.9)])
        for k in range(int(rng.integers(40, 160))):
            model.step(t, q, v, w, ctrl)
        q[7] += 0.9 * np.sign(ctrl[1])                              # push the steering wheel past its +-1 limit half the time
        vel = rng.normal(size=29) * 0.3
        eps = 1e-6
        M, qfs, J, D, R, aref, fl, ty = model.constraint_problem(t, q, np.zeros(29), ctrl); p0 = model.last_pos.copy()
        _, _, Jp, *_rest, typ = model.constraint_problem(t, _integrate_pos(q, vel, eps), np.zeros(29), ctrl); pp = model.last_pos.copy()
        _, _, Jm, *_rest, tym = model.constraint_problem(t, _integrate_pos(q, vel, -eps), np.zeros(29), ctrl); pm = model.last_pos.copy()
        if len(pp) != len(p0) or len(pm) != len(p0) or not (np.array_equal(typ, ty) and np.array_equal(tym, ty)):
            continue                                                # a row appeared / vanished inside the stencil
        fd = (pp - pm) / (2 * eps)
        for i in range(len(ty)):
            if ty[i] in (0, 2):
                assert abs(fd[i] - J[i] @ vel) < 1e-6 * (1 + abs(fd[i])), (trial, i, ty[i], fd[i], J[i] @ vel)
                checked[int(ty[i])] += 1
        con = np.nonzero(ty == 3)[0]
        for c in range(0, len(con), 4):                             # rows Jn +- mu Jt1, Jn +- mu Jt2: the mean is Jn
            jn = J[con[c:c + 4]].mean(0)
            assert abs(fd[con[c]] - jn @ vel) < 1e-5 * (1 + abs(fd[con[c]])), (trial, c, fd[con[c]], jn @ vel)
            checked[3] += 1
    assert checked[0] >= 20 and checked[2] >= 20 and checked[3] >= 30, checked


def test_row_regularisers_and_references_follow_the_documented_formulas(model):
    """D = 1/R and aref of every exported row recomputed in numpy from the formulas of SURVEY B.7 (impedance sigmoid
    d(x), R = (1 - d)/d * diagApprox, K = 1/(dmax^2 tc^2), B = 2/(dmax tc), pyramidal R = 2 mu^2 R_normal) and the model
    constants (dof_invweight0, body_invweight0) -- an independent statement of the softness of every constraint."""
    c = model.constants()
    dinv, binv = c["dof_invweight0"], c["body_invweight0"]
    dmax, width, mid, tc = 0.95, 0.001, 0.5, 0.02

    def imp(d0, pos):
        x = abs(pos) / width
        if x >= 1: return dmax
        if x == 0: return d0
        y = x * x / mid if x <= mid else 1 - (1 - x) ** 2 / (1 - mid)
        return d0 + y * (dmax - d0)
    K, B = 1 / (dmax * dmax * tc * tc), 2 / (dmax * tc)
    rng = np.random.default_rng(6)
    wheel_body = {7: 3, 13: 5, 19: 7, 24: 9}                       # suspension dof -> wheel body id (depth-first numbering)
    rows = 0
    for trial in range(10):
        q, v, w = model.reset(rng.uniform(5, 35), -rng.uniform(5, 35), rng.uniform(-3, 3))
        ctrl = np.array([rng.uniform(0, 4), rng.uniform(-0.9, 0.9)])
        for k in range(int(rng.integers(30, 200))):
            model.step(None, q, v, w, ctrl)
        q[7] += 0.8 * np.sign(ctrl[1])
        M, qfs, J, D, R, aref, fl, ty = model.constraint_problem(None, q, v, ctrl)
        pos = model.last_pos
        i = 0
        while i < len(ty):
            nz = np.nonzero(J[i])[0]
            if ty[i] == 0:                                           # joint equality: two dofs
                diag, d0 = dinv[nz].sum(), 0.9
            elif ty[i] in (1, 2):                                    # friction loss / limit: one dof
                assert len(nz) == 1
                diag, d0 = dinv[nz[0]], 0.9
            else:                                                    # wheel-ground contact: four pyramid rows, mixed solimp d0 = 0.45
                susp = [d for d in wheel_body if J[i][d] != 0]
                assert len(susp) == 1
                diag, d0 = binv[wheel_body[susp[0]], 0], 0.45
            d = imp(d0, pos[i])
            Rn = max(1e-15, (1 - d) * diag / d)
            vel = J[i] @ v
            if ty[i] == 3:
                mu = 0.5
                for r in range(4):
                    assert abs(R[i + r] / (2 * mu * mu * Rn) - 1) < 1e-12 and abs(D[i + r] * R[i + r] - 1) < 1e-12
                    want = -B * (J[i + r] @ v) - K * d * pos[i + r]
                    assert abs(aref[i + r] - want) <= 1e-9 * (1 + abs(want))
                i += 4; rows += 4
                continue
            assert abs(R[i] / Rn - 1) < 1e-12 and abs(D[i] * R[i] - 1) < 1e-12, (ty[i], R[i], Rn)
            want = -B * vel - (0 if ty[i] == 1 else K * d * pos[i])
            assert abs(aref[i] - want) <= 1e-9 * (1 + abs(want)), (ty[i], aref[i], want)
            i += 1; rows += 1
    assert rows > 300


def test_wheel_ground_distance_against_bruteforce_ellipsoid(model):
    """The contact distance of every wheel-ground contact row against first principles: the wheel ellipsoid
    (0.03, 0.01, 0.03; mushr.em.xml:69) is posed with numpy kinematics (car pose, wheel offset, suspension slide along
    the car's z, steering about z, spin about y; mushr.em.xml:124-173), sampled at 1e6 surface points, and the lowest
    one is compared with the row's `pos` (distance to the plane z = 0.01, mushr.em.xml:94)."""
    rng = np.random.default_rng(8)
    th, ph = np.meshgrid(np.linspace(0, np.pi, 1000), np.linspace(0, 2 * np.pi, 1000))
    unit = np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], -1).reshape(-1, 3)
    ell = unit * np.array([0.03, 0.01, 0.03])
    wheels = [(0.06925, 0.0575, 8, 9, 10, 7), (0.06925, -0.0575, 15, 16, 17, 13), (-0.079, 0.0575, 22, None, 23, 19),
              (-0.079, -0.0575, 28, None, 29, 24)]                                   # x, y, q susp, q steer, q spin, dof susp

    def rot(q):
        w, x, y, z = q / np.linalg.norm(q)
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                         [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                         [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    checked = 0
    for trial in range(6):
        q, v, w = model.reset(rng.uniform(5, 35), -rng.uniform(5, 35), rng.uniform(-3, 3))
        ctrl = np.array([rng.uniform(0, 4), rng.uniform(-0.9, 0.9)])
        for k in range(int(rng.integers(20, 150))):
            model.step(None, q, v, w, ctrl)
        M, qfs, J, D, R, aref, fl, ty = model.constraint_problem(None, q, v, ctrl)
        pos = model.last_pos
        R1 = rot(q[3:7])
        con = np.nonzero(ty == 3)[0]
        for c in range(0, len(con), 4):
            i = con[c]
            (wx, wy, qs, qst, qsp, dsusp), = [wh for wh in wheels if J[i][wh[5]] != 0]
            a = q[qst] if qst is not None else 0.0
            Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
            b = q[qsp]
            Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
            Rw = R1 @ Rz @ Ry
            pw = q[0:3] + R1 @ np.array([wx, wy, 0.0244 + q[qs]])
            lowest = (ell @ Rw.T + pw)[:, 2].min()
            assert abs((lowest - 0.01) - pos[i]) < 2e-7, (trial, c, lowest - 0.01, pos[i])
            # closed form of the same support function
            assert abs(pw[2] - np.sqrt(((np.array([0.03, 0.01, 0.03]) * Rw[2]) ** 2).sum()) - 0.01 - pos[i]) < 1e-12
            checked += 1
    assert checked >= 12


def test_euler_update_against_numpy_restatement(model, otracks):
    """mj_Euler with implicit joint damping and mj_integratePos restated in numpy from the exported
    problem: qacc' = (M + h diag(damping))^-1 (qfrc_smooth + J' f(qacc)), qvel += h qacc', qpos (+)= h qvel (new velocity,
    quaternions by the exponential map) -- compared with what the oracle's step returns."""
    h = 0.004
    damping = np.zeros(29)
    damping[[7, 13, 19, 24]] = 12.5                                 # suspension slides   (mushr.em.xml:63)
    damping[[6, 8, 14]] = 0.1                                       # steering hinges     (:78)
    damping[[9, 15, 20, 25]] = 0.01                                 # throttle hinges     (:81)
    rng = np.random.default_rng(10)
    t = otracks["track"]
    worst = 0
    for trial in range(5):
        q, v, w = model.reset(rng.uniform(5, 35), -rng.uniform(5, 35), rng.uniform(-3, 3))
        ctrl = np.array([rng.uniform(0, 4), rng.uniform(-0.9, 0.9)])
        for k in range(120):
            if k % 10 == 0:
                M, qfs, J, D, R, aref, fl, ty = model.constraint_problem(t, q, v, ctrl)
            q0, v0 = q.copy(), v.copy()
            model.step(t, q, v, w, ctrl)
            if k % 10 == 0:
                a = w                                               # qacc_warmstart <- the solver's qacc
                r = J @ a - aref
                f = np.where(ty == 0, -D * r, 0.0)
                one = ((ty == 2) | (ty == 3)) & (r < 0)
                f = np.where(one, -D * r, f)
                fr = ty == 1
                knee = R * fl
                f = np.where(fr & (np.abs(r) < knee), -D * r, f)
                f = np.where(fr & (np.abs(r) >= knee), -np.sign(r) * fl, f)
                qa = np.linalg.solve(M + h * np.diag(damping), qfs + J.T @ f)
                v1 = v0 + h * qa
                q1 = _integrate_pos(q0, v1, h)
                worst = max(worst, np.abs(v1 - v).max() / (1 + np.abs(v).max()), np.abs(q1 - q).max())
    assert worst < 1e-8, worst                                      # measured 1.3e-10 (D up to 1e6 times rounding of the residuals)


# ---------------------------------------------------------------------------------------------- N-car worlds (oracle only)
def test_world_step_of_one_car_is_the_single_car_step(model):
    """The world-level solver (dynamic sizes, one Newton problem over all cars) restricted to one car reproduces fto_step
    bit for bit: same rows, same order of operations."""
    q, v, w = model.reset(3.0, -4.0, 0.5)
    ctrl = np.array([2.0, 0.3])
    for k in range(120):
        q1, v1, w1 = q.copy(), v.copy(), w.copy()
        _, i1 = model.step(None, q1, v1, w1, ctrl)
        Q, V, W = q[None].copy(), v[None].copy(), w[None].copy()
        _, iw = model.world_step(None, Q, V, W, ctrl[None])
        assert np.array_equal(Q[0], q1) and np.array_equal(V[0], v1) and np.array_equal(W[0], w1) and iw[0] == i1[0]
        q, v, w = q1, v1, w1


def test_world_of_distant_cars_agrees_with_independent_cars(model):
    """Cars that do not touch are coupled only through the shared line search and stopping rule of the world's single
    Newton problem: both converge to the same minimisers (deviation far below the 1e-5 bar; measured 2e-13)."""
    n = 3
    Q = np.zeros((n, 34)); V = np.zeros((n, 29)); W = np.zeros((n, 29))
    U = np.array([[2.0, 0.3], [3.0, -0.5], [1.0, 0.0]])
    for c in range(n):
        Q[c], V[c], W[c] = model.reset(5.0 * c, -3.0, 0.3 * c)
    Qi, Vi, Wi = Q.copy(), V.copy(), W.copy()
    for k in range(150):
        _, info = model.world_step(None, Q, V, W, U)
        assert info[2] == 0
        for c in range(n):
            model.step(None, Qi[c], Vi[c], Wi[c], U[c])
        np.testing.assert_allclose(Q, Qi, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(V, Vi, rtol=1e-9, atol=1e-9)


def test_world_with_touching_cars(model):
    """Car-car contacts (this framework's definition: hull vertex of A inside the hull box of B): the rear car drives into
    the stationary front car and pushes it; the coupled world problem's solution equals the independent numpy minimiser;
    the contact rows act with opposite sign on the two cars' translation dofs; a shadowed car is passed through."""
    def spawn():
        Q = np.zeros((2, 34)); V = np.zeros((2, 29)); W = np.zeros((2, 29))
        Q[0], V[0], W[0] = model.reset(10.0, -10.0, 0.0)
        Q[1], V[1], W[1] = model.reset(10.17, -10.01, 0.1)
        return Q, V, W
    U = np.array([[2.0, 0.0], [0.0, 0.0]])
    Q, V, W = spawn()
    seen = 0
    for k in range(80):
        if k in (3, 30, 70):
            M, qfs, J, D, R, aref, fl, ty, ncc = model.world_problem(None, Q, V, U)
            assert ncc >= 1
            cc = J[len(ty) - 4 * ncc:]                               # car-car rows are last
            assert np.abs(cc[:, 0:3] + cc[:, 29:32]).max() < 1e-12   # action = -reaction on the translation dofs
            Q2, V2, W2 = Q.copy(), V.copy(), W.copy()
            model.world_step(None, Q2, V2, W2, U)
            a = _independent_minimiser(M, qfs, J, D, R, aref, fl, ty)
            assert np.abs(a - W2.ravel()).max() < 1e-7 * (1 + np.abs(a).max())
        _, info = model.world_step(None, Q, V, W, U)
        seen += info[2] > 0
    assert seen > 40 and V[1, 0] > 0.1 and Q[1, 0] - Q[0, 0] > 0.15      # the front car is being pushed, not run through
    Q, V, W = spawn()
    for k in range(80):
        _, info = model.world_step(None, Q, V, W, U, shadowed=[0, 1])
        assert info[2] == 0
    assert abs(V[1, 0]) < 1e-2 and Q[0, 0] > 10.05                        # the shadowed car is not pushed; the other drove on


def test_quad_kernel_source_as_a_warp_of_quads_is_deadlock_free_and_bit_identical(host_quad_kernel):
    """Four quads as ONE warp on the host, every collective a barrier over all 16 threads (what full-mask shuffles and
    __syncthreads_or demand on the device), staged solve with cars that need different numbers of Newton rounds: must
    finish (a quad leaving before a trailing collective hangs it -- the bug that once cost a 20-minute GPU hang is caught
    here) and reproduce the single-quad results bit for bit."""
    import sys
    script = os.path.join(ROOT, "tests", "host_harness", "warp_lockstep_check.py")
    lib = os.path.join(ROOT, "tests", "host_harness", "libstep_quad_host.so")
    out = subprocess.run([sys.executable, script, ROOT, lib, "60"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.startswith("ok") and int(out.stdout.split()[1]) > 100          # suspensions really happened


def _host_walls(host_quad_kernel, track_name="track"):
    """The product's own geometry blob (ftgp_geom_blob, host memory) handed to the host build of the quad kernel."""
    import ft_grandprix_b200 as ft
    t = ft.Track.bundled(track_name)
    lib = ft._lib.load()
    arr = (C.c_void_p * 1)(t._ptr)
    n = lib.ftgp_geom_blob(arr, None, 1, None, 0)
    blob = np.zeros(n, dtype=np.uint32)
    assert lib.ftgp_geom_blob(arr, None, 1, P(blob), n) == n
    out4 = np.zeros(4, dtype=np.int32); size = np.zeros(2)
    assert lib.ftgp_blob_track_view(P(blob), 0, P(out4), P(size)) == 0
    host_quad_kernel.hq_set_walls.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
    base = blob.ctypes.data
    host_quad_kernel.hq_set_walls(C.c_void_p(base + 4 * int(out4[0])), C.c_void_p(base + 4 * int(out4[1])), int(out4[2]), int(out4[3]),
                                  float(size[0]), float(size[1]))
    return blob, t                  # keep the blob alive while the harness points into it


def test_quad_kernel_wall_and_ground_contacts_match_oracle(model, otracks, host_quad_kernel):
    """The product's contact rules (csrc/hfield_contact.cuh + quad_prepare: wheel ellipsoids, chassis hull vertices and the
    lidar cylinder against walls and ground) in the host build of the quad kernel vs the oracle's independent statement
    (oracle/step.c car_contacts), single steps along trajectories that (a) run head-on into walls, (b) slide sideways into
    them so that the wheels touch first, (c) land upside down."""
    blob, t = _host_walls(host_quad_kernel)
    try:
        ot = otracks["track"]
        path = t.path
        rng = np.random.default_rng(11)
        seen = {"wheel_wall": 0, "body_wall": 0, "ground": 0}
        worst = 0.0
        cases = []
        for k in (10, 22, 37, 64, 81):
            d = path[k + 1] - path[k]
            yaw = float(np.arctan2(d[1], d[0]))
            cases.append(("head-on", path[k], yaw + rng.uniform(0.5, 1.2) * rng.choice([-1, 1]), np.array([4.0, 0.0]), None, 0.0))
            cases.append(("sideways", path[k], yaw, np.array([0.0, 0.0]), 2.0 * np.array([-np.sin(yaw), np.cos(yaw)]) * rng.choice([-1, 1]), 0.0))
            cases.append(("flipped", path[k], yaw, np.array([1.0, 0.2]), None, np.pi + rng.normal(0, 0.2)))
        for kind, xy, yaw, ctrl, vlat, roll in cases:
            q, v, w = model.reset(float(xy[0]), float(xy[1]), yaw)
            if roll:
                q[2] = 0.12
                q[3:7] = [np.cos(roll / 2) * np.cos(yaw / 2), np.sin(roll / 2) * np.cos(yaw / 2), np.sin(roll / 2) * np.sin(yaw / 2), np.cos(roll / 2) * np.sin(yaw / 2)]
            for k in range(700):
                if vlat is not None and k == 40:
                    v[0:2] = vlat                                                # shove the settled car sideways
                qh, vh, wh = q.copy(), v.copy(), w.copy()
                info = np.zeros(4, dtype=np.int32)
                host_quad_kernel.hq_step_ghost(P(qh), P(vh), P(wh), P(ctrl), 1, 1, P(info), 0, 0)
                rc, oi = model.step(ot, q, v, w, ctrl)
                assert info[1] == oi[2] and info[2] == oi[3], (kind, k, info, oi)    # wheel-ground and wall contact counts
                err = max(np.abs(qh - q).max(), np.abs(vh - v).max() * 1e-2)
                worst = max(worst, err)
                assert err < 1e-9, (kind, k, err, oi)
                seen["wheel_wall"] += int(oi[6]); seen["body_wall"] += int(oi[3] - oi[6]); seen["ground"] += int(oi[5])
        assert seen["wheel_wall"] > 30 and seen["body_wall"] > 30 and seen["ground"] > 100, seen
    finally:
        host_quad_kernel.hq_set_walls(None, None, 0, 0, 1.0, 1.0)


def test_quad_kernel_staged_solve_with_wall_contacts_is_bit_identical(model, otracks, host_quad_kernel):
    """suspend / resume carries the wall contacts (body + wheel) through the record: staged == unstaged, bit for bit"""
    blob, t = _host_walls(host_quad_kernel)
    try:
        host_quad_kernel.hq_step_staged.argtypes = [C.c_void_p] * 4 + [C.c_long, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        ot = otracks["track"]
        ctrl = np.array([0.0, 0.0])
        nsus = ncon = 0
        for kp in (10, 22, 37, 64, 81):
            for sgn in (-1.0, 1.0):
                d = t.path[kp + 1] - t.path[kp]
                yaw = float(np.arctan2(d[1], d[0]))
                q, v, w = model.reset(float(t.path[kp, 0]), float(t.path[kp, 1]), yaw)
                for k in range(350):
                    if k == 40:
                        v[0:2] = sgn * 2.5 * np.array([-np.sin(yaw), np.cos(yaw)])
                    rc, oi = model.step(ot, q.copy(), v.copy(), w.copy(), ctrl)
                    if oi[3] > 0:                                                # a wall contact this tick: staged vs unstaged
                        qa, va, wa = q.copy(), v.copy(), w.copy(); qb, vb, wb = q.copy(), v.copy(), w.copy()
                        ia = np.zeros(4, dtype=np.int32); ib = np.zeros(4, dtype=np.int32); ns = C.c_int(0)
                        host_quad_kernel.hq_step_ghost(P(qa), P(va), P(wa), P(ctrl), 1, 1, P(ia), 0, 0)
                        host_quad_kernel.hq_step_staged(P(qb), P(vb), P(wb), P(ctrl), 1, P(ib), 1, 1, C.byref(ns))
                        assert np.array_equal(qa, qb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(ia, ib), (kp, k)
                        assert ia[2] == oi[3]
                        ncon += 1; nsus += ns.value
                    model.step(ot, q, v, w, ctrl)
        assert ncon > 20 and nsus > 10, (ncon, nsus)
    finally:
        host_quad_kernel.hq_set_walls(None, None, 0, 0, 1.0, 1.0)


def test_quad_kernel_bubble_wrap_softener_contacts_match_oracle(model, otracks, host_quad_kernel, host_kernel):
    """Option bubble_wrap (custom.py:970-972,1041-1055): the softener spheres (radius 0.0435 > the wheel's 0.03) touch
    the walls before the wheels do.  Host build of the quad kernel vs the oracle, sideways slides into walls; also the
    compile-time constant body_invweight0 of the softener bodies."""
    blob, t = _host_walls(host_quad_kernel)
    bw = oracle_model_with_bubble_wrap = model.__class__()
    bw.set_bubble_wrap(True)
    host_quad_kernel.hq_set_bubble_wrap(1)
    try:
        ot = otracks["track"]
        rng = np.random.default_rng(13)
        soft = total = 0
        for kp in (10, 22, 37, 64, 81):
            d = t.path[kp + 1] - t.path[kp]
            yaw = float(np.arctan2(d[1], d[0]))
            q, v, w = bw.reset(float(t.path[kp, 0]), float(t.path[kp, 1]), yaw)
            ctrl = np.array([0.5, 0.0])
            sgn = rng.choice([-1.0, 1.0])
            for k in range(450):
                if k == 40:
                    v[0:2] = sgn * 2.5 * np.array([-np.sin(yaw), np.cos(yaw)])
                qh, vh, wh = q.copy(), v.copy(), w.copy()
                info = np.zeros(4, dtype=np.int32)
                host_quad_kernel.hq_step_ghost(P(qh), P(vh), P(wh), P(ctrl), 1, 1, P(info), 0, 0)
                rc, oi = bw.step(ot, q, v, w, ctrl)
                assert info[1] == oi[2] and info[2] == oi[3], (kp, k, info, oi)
                err = max(np.abs(qh - q).max(), np.abs(vh - v).max() * 1e-2)
                assert err < 1e-9, (kp, k, err, oi)
                soft += int(oi[7]); total += int(oi[3])
        assert soft > 50 and total > soft, (soft, total)
        # without the option the same slide has no softener contacts
        q, v, w = model.reset(float(t.path[10, 0]), float(t.path[10, 1]), 0.3)
        _, oi = model.step(ot, q, v, w, np.zeros(2))
        assert oi[7] == 0
    finally:
        host_quad_kernel.hq_set_bubble_wrap(0)
        host_quad_kernel.hq_set_walls(None, None, 0, 0, 1.0, 1.0)


@pytest.fixture(scope="module")
def host_world_kernel():
    """The product's world step (csrc/mushr_world.cuh: per-car block-arrow factors + Woodbury over the car-car rows)
    compiled for the host."""
    src = os.path.join(ROOT, "tests", "host_harness", "world_host.cpp")
    out = os.path.join(ROOT, "tests", "host_harness", "libworld_host.so")
    deps = [src] + [os.path.join(ROOT, "ft_grandprix_b200", "csrc", f) for f in ("mushr_world.cuh", "mushr_step_quad.cuh", "mushr_step.cuh", "mushr_consts.h", "mushr_mesh.h", "hfield_contact.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", out, src])
    lib = C.CDLL(out)
    lib.hw_world_step.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p]
    lib.hw_world_has_contact.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    return lib


def test_product_world_step_matches_the_oracle_world_solver(model, host_world_kernel):
    """BASELINE config 5 physics (f1): cars of one world that touch are ONE Newton problem (one direction, one step length,
    one stopping rule).  The product solves it from the per-car block-arrow factors with the Woodbury identity; the oracle
    with a dense Cholesky over 29 N dofs.  Same contact definition, independent code: trajectories of pile-ups of 2..6
    cars agree to 1e-9 per step, with identical Newton iteration counts and contact counts."""
    rng = np.random.default_rng(21)
    total_cc = 0
    for ncars, steps in ((2, 120), (3, 100), (6, 80)):
        Q = np.zeros((ncars, 34)); V = np.zeros((ncars, 29)); W = np.zeros((ncars, 29)); U = np.zeros((ncars, 2))
        for c in range(ncars):                                            # a queue of cars; the rear ones drive into the front ones
            Q[c], V[c], W[c] = model.reset(10.0 + 0.19 * c + rng.normal(0, 0.005), -10.0 + rng.normal(0, 0.01), rng.normal(0, 0.08))
            # (cars spawned at exactly the same height and level have hull vertices exactly ON the other car's box faces:
            # inside / outside is then a matter of the last bit; give every car its own height and a small tilt)
            Q[c, 2] = rng.uniform(0.0, 0.004)
            Q[c, 3:7] += rng.normal(0, 0.01, 4); Q[c, 3:7] /= np.linalg.norm(Q[c, 3:7])
            U[c] = [max(0.0, 3.0 - 0.8 * c), rng.normal(0, 0.1)]
        worst = 0.0
        for k in range(steps):
            Qh, Vh, Wh = Q.copy(), V.copy(), W.copy()
            info = np.zeros(4, dtype=np.int32)
            assert host_world_kernel.hw_world_step(P(Qh), P(Vh), P(Wh), P(U), ncars, None, P(info)) == 0
            flagged = host_world_kernel.hw_world_has_contact(P(Q), ncars, None)
            _, oi = model.world_step(None, Q, V, W, U)
            assert info[1] == oi[2], (ncars, k, info, oi)                # car-car contacts
            assert flagged == (oi[2] > 0)                                  # the cheap pre-test agrees with the contact list
            assert info[0] == oi[0], (ncars, k, info, oi)                # Newton iterations
            err = max(np.abs(Qh - Q).max(), np.abs(Vh - V).max() * 1e-2, np.abs(Wh - W).max() * 1e-4)
            worst = max(worst, err)
            assert err < 1e-9, (ncars, k, err, oi)
            total_cc += int(oi[2])
    assert total_cc > 200


def test_product_world_step_without_contacts_and_with_shadowed_cars(model, host_world_kernel):
    """a world of distant cars == the oracle's world step (shared line search, no coupling rows); shadowed cars are passed through"""
    Q = np.zeros((4, 34)); V = np.zeros((4, 29)); W = np.zeros((4, 29))
    for c in range(4):
        Q[c], V[c], W[c] = model.reset(10.0 + 2.0 * c, -10.0, 0.3 * c)
    U = np.array([[2.0, 0.1], [1.0, -0.2], [3.0, 0.0], [0.0, 0.0]])
    for k in range(60):
        Qh, Vh, Wh = Q.copy(), V.copy(), W.copy()
        info = np.zeros(4, dtype=np.int32)
        host_world_kernel.hw_world_step(P(Qh), P(Vh), P(Wh), P(U), 4, None, P(info))
        model.world_step(None, Q, V, W, U)
        assert info[1] == 0 and np.abs(Qh - Q).max() < 1e-10 and np.abs(Vh - V).max() < 1e-8
    Q = np.zeros((2, 34)); V = np.zeros((2, 29)); W = np.zeros((2, 29))
    Q[0], V[0], W[0] = model.reset(10.0, -10.0, 0.0)
    Q[1], V[1], W[1] = model.reset(10.17, -10.01, 0.1)
    U = np.array([[2.0, 0.0], [0.0, 0.0]])
    sh = np.array([0, 1], dtype=np.uint8)
    for k in range(60):
        Qh, Vh, Wh = Q.copy(), V.copy(), W.copy()
        info = np.zeros(4, dtype=np.int32)
        host_world_kernel.hw_world_step(P(Qh), P(Vh), P(Wh), P(U), 2, P(sh), P(info))
        _, oi = model.world_step(None, Q, V, W, U, shadowed=[0, 1])
        assert info[1] == 0 and oi[2] == 0 and np.abs(Qh - Q).max() < 1e-10


def test_warp_of_quads_with_walls_and_shadowed_cars_is_deadlock_free(host_quad_kernel):
    """Four quads as one warp (every collective a barrier over all 16 threads), cars sliding into walls, option
    bubble_wrap on, some quads stepping shadowed cars (no walls): a collective reached only by the quads with walls hangs
    here -- in round 2 exactly that hung a B200 until the job's time limit.  Results equal the cars stepped alone."""
    import sys
    script = os.path.join(ROOT, "tests", "host_harness", "warp_walls_check.py")
    lib = os.path.join(ROOT, "tests", "host_harness", "libstep_quad_host.so")
    out = subprocess.run([sys.executable, script, ROOT, lib, "160"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.startswith("ok") and int(out.stdout.split()[1]) > 50          # wall contacts really happened


def test_root_only_solves_of_the_coupled_direction_equal_the_full_arrow_solve(host_world_kernel):
    """A car-car contact row touches only the six chassis dofs, so the Woodbury pieces of the coupled-world direction are
    right-hand sides on the root dofs: arrow_root_inverse6 (root rows of the first six columns of H^-1) and
    arrow_solve_root_rhs (H^-1 (b, 0)) skip the chain forward pass.  Both must equal the full block-arrow solve."""
    rng = np.random.default_rng(8)
    NP, NR, NC = 31, 7, 6
    out = np.zeros(2)
    for trial in range(20):
        A = np.zeros((NP, NP))
        G = rng.normal(size=(NR, NR)); A[:NR, :NR] = G @ G.T + NR * np.eye(NR)
        for w in range(4):
            o = NR + NC * w
            G = rng.normal(size=(NC, NC)); A[o:o + NC, o:o + NC] = G @ G.T + NC * np.eye(NC)
            B = 0.3 * rng.normal(size=(NC, NR)); A[o:o + NC, :NR] = B; A[:NR, o:o + NC] = B.T
        assert np.linalg.eigvalsh(A).min() > 0
        br = rng.normal(size=NR); br[6] = 0.0
        host_world_kernel.hw_arrow_root_check(P(A), P(br), P(out))
        assert out[0] < 1e-13 and out[1] < 1e-13, (trial, out)
