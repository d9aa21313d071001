// mushr_world.cuh -- one mj_step of an N-car WORLD (N <= 8) whose cars touch each other: BASELINE config 5's physics.
//
// In the reference every car lives in ONE MjModel (template/mushr.em.xml:95, template/cars/cars.json), so
// mujoco.mj_step (ft_grandprix/custom.py:1425) solves ONE constraint problem over 29 N dofs: a chassis-chassis contact
// couples the two cars' free joints, and the Newton solver has one search direction, one step length and one stopping
// rule for the whole model.  The fast path (mushr_step_quad.cuh) treats cars as independent, which is exact as long as no
// car touches another; worlds with a car-car contact this tick take THIS path instead (step.cu: world_detect_kernel
// flags them, world_step_kernel runs them, one warp per world).
//
// Formulation (B200-first, not MuJoCo's dense model-wide Cholesky): the Hessian is
//     H = blockdiag(H_1 .. H_N) + U' D U,      H_c = M_c + J_c' D_c J_c  (block-arrow, factorised per car as before),
// with U the (few) active pyramid rows of the car-car contacts, each touching only the six chassis dofs of two cars.  The
// Newton direction follows from the per-car factors by the Woodbury identity:
//     d0_c = H_c^-1 g_c,   Z = H0^-1 U',   G = D^-1 + U Z (rows x rows, <= 64),   y = G^-1 (U d0),   d = d0 - Z y.
// Cost along the line, step length and termination are world sums (mj_solPrimal over the model).
//
// Car-car contacts are this framework's definition (as in the oracle, oracle/step.c world_assemble, written there from
// independent code): a chassis hull vertex of car A strictly inside the bounding box of car B's hull gives one condim-3
// contact, normal = B's box face of least penetration pointing out of B, dist = -penetration, mu = 1, default
// solref / solimp, rows J = J_A(p) - J_B(p).
//
// Work split: lane c of the world's group handles car c (position stage, per-car factor and solves); the small coupled
// pieces are computed redundantly by every lane from the shared workspace, so control flow is uniform.  `Comm` gives
// lane / nlanes / sync(); on the host (tests) nlanes = 1 and the loops run sequentially.
#pragma once
#include "mushr_step.cuh"
#include "hfield_contact.cuh"

namespace ftgp {
namespace mushr {

constexpr int WMAXCARS = 8, WMAXCC = 16, WMAXROWS = 4 * WMAXCC, MAXBODY = 8;

struct CarWork {
    Arrow M, H;
    Rows r;
    Solver s;
    double qfrc_smooth[NP], qacc_smooth[NP], qfrc_con[NP], v[NP], wa[NP], d0[NP];
    double p1[3], R1[9], com[3], cdof6[6][6];
    LsCtx ls;
    double cost_w, cost_s;          // cost of the warm start / of qacc_smooth (this car's share)
    int nwheel, nwall, nground;
    bool shadowed;
};
struct CcContact {
    int a, b;
    double dist, Ja[3][6], Jb[3][6], D, aref[4];
    double dxa[3], dsa[3];          // line search: J x and J s (both cars summed)
    unsigned active;                // pyramid rows with jar < 0 at the current point
};
struct WorldWork {
    CarWork car[WMAXCARS];
    CcContact cc[WMAXCC];
    int ncc, nrows;
    int row_cc[WMAXROWS], row_rr[WMAXROWS];
    double S6[WMAXCARS][6][NR];     // root rows of columns 0-5 of H_c^-1: every row of U touches only the six chassis dofs of its two cars
    int touched[WMAXCARS];          // the car takes part in an active car-car row
    double G[WMAXROWS * WMAXROWS], y[WMAXROWS];
    double lsq[WMAXCARS][3];        // line search: each car's share of the quadratic along the line (written by its lane)
};

struct SeqComm {                    // host: one "lane" does everything
    int lane = 0, nlanes = 1;
    FT_HD void sync() const {}
};

// ---- the full contact set of one car, thread-per-car flavour (same rules and order as quad_prepare / the oracle)
template <class WallFn>
FT_HDN void car_contacts(const ModelConsts& mc, const Kin& k, const WallFn& walls, Rows& r, int& nwall, int& nground) {
    nwall = nground = 0;
    const bool wall_on = walls.enabled();
    auto fill = [&](Contact& c, const QWallHit& h, int wheel, int nchain, double dmin, double tran) {
        c.dist = h.dist; c.mu = 1.0; c.dmin = dmin; c.wheel = wheel; c.nchain = nchain; c.tran = tran;
        const double off[3] = {h.pnt[0] - k.com[0], h.pnt[1] - k.com[1], h.pnt[2] - k.com[2]};
        for (int col = 0; col < 12; col++) {
            double jp[3] = {0, 0, 0};
            if (col < 6 + nchain) {
                const int p = col < 6 ? col : NR + NC * wheel + (col - 6);
                cross3(jp, k.cdof[p], off);
                for (int a = 0; a < 3; a++) jp[a] += k.cdof[p][3 + a];
            }
            c.J[0][col] = dot3(h.nrm, jp); c.J[1][col] = dot3(h.t1, jp); c.J[2][col] = dot3(h.t2, jp);
        }
    };
    if (wall_on) {
        const double wsz[3] = {WS0, WS1, WS2}, ssz[1] = {MUSHR_SOFTENER_RADIUS};
        for (int w = 0; w < 4; w++) {
            QWallHit h;
            if (walls.convex(HF_ELLIPSOID, wsz, 0.03, k.pw[w], k.Rw[w], h)) { fill(r.con[r.ncon++], h, w, 3, 0.45, mc.wheel_invweight0[w]); nwall++; }
        }
        if (walls.bubble_wrap())
            for (int w = 0; w < 4; w++) {
                QWallHit h;
                if (walls.convex(HF_SPHERE, ssz, MUSHR_SOFTENER_RADIUS, k.ps[w], k.Rs[w], h)) { fill(r.con[r.ncon++], h, w, 6, 0.9, mc.soft_invweight0[w]); nwall++; }
            }
    }
    const double hull[MUSHR_CHASSIS_NHULL][3] = MUSHR_CHASSIS_HULL;
    const double csz[2] = {0.03, 0.015}, cloc[3] = {-0.0525, 0.0, 0.065 - 0.015 / 2}, down[3] = {0, 0, -1};
    int nb = 0;
    for (int v = 0; v < MUSHR_CHASSIS_NHULL; v++) {
        double p[3];
        mat_vec3(p, k.R1, hull[v]);
        for (int a = 0; a < 3; a++) p[a] += k.p1[a];
        QWallHit h;
        if (wall_on && nb < MAXBODY && walls.vertex(p, h)) { fill(r.con[r.ncon++], h, -1, 0, 0.9, mc.chassis_invweight0); nb++; nwall++; }
        if (nb < MAXBODY && ground_probe(p, h)) { fill(r.con[r.ncon++], h, -1, 0, 0.9, mc.chassis_invweight0); nb++; nground++; }
    }
    double pc[3], sp[3];
    mat_vec3(pc, k.R1, cloc);
    for (int a = 0; a < 3; a++) pc[a] += k.p1[a];
    QWallHit h;
    if (wall_on && nb < MAXBODY && walls.convex(HF_CYLINDER, csz, 0.0336, pc, k.R1, h)) { fill(r.con[r.ncon++], h, -1, 0, 0.9, mc.chassis_invweight0); nb++; nwall++; }
    hf_support(HF_CYLINDER, csz, pc, k.R1, down, sp);
    if (nb < MAXBODY && ground_probe(sp, h)) { fill(r.con[r.ncon++], h, -1, 0, 0.9, mc.chassis_invweight0); nb++; nground++; }
}

// ---- position + velocity stage of one car: everything up to qacc_smooth (as step_car)
template <class WallFn>
FT_HDN void world_car_prepare(const ModelConsts& mc, const double* qpos, const double* qvel, const double* warm, const double* ctrl,
                              const WallFn& walls, CarWork& c, Kin& k) {
    for (int p = 0; p < NP; p++) { const int d = p2d(p); c.v[p] = d >= 0 ? qvel[d] : 0.0; c.wa[p] = d >= 0 ? warm[d] : 0.0; }
    kinematics(mc, qpos, k);
    mass_matrix(k, c.M);
    c.r.ncon = 0;
    wheel_contacts(mc, k, c.r);
    c.nwheel = c.r.ncon;
    car_contacts(mc, k, walls, c.r, c.nwall, c.nground);
    make_rows(mc, qpos, c.v, c.r);
    bias_forces(k, c.v, c.qfrc_smooth);
    for (int p = 0; p < NP; p++) c.qfrc_smooth[p] = -c.qfrc_smooth[p] - dof_damping(p) * c.v[p];
    for (int w = 0; w < 4; w++) c.qfrc_smooth[NR + NC * w] += -500.0 * (qpos[chain_q(w)] - (-0.015));
    for (int w = 2; w < 4; w++) c.qfrc_smooth[NR + NC * w + 1] = 0;
    c.qfrc_smooth[6] += 20.0 * ctrl[1] - 20.0 * qpos[7];
    double tv = 0;
    for (int w = 0; w < 4; w++) tv += 0.25 * c.v[NR + NC * w + 2];
    double f = 100.0 * ctrl[0] - 100.0 * (0.04 * tv);
    f = f > 500.0 ? 500.0 : (f < -500.0 ? -500.0 : f);
    for (int w = 0; w < 4; w++) c.qfrc_smooth[NR + NC * w + 2] += 0.04 * 0.25 * f;
    c.H = c.M;
    arrow_factor(c.H);
    for (int p = 0; p < NP; p++) c.qacc_smooth[p] = c.qfrc_smooth[p];
    arrow_solve(c.H, c.qacc_smooth);
    for (int a = 0; a < 3; a++) { c.p1[a] = k.p1[a]; c.com[a] = k.com[a]; }
    for (int a = 0; a < 9; a++) c.R1[a] = k.R1[a];
    for (int i = 0; i < 6; i++) for (int a = 0; a < 6; a++) c.cdof6[i][a] = k.cdof[i][a];
}

// ---- car-car contacts of the world (every lane computes the same list)
FT_HDN void world_detect(const ModelConsts& mc, int ncars, const double* const* qvel, WorldWork& W) {
    const double hull[MUSHR_CHASSIS_NHULL][3] = MUSHR_CHASSIS_HULL;
    double lo[3] = {1e9, 1e9, 1e9}, hi[3] = {-1e9, -1e9, -1e9};
    for (int v = 0; v < MUSHR_CHASSIS_NHULL; v++) for (int a = 0; a < 3; a++) { lo[a] = fmin(lo[a], hull[v][a]); hi[a] = fmax(hi[a], hull[v][a]); }
    W.ncc = 0;
    for (int A = 0; A < ncars; A++) for (int B = 0; B < ncars; B++) {
        if (A == B || W.car[A].shadowed || W.car[B].shadowed) continue;
        const CarWork& ca = W.car[A]; const CarWork& cb = W.car[B];
        {   // both hulls lie within 0.125 m of their body origins: far pairs have no vertex in the other's box
            const double dx = ca.p1[0] - cb.p1[0], dy = ca.p1[1] - cb.p1[1], dz = ca.p1[2] - cb.p1[2];
            if (dx * dx + dy * dy + dz * dz > 0.26 * 0.26) continue;
        }
        for (int v = 0; v < MUSHR_CHASSIS_NHULL && W.ncc < WMAXCC; v++) {
            double pw[3], rel[3], ql[3];
            mat_vec3(pw, ca.R1, hull[v]);
            for (int a = 0; a < 3; a++) { pw[a] += ca.p1[a]; rel[a] = pw[a] - cb.p1[a]; }
            for (int a = 0; a < 3; a++) ql[a] = cb.R1[a] * rel[0] + cb.R1[3 + a] * rel[1] + cb.R1[6 + a] * rel[2];
            if (ql[0] <= lo[0] || ql[0] >= hi[0] || ql[1] <= lo[1] || ql[1] >= hi[1] || ql[2] <= lo[2] || ql[2] >= hi[2]) continue;
            int axis = 0; double depth = 1e9, sign = 1;
            for (int a = 0; a < 3; a++) {
                if (ql[a] - lo[a] < depth) { depth = ql[a] - lo[a]; axis = a; sign = -1; }
                if (hi[a] - ql[a] < depth) { depth = hi[a] - ql[a]; axis = a; sign = 1; }
            }
            CcContact& c = W.cc[W.ncc++];
            c.a = A; c.b = B; c.dist = -depth;
            QWallHit h;
            for (int a = 0; a < 3; a++) h.nrm[a] = sign * cb.R1[3 * a + axis];
            h.dist = -depth;
            hf_frame(h, pw);                                            // tangents; pnt = pw + n depth / 2
            double vel[3] = {0, 0, 0};
            for (int side = 0; side < 2; side++) {
                const CarWork& cc_ = side == 0 ? ca : cb;
                const double* qv = qvel[side == 0 ? A : B];
                const double off[3] = {h.pnt[0] - cc_.com[0], h.pnt[1] - cc_.com[1], h.pnt[2] - cc_.com[2]};
                for (int col = 0; col < 6; col++) {
                    double jp[3];
                    cross3(jp, cc_.cdof6[col], off);
                    for (int a = 0; a < 3; a++) jp[a] += cc_.cdof6[col][3 + a];
                    const double sg = side == 0 ? 1.0 : -1.0;
                    const double j0 = sg * dot3(h.nrm, jp), j1 = sg * dot3(h.t1, jp), j2 = sg * dot3(h.t2, jp);
                    if (side == 0) { c.Ja[0][col] = j0; c.Ja[1][col] = j1; c.Ja[2][col] = j2; }
                    else { c.Jb[0][col] = j0; c.Jb[1][col] = j1; c.Jb[2][col] = j2; }
                    vel[0] += j0 * qv[col]; vel[1] += j1 * qv[col]; vel[2] += j2 * qv[col];
                }
            }
            double K, Bc, imp, R;
            kbi(0.9, c.dist, 2 * mc.chassis_invweight0, K, Bc, imp, R);
            double Rpy = 2 * R; if (Rpy < MINVAL) Rpy = MINVAL;          // mu = 1
            c.D = 1 / Rpy;
            for (int rr = 0; rr < 4; rr++) {
                const double sg = (rr & 1) ? -1.0 : 1.0;
                c.aref[rr] = -Bc * (vel[0] + sg * vel[1 + (rr >> 1)]) - K * imp * c.dist;
            }
            c.active = 0;
        }
    }
}

// J x of a car-car contact in its frame, x = per-car padded vectors selected by `which`
FT_HD void cc_dots(const CcContact& c, const double* xa, const double* xb, double* d3) {
    for (int a = 0; a < 3; a++) {
        double s = 0;
        for (int col = 0; col < 6; col++) s += c.Ja[a][col] * xa[col] + c.Jb[a][col] * xb[col];
        d3[a] = s;
    }
}

// cost of the car-car rows at the vectors selected by `sel` (0: qacc, 1: warm start, 2: qacc_smooth); optionally forces
// into the cars' qfrc_con / grad and the active masks
FT_HDN double cc_cost(WorldWork& W, int sel, bool apply) {
    double cost = 0;
    for (int k = 0; k < W.ncc; k++) {
        CcContact& c = W.cc[k];
        CarWork& A = W.car[c.a]; CarWork& B = W.car[c.b];
        const double* xa = sel == 0 ? A.s.qacc : (sel == 1 ? A.wa : A.qacc_smooth);
        const double* xb = sel == 0 ? B.s.qacc : (sel == 1 ? B.wa : B.qacc_smooth);
        double d3[3];
        cc_dots(c, xa, xb, d3);
        unsigned act = 0;
        for (int rr = 0; rr < 4; rr++) {
            const double sg = (rr & 1) ? -1.0 : 1.0; const int ta = 1 + (rr >> 1);
            const double jar = d3[0] + sg * d3[ta] - c.aref[rr];
            if (jar >= 0) continue;
            cost += 0.5 * c.D * jar * jar;
            act |= 1u << rr;
            if (apply) {
                const double f = -c.D * jar;
                for (int col = 0; col < 6; col++) {
                    const double fa = (c.Ja[0][col] + sg * c.Ja[ta][col]) * f, fb = (c.Jb[0][col] + sg * c.Jb[ta][col]) * f;
                    A.qfrc_con[col] += fa; A.s.grad[col] -= fa;
                    B.qfrc_con[col] += fb; B.s.grad[col] -= fb;
                }
            }
        }
        if (apply) c.active = act;
    }
    return cost;
}

struct WorldInfo { int iters, ncc, reset; };

// ---- the step.  qpos / qvel / warm / ctrl: per-car pointers; walls[c]: that car's walls (disabled for a shadowed car)
template <class Comm, class WallFn>
FT_HDN void world_step(const Comm& cm, const ModelConsts& mc, int ncars, double* const* qpos, double* const* qvel, double* const* warm,
                       const double* const* ctrl, const WallFn* walls, const bool* shadowed, WorldWork& W, Kin& kin, WorldInfo& info) {
    info.iters = 0; info.ncc = 0; info.reset = 0;
    // mj_checkPos / mj_checkVel: a bad value anywhere resets the whole data
    bool bad = false;
    for (int c = 0; c < ncars; c++) {
        for (int i = 0; i < NQ; i++) bad |= bad_value(qpos[c][i]);
        for (int i = 0; i < NV; i++) bad |= bad_value(qvel[c][i]);
    }
    cm.sync();
    if (bad) {
        info.reset = 1;
        for (int c = cm.lane; c < ncars; c += cm.nlanes) reset_state(qpos[c], qvel[c], warm[c]);
        cm.sync();
    }
    for (int c = cm.lane; c < ncars; c += cm.nlanes) {
        W.car[c].shadowed = shadowed && shadowed[c];
        world_car_prepare(mc, qpos[c], qvel[c], warm[c], ctrl[c], walls[c], W.car[c], kin);     // kin: this lane's scratch
    }
    cm.sync();
    if (cm.lane == 0) {
        const double* qv[WMAXCARS];
        for (int c = 0; c < ncars; c++) qv[c] = qvel[c];
        world_detect(mc, ncars, qv, W);
    }
    cm.sync();
    info.ncc = W.ncc;
    // ---- warm start if its cost beats qacc_smooth's, for the whole world (mj_fwdConstraint)
    for (int c = cm.lane; c < ncars; c += cm.nlanes) {
        CarWork& C = W.car[c];
        arrow_mul(C.M, C.wa, C.s.Ma);
        double cw = rows_cost<false, false>(C.r, C.wa, nullptr, nullptr);
        for (int p = 0; p < NP; p++) cw += 0.5 * (C.s.Ma[p] - C.qfrc_smooth[p]) * (C.wa[p] - C.qacc_smooth[p]);
        C.cost_w = cw;
        C.cost_s = rows_cost<false, false>(C.r, C.qacc_smooth, nullptr, nullptr);
    }
    cm.sync();
    double cw = 0, cs = 0;
    for (int c = 0; c < ncars; c++) { cw += W.car[c].cost_w; cs += W.car[c].cost_s; }
    cw += cc_cost(W, 1, false); cs += cc_cost(W, 2, false);
    const bool cold = cw > cs;
    for (int c = cm.lane; c < ncars; c += cm.nlanes) {
        CarWork& C = W.car[c];
        if (cold) for (int p = 0; p < NP; p++) { C.s.qacc[p] = C.qacc_smooth[p]; C.s.Ma[p] = C.qfrc_smooth[p]; }
        else for (int p = 0; p < NP; p++) C.s.qacc[p] = C.wa[p];
    }
    cm.sync();
    const double scale = 1.0 / (mc.meaninertia * NV * ncars);
    double cost = 0, gnorm2 = 0;
    // cost, forces, gradient at the current point (per car, then the car-car rows on top)
    auto evaluate = [&]() {
        for (int c = cm.lane; c < ncars; c += cm.nlanes) newton_evaluate(W.car[c].r, W.car[c].s, W.car[c].qfrc_smooth, W.car[c].qacc_smooth, W.car[c].qfrc_con);
        cm.sync();
        double ccost = 0;
        if (cm.lane == 0) ccost = cc_cost(W, 0, true);
        cm.sync();
        if (cm.lane != 0) ccost = cc_cost(W, 0, false);                  // (same number, no side effects)
        cost = ccost; gnorm2 = 0;
        for (int c = 0; c < ncars; c++) {
            cost += W.car[c].s.cost;
            for (int p = 0; p < NP; p++) gnorm2 += W.car[c].s.grad[p] * W.car[c].s.grad[p];
        }
    };
    // Newton direction by Woodbury over the active car-car rows.  U's rows are non-zero only on the six chassis dofs of their
    // two cars: G = D^-1 + sum_c u_rc' S_c u_sc needs only S_c, the 6 x 6 corner of H_c^-1 (root-only solves with the Schur
    // factor, no chain work), and d_c = d0_c - H_c^-1 (w_c, 0) with w_c = sum_r y_r u_rc is one more root-only right-hand side.  Rows of G are spread over the lanes; its Cholesky is right-looking with the rows spread likewise.
    auto row_u = [&](int r, int side, double* u) {
        const CcContact& k = W.cc[W.row_cc[r]];
        const double sg = (W.row_rr[r] & 1) ? -1.0 : 1.0; const int ta = 1 + (W.row_rr[r] >> 1);
        for (int col = 0; col < 6; col++) u[col] = side == 0 ? k.Ja[0][col] + sg * k.Ja[ta][col] : k.Jb[0][col] + sg * k.Jb[ta][col];
    };
    auto direction = [&]() {
        if (cm.lane == 0) {
            W.nrows = 0;
            for (int c = 0; c < ncars; c++) W.touched[c] = 0;
            for (int k = 0; k < W.ncc; k++) for (int rr = 0; rr < 4; rr++) if (W.cc[k].active >> rr & 1u) {
                W.row_cc[W.nrows] = k; W.row_rr[W.nrows] = rr; W.nrows++;
                W.touched[W.cc[k].a] = 1; W.touched[W.cc[k].b] = 1;
            }
        }
        cm.sync();
        const int n = W.nrows;
        for (int c = cm.lane; c < ncars; c += cm.nlanes) {
            CarWork& C = W.car[c];
            C.H = C.M;
            rows_cost<false, true>(C.r, C.s.qacc, nullptr, &C.H);
            arrow_factor(C.H);
            for (int p = 0; p < NP; p++) C.d0[p] = C.s.grad[p];
            arrow_solve(C.H, C.d0);
            if (W.touched[c]) arrow_root_inverse6(C.H, W.S6[c]);
        }
        cm.sync();
        if (n > 0) {
            for (int r = cm.lane; r < n; r += cm.nlanes) {                // row r of G and of the right-hand side U d0
                const CcContact& kr = W.cc[W.row_cc[r]];
                double ua[6], ub[6], sa[6], sb[6], rhs = 0;
                row_u(r, 0, ua); row_u(r, 1, ub);
                for (int col = 0; col < 6; col++) rhs += ua[col] * W.car[kr.a].d0[col] + ub[col] * W.car[kr.b].d0[col];
                W.y[r] = rhs;
                // sa = S_a ua, sb = S_b ub (S symmetric: H_c^-1[i][j] = S6[i][j])
                for (int i = 0; i < 6; i++) {
                    double ta_ = 0, tb_ = 0;
                    for (int j = 0; j < 6; j++) { ta_ += W.S6[kr.a][j][i] * ua[j]; tb_ += W.S6[kr.b][j][i] * ub[j]; }
                    sa[i] = ta_; sb[i] = tb_;
                }
                for (int s2 = 0; s2 < n; s2++) {
                    const CcContact& ks = W.cc[W.row_cc[s2]];
                    double va[6], vb[6], g = r == s2 ? 1.0 / kr.D : 0.0;
                    row_u(s2, 0, va); row_u(s2, 1, vb);
                    if (ks.a == kr.a) for (int i = 0; i < 6; i++) g += sa[i] * va[i];
                    if (ks.b == kr.a) for (int i = 0; i < 6; i++) g += sa[i] * vb[i];
                    if (ks.a == kr.b) for (int i = 0; i < 6; i++) g += sb[i] * va[i];
                    if (ks.b == kr.b) for (int i = 0; i < 6; i++) g += sb[i] * vb[i];
                    W.G[r * n + s2] = g;
                }
            }
            cm.sync();
            for (int j = 0; j < n; j++) {                                // G = L L' (SPD: D^-1 plus a Gram matrix)
                double d = W.G[j * n + j];
                if (d < MINVAL) d = MINVAL;
                d = sqrt(d);
                cm.sync();                                               // (everybody has read G[j][j])
                if (cm.lane == j % cm.nlanes) W.G[j * n + j] = d;
                for (int i = j + 1 + cm.lane; i < n; i += cm.nlanes) W.G[i * n + j] /= d;
                cm.sync();
                for (int i = j + 1 + cm.lane; i < n; i += cm.nlanes) {
                    const double lij = W.G[i * n + j];
                    for (int q = j + 1; q <= i; q++) W.G[i * n + q] -= lij * W.G[q * n + j];
                }
                cm.sync();
            }
            if (cm.lane == 0) {                                          // y = G^-1 (U d0)
                for (int i = 0; i < n; i++) { double t = W.y[i]; for (int q = 0; q < i; q++) t -= W.G[i * n + q] * W.y[q]; W.y[i] = t / W.G[i * n + i]; }
                for (int i = n - 1; i >= 0; i--) { double t = W.y[i]; for (int q = i + 1; q < n; q++) t -= W.G[q * n + i] * W.y[q]; W.y[i] = t / W.G[i * n + i]; }
            }
            cm.sync();
        }
        for (int c = cm.lane; c < ncars; c += cm.nlanes) {
            CarWork& C = W.car[c];
            if (n > 0 && W.touched[c]) {
                double w6[6] = {0, 0, 0, 0, 0, 0}, u[6];
                for (int r = 0; r < n; r++) {
                    const CcContact& k = W.cc[W.row_cc[r]];
                    if (k.a == c) { row_u(r, 0, u); for (int i = 0; i < 6; i++) w6[i] += W.y[r] * u[i]; }
                    if (k.b == c) { row_u(r, 1, u); for (int i = 0; i < 6; i++) w6[i] += W.y[r] * u[i]; }
                }
                double dl[NP];                                           // H_c^-1 (w, 0): a right-hand side on the root dofs only
                for (int p = 0; p < NP; p++) dl[p] = p < 6 ? w6[p] : 0.0;
                arrow_solve_root_rhs(C.H, dl);
                for (int p = 0; p < NP; p++) C.d0[p] -= dl[p];
            }
            for (int p = 0; p < NP; p++) C.s.search[p] = -C.d0[p];
        }
        cm.sync();
    };
    // (every lane evaluates its own car's rows, the shares are summed by everybody: evaluating all eight cars in every lane
    // made the line search 8x the cost of the factorisations)
    auto ls_world = [&](LsPoint& pt, double alpha) {
        for (int c = cm.lane; c < ncars; c += cm.nlanes) {
            double a0 = 0, a1 = 0, a2 = 0;
            ls_quad(W.car[c].ls, alpha, a0, a1, a2);
            W.lsq[c][0] = a0; W.lsq[c][1] = a1; W.lsq[c][2] = a2;
        }
        cm.sync();
        double q0 = 0, q1 = 0, q2 = 0;
        for (int c = 0; c < ncars; c++) { q0 += W.lsq[c][0]; q1 += W.lsq[c][1]; q2 += W.lsq[c][2]; }
        for (int k = 0; k < W.ncc; k++) {
            const CcContact& c = W.cc[k];
            for (int rr = 0; rr < 4; rr++) {
                const double sg = (rr & 1) ? -1.0 : 1.0; const int ta = 1 + (rr >> 1);
                const double jar = c.dxa[0] + sg * c.dxa[ta] - c.aref[rr], jv = c.dsa[0] + sg * c.dsa[ta];
                if (jar + alpha * jv < 0) { q0 += 0.5 * c.D * jar * jar; q1 += c.D * jar * jv; q2 += 0.5 * c.D * jv * jv; }
            }
        }
        cm.sync();                                                       // (everybody has read the shares before they are overwritten)
        ls_point(pt, alpha, q0, q1, q2);
    };
    auto line_search_world = [&]() -> double {
        for (int c = cm.lane; c < ncars; c += cm.nlanes) {
            CarWork& C = W.car[c];
            arrow_mul(C.M, C.s.search, C.s.Mv);
            LsCtx& L = C.ls;
            L.r = &C.r; L.x = C.s.qacc; L.s = C.s.search;
            L.qg0 = C.s.gauss; L.qg1 = 0; L.qg2 = 0;
            for (int p = 0; p < NP; p++) { L.qg1 += C.s.search[p] * (C.s.Ma[p] - C.qfrc_smooth[p]); L.qg2 += 0.5 * C.s.search[p] * C.s.Mv[p]; }
            for (int k = 0; k < C.r.ncon; k++) { contact_dots(C.r.con[k], C.s.qacc, L.cdx[k]); contact_dots(C.r.con[k], C.s.search, L.cds[k]); }
        }
        if (cm.lane == 0)
            for (int k = 0; k < W.ncc; k++) {
                CcContact& c = W.cc[k];
                cc_dots(c, W.car[c.a].s.qacc, W.car[c.b].s.qacc, c.dxa);
                cc_dots(c, W.car[c.a].s.search, W.car[c.b].s.search, c.dsa);
            }
        cm.sync();
        double snorm = 0;
        for (int c = 0; c < ncars; c++) for (int p = 0; p < NP; p++) snorm += W.car[c].s.search[p] * W.car[c].s.search[p];
        snorm = sqrt(snorm);
        if (snorm < MINVAL) return 0.0;
        const double gtol = SOLVER_TOL * LS_TOL * snorm / scale;
        LsPoint p0, p1, p2, pm, a1, a2;
        int it = 0;
        ls_world(p0, 0);
        ls_world(p1, p0.alpha - p0.d0 / p0.d1);
        if (p0.cost < p1.cost) p1 = p0;
        if (fabs(p1.d0) < gtol) return p1.alpha;
        const double dir = p1.d0 < 0 ? 1.0 : -1.0;
        bool p2update = false;
        p2 = p1;
        while (p1.d0 * dir <= -gtol && it < LS_ITER) {
            p2 = p1; p2update = true;
            ls_world(p1, p1.alpha - p1.d0 / p1.d1); it++;
            if (fabs(p1.d0) < gtol) return p1.alpha;
        }
        if (it >= LS_ITER || !p2update) return p1.alpha;
        while (it < LS_ITER) {
            ls_world(pm, 0.5 * (p1.alpha + p2.alpha)); it++;
            ls_world(a1, p1.alpha - p1.d0 / p1.d1);
            ls_world(a2, p2.alpha - p2.d0 / p2.d1);
            if (fabs(a1.d0) < gtol) return a1.alpha;
            if (fabs(a2.d0) < gtol) return a2.alpha;
            if (fabs(pm.d0) < gtol) return pm.alpha;
            bool b1 = false, b2 = false;
            double lo = fmin(p1.alpha, p2.alpha), hi = fmax(p1.alpha, p2.alpha);
            for (int cnd = 0; cnd < 3; cnd++) {
                const LsPoint& q = cnd == 0 ? a1 : (cnd == 1 ? a2 : pm);
                if (q.alpha <= lo || q.alpha >= hi) continue;
                if ((q.d0 < 0) == (p1.d0 < 0)) { p1 = q; b1 = true; } else { p2 = q; b2 = true; }
                lo = fmin(p1.alpha, p2.alpha); hi = fmax(p1.alpha, p2.alpha);
            }
            if (!b1 && !b2) break;
        }
        return p1.cost <= p2.cost ? p1.alpha : p2.alpha;
    };

    evaluate();
    direction();
    int iter = 0;
    while (iter < SOLVER_ITER) {
        const double alpha = line_search_world();
        cm.sync();
        if (alpha == 0) break;
        for (int c = cm.lane; c < ncars; c += cm.nlanes) {
            Solver& s = W.car[c].s;
            for (int p = 0; p < NP; p++) { s.qacc[p] += alpha * s.search[p]; s.Ma[p] += alpha * s.Mv[p]; }
        }
        cm.sync();
        const double oldcost = cost;
        evaluate();
        iter++;
        direction();                                                     // (MuJoCo factorises before the test; the direction is unused if it ends the loop)
        if (scale * (oldcost - cost) < SOLVER_TOL || scale * sqrt(gnorm2) < SOLVER_TOL) break;
    }
    info.iters = iter;
    bool badacc = false;
    for (int c = 0; c < ncars; c++) for (int p = 0; p < NP; p++) badacc |= bad_value(W.car[c].s.qacc[p]);
    cm.sync();
    if (badacc) {                                                        // mj_checkAcc
        info.reset = 1;
        for (int c = cm.lane; c < ncars; c += cm.nlanes) reset_state(qpos[c], qvel[c], warm[c]);
        cm.sync();
        return;
    }
    // ---- mj_Euler with implicit joint damping + mj_integratePos, per car
    for (int c = cm.lane; c < ncars; c += cm.nlanes) {
        CarWork& C = W.car[c];
        C.H = C.M;
        C.H.R[tri(6, 6)] += TIMESTEP * dof_damping(6);
        for (int w = 0; w < 4; w++) for (int l = 0; l < 3; l++) if (!(l == 1 && !front(w))) C.H.W[w][tri(l, l)] += TIMESTEP * dof_damping(NR + NC * w + l);
        arrow_factor(C.H);
        double qa[NP];
        for (int p = 0; p < NP; p++) qa[p] = C.qfrc_smooth[p] + C.qfrc_con[p];
        arrow_solve(C.H, qa);
        double* q = qpos[c]; double* v = qvel[c]; double* wm = warm[c];
        for (int p = 0; p < NP; p++) { const int d = p2d(p); if (d >= 0) { wm[d] = C.s.qacc[p]; v[d] += TIMESTEP * qa[p]; } }
        for (int a = 0; a < 3; a++) q[a] += TIMESTEP * v[a];
        quat_integrate(q + 3, v + 3, TIMESTEP);
        q[7] += TIMESTEP * v[6];
        for (int w = 0; w < 4; w++) {
            const int qa0 = chain_q(w), d0 = chain_d(w), nh = front(w) ? 3 : 2;
            for (int l = 0; l < nh; l++) q[qa0 + l] += TIMESTEP * v[d0 + l];
            quat_integrate(q + qa0 + nh, v + d0 + nh, TIMESTEP);
        }
    }
    cm.sync();
}

// does any car of the world touch another one?  (cheap pre-test on the poses alone: hull vertex of A inside B's hull box)
FT_HDN bool world_has_contact(int ncars, const double* const* qpos, const bool* shadowed) {
    const double hull[MUSHR_CHASSIS_NHULL][3] = MUSHR_CHASSIS_HULL;
    double lo[3] = {1e9, 1e9, 1e9}, hi[3] = {-1e9, -1e9, -1e9};
    for (int v = 0; v < MUSHR_CHASSIS_NHULL; v++) for (int a = 0; a < 3; a++) { lo[a] = fmin(lo[a], hull[v][a]); hi[a] = fmax(hi[a], hull[v][a]); }
    double R[WMAXCARS][9];
    for (int c = 0; c < ncars; c++) { double q[4] = {qpos[c][3], qpos[c][4], qpos[c][5], qpos[c][6]}; quat_norm(q); quat2mat(R[c], q); }
    for (int A = 0; A < ncars; A++) for (int B = 0; B < ncars; B++) {
        if (A == B || (shadowed && (shadowed[A] || shadowed[B]))) continue;
        const double dx = qpos[A][0] - qpos[B][0], dy = qpos[A][1] - qpos[B][1], dz = qpos[A][2] - qpos[B][2];
        if (dx * dx + dy * dy + dz * dz > 0.26 * 0.26) continue;          // two hulls of radius < 0.125 about their origins
        for (int v = 0; v < MUSHR_CHASSIS_NHULL; v++) {
            double pw[3], rel[3], ql[3];
            mat_vec3(pw, R[A], hull[v]);
            for (int a = 0; a < 3; a++) rel[a] = pw[a] + qpos[A][a] - qpos[B][a];
            for (int a = 0; a < 3; a++) ql[a] = R[B][a] * rel[0] + R[B][3 + a] * rel[1] + R[B][6 + a] * rel[2];
            if (ql[0] > lo[0] && ql[0] < hi[0] && ql[1] > lo[1] && ql[1] < hi[1] && ql[2] > lo[2] && ql[2] < hi[2]) return true;
        }
    }
    return false;
}

}  // namespace mushr
}  // namespace ftgp
