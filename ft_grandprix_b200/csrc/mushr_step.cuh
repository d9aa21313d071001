// mushr_step.cuh -- one mj_step for one car of template/mushr.em.xml, specialised for its fixed
// topology.  Replaces mujoco.mj_step (ft_grandprix/custom.py:1425).
//
// B200-first design (not a port of MuJoCo's generic engine):
//   * the kinematic tree is compiled into the code: a 7-dof "root" (free joint + virtual steering-wheel
//     hinge, dofs 0-6) and four 6-slot wheel chains (suspension slide, steering hinge [front only; a
//     unit dummy slot at the rear], throttle hinge, 3 ball dofs of the 1e-5 kg softener body);
//   * mass matrix M and Newton Hessian H = M + J^T D J are stored and factorised as BLOCK-ARROW
//     matrices (root 7x7 + four 6x6 chain blocks + four 6x7 borders, 274 doubles) instead of dense
//     29x29 (841): every constraint of this model (joint equalities, friction loss, limits, wheel
//     and chassis contacts) keeps that sparsity, so one factorisation costs ~2.7 kflop instead of 8.1;
//   * constraint rows are implicit: friction / limit / equality rows are (dof, sign, D, aref) tuples,
//     contact rows keep a 3x9 Jacobian (6 chassis dofs + 3 chain dofs);
//   * fp64 throughout (MuJoCo's mjtNum), state rows of 34/29/29/2 doubles.  In THIS file one thread advances one car:
//     it is the plain statement of the arithmetic (the source the CPU tests compile for the host, and what evaluates the
//     model's compile-time constants); the production mapping, four lanes per car, is mushr_step_quad.cuh and shares these helpers.
// Everything is __host__ __device__ so the same source is unit-tested on the CPU build box (tests/
// compile it with g++) before it ever runs on a B200.
#pragma once
#include <math.h>
#include <stdint.h>
#include "mushr_mesh.h"

#if defined(__CUDACC__)
#define FT_HD __host__ __device__ __forceinline__
#define FT_HDN __host__ __device__
#else
#define FT_HD inline
#define FT_HDN inline
#endif

namespace ftgp {
namespace mushr {

constexpr int NV = 29, NQ = 34;
constexpr int NR = 7;            // root dofs
constexpr int NC = 6;            // chain slots per wheel
constexpr int NP = NR + 4 * NC;  // padded dof space (31): rear wheels carry a dummy steering slot
constexpr int MAXCON = 20;       // 4 wheel-ground + 4 wheel-wall + 4 softener-wall + up to 8 contacts of the car body's geoms
constexpr int NBODY = 11;

constexpr double TIMESTEP = 0.004;       // mushr.em.xml:30
constexpr double GRAV = 9.81;
constexpr double MINVAL = 1e-15;         // mjMINVAL
constexpr double SOLVER_TOL = 1e-8, LS_TOL = 0.01;
constexpr int SOLVER_ITER = 100, LS_ITER = 50;
constexpr double PLANE_Z = 0.01;         // mushr.em.xml:94

// compile-time constants that MuJoCo derives when it compiles the model: filled on the host
// at library start-up by evaluating this same code at qpos0 (see model_constants() in step.cu)
struct ModelConsts {
    double dof_invweight0[NP];   // padded dof space
    double wheel_invweight0[4];  // body_invweight0[wheel body].translational
    double soft_invweight0[4];   // body_invweight0[softener body].translational (option bubble_wrap)
    double chassis_invweight0;   // body_invweight0[car body].translational
    double meaninertia;
    double mass1, ipos1[3], inertia1[9];   // car body = chassis mesh + lidar cylinder
};

// ---- fixed topology helpers -------------------------------------------------------------------------
FT_HD constexpr int chain_q(int w) { return w == 0 ? 8 : w == 1 ? 15 : w == 2 ? 22 : 28; }      // qpos address of the suspension
FT_HD constexpr int chain_d(int w) { return w == 0 ? 7 : w == 1 ? 13 : w == 2 ? 19 : 24; }      // dof address of the suspension
FT_HD constexpr bool front(int w) { return w < 2; }
// padded index p (root 0..6, chain 7 + 6 w + l) -> actual dof, or -1 for the rear dummy steering slot
FT_HD int p2d(int p) {
    if (p < NR) return p;
    int w = (p - NR) / NC, l = (p - NR) % NC;
    if (front(w)) return chain_d(w) + l;
    return l == 0 ? chain_d(w) : (l == 1 ? -1 : chain_d(w) + l - 1);
}
// body positions in the car frame (mushr.em.xml:120,124,137,150,162; wheels scaled by mushr_scale 0.5)
FT_HD double wheel_x(int w) { return front(w) ? 0.5 * 0.1385 : 0.5 * -0.158; }
FT_HD double wheel_y(int w) { return (w & 1) ? 0.5 * -0.115 : 0.5 * 0.115; }
constexpr double WHEEL_Z = 0.5 * 0.0488;
constexpr double SW_X = 0.1385, SW_Z = 0.0488;
constexpr double WHEEL_MASS = 0.498952, SW_MASS = 0.01, SOFT_MASS = 0.00001;      // :69,122,66
constexpr double WS0 = 0.03, WS1 = 0.01, WS2 = 0.03;                              // ellipsoid semi-axes :69

// ---- small vector helpers ---------------------------------------------------------------------------
FT_HD double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
FT_HD void cross3(double* r, const double* a, const double* b) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
FT_HD void quat_norm(double* q) {
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; } else { double i = 1.0 / n; q[0] *= i; q[1] *= i; q[2] *= i; q[3] *= i; }
}
FT_HD void quat2mat(double* m, const double* q) {
    double w = q[0], x = q[1], y = q[2], z = q[3];
    m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
    m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
    m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
FT_HD void quat_mul(double* r, const double* a, const double* b) {
    double t0 = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], t1 = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    double t2 = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], t3 = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3;
}
FT_HD void mat_mul3(double* r, const double* a, const double* b) {
    double t[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
    for (int i = 0; i < 9; i++) r[i] = t[i];
}
FT_HD void mat_vec3(double* r, const double* m, const double* v) {
    double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2], z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
// mju_quatIntegrate
FT_HD void quat_integrate(double* q, const double* vel, double h) {
    double ax = vel[0], ay = vel[1], az = vel[2];
    double n = sqrt(ax * ax + ay * ay + az * az);
    if (n < MINVAL) { ax = 1; ay = az = 0; n = 0; } else { double i = 1.0 / n; ax *= i; ay *= i; az *= i; }
    double ang = h * n, qr[4];
    if (ang == 0) { qr[0] = 1; qr[1] = qr[2] = qr[3] = 0; }
    else { double s = sin(0.5 * ang); qr[0] = cos(0.5 * ang); qr[1] = ax * s; qr[2] = ay * s; qr[3] = az * s; }
    quat_norm(q);
    quat_mul(q, q, qr);
}
// spatial vectors are (angular, linear) about the car's centre of mass, world axes (MuJoCo's "com frame")
FT_HD void cross_motion(double* r, const double* vel, const double* v) {
    double a[3], b[3], c[3];
    cross3(a, vel, v); cross3(b, vel, v + 3); cross3(c, vel + 3, v);
    r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
FT_HD void cross_force(double* r, const double* vel, const double* f) {
    double a[3], b[3], c[3];
    cross3(a, vel, f); cross3(b, vel + 3, f + 3); cross3(c, vel, f + 3);
    r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
// 10-number spatial inertia: Ixx Iyy Izz Ixy Ixz Iyz  m*cx m*cy m*cz  m
FT_HD void inert_mul(double* r, const double* i, const double* v) {
    r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
    r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
    r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
    r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
    r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
    r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
// world inertia R I R^T (I full 3x3, body frame, about the body CoM) shifted to the reference point
FT_HD void inert_com(double* ci, const double* I, const double* R, const double* d, double m) {
    double T[9], W[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) T[3 * r + c] = R[3 * r] * I[c] + R[3 * r + 1] * I[3 + c] + R[3 * r + 2] * I[6 + c];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) W[3 * r + c] = T[3 * r] * R[3 * c] + T[3 * r + 1] * R[3 * c + 1] + T[3 * r + 2] * R[3 * c + 2];
    ci[0] = W[0] + m * (d[1] * d[1] + d[2] * d[2]); ci[1] = W[4] + m * (d[0] * d[0] + d[2] * d[2]); ci[2] = W[8] + m * (d[0] * d[0] + d[1] * d[1]);
    ci[3] = W[1] - m * d[0] * d[1]; ci[4] = W[2] - m * d[0] * d[2]; ci[5] = W[5] - m * d[1] * d[2];
    ci[6] = m * d[0]; ci[7] = m * d[1]; ci[8] = m * d[2]; ci[9] = m;
}
FT_HD void inert_com_diag(double* ci, double ix, double iy, double iz, const double* R, const double* d, double m) {
    double I[9] = {ix, 0, 0, 0, iy, 0, 0, 0, iz};
    inert_com(ci, I, R, d, m);
}

// ---- block-arrow symmetric matrix ---------------------------------------------------------------------
struct Arrow {
    double R[28];          // root block, lower triangle: R[i (i+1)/2 + j], j <= i < 7
    double W[4][21];       // chain blocks, lower triangle 6x6
    double B[4][NC][NR];   // borders: chain slot l (row) x root dof (col)
};
FT_HD int tri(int i, int j) { return i * (i + 1) / 2 + j; }
FT_HD double inv_sqrt(double d) {
#if defined(__CUDA_ARCH__)
    return rsqrt(d);
#else
    return 1.0 / sqrt(d);
#endif
}

// y = A x in padded space
FT_HDN void arrow_mul(const Arrow& A, const double* x, double* y) {
    for (int i = 0; i < NR; i++) {
        double s = 0;
        for (int j = 0; j < NR; j++) s += A.R[i >= j ? tri(i, j) : tri(j, i)] * x[j];
        y[i] = s;
    }
    for (int w = 0; w < 4; w++) {
        const double* xc = x + NR + NC * w; double* yc = y + NR + NC * w;
        for (int l = 0; l < NC; l++) {
            double s = 0;
            for (int k = 0; k < NC; k++) s += A.W[w][l >= k ? tri(l, k) : tri(k, l)] * xc[k];
            for (int j = 0; j < NR; j++) { s += A.B[w][l][j] * x[j]; }
            yc[l] = s;
        }
        for (int j = 0; j < NR; j++) { double s = 0; for (int l = 0; l < NC; l++) s += A.B[w][l][j] * xc[l]; y[j] += s; }
    }
}
// in-place Cholesky: W_w = L_w L_w^T, B_w <- Y_w = L_w^{-1} B_w, R <- chol(R - sum_w Y_w^T Y_w).
// The diagonals of the factors hold 1 / L_jj (every later use divides by them).
FT_HDN void arrow_factor(Arrow& A) {
    for (int w = 0; w < 4; w++) {
        double* L = A.W[w];
        for (int j = 0; j < NC; j++) {
            double d = L[tri(j, j)];
            for (int k = 0; k < j; k++) d -= L[tri(j, k)] * L[tri(j, k)];
            if (d < MINVAL) d = MINVAL;
            const double id = inv_sqrt(d);
            L[tri(j, j)] = id;
            for (int i = j + 1; i < NC; i++) {
                double s = L[tri(i, j)];
                for (int k = 0; k < j; k++) s -= L[tri(i, k)] * L[tri(j, k)];
                L[tri(i, j)] = s * id;
            }
        }
        for (int c = 0; c < NR; c++)
            for (int l = 0; l < NC; l++) {
                double s = A.B[w][l][c];
                for (int k = 0; k < l; k++) s -= L[tri(l, k)] * A.B[w][k][c];
                A.B[w][l][c] = s * L[tri(l, l)];
            }
        for (int i = 0; i < NR; i++)
            for (int j = 0; j <= i; j++) {
                double s = 0;
                for (int l = 0; l < NC; l++) s += A.B[w][l][i] * A.B[w][l][j];
                A.R[tri(i, j)] -= s;
            }
    }
    double* L = A.R;
    for (int j = 0; j < NR; j++) {
        double d = L[tri(j, j)];
        for (int k = 0; k < j; k++) d -= L[tri(j, k)] * L[tri(j, k)];
        if (d < MINVAL) d = MINVAL;
        const double id = inv_sqrt(d);
        L[tri(j, j)] = id;
        for (int i = j + 1; i < NR; i++) {
            double s = L[tri(i, j)];
            for (int k = 0; k < j; k++) s -= L[tri(i, k)] * L[tri(j, k)];
            L[tri(i, j)] = s * id;
        }
    }
}
// x <- A^{-1} x with the factor from arrow_factor
FT_HDN void arrow_solve(const Arrow& A, double* x) {
    for (int w = 0; w < 4; w++) {                       // forward: z_w = L_w^{-1} x_w ; x_r -= Y_w^T z_w
        const double* L = A.W[w]; double* xc = x + NR + NC * w;
        for (int l = 0; l < NC; l++) {
            double s = xc[l];
            for (int k = 0; k < l; k++) s -= L[tri(l, k)] * xc[k];
            xc[l] = s * L[tri(l, l)];
        }
        for (int j = 0; j < NR; j++) { double s = 0; for (int l = 0; l < NC; l++) s += A.B[w][l][j] * xc[l]; x[j] -= s; }
    }
    const double* L = A.R;
    for (int i = 0; i < NR; i++) { double s = x[i]; for (int k = 0; k < i; k++) s -= L[tri(i, k)] * x[k]; x[i] = s * L[tri(i, i)]; }
    for (int i = NR - 1; i >= 0; i--) { double s = x[i]; for (int k = i + 1; k < NR; k++) s -= L[tri(k, i)] * x[k]; x[i] = s * L[tri(i, i)]; }
    for (int w = 0; w < 4; w++) {                       // backward: x_w = L_w^{-T} (z_w - Y_w x_r)
        const double* Lw = A.W[w]; double* xc = x + NR + NC * w;
        for (int l = 0; l < NC; l++) { double s = 0; for (int j = 0; j < NR; j++) s += A.B[w][l][j] * x[j]; xc[l] -= s; }
        for (int l = NC - 1; l >= 0; l--) { double s = xc[l]; for (int k = l + 1; k < NC; k++) s -= Lw[tri(k, l)] * xc[k]; xc[l] = s * Lw[tri(l, l)]; }
    }
}

// Right-hand sides that live on the root dofs only (a car-car contact row touches nothing but the six chassis dofs): the
// forward pass over the chains has nothing to do, so  x_r = S b_r  with S the inverse of the root Schur complement (two
// 7 x 7 triangular solves) and  x_w = -L_w^{-T} (Y_w x_r).
// S6[i][j] = (A^{-1})[j][i] for the root unit vectors i < 6, root rows j < NR
FT_HDN void arrow_root_inverse6(const Arrow& A, double S6[6][NR]) {
    const double* L = A.R;
    for (int c = 0; c < 6; c++) {
        double* x = S6[c];
        for (int i = 0; i < NR; i++) x[i] = i == c ? 1.0 : 0.0;
        for (int i = c; i < NR; i++) { double s = x[i]; for (int k = c; k < i; k++) s -= L[tri(i, k)] * x[k]; x[i] = s * L[tri(i, i)]; }
        for (int i = NR - 1; i >= 0; i--) { double s = x[i]; for (int k = i + 1; k < NR; k++) s -= L[tri(k, i)] * x[k]; x[i] = s * L[tri(i, i)]; }
    }
}
// x <- A^{-1} (b_r, 0): b_r in x[0..NR), chain parts of x are overwritten
FT_HDN void arrow_solve_root_rhs(const Arrow& A, double* x) {
    const double* L = A.R;
    for (int i = 0; i < NR; i++) { double s = x[i]; for (int k = 0; k < i; k++) s -= L[tri(i, k)] * x[k]; x[i] = s * L[tri(i, i)]; }
    for (int i = NR - 1; i >= 0; i--) { double s = x[i]; for (int k = i + 1; k < NR; k++) s -= L[tri(k, i)] * x[k]; x[i] = s * L[tri(i, i)]; }
    for (int w = 0; w < 4; w++) {
        const double* Lw = A.W[w]; double* xc = x + NR + NC * w;
        for (int l = 0; l < NC; l++) { double s = 0; for (int j = 0; j < NR; j++) s += A.B[w][l][j] * x[j]; xc[l] = -s; }
        for (int l = NC - 1; l >= 0; l--) { double s = xc[l]; for (int k = l + 1; k < NC; k++) s -= Lw[tri(k, l)] * xc[k]; xc[l] = s * Lw[tri(l, l)]; }
    }
}

// ---- per-step workspace ---------------------------------------------------------------------------------
struct Contact {
    double J[3][12];    // rows: normal, tangent1, tangent2; cols: chassis dofs 0-5, chain slots 0-5 (susp, steer, throttle, ball)
    double dist, mu, dmin;      // solimp d0 (0.45 with a wheel, else 0.9)
    int wheel;          // 0..3, or -1: contact on the car body itself (chain columns unused)
    int nchain;         // chain columns in use: 3 (wheel body), 6 (softener body behind the ball joint), 0 (car body)
    double tran;        // body_invweight0 translational
};

struct Kin {
    double p1[3], R1[9];
    double com[3];
    double cinert1[10], cinert_sw[10], cinert_w[4][10], cinert_s[4][10];
    double cdof[NP][6];         // padded dof space; rows 0-2 are implicit unit translations but stored for uniformity
    double pw[4][3], Rw[4][9];  // wheel position / orientation
    double ps[4][3], Rs[4][9];  // softener sphere centre / softener body orientation
};

struct Rows {
    // equality rows (2): slot (w = 0,1; l = 1) minus der * root dof 6
    double eq_D[2], eq_aref[2], eq_der[2];
    // friction-loss rows: one per padded dof p >= 6 that has frictionloss (dummy slots have floss = 0 -> skipped)
    double fr_D[NP], fr_aref[NP], fr_Rf[NP], fr_f[NP];
    // limit rows: p = 6, chain slots 0 (suspension) and 1 (front steering); sign = J entry (+1 lower, -1 upper), 0 = inactive
    double lim_D[9], lim_aref[9]; int lim_sign[9];
    // contacts
    int ncon; Contact con[MAXCON];
    double con_D[MAXCON], con_aref[MAXCON];      // D identical for the 4 pyramid rows; aref per contact row below
    double con_arefr[MAXCON][4];
};
FT_HD int lim_p(int k) { return k == 0 ? 6 : (k <= 4 ? NR + NC * (k - 1) : NR + NC * (k - 5) + 1); }   // k: 0 sw, 1-4 susp, 5-6 front steer
constexpr int NLIM = 7;

// ========================================================================================================
// position stage
// ========================================================================================================
FT_HDN void kinematics(const ModelConsts& mc, const double* qpos, Kin& k) {
    double q1[4] = {qpos[3], qpos[4], qpos[5], qpos[6]};
    quat_norm(q1);
    quat2mat(k.R1, q1);
    k.p1[0] = qpos[0]; k.p1[1] = qpos[1]; k.p1[2] = qpos[2];
    const double* R1 = k.R1;
    const double zax[3] = {R1[2], R1[5], R1[8]};
    // steering wheel body (hinge about the car's z axis)
    double p2[3], c2[3] = {SW_X, 0, SW_Z};
    mat_vec3(p2, R1, c2);
    for (int a = 0; a < 3; a++) p2[a] += k.p1[a];
    double R2[9];
    { double c = cos(qpos[7]), s = sin(qpos[7]); double Rz[9] = {c, -s, 0, s, c, 0, 0, 0, 1}; mat_mul3(R2, R1, Rz); }
    // wheels and softener bodies
    double Rs[4][9], ps[4][3];
    const double sc[3] = MUSHR_SOFTENER_CENTER;
    double Rsteer[4][9];
    for (int w = 0; w < 4; w++) {
        const int qa = chain_q(w);
        double c[3] = {wheel_x(w), wheel_y(w), WHEEL_Z + qpos[qa]};          // slide along the car's z (qpos0 = 0)
        mat_vec3(k.pw[w], R1, c);
        for (int a = 0; a < 3; a++) k.pw[w][a] += k.p1[a];
        double thr;
        if (front(w)) {
            double cs = cos(qpos[qa + 1]), sn = sin(qpos[qa + 1]);
            double Rz[9] = {cs, -sn, 0, sn, cs, 0, 0, 0, 1};
            mat_mul3(Rsteer[w], R1, Rz);
            thr = qpos[qa + 2];
        } else { for (int a = 0; a < 9; a++) Rsteer[w][a] = R1[a]; thr = qpos[qa + 1]; }
        { double cs = cos(thr), sn = sin(thr); double Ry[9] = {cs, 0, sn, 0, 1, 0, -sn, 0, cs}; mat_mul3(k.Rw[w], Rsteer[w], Ry); }
        const int qb = front(w) ? qa + 3 : qa + 2;
        double qs[4] = {qpos[qb], qpos[qb + 1], qpos[qb + 2], qpos[qb + 3]}, Rb[9];
        quat_norm(qs); quat2mat(Rb, qs);
        mat_mul3(Rs[w], k.Rw[w], Rb);
        double t[3];
        mat_vec3(t, Rs[w], sc);
        for (int a = 0; a < 3; a++) { ps[w][a] = k.pw[w][a] + t[a]; k.ps[w][a] = ps[w][a]; }
        for (int a = 0; a < 9; a++) k.Rs[w][a] = Rs[w][a];
    }
    // centre of mass of the whole car (subtree_com of the root body)
    double xi1[3];
    mat_vec3(xi1, R1, mc.ipos1);
    for (int a = 0; a < 3; a++) xi1[a] += k.p1[a];
    const double mtot = mc.mass1 + SW_MASS + 4 * (WHEEL_MASS + SOFT_MASS);
    for (int a = 0; a < 3; a++) {
        double s = mc.mass1 * xi1[a] + SW_MASS * p2[a];
        for (int w = 0; w < 4; w++) s += WHEEL_MASS * k.pw[w][a] + SOFT_MASS * ps[w][a];
        k.com[a] = s / mtot;
    }
    // spatial inertias about the com (mju_inertCom)
    double d[3];
    for (int a = 0; a < 3; a++) d[a] = xi1[a] - k.com[a];
    inert_com(k.cinert1, mc.inertia1, R1, d, mc.mass1);
    const double e0 = (WS1 * WS1 + WS2 * WS2) / 5, e1 = (WS0 * WS0 + WS2 * WS2) / 5, e2 = (WS0 * WS0 + WS1 * WS1) / 5;
    for (int a = 0; a < 3; a++) d[a] = p2[a] - k.com[a];
    inert_com_diag(k.cinert_sw, SW_MASS * e0, SW_MASS * e1, SW_MASS * e2, R2, d, SW_MASS);
    const double is = 0.4 * SOFT_MASS * MUSHR_SOFTENER_RADIUS * MUSHR_SOFTENER_RADIUS;
    for (int w = 0; w < 4; w++) {
        for (int a = 0; a < 3; a++) d[a] = k.pw[w][a] - k.com[a];
        inert_com_diag(k.cinert_w[w], WHEEL_MASS * e0, WHEEL_MASS * e1, WHEEL_MASS * e2, k.Rw[w], d, WHEEL_MASS);
        for (int a = 0; a < 3; a++) d[a] = ps[w][a] - k.com[a];
        inert_com_diag(k.cinert_s[w], is, is, is, Rs[w], d, SOFT_MASS);
    }
    // motion axes (mj_comPos): (axis, axis x (com - anchor)) for rotations, (0, axis) for translations
    for (int p = 0; p < NP; p++) for (int a = 0; a < 6; a++) k.cdof[p][a] = 0;
    double off[3];
    for (int c = 0; c < 3; c++) k.cdof[c][3 + c] = 1;
    for (int a = 0; a < 3; a++) off[a] = k.com[a] - k.p1[a];
    for (int c = 0; c < 3; c++) { double ax[3] = {R1[c], R1[3 + c], R1[6 + c]}; for (int a = 0; a < 3; a++) k.cdof[3 + c][a] = ax[a]; cross3(k.cdof[3 + c] + 3, ax, off); }
    for (int a = 0; a < 3; a++) off[a] = k.com[a] - p2[a];
    for (int a = 0; a < 3; a++) k.cdof[6][a] = zax[a];
    cross3(k.cdof[6] + 3, zax, off);
    for (int w = 0; w < 4; w++) {
        double (*cd)[6] = &k.cdof[NR + NC * w];
        for (int a = 0; a < 3; a++) off[a] = k.com[a] - k.pw[w][a];
        for (int a = 0; a < 3; a++) cd[0][3 + a] = zax[a];                                  // suspension slide
        if (front(w)) { for (int a = 0; a < 3; a++) cd[1][a] = zax[a]; cross3(cd[1] + 3, zax, off); }   // steering hinge
        { double ay[3] = {Rsteer[w][1], Rsteer[w][4], Rsteer[w][7]}; for (int a = 0; a < 3; a++) cd[2][a] = ay[a]; cross3(cd[2] + 3, ay, off); }
        for (int c = 0; c < 3; c++) { double ax[3] = {Rs[w][c], Rs[w][3 + c], Rs[w][6 + c]}; for (int a = 0; a < 3; a++) cd[3 + c][a] = ax[a]; cross3(cd[3 + c] + 3, ax, off); }
    }
}

// joint parameters in padded dof space
FT_HD double dof_armature(int p) { if (p < 6) return 0; if (p == 6) return 0.0002; int l = (p - NR) % NC; return l == 0 ? 0.01 : l == 1 ? 0.0002 : l == 2 ? 0.01 : 0.0; }
FT_HD double dof_damping(int p) { if (p < 6) return 0; if (p == 6) return 0.1; int l = (p - NR) % NC; return l == 0 ? 12.5 : l == 1 ? 0.1 : l == 2 ? 0.01 : 0.0; }
FT_HD double dof_floss(int p) { if (p < 6) return 0; if (p == 6) return 0.01; int l = (p - NR) % NC; return l == 0 ? 0.001 : l == 1 ? 0.01 : l == 2 ? 0.001 : 0.25; }
FT_HD bool dof_dummy(int p) { return p >= NR && !front((p - NR) / NC) && (p - NR) % NC == 1; }

// composite-rigid-body mass matrix in block-arrow form (mj_crb)
FT_HDN void mass_matrix(const Kin& k, Arrow& M) {
    double crb1[10], crbw[4][10];
    for (int a = 0; a < 10; a++) crb1[a] = k.cinert1[a] + k.cinert_sw[a];
    for (int w = 0; w < 4; w++) for (int a = 0; a < 10; a++) { crbw[w][a] = k.cinert_w[w][a] + k.cinert_s[w][a]; crb1[a] += crbw[w][a]; }
    double buf[6];
    for (int i = 0; i < 6; i++) {
        inert_mul(buf, crb1, k.cdof[i]);
        for (int j = 0; j <= i; j++) { double s = 0; for (int a = 0; a < 6; a++) s += k.cdof[j][a] * buf[a]; M.R[tri(i, j)] = s; }
    }
    inert_mul(buf, k.cinert_sw, k.cdof[6]);
    for (int j = 0; j <= 6; j++) { double s = 0; for (int a = 0; a < 6; a++) s += k.cdof[j][a] * buf[a]; M.R[tri(6, j)] = s; }
    M.R[tri(6, 6)] += dof_armature(6);
    for (int w = 0; w < 4; w++) {
        const double (*cd)[6] = &k.cdof[NR + NC * w];
        for (int l = 0; l < NC; l++) {
            if (l == 1 && !front(w)) {                   // rear dummy steering slot: unit diagonal, no coupling
                for (int kk = 0; kk < l; kk++) M.W[w][tri(l, kk)] = 0;
                M.W[w][tri(l, l)] = 1;
                for (int j = 0; j < NR; j++) M.B[w][l][j] = 0;
                continue;
            }
            inert_mul(buf, l < 3 ? crbw[w] : k.cinert_s[w], cd[l]);
            for (int kk = 0; kk <= l; kk++) {
                double s = 0;
                if (!(kk == 1 && !front(w))) for (int a = 0; a < 6; a++) s += cd[kk][a] * buf[a];
                M.W[w][tri(l, kk)] = s;
            }
            M.W[w][tri(l, l)] += dof_armature(NR + NC * w + l);
            for (int j = 0; j < 6; j++) { double s = 0; for (int a = 0; a < 6; a++) s += k.cdof[j][a] * buf[a]; M.B[w][l][j] = s; }
            M.B[w][l][6] = 0;
        }
    }
}

// ========================================================================================================
// velocity stage: bias forces by recursive Newton-Euler (mj_comVel + mj_rne), padded dof space
// ========================================================================================================
FT_HDN void bias_forces(const Kin& k, const double* v, double* bias) {
    double cv1[6] = {0, 0, 0, v[0], v[1], v[2]};
    double cacc1[6] = {0, 0, 0, 0, 0, GRAV};
    double dd[6];
    // free joint: cdof_dot of the three rotational dofs uses the velocity after the translations only
    double cvr[6] = {cv1[0], cv1[1], cv1[2], cv1[3], cv1[4], cv1[5]};
    for (int c = 0; c < 3; c++) {
        cross_motion(dd, cv1, k.cdof[3 + c]);
        for (int a = 0; a < 6; a++) { cacc1[a] += dd[a] * v[3 + c]; cvr[a] += k.cdof[3 + c][a] * v[3 + c]; }
    }
    for (int a = 0; a < 6; a++) cv1[a] = cvr[a];
    double t[6], t2[6], cfrc1[6], f[6];
    inert_mul(cfrc1, k.cinert1, cacc1);
    inert_mul(t, k.cinert1, cv1); cross_force(t2, cv1, t);
    for (int a = 0; a < 6; a++) cfrc1[a] += t2[a];
    // steering wheel
    {
        double cv[6], ca[6];
        cross_motion(dd, cv1, k.cdof[6]);
        for (int a = 0; a < 6; a++) { ca[a] = cacc1[a] + dd[a] * v[6]; cv[a] = cv1[a] + k.cdof[6][a] * v[6]; }
        inert_mul(f, k.cinert_sw, ca);
        inert_mul(t, k.cinert_sw, cv); cross_force(t2, cv, t);
        for (int a = 0; a < 6; a++) f[a] += t2[a];
        double s = 0; for (int a = 0; a < 6; a++) s += k.cdof[6][a] * f[a];
        bias[6] = s;
        for (int a = 0; a < 6; a++) cfrc1[a] += f[a];
    }
    for (int w = 0; w < 4; w++) {
        const double (*cd)[6] = &k.cdof[NR + NC * w];
        const double* vc = v + NR + NC * w; double* bc = bias + NR + NC * w;
        double cv[6], ca[6];
        for (int a = 0; a < 6; a++) { cv[a] = cv1[a]; ca[a] = cacc1[a]; }
        for (int l = 0; l < 3; l++) {
            if (l == 1 && !front(w)) continue;
            cross_motion(dd, cv, cd[l]);
            for (int a = 0; a < 6; a++) { ca[a] += dd[a] * vc[l]; cv[a] += cd[l][a] * vc[l]; }
        }
        double fw[6];
        inert_mul(fw, k.cinert_w[w], ca);
        inert_mul(t, k.cinert_w[w], cv); cross_force(t2, cv, t);
        for (int a = 0; a < 6; a++) fw[a] += t2[a];
        // softener body behind the ball joint
        double cvs[6], cas[6];
        for (int a = 0; a < 6; a++) { cvs[a] = cv[a]; cas[a] = ca[a]; }
        for (int c = 0; c < 3; c++) {
            cross_motion(dd, cv, cd[3 + c]);
            for (int a = 0; a < 6; a++) { cas[a] += dd[a] * vc[3 + c]; cvs[a] += cd[3 + c][a] * vc[3 + c]; }
        }
        double fs[6];
        inert_mul(fs, k.cinert_s[w], cas);
        inert_mul(t, k.cinert_s[w], cvs); cross_force(t2, cvs, t);
        for (int a = 0; a < 6; a++) { fs[a] += t2[a]; fw[a] += fs[a]; cfrc1[a] += fw[a]; }
        for (int l = 0; l < NC; l++) {
            double s = 0;
            const double* ff = l < 3 ? fw : fs;
            for (int a = 0; a < 6; a++) s += cd[l][a] * ff[a];
            bc[l] = (l == 1 && !front(w)) ? 0.0 : s;
        }
    }
    for (int i = 0; i < 6; i++) { double s = 0; for (int a = 0; a < 6; a++) s += k.cdof[i][a] * cfrc1[a]; bias[i] = s; }
}

// ========================================================================================================
// constraints
// ========================================================================================================
// impedance d(x) for |pos - margin| (getimpedance)
FT_HD double impedance(double dmin, double dmax, double width, double mid, double power, double pos) {
    double x = fabs(pos / width);
    if (x >= 1) return dmax;
    if (x == 0) return dmin;
    double y;                                           // power == 2 for every solimp of this model
    if (power == 2.0) y = x <= mid ? x * x / mid : 1 - (1 - x) * (1 - x) / (1 - mid);
    else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
    else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
    return dmin + y * (dmax - dmin);
}
// default solref (0.02, 1) with refsafe: timeconst = max(0.02, 2 h) = 0.02; all solimp here have dmax 0.95,
// width 0.001, midpoint 0.5, power 2
FT_HD void kbi(double dmin, double pos, double diag, double& K, double& B, double& imp, double& R) {
    const double dmax = 0.95, tc = 0.02, dr = 1.0;
    imp = impedance(dmin, dmax, 0.001, 0.5, 2.0, pos);
    R = (1 - imp) * diag / imp; if (R < MINVAL) R = MINVAL;
    K = 1 / (dmax * dmax * tc * tc * dr * dr); B = 2 / (dmax * tc);
}

// wheel ellipsoid vs ground plane (mjc_PlaneConvex + ellipsoid support point), contact Jacobian rows
FT_HDN void wheel_contacts(const ModelConsts& mc, const Kin& k, Rows& r) {
    for (int w = 0; w < 4; w++) {
        const double* R = k.Rw[w];
        double dl[3] = {-R[6], -R[7], -R[8]};                      // -n (n = +z) in the wheel frame
        double s[3] = {WS0 * dl[0], WS1 * dl[1], WS2 * dl[2]};
        double nn = sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
        s[0] = WS0 * s[0] / nn; s[1] = WS1 * s[1] / nn; s[2] = WS2 * s[2] / nn;
        double sw[3];
        mat_vec3(sw, R, s);
        for (int a = 0; a < 3; a++) sw[a] += k.pw[w][a];
        double dist = sw[2] - PLANE_Z;
        if (dist > 0 || r.ncon >= MAXCON) continue;
        Contact& c = r.con[r.ncon++];
        c.dist = dist; c.mu = 0.5; c.dmin = 0.45; c.wheel = w; c.nchain = 3; c.tran = mc.wheel_invweight0[w];
        double off[3] = {sw[0] - k.com[0], sw[1] - k.com[1], sw[2] - 0.5 * dist - k.com[2]};   // contact point - com
        // frame: n = (0,0,1), t1 = (0,1,0), t2 = (-1,0,0)  (mju_makeFrame)
        for (int col = 0; col < 9; col++) {
            int p = col < 6 ? col : NR + NC * w + (col - 6);
            double jp[3] = {0, 0, 0};
            if (!(col == 7 && !front(w))) {
                cross3(jp, k.cdof[p], off);
                for (int a = 0; a < 3; a++) jp[a] += k.cdof[p][3 + a];
            }
            c.J[0][col] = jp[2]; c.J[1][col] = jp[1]; c.J[2][col] = -jp[0];
        }
    }
}

FT_HD double poly_val(int w, double x) { const double s = w == 0 ? 1.0 : -1.0; return x * (1 + x * (s * 0.375 + x * (0.140625 + x * (-s * 0.0722656)))); }
FT_HD double poly_der(int w, double x) { const double s = w == 0 ? 1.0 : -1.0; return 1 + x * (2 * s * 0.375 + x * (3 * 0.140625 + x * (-4 * s * 0.0722656))); }

// qp / v in padded space.  Fills D, aref for every row.
FT_HDN void make_rows(const ModelConsts& mc, const double* qpos, const double* v, Rows& r) {
    double K, B, imp, R;
    // joint equalities (mushr.em.xml:185-186)
    for (int w = 0; w < 2; w++) {
        const double x = qpos[7], q1 = qpos[chain_q(w) + 1];
        const double pos = q1 - poly_val(w, x), der = poly_der(w, x);
        const int p1 = NR + NC * w + 1;
        kbi(0.9, pos, mc.dof_invweight0[p1] + mc.dof_invweight0[6], K, B, imp, R);
        r.eq_der[w] = der; r.eq_D[w] = 1 / R;
        r.eq_aref[w] = -B * (v[p1] - der * v[6]) - K * imp * pos;
    }
    // friction loss: pos = 0 -> imp = d0 = 0.9, K = 0
    for (int p = 6; p < NP; p++) {
        const double f = dof_dummy(p) ? 0.0 : dof_floss(p);
        r.fr_f[p] = f;
        if (f <= 0) { r.fr_D[p] = 0; r.fr_aref[p] = 0; r.fr_Rf[p] = 0; continue; }
        kbi(0.9, 0.0, mc.dof_invweight0[p], K, B, imp, R);
        r.fr_D[p] = 1 / R; r.fr_Rf[p] = R * f; r.fr_aref[p] = -B * v[p];
    }
    // joint limits (margin 0): steering wheel +-1, suspensions [-0.03, 0], front steering +-1
    for (int kk = 0; kk < NLIM; kk++) {
        double q, lo, hi;
        if (kk == 0) { q = qpos[7]; lo = -1; hi = 1; }
        else if (kk <= 4) { q = qpos[chain_q(kk - 1)]; lo = -0.03; hi = 0; }
        else { q = qpos[chain_q(kk - 5) + 1]; lo = -1; hi = 1; }
        const int p = lim_p(kk);
        r.lim_sign[kk] = 0; r.lim_D[kk] = 0; r.lim_aref[kk] = 0;
        double dist; int sign;
        if (q - lo < 0) { dist = q - lo; sign = 1; } else if (hi - q < 0) { dist = hi - q; sign = -1; } else continue;
        kbi(0.9, dist, mc.dof_invweight0[p], K, B, imp, R);
        r.lim_sign[kk] = sign; r.lim_D[kk] = 1 / R; r.lim_aref[kk] = -B * (sign * v[p]) - K * imp * dist;
    }
    // pyramidal contacts: rows Jn +- mu Jt1, Jn +- mu Jt2; R = 2 mu^2 R_normal for all four
    for (int c = 0; c < r.ncon; c++) {
        const Contact& ct = r.con[c];
        kbi(ct.dmin, ct.dist, ct.tran, K, B, imp, R);
        double Rpy = 2 * ct.mu * ct.mu * R; if (Rpy < MINVAL) Rpy = MINVAL;
        r.con_D[c] = 1 / Rpy;
        double vel[3];
        for (int a = 0; a < 3; a++) {
            double s = 0;
            for (int col = 0; col < 6; col++) s += ct.J[a][col] * v[col];
            for (int col = 0; col < ct.nchain; col++) s += ct.J[a][6 + col] * v[NR + NC * ct.wheel + col];
            vel[a] = s;
        }
        for (int rr = 0; rr < 4; rr++) {
            double sg = (rr & 1) ? -1.0 : 1.0;
            double vr = vel[0] + sg * ct.mu * vel[1 + (rr >> 1)];
            r.con_arefr[c][rr] = -B * vr - K * imp * ct.dist;
        }
    }
}

// ========================================================================================================
// Newton solver on the block-arrow Hessian
// ========================================================================================================
// contact row value J_row . x
FT_HD void contact_dots(const Contact& ct, const double* x, double* d3) {
    for (int a = 0; a < 3; a++) {
        double s = 0;
        for (int col = 0; col < 6; col++) s += ct.J[a][col] * x[col];
        if (ct.nchain > 0) { const double* xc = x + NR + NC * ct.wheel; for (int col = 0; col < ct.nchain; col++) s += ct.J[a][6 + col] * xc[col]; }
        d3[a] = s;
    }
}

struct Solver {
    double qacc[NP], Ma[NP], grad[NP], search[NP], Mv[NP];
    double cost, gauss;
};

// cost of the constraint rows at acceleration x, optionally accumulating qfrc_constraint = J^T force and H += J^T D J (active rows)
template <bool WITH_FORCE, bool WITH_H>
FT_HDN double rows_cost(const Rows& r, const double* x, double* qfrc, Arrow* H) {
    double cost = 0;
    for (int w = 0; w < 2; w++) {                                        // equality: always quadratic
        const int p1 = NR + NC * w + 1;
        const double jar = x[p1] - r.eq_der[w] * x[6] - r.eq_aref[w], D = r.eq_D[w];
        cost += 0.5 * D * jar * jar;
        if (WITH_FORCE) { const double f = -D * jar; qfrc[p1] += f; qfrc[6] -= r.eq_der[w] * f; }
        if (WITH_H) { H->W[w][tri(1, 1)] += D; H->B[w][1][6] -= D * r.eq_der[w]; H->R[tri(6, 6)] += D * r.eq_der[w] * r.eq_der[w]; }
    }
    for (int p = 6; p < NP; p++) {                                       // friction loss: quadratic inside +-R f, linear outside
        const double f = r.fr_f[p];
        if (f <= 0) continue;
        const double jar = x[p] - r.fr_aref[p], Rf = r.fr_Rf[p], D = r.fr_D[p];
        if (jar <= -Rf) { cost += -0.5 * Rf * f - f * jar; if (WITH_FORCE) qfrc[p] += f; }
        else if (jar >= Rf) { cost += -0.5 * Rf * f + f * jar; if (WITH_FORCE) qfrc[p] -= f; }
        else {
            cost += 0.5 * D * jar * jar;
            if (WITH_FORCE) qfrc[p] += -D * jar;
            if (WITH_H) { if (p == 6) H->R[tri(6, 6)] += D; else { int w = (p - NR) / NC, l = (p - NR) % NC; H->W[w][tri(l, l)] += D; } }
        }
    }
    for (int kk = 0; kk < NLIM; kk++) {                                  // limits: active when jar < 0
        const int sg = r.lim_sign[kk];
        if (!sg) continue;
        const int p = lim_p(kk);
        const double jar = sg * x[p] - r.lim_aref[kk], D = r.lim_D[kk];
        if (jar < 0) {
            cost += 0.5 * D * jar * jar;
            if (WITH_FORCE) qfrc[p] += sg * (-D * jar);
            if (WITH_H) { if (p == 6) H->R[tri(6, 6)] += D; else { int w = (p - NR) / NC, l = (p - NR) % NC; H->W[w][tri(l, l)] += D; } }
        }
    }
    for (int c = 0; c < r.ncon; c++) {                                   // pyramidal contact rows: active when jar < 0
        const Contact& ct = r.con[c];
        double d3[3];
        contact_dots(ct, x, d3);
        const double D = r.con_D[c];
        for (int rr = 0; rr < 4; rr++) {
            const double sg = (rr & 1) ? -1.0 : 1.0; const int ta = 1 + (rr >> 1);
            const double jar = d3[0] + sg * ct.mu * d3[ta] - r.con_arefr[c][rr];
            if (jar >= 0) continue;
            cost += 0.5 * D * jar * jar;
            if (WITH_FORCE || WITH_H) {
                double Jr[12];
                for (int col = 0; col < 6 + ct.nchain; col++) Jr[col] = ct.J[0][col] + sg * ct.mu * ct.J[ta][col];
                if (WITH_FORCE) {
                    const double f = -D * jar;
                    for (int col = 0; col < 6; col++) qfrc[col] += Jr[col] * f;
                    for (int col = 0; col < ct.nchain; col++) qfrc[NR + NC * ct.wheel + col] += Jr[6 + col] * f;
                }
                if (WITH_H) {
                    for (int i = 0; i < 6; i++) for (int j = 0; j <= i; j++) H->R[tri(i, j)] += D * Jr[i] * Jr[j];
                    if (ct.nchain > 0) {
                        const int w = ct.wheel;
                        for (int l = 0; l < ct.nchain; l++) {
                            for (int kk = 0; kk <= l; kk++) H->W[w][tri(l, kk)] += D * Jr[6 + l] * Jr[6 + kk];
                            for (int j = 0; j < 6; j++) H->B[w][l][j] += D * Jr[6 + l] * Jr[j];
                        }
                    }
                }
            }
        }
    }
    return cost;
}

// line-search evaluation: total cost and its first two derivatives at qacc + alpha * search
struct LsPoint { double alpha, cost, d0, d1; };
struct LsCtx { const Rows* r; const double* x; const double* s; double qg0, qg1, qg2; double cdx[MAXCON][3], cds[MAXCON][3]; };

// the car's share of the cost along the line, as the coefficients of a quadratic in alpha (added to q0, q1, q2)
FT_HDN void ls_quad(const LsCtx& c, double alpha, double& q0, double& q1, double& q2) {
    const Rows& r = *c.r; const double* x = c.x; const double* s = c.s;
    q0 += c.qg0; q1 += c.qg1; q2 += c.qg2;
    for (int w = 0; w < 2; w++) {
        const int p1 = NR + NC * w + 1;
        const double jar = x[p1] - r.eq_der[w] * x[6] - r.eq_aref[w], jv = s[p1] - r.eq_der[w] * s[6], D = r.eq_D[w];
        q0 += 0.5 * D * jar * jar; q1 += D * jar * jv; q2 += 0.5 * D * jv * jv;
    }
    for (int p = 6; p < NP; p++) {
        const double f = r.fr_f[p];
        if (f <= 0) continue;
        const double jar = x[p] - r.fr_aref[p], jv = s[p], Rf = r.fr_Rf[p], D = r.fr_D[p];
        const double xx = jar + alpha * jv;
        if (xx <= -Rf) { q0 += f * (-0.5 * Rf - jar); q1 += -f * jv; }
        else if (xx >= Rf) { q0 += f * (-0.5 * Rf + jar); q1 += f * jv; }
        else { q0 += 0.5 * D * jar * jar; q1 += D * jar * jv; q2 += 0.5 * D * jv * jv; }
    }
    for (int kk = 0; kk < NLIM; kk++) {
        const int sg = r.lim_sign[kk];
        if (!sg) continue;
        const int p = lim_p(kk);
        const double jar = sg * x[p] - r.lim_aref[kk], jv = sg * s[p], D = r.lim_D[kk];
        if (jar + alpha * jv < 0) { q0 += 0.5 * D * jar * jar; q1 += D * jar * jv; q2 += 0.5 * D * jv * jv; }
    }
    for (int cc = 0; cc < r.ncon; cc++) {
        const Contact& ct = r.con[cc];
        const double* dx = c.cdx[cc]; const double* ds = c.cds[cc];
        const double D = r.con_D[cc];
        for (int rr = 0; rr < 4; rr++) {
            const double sg = (rr & 1) ? -1.0 : 1.0; const int ta = 1 + (rr >> 1);
            const double jar = dx[0] + sg * ct.mu * dx[ta] - r.con_arefr[cc][rr], jv = ds[0] + sg * ct.mu * ds[ta];
            if (jar + alpha * jv < 0) { q0 += 0.5 * D * jar * jar; q1 += D * jar * jv; q2 += 0.5 * D * jv * jv; }
        }
    }
}
FT_HD void ls_point(LsPoint& pt, double alpha, double q0, double q1, double q2) {
    pt.alpha = alpha; pt.cost = alpha * alpha * q2 + alpha * q1 + q0;
    pt.d0 = 2 * alpha * q2 + q1; pt.d1 = 2 * q2;
    if (pt.d1 <= 0) pt.d1 = MINVAL;
}
FT_HDN void ls_eval(const LsCtx& c, LsPoint& pt, double alpha) {
    double q0 = 0, q1 = 0, q2 = 0;
    ls_quad(c, alpha, q0, q1, q2);
    ls_point(pt, alpha, q0, q1, q2);
}

// exact line search (PrimalSearch): Newton on the derivative, then bracketing
FT_HDN double line_search(const Rows& r, const Arrow& M, Solver& s, const double* qfrc_smooth, double scale) {
    double snorm = 0;
    for (int p = 0; p < NP; p++) snorm += s.search[p] * s.search[p];
    snorm = sqrt(snorm);
    if (snorm < MINVAL) return 0;
    arrow_mul(M, s.search, s.Mv);
    LsCtx c; c.r = &r; c.x = s.qacc; c.s = s.search;
    c.qg0 = s.gauss; c.qg1 = 0; c.qg2 = 0;
    for (int p = 0; p < NP; p++) { c.qg1 += s.search[p] * (s.Ma[p] - qfrc_smooth[p]); c.qg2 += 0.5 * s.search[p] * s.Mv[p]; }
    for (int cc = 0; cc < r.ncon; cc++) { contact_dots(r.con[cc], s.qacc, c.cdx[cc]); contact_dots(r.con[cc], s.search, c.cds[cc]); }
    const double gtol = SOLVER_TOL * LS_TOL * snorm / scale;
    LsPoint p0, p1, p2, pm, a1, a2;
    int it = 0;
    ls_eval(c, p0, 0);
    ls_eval(c, p1, p0.alpha - p0.d0 / p0.d1);
    if (p0.cost < p1.cost) p1 = p0;
    if (fabs(p1.d0) < gtol) return p1.alpha;
    const double dir = p1.d0 < 0 ? 1.0 : -1.0;
    bool p2update = false;
    p2 = p1;
    while (p1.d0 * dir <= -gtol && it < LS_ITER) {
        p2 = p1; p2update = true;
        ls_eval(c, p1, p1.alpha - p1.d0 / p1.d1); it++;
        if (fabs(p1.d0) < gtol) return p1.alpha;
    }
    if (it >= LS_ITER || !p2update) return p1.alpha;
    while (it < LS_ITER) {
        ls_eval(c, pm, 0.5 * (p1.alpha + p2.alpha)); it++;
        ls_eval(c, a1, p1.alpha - p1.d0 / p1.d1);
        ls_eval(c, a2, p2.alpha - p2.d0 / p2.d1);
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
}

// gradient + Newton direction at the current point: grad = Ma - qfrc_smooth - J^T f ; search = -H^{-1} grad
// cost, constraint force J^T f and gradient at the current point
FT_HDN void newton_evaluate(const Rows& r, Solver& s, const double* qfrc_smooth, const double* qacc_smooth, double* qfrc_con) {
    for (int p = 0; p < NP; p++) qfrc_con[p] = 0;
    double cc = rows_cost<true, false>(r, s.qacc, qfrc_con, nullptr);
    double g = 0;
    for (int p = 0; p < NP; p++) g += (s.Ma[p] - qfrc_smooth[p]) * (s.qacc[p] - qacc_smooth[p]);
    s.gauss = 0.5 * g; s.cost = cc + s.gauss;
    for (int p = 0; p < NP; p++) s.grad[p] = s.Ma[p] - qfrc_smooth[p] - qfrc_con[p];
}
// Newton direction: search = -H^{-1} grad with H = M + J^T D J over the quadratic rows at the current point
FT_HDN void newton_direction(const Rows& r, const Arrow& M, Solver& s, Arrow& H) {
    H = M;
    rows_cost<false, true>(r, s.qacc, nullptr, &H);
    for (int p = 0; p < NP; p++) s.search[p] = s.grad[p];
    arrow_factor(H);
    arrow_solve(H, s.search);
    for (int p = 0; p < NP; p++) s.search[p] = -s.search[p];
}

// ========================================================================================================
// the step
// ========================================================================================================
struct StepInfo { int iters, ncon_wheel, ncon_wall, reset, ncon_ground, near_wall; };   // ncon_wall: wheels + chassis + lidar cylinder against walls

FT_HD bool bad_value(double x) { return !(x <= 1e10 && x >= -1e10); }

FT_HDN void reset_state(double* qpos, double* qvel, double* warm) {
    for (int i = 0; i < NQ; i++) qpos[i] = 0;
    qpos[1] = 2.0; qpos[3] = 1.0; qpos[11] = 1.0; qpos[18] = 1.0; qpos[24] = 1.0; qpos[30] = 1.0;
    for (int i = 0; i < NV; i++) { qvel[i] = 0; warm[i] = 0; }
}

// optional hook: extra chassis contacts (walls) are appended by the caller through this functor type
struct NoWalls { FT_HD void operator()(const ModelConsts&, const Kin&, Rows&) const {} };

template <class WallFn>
FT_HDN void step_car(const ModelConsts& mc, double* qpos, double* qvel, double* warm, const double* ctrl,
                     const WallFn& walls, StepInfo& info) {
    info.reset = 0;
    for (int i = 0; i < NQ; i++) if (bad_value(qpos[i])) info.reset = 1;
    for (int i = 0; i < NV; i++) if (bad_value(qvel[i])) info.reset = 1;
    if (info.reset) reset_state(qpos, qvel, warm);                     // mj_checkPos / mj_checkVel
    // ---- padded velocity / warm start
    double v[NP], wa[NP];
    for (int p = 0; p < NP; p++) { int d = p2d(p); v[p] = d >= 0 ? qvel[d] : 0.0; wa[p] = d >= 0 ? warm[d] : 0.0; }
    // ---- position stage
    // the position-stage scratch (Kin) is dead before the solver state (H, Solver) is first written: share the storage
    struct SolverState { Arrow H; Solver s; };
    constexpr size_t SCRATCH = sizeof(Kin) > sizeof(SolverState) ? sizeof(Kin) : sizeof(SolverState);
    double scratch[SCRATCH / sizeof(double)];
    Kin& k = *reinterpret_cast<Kin*>(scratch);
    kinematics(mc, qpos, k);
    Arrow M;
    mass_matrix(k, M);
    Rows r;
    r.ncon = 0;
    wheel_contacts(mc, k, r);
    info.ncon_wheel = r.ncon;
    walls(mc, k, r);
    info.ncon_wall = r.ncon - info.ncon_wheel;
    make_rows(mc, qpos, v, r);
    // ---- smooth forces: passive + actuation - bias
    double qfrc_smooth[NP], qacc_smooth[NP];
    bias_forces(k, v, qfrc_smooth);
    for (int p = 0; p < NP; p++) qfrc_smooth[p] = -qfrc_smooth[p] - dof_damping(p) * v[p];
    for (int w = 0; w < 4; w++) qfrc_smooth[NR + NC * w] += -500.0 * (qpos[chain_q(w)] - (-0.015));     // suspension spring :63
    for (int w = 2; w < 4; w++) qfrc_smooth[NR + NC * w + 1] = 0;                                         // dummy slots
    {
        // turn = <position kp=20> on the steering-wheel hinge (:179); forward = <velocity kv=100 gear=0.04
        // forcerange=+-500> on the 0.25-weighted throttle tendon (:180,191-196)
        qfrc_smooth[6] += 20.0 * ctrl[1] - 20.0 * qpos[7];
        double tv = 0;
        for (int w = 0; w < 4; w++) tv += 0.25 * v[NR + NC * w + 2];
        double f = 100.0 * ctrl[0] - 100.0 * (0.04 * tv);
        f = f > 500.0 ? 500.0 : (f < -500.0 ? -500.0 : f);
        for (int w = 0; w < 4; w++) qfrc_smooth[NR + NC * w + 2] += 0.04 * 0.25 * f;
    }
    SolverState& so = *reinterpret_cast<SolverState*>(scratch);
    Arrow& H = so.H;
    H = M;
    arrow_factor(H);
    for (int p = 0; p < NP; p++) qacc_smooth[p] = qfrc_smooth[p];
    arrow_solve(H, qacc_smooth);
    // ---- Newton: warm start if its cost beats qacc_smooth's (mj_fwdConstraint)
    Solver& s = so.s;
    double qfrc_con[NP];
    {
        arrow_mul(M, wa, s.Ma);
        double cw = rows_cost<false, false>(r, wa, nullptr, nullptr);
        for (int p = 0; p < NP; p++) cw += 0.5 * (s.Ma[p] - qfrc_smooth[p]) * (wa[p] - qacc_smooth[p]);
        double cs = rows_cost<false, false>(r, qacc_smooth, nullptr, nullptr);
        if (cw > cs) { for (int p = 0; p < NP; p++) { s.qacc[p] = qacc_smooth[p]; s.Ma[p] = qfrc_smooth[p]; } }
        else for (int p = 0; p < NP; p++) s.qacc[p] = wa[p];
    }
    const double scale = 1.0 / (mc.meaninertia * NV);
    newton_evaluate(r, s, qfrc_smooth, qacc_smooth, qfrc_con);
    newton_direction(r, M, s, H);
    int iter = 0;
    while (iter < SOLVER_ITER) {
        const double alpha = line_search(r, M, s, qfrc_smooth, scale);
        if (alpha == 0) break;
        for (int p = 0; p < NP; p++) { s.qacc[p] += alpha * s.search[p]; s.Ma[p] += alpha * s.Mv[p]; }
        const double oldcost = s.cost;
        newton_evaluate(r, s, qfrc_smooth, qacc_smooth, qfrc_con);
        double gn = 0;
        for (int p = 0; p < NP; p++) gn += s.grad[p] * s.grad[p];
        iter++;
        // MuJoCo factorises H before this test; the direction is unused when the test ends the loop, so the
        // factorisation is skipped then (same qacc, one block-arrow Cholesky less per step)
        if (scale * (oldcost - s.cost) < SOLVER_TOL || scale * sqrt(gn) < SOLVER_TOL) break;
        newton_direction(r, M, s, H);
    }
    info.iters = iter;
    bool badacc = false;
    for (int p = 0; p < NP; p++) if (bad_value(s.qacc[p])) badacc = true;
    if (badacc) { reset_state(qpos, qvel, warm); info.reset = 1; return; }      // mj_checkAcc
    // ---- mj_Euler with implicit joint damping: (M + h diag(b)) qacc' = qfrc_smooth + qfrc_constraint
    H = M;
    H.R[tri(6, 6)] += TIMESTEP * dof_damping(6);
    for (int w = 0; w < 4; w++) for (int l = 0; l < 3; l++) if (!(l == 1 && !front(w))) H.W[w][tri(l, l)] += TIMESTEP * dof_damping(NR + NC * w + l);
    arrow_factor(H);
    double qa[NP];
    for (int p = 0; p < NP; p++) qa[p] = qfrc_smooth[p] + qfrc_con[p];
    arrow_solve(H, qa);
    for (int p = 0; p < NP; p++) { int d = p2d(p); if (d >= 0) { warm[d] = s.qacc[p]; qvel[d] += TIMESTEP * qa[p]; } }
    // ---- mj_integratePos with the new velocity
    for (int a = 0; a < 3; a++) qpos[a] += TIMESTEP * qvel[a];
    quat_integrate(qpos + 3, qvel + 3, TIMESTEP);
    qpos[7] += TIMESTEP * qvel[6];
    for (int w = 0; w < 4; w++) {
        const int qa0 = chain_q(w), d0 = chain_d(w);
        const int nh = front(w) ? 3 : 2;
        for (int l = 0; l < nh; l++) qpos[qa0 + l] += TIMESTEP * qvel[d0 + l];
        quat_integrate(qpos + qa0 + nh, qvel + d0 + nh, TIMESTEP);
    }
}

}  // namespace mushr
}  // namespace ftgp
