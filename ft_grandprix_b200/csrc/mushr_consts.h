// mushr_consts.h -- host-side evaluation of the constants MuJoCo derives when it compiles the model
// (dof_invweight0, body_invweight0, stat.meaninertia: SURVEY.md B.7), by running the same kinematics /
// mass-matrix code as the kernel at qpos0.  Host only; the result is uploaded to __constant__ memory.
#pragma once
#include "mushr_step.cuh"

namespace ftgp {
namespace mushr {

inline ModelConsts model_constants() {
    ModelConsts mc{};
    // car body = chassis mesh (explicit mass, mushr.em.xml:119) + lidar cylinder (density 1000, :108)
    {
        const double cm = MUSHR_CHASSIS_MASS, cc[3] = MUSHR_CHASSIS_COM, cI[9] = MUSHR_CHASSIS_INERTIA;
        const double lr = 0.030, lh = 0.015;
        const double lm = 1000.0 * 3.14159265358979323846 * lr * lr * 2 * lh;
        const double lc[3] = {-0.0525, 0.0, 0.065 - lh / 2};
        const double lxx = lm * (3 * lr * lr + 4 * lh * lh) / 12, lzz = lm * lr * lr / 2;
        mc.mass1 = cm + lm;
        for (int a = 0; a < 3; a++) mc.ipos1[a] = (cm * cc[a] + lm * lc[a]) / mc.mass1;
        for (int a = 0; a < 9; a++) mc.inertia1[a] = cI[a];
        mc.inertia1[0] += lxx; mc.inertia1[4] += lxx; mc.inertia1[8] += lzz;
        for (int g = 0; g < 2; g++) {
            const double m = g ? lm : cm; const double* c = g ? lc : cc;
            double d[3] = {c[0] - mc.ipos1[0], c[1] - mc.ipos1[1], c[2] - mc.ipos1[2]};
            double d2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) mc.inertia1[3 * a + b] += m * ((a == b ? d2 : 0.0) - d[a] * d[b]);
        }
    }
    double qpos0[NQ] = {0};
    qpos0[1] = 2.0; qpos0[3] = 1.0; qpos0[11] = qpos0[18] = qpos0[24] = qpos0[30] = 1.0;
    Kin k;
    kinematics(mc, qpos0, k);
    Arrow M;
    mass_matrix(k, M);
    double tr = 0;
    tr += M.R[tri(0, 0)] + M.R[tri(1, 1)] + M.R[tri(2, 2)] + M.R[tri(3, 3)] + M.R[tri(4, 4)] + M.R[tri(5, 5)] + M.R[tri(6, 6)];
    for (int w = 0; w < 4; w++) for (int l = 0; l < NC; l++) if (!(l == 1 && !front(w))) tr += M.W[w][tri(l, l)];
    mc.meaninertia = tr / NV;
    Arrow L = M;
    arrow_factor(L);
    static double Minv[NP][NP];
    for (int i = 0; i < NP; i++) {
        double e[NP] = {0}; e[i] = 1;
        arrow_solve(L, e);
        for (int j = 0; j < NP; j++) Minv[j][i] = e[j];
    }
    for (int p = 0; p < NP; p++) mc.dof_invweight0[p] = Minv[p][p];
    auto avg3 = [&](int p) { double a = (Minv[p][p] + Minv[p + 1][p + 1] + Minv[p + 2][p + 2]) / 3; mc.dof_invweight0[p] = mc.dof_invweight0[p + 1] = mc.dof_invweight0[p + 2] = a; };
    avg3(0); avg3(3);
    for (int w = 0; w < 4; w++) avg3(NR + NC * w + 3);
    // body_invweight0 (translational): trace(Jp Minv Jp^T) / 3 at the body's CoM
    auto tran_at = [&](const double* point, int w, bool soft) {
        double J[3][NP] = {{0}};
        double off[3] = {point[0] - k.com[0], point[1] - k.com[1], point[2] - k.com[2]};
        for (int col = 0; col < (w >= 0 ? (soft ? 12 : 9) : 6); col++) {
            int p = col < 6 ? col : NR + NC * w + (col - 6);
            if (col == 7 && !front(w)) continue;
            double jp[3];
            cross3(jp, k.cdof[p], off);
            for (int a = 0; a < 3; a++) J[a][p] = jp[a] + k.cdof[p][3 + a];
        }
        double t = 0;
        for (int a = 0; a < 3; a++) for (int i = 0; i < NP; i++) if (J[a][i] != 0) for (int j = 0; j < NP; j++) t += J[a][i] * Minv[i][j] * J[a][j];
        return t / 3 < MINVAL ? MINVAL : t / 3;
    };
    double xi1[3];
    mat_vec3(xi1, k.R1, mc.ipos1);
    for (int a = 0; a < 3; a++) xi1[a] += k.p1[a];
    mc.chassis_invweight0 = tran_at(xi1, -1, false);
    for (int w = 0; w < 4; w++) mc.wheel_invweight0[w] = tran_at(k.pw[w], w, false);
    for (int w = 0; w < 4; w++) {                              // softener body: sphere centre behind the ball joint (qpos0: ball = identity)
        const double sc[3] = MUSHR_SOFTENER_CENTER;
        double t[3], ps[3];
        mat_vec3(t, k.Rw[w], sc);
        for (int a = 0; a < 3; a++) ps[a] = k.pw[w][a] + t[a];
        mc.soft_invweight0[w] = tran_at(ps, w, true);
    }
    return mc;
}

}  // namespace mushr
}  // namespace ftgp
