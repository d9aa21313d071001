// mushr_step_quad.cuh -- QUAD-PER-CAR version of the vehicle step: four lanes advance one car, one lane per
// wheel chain (same arithmetic as mushr_step.cuh, i.e. mj_step for template/mushr.em.xml; replaces
// mujoco.mj_step at ft_grandprix/custom.py:1425).
//
// Why this mapping (profiles/ncu_summary_r03.md): thread-per-car needs 16.7 KB of scratch per thread, which lives
// in local memory and goes through DRAM 120x more often than the algorithmic state (FP64 pipe 5 % busy);
// warp-per-car keeps the state on chip but executes 5.4x the instructions (serial leader sections, lanes idle).
// The model itself says how to split a car: a 7-dof root plus FOUR IDENTICAL 6-slot wheel chains that are
// coupled only through the root (block-arrow M and H).  So lane w of a quad owns
//   * the chain block W_w (6x6), its border B_w (6x7), its wheel-ground contact, its share of the chassis
//     contacts, its friction-loss / limit / equality rows, and the chain part of every dof vector;
//   * a replica of the 7-entry root part of every dof vector and of the 7x7 root block.
// Chain work (the bulk: Cholesky of W_w, Y_w = L^-1 B_w, Schur term Y_w^T Y_w, J^T D J of the contact rows,
// kinematics / CRB / RNE of the wheel and softener bodies) runs 4-wide with no communication; the root block is
// reduced across the quad with two xor-shuffles per value and then handled redundantly by the four lanes, which
// keeps every loop bound and branch quad-uniform (sums are bit-identical in the four lanes).
// A warp holds 8 cars; per-lane state is ~1/4 of a car: the factorisation runs entirely in registers, M and the
// wheel contact Jacobian sit in shared memory ([slot][thread] layout: conflict-free), the rest is a ~1 KB frame.
//
// The code is __host__ __device__ over a communicator policy Q (device: shuffles inside the quad; host tests:
// four OS threads and a barrier, tests/host_harness/step_quad_host.cpp), so its arithmetic is checked against the
// oracle on the CPU build box before it runs on a B200.
#pragma once
#include "mushr_step.cuh"

#if defined(__CUDACC__)
#define FT_QN __host__ __device__ __noinline__
#else
#define FT_QN __attribute__((noinline))
#endif

namespace ftgp {
namespace mushr {

// ---- shared-memory slots (doubles).  P: private to the lane, C: one copy per car --------------------------------
constexpr int QP_MW = 0;             // 21  chain block of M, lower triangle
constexpr int QP_MB = 21;            // 36  border of M: slot l x root dof j < 6 (column 6 of M's border is zero)
constexpr int QP_CJ = 57;            // 18  wheel-ground contact Jacobian, row a x (3 root rotations, 3 chain slots)
constexpr int QP_N = 75;
constexpr int QC_MR = 0;             // 28  root block of M, lower triangle
constexpr int QC_N = 28;

template <int PS_, int CS_>
struct QuadMem {                     // PS: stride between slots of private data (threads per CTA), CS: cars per CTA
    static constexpr int PS = PS_, CS = CS_;
    double* priv; double* shr; int w;
    FT_HD double& P(int i) const { return priv[i * PS]; }
    FT_HD double& C(int i) const { return shr[i * CS]; }
    FT_HD int lane() const { return w; }
};

#if defined(__CUDACC__)
template <int PS_, int CS_>
struct QuadDev : QuadMem<PS_, CS_> {
    unsigned mask;
    __device__ __forceinline__ double sum(double v) const {
        v += __shfl_xor_sync(mask, v, 1);
        v += __shfl_xor_sync(mask, v, 2);
        return v;
    }
    __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(mask, p) >> (__ffs(mask) - 1)) & 0xFu; }
    __device__ __forceinline__ bool any(bool p) const { return __any_sync(mask, p) != 0; }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
};
#endif

// ---- per-lane solver state ------------------------------------------------------------------------------------
constexpr int QMAXCH = 2;            // chassis contacts one lane can own (MAXCON = 8 over four lanes)
struct QRows {
    double fr_D[NC], fr_Rf[NC], fr_f[NC], fr_aref[NC];     // friction loss of the chain slots (f = 0: no row)
    double eq_D, eq_aref, eq_der;                           // Ackermann equality of a front chain (D = 0: none)
    double lim_D[2], lim_aref[2], lim_sign[2];              // suspension, front steering (sign 0: inactive)
    double fr6_D, fr6_Rf, fr6_f, fr6_aref;                  // rows of root dof 6 (steering wheel): lane 0 only
    double lim6_D, lim6_aref, lim6_sign;
    double wc_D, wc_aref[4];                                // wheel-ground contact, D = 0: none
    int nch;                                                // chassis (wall) contacts owned by this lane
    double ch_D[QMAXCH], ch_aref[QMAXCH][4], ch_J[QMAXCH][3][6];
};
struct QCar {
    QRows r;
    double qfs_r[NR], qfs_c[NC], qas_r[NR], qas_c[NC];
    double x_r[NR], x_c[NC], Ma_r[NR], Ma_c[NC], g_r[NR], g_c[NC], s_r[NR], s_c[NC], Mv_r[NR], Mv_c[NC];
    double fc_r[NR], fc_c[NC];
    double cost, gauss;
    unsigned mask;                   // which rows are in their quadratic zone at x (bits below)
};
constexpr int QB_FR = 0, QB_FR6 = 6, QB_LIM = 7, QB_LIM6 = 9, QB_WC = 10, QB_CH = 14;
constexpr double WC_MU = 0.5, CH_MU = 1.0;

FT_HD int popc4(unsigned m) { return (int)((m & 1u) + ((m >> 1) & 1u) + ((m >> 2) & 1u) + ((m >> 3) & 1u)); }
FT_HD double dot6q(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5]; }

// wheel contact Jacobian row a (0 normal, 1, 2 tangents) in the lane's 9 columns (root 0-5, chain slots 0-2):
// frame n = (0,0,1), t1 = (0,1,0), t2 = (-1,0,0); the translation columns are constants
template <class Q>
FT_HD void wc_jac(const Q& qd, double J[3][9]) {
    J[0][0] = 0; J[0][1] = 0; J[0][2] = 1;
    J[1][0] = 0; J[1][1] = 1; J[1][2] = 0;
    J[2][0] = -1; J[2][1] = 0; J[2][2] = 0;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int k = 0; k < 6; k++) J[a][3 + k] = qd.P(QP_CJ + 6 * a + k);
}
FT_HD void wc_dots(const double J[3][9], const double* xr, const double* xc, double* d3) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
        double s = 0;
#pragma unroll
        for (int col = 0; col < 6; col++) s += J[a][col] * xr[col];
#pragma unroll
        for (int col = 0; col < 3; col++) s += J[a][6 + col] * xc[col];
        d3[a] = s;
    }
}
FT_HD void ch_dots(const double J[3][6], const double* xr, double* d3) {
    for (int a = 0; a < 3; a++) d3[a] = dot6q(J[a], xr);
}

// ---- cost of this lane's rows at (xr, xc): forces J^T f into fr (root, partial) / fc (chain), zone mask ----------
template <class Q>
FT_QN double rows_eval(const Q& qd, const QRows& r, const double* xr, const double* xc, double* fr, double* fc, unsigned& mask_out) {
    double cost = 0;
    unsigned mask = 0;
    for (int i = 0; i < NR; i++) fr[i] = 0;
    for (int l = 0; l < NC; l++) fc[l] = 0;
    if (r.eq_D > 0) {                                                    // equality: always quadratic
        const double jar = xc[1] - r.eq_der * xr[6] - r.eq_aref, D = r.eq_D;
        cost += 0.5 * D * jar * jar;
        const double f = -D * jar;
        fc[1] += f; fr[6] -= r.eq_der * f;
    }
#pragma unroll
    for (int l = 0; l < NC; l++) {                                       // friction loss
        const double f = r.fr_f[l];
        if (f <= 0) continue;
        const double jar = xc[l] - r.fr_aref[l], Rf = r.fr_Rf[l], D = r.fr_D[l];
        if (jar <= -Rf) { cost += -0.5 * Rf * f - f * jar; fc[l] += f; }
        else if (jar >= Rf) { cost += -0.5 * Rf * f + f * jar; fc[l] -= f; }
        else { cost += 0.5 * D * jar * jar; fc[l] += -D * jar; mask |= 1u << (QB_FR + l); }
    }
    if (r.fr6_f > 0) {
        const double f = r.fr6_f, jar = xr[6] - r.fr6_aref, Rf = r.fr6_Rf, D = r.fr6_D;
        if (jar <= -Rf) { cost += -0.5 * Rf * f - f * jar; fr[6] += f; }
        else if (jar >= Rf) { cost += -0.5 * Rf * f + f * jar; fr[6] -= f; }
        else { cost += 0.5 * D * jar * jar; fr[6] += -D * jar; mask |= 1u << QB_FR6; }
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {                                        // limits: active when jar < 0
        const double sg = r.lim_sign[k];
        if (sg == 0) continue;
        const double jar = sg * xc[k] - r.lim_aref[k], D = r.lim_D[k];
        if (jar < 0) { cost += 0.5 * D * jar * jar; fc[k] += sg * (-D * jar); mask |= 1u << (QB_LIM + k); }
    }
    if (r.lim6_sign != 0) {
        const double sg = r.lim6_sign, jar = sg * xr[6] - r.lim6_aref, D = r.lim6_D;
        if (jar < 0) { cost += 0.5 * D * jar * jar; fr[6] += sg * (-D * jar); mask |= 1u << QB_LIM6; }
    }
    if (r.wc_D > 0) {                                                    // pyramidal rows of the wheel contact
        double J[3][9], d3[3];
        wc_jac(qd, J);
        wc_dots(J, xr, xc, d3);
        const double D = r.wc_D;
#pragma unroll
        for (int rr = 0; rr < 4; rr++) {
            const double sg = (rr & 1) ? -WC_MU : WC_MU; const int ta = 1 + (rr >> 1);
            const double jar = d3[0] + sg * d3[ta] - r.wc_aref[rr];
            if (jar >= 0) continue;
            cost += 0.5 * D * jar * jar;
            mask |= 1u << (QB_WC + rr);
            const double f = -D * jar;
#pragma unroll
            for (int col = 0; col < 6; col++) fr[col] += (J[0][col] + sg * J[ta][col]) * f;
#pragma unroll
            for (int col = 0; col < 3; col++) fc[col] += (J[0][6 + col] + sg * J[ta][6 + col]) * f;
        }
    }
    for (int s = 0; s < r.nch; s++) {                                    // chassis contacts (walls): root dofs only
        double d3[3];
        ch_dots(r.ch_J[s], xr, d3);
        const double D = r.ch_D[s];
        for (int rr = 0; rr < 4; rr++) {
            const double sg = (rr & 1) ? -CH_MU : CH_MU; const int ta = 1 + (rr >> 1);
            const double jar = d3[0] + sg * d3[ta] - r.ch_aref[s][rr];
            if (jar >= 0) continue;
            cost += 0.5 * D * jar * jar;
            mask |= 1u << (QB_CH + 4 * s + rr);
            const double f = -D * jar;
            for (int col = 0; col < 6; col++) fr[col] += (r.ch_J[s][0][col] + sg * r.ch_J[s][ta][col]) * f;
        }
    }
    mask_out = mask;
    return cost;
}

// y = M x with M in shared memory; root part replicated
template <class Q>
FT_QN void quad_mul(const Q& qd, const double* xr, const double* xc, double* yr, double* yc) {
    double part[6];
#pragma unroll
    for (int j = 0; j < 6; j++) part[j] = 0;
#pragma unroll
    for (int l = 0; l < NC; l++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < NC; k++) s += qd.P(QP_MW + (l >= k ? tri(l, k) : tri(k, l))) * xc[k];
#pragma unroll
        for (int j = 0; j < 6; j++) { const double b = qd.P(QP_MB + 6 * l + j); s += b * xr[j]; part[j] += b * xc[l]; }
        yc[l] = s;
    }
#pragma unroll
    for (int i = 0; i < NR; i++) {
        double s = 0;
#pragma unroll
        for (int j = 0; j < NR; j++) s += qd.C(QC_MR + (i >= j ? tri(i, j) : tri(j, i))) * xr[j];
        if (i < 6) s += qd.sum(part[i]);
        yr[i] = s;
    }
}

// gradient, cost and constraint force at c.x (needs c.Ma = M c.x)
template <class Q>
FT_QN void quad_evaluate(const Q& qd, QCar& c) {
    double fr[NR];
    unsigned mask;
    double cc = rows_eval(qd, c.r, c.x_r, c.x_c, fr, c.fc_c, mask);
    c.mask = mask;
    double g = 0;
#pragma unroll
    for (int l = 0; l < NC; l++) g += (c.Ma_c[l] - c.qfs_c[l]) * (c.x_c[l] - c.qas_c[l]);
    cc = qd.sum(cc); g = qd.sum(g);
#pragma unroll
    for (int i = 0; i < NR; i++) { c.fc_r[i] = qd.sum(fr[i]); g += (c.Ma_r[i] - c.qfs_r[i]) * (c.x_r[i] - c.qas_r[i]); }
    c.gauss = 0.5 * g; c.cost = cc + c.gauss;
#pragma unroll
    for (int i = 0; i < NR; i++) c.g_r[i] = c.Ma_r[i] - c.qfs_r[i] - c.fc_r[i];
#pragma unroll
    for (int l = 0; l < NC; l++) c.g_c[l] = c.Ma_c[l] - c.qfs_c[l] - c.fc_c[l];
}

// (b_r, b_c) <- A^-1 (b_r, b_c) with A = M (mode 0), M + J^T D J over the rows flagged in c.mask (mode 1),
// M + h diag(damping) (mode 2).  The whole factorisation lives in registers: chain Cholesky, Y = L^-1 B, the lane's
// share of the Schur complement, one 28-value reduction across the quad, root Cholesky (replicated), solve.
template <class Q>
FT_QN void quad_factor_solve(const Q& qd, const QCar& c, int mode, double* br, double* bc) {
    const int w = qd.lane();
    double W[21], B[NC][NR], Pp[28];
#pragma unroll
    for (int i = 0; i < 21; i++) W[i] = qd.P(QP_MW + i);
#pragma unroll
    for (int l = 0; l < NC; l++) {
#pragma unroll
        for (int j = 0; j < 6; j++) B[l][j] = qd.P(QP_MB + 6 * l + j);
        B[l][6] = 0;
    }
#pragma unroll
    for (int i = 0; i < 28; i++) Pp[i] = 0;
    if (mode == 2) {
        W[tri(0, 0)] += TIMESTEP * 12.5;
        if (front(w)) W[tri(1, 1)] += TIMESTEP * 0.1;
        W[tri(2, 2)] += TIMESTEP * 0.01;
        if (w == 0) Pp[tri(6, 6)] += TIMESTEP * 0.1;
    } else if (mode == 1) {
        const QRows& r = c.r; const unsigned mask = c.mask;
        if (r.eq_D > 0) { W[tri(1, 1)] += r.eq_D; B[1][6] -= r.eq_D * r.eq_der; Pp[tri(6, 6)] += r.eq_D * r.eq_der * r.eq_der; }
#pragma unroll
        for (int l = 0; l < NC; l++) if (mask >> (QB_FR + l) & 1u) W[tri(l, l)] += r.fr_D[l];
        if (mask >> QB_FR6 & 1u) Pp[tri(6, 6)] += r.fr6_D;
#pragma unroll
        for (int k = 0; k < 2; k++) if (mask >> (QB_LIM + k) & 1u) W[tri(k, k)] += r.lim_D[k];
        if (mask >> QB_LIM6 & 1u) Pp[tri(6, 6)] += r.lim6_D;
        if (mask >> QB_WC & 0xFu) {
            double J[3][9];
            wc_jac(qd, J);
            const double D = r.wc_D;
#pragma unroll
            for (int rr = 0; rr < 4; rr++) {
                if (!(mask >> (QB_WC + rr) & 1u)) continue;
                const double sg = (rr & 1) ? -WC_MU : WC_MU; const int ta = 1 + (rr >> 1);
                double Jr[9];
#pragma unroll
                for (int col = 0; col < 9; col++) Jr[col] = J[0][col] + sg * J[ta][col];
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    const double di = D * Jr[i];
#pragma unroll
                    for (int j = 0; j <= i; j++) Pp[tri(i, j)] += di * Jr[j];
                }
#pragma unroll
                for (int l = 0; l < 3; l++) {
                    const double dl = D * Jr[6 + l];
#pragma unroll
                    for (int k = 0; k <= l; k++) W[tri(l, k)] += dl * Jr[6 + k];
#pragma unroll
                    for (int j = 0; j < 6; j++) B[l][j] += dl * Jr[j];
                }
            }
        }
        for (int s = 0; s < r.nch; s++) {
            const double D = r.ch_D[s];
            for (int rr = 0; rr < 4; rr++) {
                if (!(mask >> (QB_CH + 4 * s + rr) & 1u)) continue;
                const double sg = (rr & 1) ? -CH_MU : CH_MU; const int ta = 1 + (rr >> 1);
                double Jr[6];
#pragma unroll
                for (int col = 0; col < 6; col++) Jr[col] = r.ch_J[s][0][col] + sg * r.ch_J[s][ta][col];
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    const double di = D * Jr[i];
#pragma unroll
                    for (int j = 0; j <= i; j++) Pp[tri(i, j)] += di * Jr[j];
                }
            }
        }
    }
    // chain block: W = L L^T, diagonal keeps 1 / L_jj
#pragma unroll
    for (int j = 0; j < NC; j++) {
        double d = W[tri(j, j)];
#pragma unroll
        for (int k = 0; k < j; k++) d -= W[tri(j, k)] * W[tri(j, k)];
        if (d < MINVAL) d = MINVAL;
        const double id = inv_sqrt(d);
        W[tri(j, j)] = id;
#pragma unroll
        for (int i = j + 1; i < NC; i++) {
            double s = W[tri(i, j)];
#pragma unroll
            for (int k = 0; k < j; k++) s -= W[tri(i, k)] * W[tri(j, k)];
            W[tri(i, j)] = s * id;
        }
    }
    // Y = L^-1 B and the right-hand side's chain part z = L^-1 b_c
    double z[NC];
#pragma unroll
    for (int l = 0; l < NC; l++) {
#pragma unroll
        for (int col = 0; col < NR; col++) {
            double s = B[l][col];
#pragma unroll
            for (int k = 0; k < l; k++) s -= W[tri(l, k)] * B[k][col];
            B[l][col] = s * W[tri(l, l)];
        }
        double s = bc[l];
#pragma unroll
        for (int k = 0; k < l; k++) s -= W[tri(l, k)] * z[k];
        z[l] = s * W[tri(l, l)];
    }
    // lane's share of the Schur complement and of the root right-hand side, reduced across the quad
    double R[28], xr[NR];
#pragma unroll
    for (int i = 0; i < NR; i++) {
#pragma unroll
        for (int j = 0; j <= i; j++) {
            double s = Pp[tri(i, j)];
#pragma unroll
            for (int l = 0; l < NC; l++) s -= B[l][i] * B[l][j];
            R[tri(i, j)] = qd.C(QC_MR + tri(i, j)) + qd.sum(s);
        }
        double s = 0;
#pragma unroll
        for (int l = 0; l < NC; l++) s += B[l][i] * z[l];
        xr[i] = br[i] - qd.sum(s);
    }
    // root block (replicated in the four lanes)
#pragma unroll
    for (int j = 0; j < NR; j++) {
        double d = R[tri(j, j)];
#pragma unroll
        for (int k = 0; k < j; k++) d -= R[tri(j, k)] * R[tri(j, k)];
        if (d < MINVAL) d = MINVAL;
        const double id = inv_sqrt(d);
        R[tri(j, j)] = id;
#pragma unroll
        for (int i = j + 1; i < NR; i++) {
            double s = R[tri(i, j)];
#pragma unroll
            for (int k = 0; k < j; k++) s -= R[tri(i, k)] * R[tri(j, k)];
            R[tri(i, j)] = s * id;
        }
    }
#pragma unroll
    for (int i = 0; i < NR; i++) { double s = xr[i]; for (int k = 0; k < i; k++) s -= R[tri(i, k)] * xr[k]; xr[i] = s * R[tri(i, i)]; }
#pragma unroll
    for (int i = NR - 1; i >= 0; i--) { double s = xr[i]; for (int k = i + 1; k < NR; k++) s -= R[tri(k, i)] * xr[k]; xr[i] = s * R[tri(i, i)]; }
    // back substitution of the chain: x_c = L^-T (z - Y x_r)
#pragma unroll
    for (int l = 0; l < NC; l++) { double s = 0; for (int j = 0; j < NR; j++) s += B[l][j] * xr[j]; z[l] -= s; }
#pragma unroll
    for (int l = NC - 1; l >= 0; l--) { double s = z[l]; for (int k = l + 1; k < NC; k++) s -= W[tri(k, l)] * z[k]; z[l] = s * W[tri(l, l)]; }
#pragma unroll
    for (int i = 0; i < NR; i++) br[i] = xr[i];
#pragma unroll
    for (int l = 0; l < NC; l++) bc[l] = z[l];
}

// ---- exact line search (PrimalSearch) ---------------------------------------------------------------------------
struct QLs { double qg0, qg1, qg2; double wdx[3], wds[3], cdx[QMAXCH][3], cds[QMAXCH][3]; };

template <class Q>
FT_QN void quad_ls_eval(const Q& qd, const QCar& c, const QLs& L, LsPoint& pt, double alpha) {
    const QRows& r = c.r;
    double q0 = 0, q1 = 0, q2 = 0;
    if (r.eq_D > 0) {
        const double jar = c.x_c[1] - r.eq_der * c.x_r[6] - r.eq_aref, jv = c.s_c[1] - r.eq_der * c.s_r[6], D = r.eq_D;
        q0 += 0.5 * D * jar * jar; q1 += D * jar * jv; q2 += 0.5 * D * jv * jv;
    }
#pragma unroll
    for (int l = 0; l < NC; l++) {
        const double f = r.fr_f[l];
        if (f <= 0) continue;
        const double jar = c.x_c[l] - r.fr_aref[l], jv = c.s_c[l], Rf = r.fr_Rf[l], D = r.fr_D[l];
        const double xx = jar + alpha * jv;
        if (xx <= -Rf) { q0 += f * (-0.5 * Rf - jar); q1 += -f * jv; }
        else if (xx >= Rf) { q0 += f * (-0.5 * Rf + jar); q1 += f * jv; }
        else { q0 += 0.5 * D * jar * jar; q1 += D * jar * jv; q2 += 0.5 * D * jv * jv; }
    }
    if (r.fr6_f > 0) {
        const double f = r.fr6_f, jar = c.x_r[6] - r.fr6_aref, jv = c.s_r[6], Rf = r.fr6_Rf, D = r.fr6_D;
        const double xx = jar + alpha * jv;
        if (xx <= -Rf) { q0 += f * (-0.5 * Rf - jar); q1 += -f * jv; }
        else if (xx >= Rf) { q0 += f * (-0.5 * Rf + jar); q1 += f * jv; }
        else { q0 += 0.5 * D * jar * jar; q1 += D * jar * jv; q2 += 0.5 * D * jv * jv; }
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const double sg = r.lim_sign[k];
        if (sg == 0) continue;
        const double jar = sg * c.x_c[k] - r.lim_aref[k], jv = sg * c.s_c[k], D = r.lim_D[k];
        if (jar + alpha * jv < 0) { q0 += 0.5 * D * jar * jar; q1 += D * jar * jv; q2 += 0.5 * D * jv * jv; }
    }
    if (r.lim6_sign != 0) {
        const double sg = r.lim6_sign, jar = sg * c.x_r[6] - r.lim6_aref, jv = sg * c.s_r[6], D = r.lim6_D;
        if (jar + alpha * jv < 0) { q0 += 0.5 * D * jar * jar; q1 += D * jar * jv; q2 += 0.5 * D * jv * jv; }
    }
    if (r.wc_D > 0) {
        const double D = r.wc_D;
#pragma unroll
        for (int rr = 0; rr < 4; rr++) {
            const double sg = (rr & 1) ? -WC_MU : WC_MU; const int ta = 1 + (rr >> 1);
            const double jar = L.wdx[0] + sg * L.wdx[ta] - r.wc_aref[rr], jv = L.wds[0] + sg * L.wds[ta];
            if (jar + alpha * jv < 0) { q0 += 0.5 * D * jar * jar; q1 += D * jar * jv; q2 += 0.5 * D * jv * jv; }
        }
    }
    for (int s = 0; s < r.nch; s++) {
        const double D = r.ch_D[s];
        for (int rr = 0; rr < 4; rr++) {
            const double sg = (rr & 1) ? -CH_MU : CH_MU; const int ta = 1 + (rr >> 1);
            const double jar = L.cdx[s][0] + sg * L.cdx[s][ta] - r.ch_aref[s][rr], jv = L.cds[s][0] + sg * L.cds[s][ta];
            if (jar + alpha * jv < 0) { q0 += 0.5 * D * jar * jar; q1 += D * jar * jv; q2 += 0.5 * D * jv * jv; }
        }
    }
    q0 = qd.sum(q0) + L.qg0; q1 = qd.sum(q1) + L.qg1; q2 = qd.sum(q2) + L.qg2;
    pt.alpha = alpha; pt.cost = alpha * alpha * q2 + alpha * q1 + q0;
    pt.d0 = 2 * alpha * q2 + q1; pt.d1 = 2 * q2;
    if (pt.d1 <= 0) pt.d1 = MINVAL;
}

template <class Q>
FT_QN double quad_line_search(const Q& qd, QCar& c, double scale) {
    double sn = 0;
#pragma unroll
    for (int l = 0; l < NC; l++) sn += c.s_c[l] * c.s_c[l];
    sn = qd.sum(sn);
#pragma unroll
    for (int i = 0; i < NR; i++) sn += c.s_r[i] * c.s_r[i];
    const double snorm = sqrt(sn);
    if (snorm < MINVAL) return 0;
    quad_mul(qd, c.s_r, c.s_c, c.Mv_r, c.Mv_c);
    QLs L;
    double g1 = 0, g2 = 0;
#pragma unroll
    for (int l = 0; l < NC; l++) { g1 += c.s_c[l] * (c.Ma_c[l] - c.qfs_c[l]); g2 += 0.5 * c.s_c[l] * c.Mv_c[l]; }
    g1 = qd.sum(g1); g2 = qd.sum(g2);
#pragma unroll
    for (int i = 0; i < NR; i++) { g1 += c.s_r[i] * (c.Ma_r[i] - c.qfs_r[i]); g2 += 0.5 * c.s_r[i] * c.Mv_r[i]; }
    L.qg0 = c.gauss; L.qg1 = g1; L.qg2 = g2;
    if (c.r.wc_D > 0) {
        double J[3][9];
        wc_jac(qd, J);
        wc_dots(J, c.x_r, c.x_c, L.wdx); wc_dots(J, c.s_r, c.s_c, L.wds);
    }
    for (int s = 0; s < c.r.nch; s++) { ch_dots(c.r.ch_J[s], c.x_r, L.cdx[s]); ch_dots(c.r.ch_J[s], c.s_r, L.cds[s]); }
    const double gtol = SOLVER_TOL * LS_TOL * snorm / scale;
    LsPoint p0, p1, p2, pm, a1, a2;
    int it = 0;
    quad_ls_eval(qd, c, L, p0, 0);
    quad_ls_eval(qd, c, L, p1, p0.alpha - p0.d0 / p0.d1);
    if (p0.cost < p1.cost) p1 = p0;
    if (fabs(p1.d0) < gtol) return p1.alpha;
    const double dir = p1.d0 < 0 ? 1.0 : -1.0;
    bool p2update = false;
    p2 = p1;
    while (p1.d0 * dir <= -gtol && it < LS_ITER) {
        p2 = p1; p2update = true;
        quad_ls_eval(qd, c, L, p1, p1.alpha - p1.d0 / p1.d1); it++;
        if (fabs(p1.d0) < gtol) return p1.alpha;
    }
    if (it >= LS_ITER || !p2update) return p1.alpha;
    while (it < LS_ITER) {
        quad_ls_eval(qd, c, L, pm, 0.5 * (p1.alpha + p2.alpha)); it++;
        quad_ls_eval(qd, c, L, a1, p1.alpha - p1.d0 / p1.d1);
        quad_ls_eval(qd, c, L, a2, p2.alpha - p2.d0 / p2.d1);
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

// ---- position + velocity stage of the lane: kinematics, M -> shared memory, bias, smooth force, contacts, rows ----
struct QWallHit { double dist, nrm[3], t1[3], t2[3], pnt[3]; };
struct QNoWalls {                    // open ground
    FT_HD bool enabled() const { return false; }
    FT_HD bool operator()(const double*, const double*, int, QWallHit&) const { return false; }
};

template <class Q, class WallFn>
FT_QN void quad_prepare(const Q& qd, const ModelConsts& mc, const double* qr, const double* qc, const double* vr, const double* vc,
                        const double* ctrl, const WallFn& walls, QCar& c, StepInfo& info) {
    const int w = qd.lane();
    const bool fr = front(w);
    qd.sync();                       // the quad is done with the previous step's shared M
    // ---- kinematics (root replicated, own wheel chain)
    double q1[4] = {qr[3], qr[4], qr[5], qr[6]}, R1[9];
    quat_norm(q1); quat2mat(R1, q1);
    const double p1[3] = {qr[0], qr[1], qr[2]};
    const double zax[3] = {R1[2], R1[5], R1[8]};
    double p2[3], R2[9];
    { const double c2[3] = {SW_X, 0, SW_Z}; mat_vec3(p2, R1, c2); for (int a = 0; a < 3; a++) p2[a] += p1[a]; }
    { const double cs = cos(qr[7]), sn = sin(qr[7]); const double Rz[9] = {cs, -sn, 0, sn, cs, 0, 0, 0, 1}; mat_mul3(R2, R1, Rz); }
    double pw[3], Rsteer[9], Rw[9], Rs[9], ps[3];
    { const double cw[3] = {wheel_x(w), wheel_y(w), WHEEL_Z + qc[0]}; mat_vec3(pw, R1, cw); for (int a = 0; a < 3; a++) pw[a] += p1[a]; }
    { const double cs = cos(qc[1]), sn = sin(qc[1]); const double Rz[9] = {cs, -sn, 0, sn, cs, 0, 0, 0, 1}; mat_mul3(Rsteer, R1, Rz); }   // rear: angle 0
    { const double cs = cos(qc[2]), sn = sin(qc[2]); const double Ry[9] = {cs, 0, sn, 0, 1, 0, -sn, 0, cs}; mat_mul3(Rw, Rsteer, Ry); }
    {
        double qs[4] = {qc[3], qc[4], qc[5], qc[6]}, Rb[9];
        quat_norm(qs); quat2mat(Rb, qs); mat_mul3(Rs, Rw, Rb);
        const double sc[3] = MUSHR_SOFTENER_CENTER; double t[3];
        mat_vec3(t, Rs, sc);
        for (int a = 0; a < 3; a++) ps[a] = pw[a] + t[a];
    }
    double xi1[3], com[3];
    mat_vec3(xi1, R1, mc.ipos1);
    const double mtot = mc.mass1 + SW_MASS + 4 * (WHEEL_MASS + SOFT_MASS);
    for (int a = 0; a < 3; a++) {
        xi1[a] += p1[a];
        com[a] = (mc.mass1 * xi1[a] + SW_MASS * p2[a] + qd.sum(WHEEL_MASS * pw[a] + SOFT_MASS * ps[a])) / mtot;
    }
    // spatial inertias about the com
    double cin1[10], cinsw[10], cinw[10], cins[10], d[3];
    const double e0 = (WS1 * WS1 + WS2 * WS2) / 5, e1 = (WS0 * WS0 + WS2 * WS2) / 5, e2 = (WS0 * WS0 + WS1 * WS1) / 5;
    const double is = 0.4 * SOFT_MASS * MUSHR_SOFTENER_RADIUS * MUSHR_SOFTENER_RADIUS;
    for (int a = 0; a < 3; a++) d[a] = xi1[a] - com[a];
    inert_com(cin1, mc.inertia1, R1, d, mc.mass1);
    for (int a = 0; a < 3; a++) d[a] = p2[a] - com[a];
    inert_com_diag(cinsw, SW_MASS * e0, SW_MASS * e1, SW_MASS * e2, R2, d, SW_MASS);
    for (int a = 0; a < 3; a++) d[a] = pw[a] - com[a];
    inert_com_diag(cinw, WHEEL_MASS * e0, WHEEL_MASS * e1, WHEEL_MASS * e2, Rw, d, WHEEL_MASS);
    for (int a = 0; a < 3; a++) d[a] = ps[a] - com[a];
    inert_com_diag(cins, is, is, is, Rs, d, SOFT_MASS);
    // motion axes: root rows 0-6 (replicated), chain slots 0-5
    double cdr[NR][6], cd[NC][6], off[3];
    for (int p = 0; p < NR; p++) for (int a = 0; a < 6; a++) cdr[p][a] = 0;
    for (int l = 0; l < NC; l++) for (int a = 0; a < 6; a++) cd[l][a] = 0;
    for (int cc = 0; cc < 3; cc++) cdr[cc][3 + cc] = 1;
    for (int a = 0; a < 3; a++) off[a] = com[a] - p1[a];
    for (int cc = 0; cc < 3; cc++) { const double ax[3] = {R1[cc], R1[3 + cc], R1[6 + cc]}; for (int a = 0; a < 3; a++) cdr[3 + cc][a] = ax[a]; cross3(cdr[3 + cc] + 3, ax, off); }
    for (int a = 0; a < 3; a++) off[a] = com[a] - p2[a];
    for (int a = 0; a < 3; a++) cdr[6][a] = zax[a];
    cross3(cdr[6] + 3, zax, off);
    for (int a = 0; a < 3; a++) off[a] = com[a] - pw[a];
    for (int a = 0; a < 3; a++) cd[0][3 + a] = zax[a];
    if (fr) { for (int a = 0; a < 3; a++) cd[1][a] = zax[a]; cross3(cd[1] + 3, zax, off); }
    { const double ay[3] = {Rsteer[1], Rsteer[4], Rsteer[7]}; for (int a = 0; a < 3; a++) cd[2][a] = ay[a]; cross3(cd[2] + 3, ay, off); }
    for (int cc = 0; cc < 3; cc++) { const double ax[3] = {Rs[cc], Rs[3 + cc], Rs[6 + cc]}; for (int a = 0; a < 3; a++) cd[3 + cc][a] = ax[a]; cross3(cd[3 + cc] + 3, ax, off); }
    // ---- composite-rigid-body mass matrix -> shared memory
    {
        double crbw[10], crb1[10], buf[6];
        for (int a = 0; a < 10; a++) { crbw[a] = cinw[a] + cins[a]; crb1[a] = cin1[a] + cinsw[a] + qd.sum(crbw[a]); }
        for (int i = 0; i < 6; i++) {
            inert_mul(buf, crb1, cdr[i]);
            for (int j = 0; j <= i; j++) { const double s = dot6q(cdr[j], buf); if (((tri(i, j)) & 3) == w) qd.C(QC_MR + tri(i, j)) = s; }
        }
        inert_mul(buf, cinsw, cdr[6]);
        for (int j = 0; j <= 6; j++) { double s = dot6q(cdr[j], buf); if (j == 6) s += dof_armature(6); if (((tri(6, j)) & 3) == w) qd.C(QC_MR + tri(6, j)) = s; }
        for (int l = 0; l < NC; l++) {
            inert_mul(buf, l < 3 ? crbw : cins, cd[l]);
            for (int kk = 0; kk <= l; kk++) {
                double s = dot6q(cd[kk], buf);
                if (kk == l) s = (l == 1 && !fr) ? 1.0 : s + dof_armature(NR + l);      // rear dummy steering slot: unit diagonal
                qd.P(QP_MW + tri(l, kk)) = s;
            }
            for (int j = 0; j < 6; j++) qd.P(QP_MB + 6 * l + j) = dot6q(cdr[j], buf);
        }
    }
    // ---- bias forces (recursive Newton-Euler), smooth force
    double bias_r[NR], bias_c[NC];
    {
        double cv1[6] = {0, 0, 0, vr[0], vr[1], vr[2]}, cacc1[6] = {0, 0, 0, 0, 0, GRAV}, dd[6];
        double cvr[6] = {cv1[0], cv1[1], cv1[2], cv1[3], cv1[4], cv1[5]};
        for (int cc = 0; cc < 3; cc++) {
            cross_motion(dd, cv1, cdr[3 + cc]);
            for (int a = 0; a < 6; a++) { cacc1[a] += dd[a] * vr[3 + cc]; cvr[a] += cdr[3 + cc][a] * vr[3 + cc]; }
        }
        for (int a = 0; a < 6; a++) cv1[a] = cvr[a];
        double t[6], t2[6], cfrc1[6], f[6];
        inert_mul(cfrc1, cin1, cacc1);
        inert_mul(t, cin1, cv1); cross_force(t2, cv1, t);
        for (int a = 0; a < 6; a++) cfrc1[a] += t2[a];
        {
            double cv[6], ca[6];
            cross_motion(dd, cv1, cdr[6]);
            for (int a = 0; a < 6; a++) { ca[a] = cacc1[a] + dd[a] * vr[6]; cv[a] = cv1[a] + cdr[6][a] * vr[6]; }
            inert_mul(f, cinsw, ca);
            inert_mul(t, cinsw, cv); cross_force(t2, cv, t);
            for (int a = 0; a < 6; a++) f[a] += t2[a];
            bias_r[6] = dot6q(cdr[6], f);
            for (int a = 0; a < 6; a++) cfrc1[a] += f[a];
        }
        double cv[6], ca[6];
        for (int a = 0; a < 6; a++) { cv[a] = cv1[a]; ca[a] = cacc1[a]; }
        for (int l = 0; l < 3; l++) {                       // rear dummy slot: zero axis and zero velocity
            cross_motion(dd, cv, cd[l]);
            for (int a = 0; a < 6; a++) { ca[a] += dd[a] * vc[l]; cv[a] += cd[l][a] * vc[l]; }
        }
        double fw[6];
        inert_mul(fw, cinw, ca);
        inert_mul(t, cinw, cv); cross_force(t2, cv, t);
        for (int a = 0; a < 6; a++) fw[a] += t2[a];
        double cvs[6], cas[6];
        for (int a = 0; a < 6; a++) { cvs[a] = cv[a]; cas[a] = ca[a]; }
        for (int cc = 0; cc < 3; cc++) {
            cross_motion(dd, cv, cd[3 + cc]);
            for (int a = 0; a < 6; a++) { cas[a] += dd[a] * vc[3 + cc]; cvs[a] += cd[3 + cc][a] * vc[3 + cc]; }
        }
        double fs[6];
        inert_mul(fs, cins, cas);
        inert_mul(t, cins, cvs); cross_force(t2, cvs, t);
        for (int a = 0; a < 6; a++) { fs[a] += t2[a]; fw[a] += fs[a]; cfrc1[a] += qd.sum(fw[a]); }
        for (int l = 0; l < NC; l++) bias_c[l] = dot6q(cd[l], l < 3 ? fw : fs);
        for (int i = 0; i < 6; i++) bias_r[i] = dot6q(cdr[i], cfrc1);
    }
    for (int i = 0; i < NR; i++) c.qfs_r[i] = -bias_r[i] - dof_damping(i) * vr[i];
    for (int l = 0; l < NC; l++) c.qfs_c[l] = -bias_c[l] - dof_damping(NR + l) * vc[l];
    c.qfs_c[0] += -500.0 * (qc[0] - (-0.015));                                            // suspension spring :63
    if (!fr) c.qfs_c[1] = 0;
    {
        c.qfs_r[6] += 20.0 * ctrl[1] - 20.0 * qr[7];                                      // <position kp=20> :179
        const double tv = qd.sum(0.25 * vc[2]);
        double f = 100.0 * ctrl[0] - 100.0 * (0.04 * tv);                                 // <velocity kv=100 gear=0.04> :180
        f = f > 500.0 ? 500.0 : (f < -500.0 ? -500.0 : f);
        c.qfs_c[2] += 0.04 * 0.25 * f;
    }
    // ---- rows
    QRows& r = c.r;
    double K, B, imp, R;
    for (int l = 0; l < NC; l++) {
        const double f = (l == 1 && !fr) ? 0.0 : dof_floss(NR + l);
        r.fr_f[l] = f; r.fr_D[l] = 0; r.fr_Rf[l] = 0; r.fr_aref[l] = 0;
        if (f <= 0) continue;
        kbi(0.9, 0.0, mc.dof_invweight0[NR + NC * w + l], K, B, imp, R);
        r.fr_D[l] = 1 / R; r.fr_Rf[l] = R * f; r.fr_aref[l] = -B * vc[l];
    }
    r.fr6_f = 0; r.fr6_D = 0; r.fr6_Rf = 0; r.fr6_aref = 0;
    r.lim6_sign = 0; r.lim6_D = 0; r.lim6_aref = 0;
    if (w == 0) {
        kbi(0.9, 0.0, mc.dof_invweight0[6], K, B, imp, R);
        r.fr6_f = dof_floss(6); r.fr6_D = 1 / R; r.fr6_Rf = R * r.fr6_f; r.fr6_aref = -B * vr[6];
        const double q = qr[7];
        double dist = 0, sign = 0;
        if (q + 1 < 0) { dist = q + 1; sign = 1; } else if (1 - q < 0) { dist = 1 - q; sign = -1; }
        if (sign != 0) {
            kbi(0.9, dist, mc.dof_invweight0[6], K, B, imp, R);
            r.lim6_sign = sign; r.lim6_D = 1 / R; r.lim6_aref = -B * (sign * vr[6]) - K * imp * dist;
        }
    }
    r.eq_D = 0; r.eq_aref = 0; r.eq_der = 0;
    if (fr) {
        const double x = qr[7], pos = qc[1] - poly_val(w, x), der = poly_der(w, x);
        kbi(0.9, pos, mc.dof_invweight0[NR + NC * w + 1] + mc.dof_invweight0[6], K, B, imp, R);
        r.eq_der = der; r.eq_D = 1 / R; r.eq_aref = -B * (vc[1] - der * vr[6]) - K * imp * pos;
    }
    for (int k = 0; k < 2; k++) {
        r.lim_sign[k] = 0; r.lim_D[k] = 0; r.lim_aref[k] = 0;
        if (k == 1 && !fr) continue;
        const double q = qc[k], lo = k == 0 ? -0.03 : -1.0, hi = k == 0 ? 0.0 : 1.0;
        double dist, sign;
        if (q - lo < 0) { dist = q - lo; sign = 1; } else if (hi - q < 0) { dist = hi - q; sign = -1; } else continue;
        kbi(0.9, dist, mc.dof_invweight0[NR + NC * w + k], K, B, imp, R);
        r.lim_sign[k] = sign; r.lim_D[k] = 1 / R; r.lim_aref[k] = -B * (sign * vc[k]) - K * imp * dist;
    }
    // ---- wheel ellipsoid vs ground plane
    r.wc_D = 0;
    for (int rr = 0; rr < 4; rr++) r.wc_aref[rr] = 0;
    {
        const double dl[3] = {-Rw[6], -Rw[7], -Rw[8]};
        double s[3] = {WS0 * dl[0], WS1 * dl[1], WS2 * dl[2]};
        const double nn = sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
        s[0] = WS0 * s[0] / nn; s[1] = WS1 * s[1] / nn; s[2] = WS2 * s[2] / nn;
        double sw[3];
        mat_vec3(sw, Rw, s);
        for (int a = 0; a < 3; a++) sw[a] += pw[a];
        const double dist = sw[2] - PLANE_Z;
        const bool on = !(dist > 0);
        if (on) {
            const double o[3] = {sw[0] - com[0], sw[1] - com[1], sw[2] - 0.5 * dist - com[2]};
            double vel[3] = {vr[2], vr[1], -vr[0]};                       // translation columns of the frame
            for (int k = 0; k < 6; k++) {
                const double* ax = k < 3 ? cdr[3 + k] : cd[k - 3];
                double jp[3];
                cross3(jp, ax, o);
                for (int a = 0; a < 3; a++) jp[a] += ax[3 + a];
                const double j0 = jp[2], j1 = jp[1], j2 = -jp[0], vk = k < 3 ? vr[3 + k] : vc[k - 3];
                qd.P(QP_CJ + k) = j0; qd.P(QP_CJ + 6 + k) = j1; qd.P(QP_CJ + 12 + k) = j2;
                vel[0] += j0 * vk; vel[1] += j1 * vk; vel[2] += j2 * vk;
            }
            kbi(0.45, dist, mc.wheel_invweight0[w], K, B, imp, R);
            double Rpy = 2 * WC_MU * WC_MU * R; if (Rpy < MINVAL) Rpy = MINVAL;
            r.wc_D = 1 / Rpy;
            for (int rr = 0; rr < 4; rr++) {
                const double sg = (rr & 1) ? -1.0 : 1.0;
                r.wc_aref[rr] = -B * (vel[0] + sg * WC_MU * vel[1 + (rr >> 1)]) - K * imp * dist;
            }
        }
        info.ncon_wheel = popc4(qd.ballot(on));
    }
    // ---- chassis hull vertices vs walls: hit i (in vertex order, capped like the thread-per-car version) goes to
    // lane i & 3, slot i >> 2
    r.nch = 0;
    info.ncon_wall = 0;
    if (walls.enabled()) {
        unsigned hits = 0;
        for (int k = 0; k < 3; k++) {
            const int v = 4 * k + w;
            QWallHit h;
            const bool hit = v < MUSHR_CHASSIS_NHULL && walls(R1, p1, v, h);
            hits |= qd.ballot(hit) << (4 * k);
        }
        if (hits) {
            const int cap = MAXCON - info.ncon_wheel;
            int rank = 0;
            for (int v = 0; v < MUSHR_CHASSIS_NHULL && rank < cap; v++) {
                if (!(hits >> v & 1u)) continue;
                if ((rank & 3) == w) {
                    QWallHit h;
                    walls(R1, p1, v, h);
                    const int s = r.nch++;
                    double o[3], vel[3] = {0, 0, 0};
                    for (int a = 0; a < 3; a++) o[a] = h.pnt[a] - com[a];
                    for (int col = 0; col < 6; col++) {
                        double jp[3];
                        cross3(jp, cdr[col], o);
                        for (int a = 0; a < 3; a++) jp[a] += cdr[col][3 + a];
                        const double j0 = dot3(h.nrm, jp), j1 = dot3(h.t1, jp), j2 = dot3(h.t2, jp);
                        r.ch_J[s][0][col] = j0; r.ch_J[s][1][col] = j1; r.ch_J[s][2][col] = j2;
                        vel[0] += j0 * vr[col]; vel[1] += j1 * vr[col]; vel[2] += j2 * vr[col];
                    }
                    kbi(0.9, h.dist, mc.chassis_invweight0, K, B, imp, R);
                    double Rpy = 2 * CH_MU * CH_MU * R; if (Rpy < MINVAL) Rpy = MINVAL;
                    r.ch_D[s] = 1 / Rpy;
                    for (int rr = 0; rr < 4; rr++) {
                        const double sg = (rr & 1) ? -1.0 : 1.0;
                        r.ch_aref[s][rr] = -B * (vel[0] + sg * CH_MU * vel[1 + (rr >> 1)]) - K * imp * h.dist;
                    }
                }
                rank++;
            }
            info.ncon_wall = rank;
        }
    }
    qd.sync();                       // M (shared part) visible to the quad
}

// ---- the step -----------------------------------------------------------------------------------------------------
template <class Q, class WallFn>
FT_HDN void step_car_quad(const Q& qd, const ModelConsts& mc, double* qpos, double* qvel, double* warm, const double* ctrl,
                          const WallFn& walls, bool live, StepInfo& info) {
    const int w = qd.lane();
    const bool fr = front(w);
    const int qa = chain_q(w), da = chain_d(w);
    // ---- state of the lane: root (replicated) + own chain in slot order (susp, steer, throttle, ball)
    double qr[8], qc[7], vr[NR], vc[NC], wr[NR], wc[NC];
    for (int i = 0; i < 8; i++) qr[i] = qpos[i];
    for (int i = 0; i < NR; i++) { vr[i] = qvel[i]; wr[i] = warm[i]; }
    if (fr) {
        for (int i = 0; i < 7; i++) qc[i] = qpos[qa + i];
        for (int l = 0; l < NC; l++) { vc[l] = qvel[da + l]; wc[l] = warm[da + l]; }
    } else {
        qc[0] = qpos[qa]; qc[1] = 0; for (int i = 2; i < 7; i++) qc[i] = qpos[qa + i - 1];
        vc[0] = qvel[da]; vc[1] = 0; wc[0] = warm[da]; wc[1] = 0;
        for (int l = 2; l < NC; l++) { vc[l] = qvel[da + l - 1]; wc[l] = warm[da + l - 1]; }
    }
    info.reset = 0; info.iters = 0;
    {
        bool bad = false;                                                  // mj_checkPos / mj_checkVel
        for (int i = 0; i < 8; i++) bad |= bad_value(qr[i]);
        for (int i = 0; i < 7; i++) bad |= bad_value(qc[i]);
        for (int i = 0; i < NR; i++) bad |= bad_value(vr[i]);
        for (int l = 0; l < NC; l++) bad |= bad_value(vc[l]);
        if (qd.any(bad)) {
            info.reset = 1;
            for (int i = 0; i < 8; i++) qr[i] = 0;
            qr[1] = 2.0; qr[3] = 1.0;
            for (int i = 0; i < 7; i++) qc[i] = 0;
            qc[3] = 1.0;
            for (int i = 0; i < NR; i++) { vr[i] = 0; wr[i] = 0; }
            for (int l = 0; l < NC; l++) { vc[l] = 0; wc[l] = 0; }
        }
    }
    QCar c;
    quad_prepare(qd, mc, qr, qc, vr, vc, ctrl, walls, c, info);
    // ---- qacc_smooth = M^-1 qfrc_smooth
    for (int i = 0; i < NR; i++) c.qas_r[i] = c.qfs_r[i];
    for (int l = 0; l < NC; l++) c.qas_c[l] = c.qfs_c[l];
    int mode = 0;
    const double scale = 1.0 / (mc.meaninertia * NV);
    // one copy of the factor/solve code serves qacc_smooth (mode 0), every Newton direction (1) and the
    // implicit-damping Euler update (2)
    double* br = c.qas_r; double* bc = c.qas_c;
    double qa_r[NR], qa_c[NC];
    for (;;) {
        quad_factor_solve(qd, c, mode, br, bc);
        if (mode == 2) break;
        bool done = false;
        if (mode == 0) {
            // warm start if its cost beats qacc_smooth's (mj_fwdConstraint)
            double fr_[NR], fc_[NC]; unsigned m_;
            quad_mul(qd, wr, wc, c.Ma_r, c.Ma_c);
            double cw = qd.sum(rows_eval(qd, c.r, wr, wc, fr_, fc_, m_)), gw = 0;
            for (int l = 0; l < NC; l++) gw += 0.5 * (c.Ma_c[l] - c.qfs_c[l]) * (wc[l] - c.qas_c[l]);
            gw = qd.sum(gw);
            for (int i = 0; i < NR; i++) gw += 0.5 * (c.Ma_r[i] - c.qfs_r[i]) * (wr[i] - c.qas_r[i]);
            cw += gw;
            const double cs = qd.sum(rows_eval(qd, c.r, c.qas_r, c.qas_c, fr_, fc_, m_));
            if (cw > cs) {
                for (int i = 0; i < NR; i++) { c.x_r[i] = c.qas_r[i]; c.Ma_r[i] = c.qfs_r[i]; }
                for (int l = 0; l < NC; l++) { c.x_c[l] = c.qas_c[l]; c.Ma_c[l] = c.qfs_c[l]; }
            } else {
                for (int i = 0; i < NR; i++) c.x_r[i] = wr[i];
                for (int l = 0; l < NC; l++) c.x_c[l] = wc[l];
            }
            quad_evaluate(qd, c);
            mode = 1;
        } else {
            for (int i = 0; i < NR; i++) c.s_r[i] = -c.s_r[i];
            for (int l = 0; l < NC; l++) c.s_c[l] = -c.s_c[l];
            const double alpha = quad_line_search(qd, c, scale);
            if (alpha == 0) done = true;
            else {
                for (int i = 0; i < NR; i++) { c.x_r[i] += alpha * c.s_r[i]; c.Ma_r[i] += alpha * c.Mv_r[i]; }
                for (int l = 0; l < NC; l++) { c.x_c[l] += alpha * c.s_c[l]; c.Ma_c[l] += alpha * c.Mv_c[l]; }
                const double oldcost = c.cost;
                quad_evaluate(qd, c);
                double gn = 0;
                for (int l = 0; l < NC; l++) gn += c.g_c[l] * c.g_c[l];
                gn = qd.sum(gn);
                for (int i = 0; i < NR; i++) gn += c.g_r[i] * c.g_r[i];
                info.iters++;
                if (scale * (oldcost - c.cost) < SOLVER_TOL || scale * sqrt(gn) < SOLVER_TOL || info.iters >= SOLVER_ITER) done = true;
            }
        }
        if (done) {
            // mj_Euler with implicit joint damping: (M + h diag(b)) qacc' = qfrc_smooth + qfrc_constraint
            mode = 2;
            for (int i = 0; i < NR; i++) qa_r[i] = c.qfs_r[i] + c.fc_r[i];
            for (int l = 0; l < NC; l++) qa_c[l] = c.qfs_c[l] + c.fc_c[l];
            br = qa_r; bc = qa_c;
        } else {
            for (int i = 0; i < NR; i++) c.s_r[i] = c.g_r[i];
            for (int l = 0; l < NC; l++) c.s_c[l] = c.g_c[l];
            br = c.s_r; bc = c.s_c;
        }
    }
    {
        bool bad = false;                                                  // mj_checkAcc
        for (int i = 0; i < NR; i++) bad |= bad_value(c.x_r[i]);
        for (int l = 0; l < NC; l++) bad |= bad_value(c.x_c[l]);
        if (qd.any(bad)) {
            info.reset = 1;
            if (!live) return;
            if (w == 0) { for (int i = 0; i < 8; i++) qpos[i] = (i == 1) ? 2.0 : (i == 3 ? 1.0 : 0.0); for (int i = 0; i < NR; i++) { qvel[i] = 0; warm[i] = 0; } }
            const int nq = fr ? 7 : 6, nd = fr ? 6 : 5;
            for (int i = 0; i < nq; i++) qpos[qa + i] = (i == nq - 4) ? 1.0 : 0.0;
            for (int i = 0; i < nd; i++) { qvel[da + i] = 0; warm[da + i] = 0; }
            return;
        }
    }
    if (!live) return;
    // ---- velocity, then mj_integratePos with the new velocity
    for (int i = 0; i < NR; i++) vr[i] += TIMESTEP * qa_r[i];
    for (int l = 0; l < NC; l++) vc[l] += TIMESTEP * qa_c[l];
    if (w == 0) {
        for (int a = 0; a < 3; a++) qr[a] += TIMESTEP * vr[a];
        quat_integrate(qr + 3, vr + 3, TIMESTEP);
        qr[7] += TIMESTEP * vr[6];
        for (int i = 0; i < 8; i++) qpos[i] = qr[i];
        for (int i = 0; i < NR; i++) { qvel[i] = vr[i]; warm[i] = c.x_r[i]; }
    }
    for (int l = 0; l < 3; l++) qc[l] += TIMESTEP * vc[l];
    quat_integrate(qc + 3, vc + 3, TIMESTEP);
    if (fr) {
        for (int i = 0; i < 7; i++) qpos[qa + i] = qc[i];
        for (int l = 0; l < NC; l++) { qvel[da + l] = vc[l]; warm[da + l] = c.x_c[l]; }
    } else {
        qpos[qa] = qc[0]; for (int i = 2; i < 7; i++) qpos[qa + i - 1] = qc[i];
        qvel[da] = vc[0]; warm[da] = c.x_c[0];
        for (int l = 2; l < NC; l++) { qvel[da + l - 1] = vc[l]; warm[da + l - 1] = c.x_c[l]; }
    }
}

}  // namespace mushr
}  // namespace ftgp
