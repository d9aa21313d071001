// mushr_step_warp.cuh -- warp-per-car version of the vehicle step (same arithmetic as mushr_step.cuh).
//
// Why: the thread-per-car kernel needs ~21 KB of per-thread scratch, which lives in local memory and
// thrashes L1/L2 (profiles/ncu_summary_r01.md: 15.9 GB DRAM traffic per launch against 97 MB of
// algorithmic state, FP64 pipe 4 % active, warps stalled on long scoreboard).  Here ONE WARP advances one
// car and the whole working set sits in registers and ~12 KB of shared memory per warp:
//   * lane p (< 31) owns padded dof p (root 0-6, chain 7 + 6 w + l): its velocity, acceleration, gradient,
//     search direction, friction-loss / limit row, mass-matrix row ... are single registers;
//   * five "leader" lanes (root, four wheels) run the short serial recursions (kinematics, spatial
//     inertias, Newton-Euler) and publish the results in shared memory;
//   * the block-arrow matrices M and H live in shared memory and are factorised cooperatively
//     (LDL^T, 8 lanes per wheel block, Schur complement entries spread over 28 lanes);
//   * contact c is owned by lane c (pyramid rows, K = J^T D J weights), its 3x9 Jacobian is in shared memory;
//   * there is no divergence between cars: a warp's control flow (Newton iterations, line search) is
//     that of its one car, decided from warp-wide sums that are bit-identical in all lanes.
#pragma once
#include "mushr_step.cuh"

namespace ftgp {
namespace mushr {

#if defined(__CUDACC__)

struct WarpShared {
    double q[NQ];
    double R1[9], p1[3], p2[3], xi1[3];
    double pw[4][3], ps[4][3];
    union {                    // the position-stage scratch is dead before H is first written
        struct {
            double axis[NP][3];
            double cdof[NP][6];
            double inert[10][10];      // 0 car body, 1 steering wheel, 2+w wheel, 6+w softener
            double cfrc[10][6];
        };
        Arrow H;
    };
    double cframe[MAXCON][9], cpnt[MAXCON][3], cdist[MAXCON], cmu[MAXCON], cdmin[MAXCON], ctran[MAXCON];
    int cwheel[MAXCON];
    double J[MAXCON][3][9];
    double conD[MAXCON], conAref[MAXCON][4], conK[MAXCON][5], conF[MAXCON][3];
    double eqD[2], eqDer[2], eqF[2];
    Arrow M;
    double X[32], V[32];
};

__device__ __noinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double dot6(const double* a, const double* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
__device__ __forceinline__ double rcp_pivot(double d) { return __drcp_rn(d < MINVAL ? MINVAL : d); }

// ---- cooperative block-arrow algebra (A in shared memory; x vectors in shared memory, padded, 32 entries)
// y_p = (A x)_p for lane p
__device__ __noinline__ double arrow_mul_w(const Arrow& A, const double* x, int p) {
    double s = 0;
    if (p < NR) {
#pragma unroll
        for (int j = 0; j < NR; j++) s += A.R[p >= j ? tri(p, j) : tri(j, p)] * x[j];
#pragma unroll
        for (int w = 0; w < 4; w++)
#pragma unroll
            for (int l = 0; l < NC; l++) s += A.B[w][l][p] * x[NR + NC * w + l];
    } else if (p < NP) {
        const int w = (p - NR) / NC, l = (p - NR) % NC;
#pragma unroll
        for (int k = 0; k < NC; k++) s += A.W[w][l >= k ? tri(l, k) : tri(k, l)] * x[NR + NC * w + k];
#pragma unroll
        for (int j = 0; j < NR; j++) s += A.B[w][l][j] * x[j];
    }
    return s;
}

// in-place LDL^T of a lower-triangular packed N x N block in shared memory; lane t (< N) owns row t.
// On exit the diagonal holds 1/d_k and the strict lower part the unit-lower factor.
template <int N>
__device__ __forceinline__ void ldl_block(double* A, int t) {
#pragma unroll
    for (int k = 0; k < N - 1; k++) {
        const double inv = rcp_pivot(A[tri(k, k)]);
        const bool act = t > k && t < N;
        double aik = 0, upd[N];
        if (act) {
            aik = A[tri(t, k)];
            const double s = aik * inv;
#pragma unroll
            for (int m = k + 1; m < N; m++) upd[m] = (m <= t) ? s * A[tri(m, k)] : 0.0;
        }
        __syncwarp();
        if (act) {
#pragma unroll
            for (int m = k + 1; m < N; m++) if (m <= t) A[tri(t, m)] -= upd[m];
            A[tri(t, k)] = aik * inv;
        }
        if (t == k) A[tri(k, k)] = inv;
        __syncwarp();
    }
    if (t == N - 1) A[tri(N - 1, N - 1)] = rcp_pivot(A[tri(N - 1, N - 1)]);
    __syncwarp();
}

// A <- factor: W_w = L D L^T (diag holds 1/d), B_w <- Y_w = L_w^{-1} B_w, R <- LDL^T(R - sum_w Y_w^T D_w^{-1} Y_w)
__device__ __noinline__ void arrow_factor_w(Arrow& A, int T) {
    const int g = T >> 3, t = T & 7;
    ldl_block<NC>(A.W[g], t);
    if (t < NR) {                                   // column t of the border, forward substitution in registers
        double y[NC];
#pragma unroll
        for (int l = 0; l < NC; l++) {
            double s = A.B[g][l][t];
#pragma unroll
            for (int k = 0; k < NC; k++) if (k < l) s -= A.W[g][tri(l, k)] * y[k];
            y[l] = s;
        }
#pragma unroll
        for (int l = 0; l < NC; l++) A.B[g][l][t] = y[l];
    }
    __syncwarp();
    if (T < 28) {                                   // Schur complement entry (i, j)
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= T) i++;
        const int j = T - i * (i + 1) / 2;
        double s = A.R[T];
#pragma unroll
        for (int w = 0; w < 4; w++)
#pragma unroll
            for (int l = 0; l < NC; l++) s -= A.B[w][l][i] * A.W[w][tri(l, l)] * A.B[w][l][j];
        A.R[T] = s;
    }
    __syncwarp();
    ldl_block<NR>(A.R, T);
}

// X <- A^{-1} X (X in shared memory)
__device__ __noinline__ void arrow_solve_w(const Arrow& A, double* X, int T) {
    const int g = T >> 3, t = T & 7;
    if (t == 0) {                                   // z_w = L_w^{-1} x_w
        double z[NC];
        double* xc = X + NR + NC * g;
#pragma unroll
        for (int l = 0; l < NC; l++) {
            double s = xc[l];
#pragma unroll
            for (int k = 0; k < NC; k++) if (k < l) s -= A.W[g][tri(l, k)] * z[k];
            z[l] = s;
        }
#pragma unroll
        for (int l = 0; l < NC; l++) xc[l] = z[l];
    }
    __syncwarp();
    if (T < NR) {                                   // x_r -= sum_w Y_w^T D_w^{-1} z_w
        double s = X[T];
#pragma unroll
        for (int w = 0; w < 4; w++)
#pragma unroll
            for (int l = 0; l < NC; l++) s -= A.B[w][l][T] * A.W[w][tri(l, l)] * X[NR + NC * w + l];
        X[T] = s;
    }
    __syncwarp();
    if (T == 0) {                                   // root LDL^T solve, serial
        double z[NR];
#pragma unroll
        for (int i = 0; i < NR; i++) {
            double s = X[i];
#pragma unroll
            for (int k = 0; k < NR; k++) if (k < i) s -= A.R[tri(i, k)] * z[k];
            z[i] = s;
        }
#pragma unroll
        for (int i = 0; i < NR; i++) z[i] *= A.R[tri(i, i)];
#pragma unroll
        for (int i = NR - 1; i >= 0; i--) {
            double s = z[i];
#pragma unroll
            for (int k = 0; k < NR; k++) if (k > i) s -= A.R[tri(k, i)] * z[k];
            z[i] = s;
        }
#pragma unroll
        for (int i = 0; i < NR; i++) X[i] = z[i];
    }
    __syncwarp();
    if (t < NC) {                                   // t_l = D^{-1} (z_l - Y_l . x_r), one lane per chain row
        double s = X[NR + NC * g + t];
#pragma unroll
        for (int j = 0; j < NR; j++) s -= A.B[g][t][j] * X[j];
        s *= A.W[g][tri(t, t)];
        X[NR + NC * g + t] = s;
    }
    __syncwarp();
    if (t == 0) {                                   // x_w = L_w^{-T} t
        double z[NC];
        double* xc = X + NR + NC * g;
#pragma unroll
        for (int l = NC - 1; l >= 0; l--) {
            double s = xc[l];
#pragma unroll
            for (int k = 0; k < NC; k++) if (k > l) s -= A.W[g][tri(k, l)] * z[k];
            z[l] = s;
        }
#pragma unroll
        for (int l = 0; l < NC; l++) xc[l] = z[l];
    }
    __syncwarp();
}

__device__ __noinline__ void arrow_copy_w(Arrow& dst, const Arrow& src, int T) {
    double* d = reinterpret_cast<double*>(&dst); const double* s = reinterpret_cast<const double*>(&src);
    constexpr int n = sizeof(Arrow) / sizeof(double);
    for (int i = T; i < n; i += 32) d[i] = s[i];
    __syncwarp();
}

// impedance with this model's fixed solimp tail (dmax 0.95, width 0.001, midpoint 0.5, power 2)
__device__ __forceinline__ void kbi_w(double dmin, double pos, double diag, double& K, double& B, double& imp, double& R) {
    const double dmax = 0.95, tc = 0.02;
    double x = fabs(pos / 0.001);
    if (x >= 1) imp = dmax;
    else if (x == 0) imp = dmin;
    else { double y = x <= 0.5 ? x * x / 0.5 : 1 - (1 - x) * (1 - x) / 0.5; imp = dmin + y * (dmax - dmin); }
    R = (1 - imp) * diag / imp; if (R < MINVAL) R = MINVAL;
    K = 1 / (dmax * dmax * tc * tc); B = 2 / (dmax * tc);
}

// per-lane constraint row state (registers)
struct LaneRows {
    double frD, frAref, frRf, frf;        // friction loss on own dof
    double limD, limAref; int limSign;    // limit on own dof
    double eqD, eqAref, eqDer; int eqw;   // equality row owned by this lane (eqw = 0/1) or -1
    double cD, cAref[4], cmu; int cact;   // contact owned by this lane (lane c < ncon)
};

// Evaluates all rows at acceleration x (own entry xp; the full vector must already be in S.X).
// Returns the lane's cost share; fp = lane's entry of J^T force; hd = quadratic-row weight to add on H's diagonal.
// Contact lanes also publish K (J^T D J weights) and F (frame forces) in shared memory.
__device__ __noinline__ double eval_rows_w(WarpShared& S, const LaneRows& r, int p, int T, int ncon, double xp,
                                              double& fp, double& hd, double* cjar3, double& eqjar) {
    double cost = 0; fp = 0; hd = 0;
    if (r.frf > 0) {
        const double jar = xp - r.frAref;
        if (jar <= -r.frRf) { cost += -0.5 * r.frRf * r.frf - r.frf * jar; fp += r.frf; }
        else if (jar >= r.frRf) { cost += -0.5 * r.frRf * r.frf + r.frf * jar; fp -= r.frf; }
        else { cost += 0.5 * r.frD * jar * jar; fp += -r.frD * jar; hd += r.frD; }
    }
    if (r.limSign) {
        const double jar = r.limSign * xp - r.limAref;
        if (jar < 0) { cost += 0.5 * r.limD * jar * jar; fp += r.limSign * (-r.limD * jar); hd += r.limD; }
    }
    if (r.eqw >= 0) {
        const double jar = xp - r.eqDer * S.X[6] - r.eqAref;
        eqjar = jar;
        cost += 0.5 * r.eqD * jar * jar;
        const double f = -r.eqD * jar;
        fp += f; hd += r.eqD;
        S.eqF[r.eqw] = f;
    }
    if (T < ncon) {
        const int wh = S.cwheel[T];
        double d3[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            double s = 0;
#pragma unroll
            for (int c = 0; c < 6; c++) s += S.J[T][a][c] * S.X[c];
            if (wh >= 0) {
#pragma unroll
                for (int c = 0; c < 3; c++) s += S.J[T][a][6 + c] * S.X[NR + NC * wh + c];
            }
            d3[a] = s; cjar3[a] = s;
        }
        double Fn = 0, F1 = 0, F2 = 0, a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        const double mu = r.cmu, D = r.cD;
#pragma unroll
        for (int rr = 0; rr < 4; rr++) {
            const double sg = (rr & 1) ? -1.0 : 1.0;
            const double jar = d3[0] + sg * mu * d3[1 + (rr >> 1)] - r.cAref[rr];
            if (jar < 0) {
                cost += 0.5 * D * jar * jar;
                const double f = -D * jar;
                Fn += f;
                if (rr < 2) F1 += sg * mu * f; else F2 += sg * mu * f;
                if (rr == 0) a0 = 1; else if (rr == 1) a1 = 1; else if (rr == 2) a2 = 1; else a3 = 1;
            }
        }
        S.conF[T][0] = Fn; S.conF[T][1] = F1; S.conF[T][2] = F2;
        S.conK[T][0] = D * (a0 + a1 + a2 + a3);          // nn
        S.conK[T][1] = D * mu * (a0 - a1);               // n,t1
        S.conK[T][2] = D * mu * mu * (a0 + a1);          // t1,t1
        S.conK[T][3] = D * mu * (a2 - a3);               // n,t2
        S.conK[T][4] = D * mu * mu * (a2 + a3);          // t2,t2
    }
    __syncwarp();
    // gather J^T f
    if (p == 6) fp -= S.eqDer[0] * S.eqF[0] + S.eqDer[1] * S.eqF[1];
    if (p < 6) {
        for (int c = 0; c < ncon; c++) fp += S.J[c][0][p] * S.conF[c][0] + S.J[c][1][p] * S.conF[c][1] + S.J[c][2][p] * S.conF[c][2];
    } else if (p >= NR && p < NP) {
        const int w = (p - NR) / NC, l = (p - NR) % NC;
        if (l < 3)
            for (int c = 0; c < ncon; c++)
                if (S.cwheel[c] == w) fp += S.J[c][0][6 + l] * S.conF[c][0] + S.J[c][1][6 + l] * S.conF[c][1] + S.J[c][2][6 + l] * S.conF[c][2];
    }
    return cost;
}

// H = M + J^T D J for the rows that eval_rows_w found quadratic (hd per lane, K per contact)
__device__ __noinline__ void assemble_H_w(WarpShared& S, const LaneRows& r, int p, int T, int ncon, double hd) {
    arrow_copy_w(S.H, S.M, T);
    // contact blocks
    if (T < 21) {                                    // root-root (i, j < 6)
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= T) i++;
        const int j = T - i * (i + 1) / 2;
        double s = 0;
        for (int c = 0; c < ncon; c++) {
            const double* K = S.conK[c];
            const double n_i = S.J[c][0][i], t1i = S.J[c][1][i], t2i = S.J[c][2][i];
            const double n_j = S.J[c][0][j], t1j = S.J[c][1][j], t2j = S.J[c][2][j];
            s += K[0] * n_i * n_j + K[1] * (n_i * t1j + t1i * n_j) + K[2] * t1i * t1j + K[3] * (n_i * t2j + t2i * n_j) + K[4] * t2i * t2j;
        }
        S.H.R[T] += s;
    }
    for (int it = T; it < 4 * 24; it += 32) {        // per wheel: 6 chain-chain entries (l >= k, both < 3) + 18 border entries
        const int w = it / 24, e = it % 24;
        for (int c = 0; c < ncon; c++) {
            if (S.cwheel[c] != w) continue;
            const double* K = S.conK[c];
            int ca, cb;                              // Jacobian columns of the two indices
            double* dst;
            if (e < 6) { int l = e < 1 ? 0 : (e < 3 ? 1 : 2); int k = e - l * (l + 1) / 2; ca = 6 + l; cb = 6 + k; dst = &S.H.W[w][tri(l, k)]; }
            else { int l = (e - 6) / 6, j = (e - 6) % 6; ca = 6 + l; cb = j; dst = &S.H.B[w][l][j]; }
            const double n_i = S.J[c][0][ca], t1i = S.J[c][1][ca], t2i = S.J[c][2][ca];
            const double n_j = S.J[c][0][cb], t1j = S.J[c][1][cb], t2j = S.J[c][2][cb];
            *dst += K[0] * n_i * n_j + K[1] * (n_i * t1j + t1i * n_j) + K[2] * t1i * t1j + K[3] * (n_i * t2j + t2i * n_j) + K[4] * t2i * t2j;
        }
    }
    __syncwarp();
    // diagonal rows + equality couplings
    if (p < NR) {
        double add = hd;
        if (p == 6) add += S.eqD[0] * S.eqDer[0] * S.eqDer[0] + S.eqD[1] * S.eqDer[1] * S.eqDer[1];
        if (add != 0) S.H.R[tri(p, p)] += add;
    } else if (p < NP) {
        const int w = (p - NR) / NC, l = (p - NR) % NC;
        if (hd != 0) S.H.W[w][tri(l, l)] += hd;
        if (r.eqw >= 0) S.H.B[w][1][6] -= r.eqD * r.eqDer;
    }
    __syncwarp();
}

struct LsTot { double alpha, cost, d0, d1; };
struct LsLane { const LaneRows* r; double x, search, eqjar, eqjv, cj3[3], cs3[3], qg0, qg1, qg2; bool contact; };

// line-search point: total cost and derivatives at qacc + a * search (all lanes call; sums are warp-wide)
__device__ __noinline__ LsTot ls_eval_w(const LsLane& L, double a) {
    const LaneRows& r = *L.r;
    const double x = L.x, search = L.search;
    double q0 = 0, q1 = 0, q2 = 0;
    if (r.frf > 0) {
        const double jar = x - r.frAref, jv = search, xx = jar + a * jv;
        if (xx <= -r.frRf) { q0 += r.frf * (-0.5 * r.frRf - jar); q1 += -r.frf * jv; }
        else if (xx >= r.frRf) { q0 += r.frf * (-0.5 * r.frRf + jar); q1 += r.frf * jv; }
        else { q0 += 0.5 * r.frD * jar * jar; q1 += r.frD * jar * jv; q2 += 0.5 * r.frD * jv * jv; }
    }
    if (r.limSign) {
        const double jar = r.limSign * x - r.limAref, jv = r.limSign * search;
        if (jar + a * jv < 0) { q0 += 0.5 * r.limD * jar * jar; q1 += r.limD * jar * jv; q2 += 0.5 * r.limD * jv * jv; }
    }
    if (r.eqw >= 0) { q0 += 0.5 * r.eqD * L.eqjar * L.eqjar; q1 += r.eqD * L.eqjar * L.eqjv; q2 += 0.5 * r.eqD * L.eqjv * L.eqjv; }
    if (L.contact) {
        for (int rr = 0; rr < 4; rr++) {
            const double sg = (rr & 1) ? -1.0 : 1.0; const int ta = 1 + (rr >> 1);
            const double jar = L.cj3[0] + sg * r.cmu * L.cj3[ta] - r.cAref[rr], jv = L.cs3[0] + sg * r.cmu * L.cs3[ta];
            if (jar + a * jv < 0) { q0 += 0.5 * r.cD * jar * jar; q1 += r.cD * jar * jv; q2 += 0.5 * r.cD * jv * jv; }
        }
    }
    q0 = warp_sum(q0) + L.qg0; q1 = warp_sum(q1) + L.qg1; q2 = warp_sum(q2) + L.qg2;
    LsTot t;
    t.alpha = a; t.cost = a * a * q2 + a * q1 + q0; t.d0 = 2 * a * q2 + q1; t.d1 = 2 * q2;
    if (t.d1 <= 0) t.d1 = MINVAL;
    return t;
}

// one car, one warp.  status bits as in ftgp.h.
template <class WallFn>
__device__ void step_car_warp(WarpShared& S, const ModelConsts& mc, double* __restrict__ qpos_g, double* __restrict__ qvel_g,
                              double* __restrict__ warm_g, const double* __restrict__ ctrl_g, const WallFn& walls,
                              int T, bool live, StepInfo& info) {
    // Every warp of the CTA passes the same BLOCK_SYNC points: the warps then run the same code region at the same
    // time, which is what keeps this 200 KB kernel inside the 32 KB instruction cache (measured: 4 independent warps
    // per CTA 9.3 ms, 16 loosely aligned warps 7.3 ms at 65,536 cars).  `live` is false for the padding warps of
    // the last CTA: they run the step of the last car but write nothing.
#define BLOCK_SYNC() __syncthreads()
    const unsigned FULL = 0xffffffffu;
    const int p = T;                                     // padded dof of this lane (31 = spare lane)
    const bool chain = p >= NR && p < NP;
    const int w = chain ? (p - NR) / NC : -1, l = chain ? (p - NR) % NC : -1;
    const bool dummy = chain && !front(w) && l == 1;
    const int d = p < NP ? p2d(p) : -1;
    // ---- load
    for (int i = T; i < NQ; i += 32) S.q[i] = qpos_g[i];
    double v = d >= 0 ? qvel_g[d] : 0.0, wa = d >= 0 ? warm_g[d] : 0.0;
    const double ctrl0 = ctrl_g[0], ctrl1 = ctrl_g[1];
    __syncwarp();
    bool isbad = bad_value(v) || (T < NQ - 32 && bad_value(S.q[32 + T])) || bad_value(S.q[T]);
    info.reset = __any_sync(FULL, isbad) ? 1 : 0;
    if (info.reset) {                                    // mj_checkPos / mj_checkVel -> mj_resetData
        __syncwarp();
        for (int i = T; i < NQ; i += 32) S.q[i] = (i == 1) ? 2.0 : ((i == 3 || i == 11 || i == 18 || i == 24 || i == 30) ? 1.0 : 0.0);
        v = 0; wa = 0;
        __syncwarp();
    }
    S.V[T] = v;
    BLOCK_SYNC();
    // ---- leaders: kinematics
    const bool root_leader = T == 0, wheel_leader = chain && l == 0;
    double Rw[9], R2[9];                                 // leader-private orientations (wheel / steering wheel)
    if (root_leader || wheel_leader) {
        double q1[4] = {S.q[3], S.q[4], S.q[5], S.q[6]}, R1[9];
        quat_norm(q1); quat2mat(R1, q1);
        const double p1[3] = {S.q[0], S.q[1], S.q[2]};
        if (root_leader) {
            for (int a = 0; a < 9; a++) S.R1[a] = R1[a];
            for (int a = 0; a < 3; a++) S.p1[a] = p1[a];
            double c2[3] = {SW_X, 0, SW_Z}, t3[3];
            mat_vec3(t3, R1, c2);
            for (int a = 0; a < 3; a++) S.p2[a] = p1[a] + t3[a];
            { double c = cos(S.q[7]), s = sin(S.q[7]); double Rz[9] = {c, -s, 0, s, c, 0, 0, 0, 1}; mat_mul3(R2, R1, Rz); }
            mat_vec3(t3, R1, mc.ipos1);
            for (int a = 0; a < 3; a++) S.xi1[a] = p1[a] + t3[a];
            for (int c = 0; c < 3; c++) { S.axis[3 + c][0] = R1[c]; S.axis[3 + c][1] = R1[3 + c]; S.axis[3 + c][2] = R1[6 + c]; S.axis[c][0] = c == 0; S.axis[c][1] = c == 1; S.axis[c][2] = c == 2; }
            S.axis[6][0] = R1[2]; S.axis[6][1] = R1[5]; S.axis[6][2] = R1[8];
        } else {
            const int qa = chain_q(w);
            double c[3] = {wheel_x(w), wheel_y(w), WHEEL_Z + S.q[qa]}, pwl[3];
            mat_vec3(pwl, R1, c);
            for (int a = 0; a < 3; a++) { pwl[a] += p1[a]; S.pw[w][a] = pwl[a]; }
            double Rst[9], thr;
            if (front(w)) {
                double cs = cos(S.q[qa + 1]), sn = sin(S.q[qa + 1]);
                double Rz[9] = {cs, -sn, 0, sn, cs, 0, 0, 0, 1};
                mat_mul3(Rst, R1, Rz); thr = S.q[qa + 2];
            } else { for (int a = 0; a < 9; a++) Rst[a] = R1[a]; thr = S.q[qa + 1]; }
            { double cs = cos(thr), sn = sin(thr); double Ry[9] = {cs, 0, sn, 0, 1, 0, -sn, 0, cs}; mat_mul3(Rw, Rst, Ry); }
            const int qb = front(w) ? qa + 3 : qa + 2;
            double qs[4] = {S.q[qb], S.q[qb + 1], S.q[qb + 2], S.q[qb + 3]}, Rb[9], Rs[9];
            quat_norm(qs); quat2mat(Rb, qs); mat_mul3(Rs, Rw, Rb);
            const double sc[3] = MUSHR_SOFTENER_CENTER;
            double t3[3];
            mat_vec3(t3, Rs, sc);
            for (int a = 0; a < 3; a++) S.ps[w][a] = pwl[a] + t3[a];
            double (*ax)[3] = &S.axis[NR + NC * w];
            for (int a = 0; a < 3; a++) { ax[0][a] = R1[2 + 3 * a]; ax[1][a] = front(w) ? R1[2 + 3 * a] : 0.0; ax[2][a] = Rst[1 + 3 * a]; }
            for (int c3 = 0; c3 < 3; c3++) for (int a = 0; a < 3; a++) ax[3 + c3][a] = Rs[c3 + 3 * a];
        }
    }
    __syncwarp();
    // ---- every lane: centre of mass, own motion axis
    double com[3];
    {
        const double mtot = mc.mass1 + SW_MASS + 4 * (WHEEL_MASS + SOFT_MASS);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            double s = mc.mass1 * S.xi1[a] + SW_MASS * S.p2[a];
#pragma unroll
            for (int ww = 0; ww < 4; ww++) s += WHEEL_MASS * S.pw[ww][a] + SOFT_MASS * S.ps[ww][a];
            com[a] = s / mtot;
        }
    }
    double c6[6] = {0, 0, 0, 0, 0, 0};
    if (p < NP && !dummy) {
        const double ax[3] = {S.axis[p][0], S.axis[p][1], S.axis[p][2]};
        const bool trans = p < 3 || (chain && l == 0);
        if (trans) { c6[3] = ax[0]; c6[4] = ax[1]; c6[5] = ax[2]; }
        else {
            const double* an = p < 6 ? S.p1 : (p == 6 ? S.p2 : S.pw[w]);
            const double off[3] = {com[0] - an[0], com[1] - an[1], com[2] - an[2]};
            c6[0] = ax[0]; c6[1] = ax[1]; c6[2] = ax[2];
            cross3(c6 + 3, ax, off);
        }
    }
    if (p < NP) for (int a = 0; a < 6; a++) S.cdof[p][a] = c6[a];
    // ---- leaders: spatial inertias about the com
    if (root_leader) {
        double dd[3] = {S.xi1[0] - com[0], S.xi1[1] - com[1], S.xi1[2] - com[2]};
        inert_com(S.inert[0], mc.inertia1, S.R1, dd, mc.mass1);
        const double e0 = (WS1 * WS1 + WS2 * WS2) / 5, e1 = (WS0 * WS0 + WS2 * WS2) / 5, e2 = (WS0 * WS0 + WS1 * WS1) / 5;
        for (int a = 0; a < 3; a++) dd[a] = S.p2[a] - com[a];
        inert_com_diag(S.inert[1], SW_MASS * e0, SW_MASS * e1, SW_MASS * e2, R2, dd, SW_MASS);
    } else if (wheel_leader) {
        const double e0 = (WS1 * WS1 + WS2 * WS2) / 5, e1 = (WS0 * WS0 + WS2 * WS2) / 5, e2 = (WS0 * WS0 + WS1 * WS1) / 5;
        double dd[3] = {S.pw[w][0] - com[0], S.pw[w][1] - com[1], S.pw[w][2] - com[2]};
        inert_com_diag(S.inert[2 + w], WHEEL_MASS * e0, WHEEL_MASS * e1, WHEEL_MASS * e2, Rw, dd, WHEEL_MASS);
        const double is = 0.4 * SOFT_MASS * MUSHR_SOFTENER_RADIUS * MUSHR_SOFTENER_RADIUS;
        for (int a = 0; a < 3; a++) dd[a] = S.ps[w][a] - com[a];
        const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        inert_com_diag(S.inert[6 + w], is, is, is, I3, dd, SOFT_MASS);
    }
    __syncwarp();
    BLOCK_SYNC();
    // ---- mass-matrix row of this lane (mj_crb)
    if (p < NP) {
        double Ib[10];
        if (p < 6) { for (int a = 0; a < 10; a++) { double s = 0; for (int b = 0; b < 10; b++) s += S.inert[b][a]; Ib[a] = s; } }
        else if (p == 6) for (int a = 0; a < 10; a++) Ib[a] = S.inert[1][a];
        else if (l < 3) for (int a = 0; a < 10; a++) Ib[a] = S.inert[2 + w][a] + S.inert[6 + w][a];
        else for (int a = 0; a < 10; a++) Ib[a] = S.inert[6 + w][a];
        double buf[6];
        inert_mul(buf, Ib, c6);
        if (p < NR) {
            for (int j = 0; j <= p; j++) S.M.R[tri(p, j)] = dot6(S.cdof[j], buf) + (j == p ? dof_armature(p) : 0.0);
        } else {
            for (int k = 0; k <= l; k++) S.M.W[w][tri(l, k)] = dot6(S.cdof[NR + NC * w + k], buf) + (k == l ? (dummy ? 1.0 : dof_armature(p)) : 0.0);
            for (int j = 0; j < 6; j++) S.M.B[w][l][j] = dot6(S.cdof[j], buf);
            S.M.B[w][l][6] = 0;
        }
    }
    // ---- leaders: Newton-Euler forces per body (mj_comVel + mj_rne)
    if (root_leader || wheel_leader) {
        double cv1[6] = {0, 0, 0, S.V[0], S.V[1], S.V[2]}, cacc1[6] = {0, 0, 0, 0, 0, GRAV}, dd[6], cvr[6];
        for (int a = 0; a < 6; a++) cvr[a] = cv1[a];
        for (int c = 0; c < 3; c++) {
            cross_motion(dd, cv1, S.cdof[3 + c]);
            for (int a = 0; a < 6; a++) { cacc1[a] += dd[a] * S.V[3 + c]; cvr[a] += S.cdof[3 + c][a] * S.V[3 + c]; }
        }
        for (int a = 0; a < 6; a++) cv1[a] = cvr[a];
        double t6[6], t2[6];
        if (root_leader) {
            double f[6];
            inert_mul(f, S.inert[0], cacc1); inert_mul(t6, S.inert[0], cv1); cross_force(t2, cv1, t6);
            for (int a = 0; a < 6; a++) S.cfrc[0][a] = f[a] + t2[a];
            double cv[6], ca[6];
            cross_motion(dd, cv1, S.cdof[6]);
            for (int a = 0; a < 6; a++) { ca[a] = cacc1[a] + dd[a] * S.V[6]; cv[a] = cv1[a] + S.cdof[6][a] * S.V[6]; }
            inert_mul(f, S.inert[1], ca); inert_mul(t6, S.inert[1], cv); cross_force(t2, cv, t6);
            for (int a = 0; a < 6; a++) S.cfrc[1][a] = f[a] + t2[a];
        } else {
            const double (*cd)[6] = &S.cdof[NR + NC * w];
            const double* vc = &S.V[NR + NC * w];
            double cv[6], ca[6];
            for (int a = 0; a < 6; a++) { cv[a] = cv1[a]; ca[a] = cacc1[a]; }
            for (int ll = 0; ll < 3; ll++) {
                if (ll == 1 && !front(w)) continue;
                cross_motion(dd, cv, cd[ll]);
                for (int a = 0; a < 6; a++) { ca[a] += dd[a] * vc[ll]; cv[a] += cd[ll][a] * vc[ll]; }
            }
            double fw[6];
            inert_mul(fw, S.inert[2 + w], ca); inert_mul(t6, S.inert[2 + w], cv); cross_force(t2, cv, t6);
            for (int a = 0; a < 6; a++) fw[a] += t2[a];
            double cvs[6], cas[6];
            for (int a = 0; a < 6; a++) { cvs[a] = cv[a]; cas[a] = ca[a]; }
            for (int c = 0; c < 3; c++) {
                cross_motion(dd, cv, cd[3 + c]);
                for (int a = 0; a < 6; a++) { cas[a] += dd[a] * vc[3 + c]; cvs[a] += cd[3 + c][a] * vc[3 + c]; }
            }
            double fs[6];
            inert_mul(fs, S.inert[6 + w], cas); inert_mul(t6, S.inert[6 + w], cvs); cross_force(t2, cvs, t6);
            for (int a = 0; a < 6; a++) { fs[a] += t2[a]; S.cfrc[6 + w][a] = fs[a]; S.cfrc[2 + w][a] = fw[a] + fs[a]; }
        }
    }
    // ---- wheel leaders: ground contact geometry (mjc_PlaneConvex with the ellipsoid support point)
    bool hit = false; double cdist = 0, cp[3] = {0, 0, 0};
    if (wheel_leader) {
        double dl[3] = {-Rw[6], -Rw[7], -Rw[8]};
        double s3[3] = {WS0 * dl[0], WS1 * dl[1], WS2 * dl[2]};
        const double nn = sqrt(s3[0] * s3[0] + s3[1] * s3[1] + s3[2] * s3[2]);
        s3[0] = WS0 * s3[0] / nn; s3[1] = WS1 * s3[1] / nn; s3[2] = WS2 * s3[2] / nn;
        double sw[3];
        mat_vec3(sw, Rw, s3);
        for (int a = 0; a < 3; a++) sw[a] += S.pw[w][a];
        cdist = sw[2] - PLANE_Z;
        hit = !(cdist > 0);
        cp[0] = sw[0]; cp[1] = sw[1]; cp[2] = sw[2] - 0.5 * cdist;
    }
    const unsigned hitmask = __ballot_sync(FULL, hit);
    int ncon = __popc(hitmask);
    if (hit) {
        const int c = __popc(hitmask & ((1u << T) - 1));
        S.cdist[c] = cdist; S.cmu[c] = 0.5; S.cdmin[c] = 0.45; S.ctran[c] = mc.wheel_invweight0[w]; S.cwheel[c] = w;
        for (int a = 0; a < 3; a++) S.cpnt[c][a] = cp[a];
        const double fr[9] = {0, 0, 1, 0, 1, 0, -1, 0, 0};
        for (int a = 0; a < 9; a++) S.cframe[c][a] = fr[a];
    }
    info.ncon_wheel = ncon;
    ncon = walls(mc, S, com, T, ncon);                  // chassis-vs-wall contacts appended (warp-cooperative)
    info.ncon_wall = ncon - info.ncon_wheel;
    __syncwarp();
    BLOCK_SYNC();
    // ---- contact Jacobians: item = (contact, column)
    for (int it = T; it < ncon * 9; it += 32) {
        const int c = it / 9, col = it % 9, wh = S.cwheel[c];
        double jp[3] = {0, 0, 0};
        if (col < 6 || wh >= 0) {
            const int pp = col < 6 ? col : NR + NC * wh + (col - 6);
            const double off[3] = {S.cpnt[c][0] - com[0], S.cpnt[c][1] - com[1], S.cpnt[c][2] - com[2]};
            cross3(jp, S.cdof[pp], off);
            for (int a = 0; a < 3; a++) jp[a] += S.cdof[pp][3 + a];
        }
        for (int a = 0; a < 3; a++) S.J[c][a][col] = S.cframe[c][3 * a] * jp[0] + S.cframe[c][3 * a + 1] * jp[1] + S.cframe[c][3 * a + 2] * jp[2];
    }
    __syncwarp();
    // ---- bias force of this lane, smooth force
    double qfs = 0;
    if (p < NP && !dummy) {
        double f[6];
        if (p < 6) { for (int a = 0; a < 6; a++) f[a] = S.cfrc[0][a] + S.cfrc[1][a] + S.cfrc[2][a] + S.cfrc[3][a] + S.cfrc[4][a] + S.cfrc[5][a]; }
        else if (p == 6) for (int a = 0; a < 6; a++) f[a] = S.cfrc[1][a];
        else if (l < 3) for (int a = 0; a < 6; a++) f[a] = S.cfrc[2 + w][a];
        else for (int a = 0; a < 6; a++) f[a] = S.cfrc[6 + w][a];
        qfs = -dot6(c6, f) - dof_damping(p) * v;
        if (chain && l == 0) qfs += -500.0 * (S.q[chain_q(w)] + 0.015);
        if (p == 6) qfs += 20.0 * ctrl1 - 20.0 * S.q[7];
        if (chain && l == 2) {
            const double tv = 0.25 * (S.V[NR + 2] + S.V[NR + NC + 2] + S.V[NR + 2 * NC + 2] + S.V[NR + 3 * NC + 2]);
            double f2 = 100.0 * ctrl0 - 100.0 * (0.04 * tv);
            f2 = f2 > 500.0 ? 500.0 : (f2 < -500.0 ? -500.0 : f2);
            qfs += 0.04 * 0.25 * f2;
        }
    }
    // ---- constraint rows of this lane
    LaneRows r;
    r.frD = r.frAref = r.frRf = r.frf = 0; r.limD = r.limAref = 0; r.limSign = 0; r.eqD = r.eqAref = r.eqDer = 0; r.eqw = -1;
    r.cD = 0; r.cmu = 0; r.cact = 0; r.cAref[0] = r.cAref[1] = r.cAref[2] = r.cAref[3] = 0;
    {
        double K, B, imp, R;
        if (p >= 6 && p < NP && !dummy) {
            const double f = dof_floss(p);
            kbi_w(0.9, 0.0, mc.dof_invweight0[p], K, B, imp, R);
            r.frf = f; r.frD = 1 / R; r.frRf = R * f; r.frAref = -B * v;
        }
        const bool haslim = p == 6 || (chain && l == 0) || (chain && l == 1 && front(w));
        if (haslim) {
            const double qv = p == 6 ? S.q[7] : S.q[chain_q(w) + l];
            const double lo = (chain && l == 0) ? -0.03 : -1.0, hi = (chain && l == 0) ? 0.0 : 1.0;
            double dist = 0; int sg = 0;
            if (qv - lo < 0) { dist = qv - lo; sg = 1; } else if (hi - qv < 0) { dist = hi - qv; sg = -1; }
            if (sg) { kbi_w(0.9, dist, mc.dof_invweight0[p], K, B, imp, R); r.limSign = sg; r.limD = 1 / R; r.limAref = -B * (sg * v) - K * imp * dist; }
        }
        if (chain && l == 1 && front(w)) {
            const double x = S.q[7], q1 = S.q[chain_q(w) + 1];
            const double pos = q1 - poly_val(w, x), der = poly_der(w, x);
            kbi_w(0.9, pos, mc.dof_invweight0[p] + mc.dof_invweight0[6], K, B, imp, R);
            r.eqw = w; r.eqDer = der; r.eqD = 1 / R; r.eqAref = -B * (v - der * S.V[6]) - K * imp * pos;
            S.eqD[w] = r.eqD; S.eqDer[w] = der;
        }
        if (T < ncon) {
            const int wh = S.cwheel[T];
            kbi_w(S.cdmin[T], S.cdist[T], S.ctran[T], K, B, imp, R);
            const double mu = S.cmu[T];
            double Rpy = 2 * mu * mu * R; if (Rpy < MINVAL) Rpy = MINVAL;
            r.cD = 1 / Rpy; r.cmu = mu;
            double vel[3];
            for (int a = 0; a < 3; a++) {
                double s = 0;
                for (int c = 0; c < 6; c++) s += S.J[T][a][c] * S.V[c];
                if (wh >= 0) for (int c = 0; c < 3; c++) s += S.J[T][a][6 + c] * S.V[NR + NC * wh + c];
                vel[a] = s;
            }
            for (int rr = 0; rr < 4; rr++) {
                const double sg = (rr & 1) ? -1.0 : 1.0;
                r.cAref[rr] = -B * (vel[0] + sg * mu * vel[1 + (rr >> 1)]) - K * imp * S.cdist[T];
            }
        }
    }
    __syncwarp();
    BLOCK_SYNC();
    // ---- qacc_smooth = M^{-1} qfrc_smooth
    arrow_copy_w(S.H, S.M, T);
    arrow_factor_w(S.H, T);
    S.X[T] = qfs;
    __syncwarp();
    arrow_solve_w(S.H, S.X, T);
    const double qas = S.X[T];
    __syncwarp();
    BLOCK_SYNC();
    // ---- warm start vs smooth start
    double x, Ma, fp, hd, cj3[3] = {0, 0, 0}, eqjar = 0;
    {
        S.X[T] = wa;
        __syncwarp();
        const double Maw = arrow_mul_w(S.M, S.X, p);
        double cw = eval_rows_w(S, r, p, T, ncon, wa, fp, hd, cj3, eqjar) + 0.5 * (Maw - qfs) * (wa - qas);
        cw = warp_sum(cw);
        __syncwarp();
        S.X[T] = qas;
        __syncwarp();
        double cs = warp_sum(eval_rows_w(S, r, p, T, ncon, qas, fp, hd, cj3, eqjar));
        __syncwarp();
        if (cw > cs) { x = qas; Ma = qfs; } else { x = wa; Ma = Maw; }
    }
    const double scale = 1.0 / (mc.meaninertia * NV);
    double cost, gauss, grad, search;
    // cost, constraint forces and gradient at x
    auto evaluate = [&]() {
        S.X[T] = x;
        __syncwarp();
        double c = eval_rows_w(S, r, p, T, ncon, x, fp, hd, cj3, eqjar);
        const double g = 0.5 * (Ma - qfs) * (x - qas);
        gauss = warp_sum(g);
        cost = warp_sum(c) + gauss;
        grad = (p < NP) ? Ma - qfs - fp : 0.0;
    };
    // Newton direction: search = -H^{-1} grad, H = M + J^T D J over the rows evaluate() found quadratic
    auto solve_direction = [&]() {
        assemble_H_w(S, r, p, T, ncon, hd);
        arrow_factor_w(S.H, T);
        S.X[T] = grad;
        __syncwarp();
        arrow_solve_w(S.H, S.X, T);
        search = -S.X[T];
        __syncwarp();
    };
    BLOCK_SYNC();
    evaluate();
    solve_direction();
    int iter = 0;
    bool active = true;
    // the loop is CTA-uniform: a warp whose car has converged idles through the remaining rounds
    while (__syncthreads_or(active && iter < SOLVER_ITER)) {
        // ---- exact line search (PrimalSearch)
        double alpha = 0;
        if (active && iter < SOLVER_ITER) {
            const double snorm = sqrt(warp_sum(search * search));
            if (snorm >= MINVAL) {
                S.X[T] = search;
                __syncwarp();
                const double Mv = arrow_mul_w(S.M, S.X, p);
                // per-lane jv
                double cs3[3] = {0, 0, 0}, eqjv = 0;
                if (r.eqw >= 0) eqjv = search - r.eqDer * S.X[6];
                if (T < ncon) {
                    const int wh = S.cwheel[T];
                    for (int a = 0; a < 3; a++) {
                        double s = 0;
                        for (int c = 0; c < 6; c++) s += S.J[T][a][c] * S.X[c];
                        if (wh >= 0) for (int c = 0; c < 3; c++) s += S.J[T][a][6 + c] * S.X[NR + NC * wh + c];
                        cs3[a] = s;
                    }
                }
                __syncwarp();
                const double qg0 = gauss, qg1 = warp_sum(search * (Ma - qfs)), qg2 = warp_sum(0.5 * search * Mv);
                const double gtol = SOLVER_TOL * LS_TOL * snorm / scale;
                LsLane L;
                L.r = &r; L.x = x; L.search = search; L.eqjar = eqjar; L.eqjv = eqjv;
                L.cj3[0] = cj3[0]; L.cj3[1] = cj3[1]; L.cj3[2] = cj3[2]; L.cs3[0] = cs3[0]; L.cs3[1] = cs3[1]; L.cs3[2] = cs3[2];
                L.qg0 = qg0; L.qg1 = qg1; L.qg2 = qg2; L.contact = T < ncon;
                auto ev = [&](double a) { return ls_eval_w(L, a); };
                LsTot p0 = ev(0), p1 = ev(p0.alpha - p0.d0 / p0.d1), p2, pm, a1, a2;
                if (p0.cost < p1.cost) p1 = p0;
                bool done = fabs(p1.d0) < gtol;
                alpha = p1.alpha;
                if (!done) {
                    int it = 0;
                    const double dir = p1.d0 < 0 ? 1.0 : -1.0;
                    bool p2update = false;
                    p2 = p1;
                    while (p1.d0 * dir <= -gtol && it < LS_ITER) {
                        p2 = p1; p2update = true;
                        p1 = ev(p1.alpha - p1.d0 / p1.d1); it++;
                        if (fabs(p1.d0) < gtol) { done = true; break; }
                    }
                    alpha = p1.alpha;
                    if (!done && it < LS_ITER && p2update) {
                        bool found = false;
                        while (it < LS_ITER) {
                            pm = ev(0.5 * (p1.alpha + p2.alpha)); it++;
                            a1 = ev(p1.alpha - p1.d0 / p1.d1);
                            a2 = ev(p2.alpha - p2.d0 / p2.d1);
                            if (fabs(a1.d0) < gtol) { alpha = a1.alpha; found = true; break; }
                            if (fabs(a2.d0) < gtol) { alpha = a2.alpha; found = true; break; }
                            if (fabs(pm.d0) < gtol) { alpha = pm.alpha; found = true; break; }
                            bool b1 = false, b2 = false;
                            double lo = fmin(p1.alpha, p2.alpha), hi = fmax(p1.alpha, p2.alpha);
                            for (int cnd = 0; cnd < 3; cnd++) {
                                const LsTot& qq = cnd == 0 ? a1 : (cnd == 1 ? a2 : pm);
                                if (qq.alpha <= lo || qq.alpha >= hi) continue;
                                if ((qq.d0 < 0) == (p1.d0 < 0)) { p1 = qq; b1 = true; } else { p2 = qq; b2 = true; }
                                lo = fmin(p1.alpha, p2.alpha); hi = fmax(p1.alpha, p2.alpha);
                            }
                            if (!b1 && !b2) break;
                        }
                        if (!found) alpha = p1.cost <= p2.cost ? p1.alpha : p2.alpha;
                    }
                }
                if (alpha != 0) { x += alpha * search; Ma += alpha * Mv; }
            }
        }
        else active = false;
        if (active) {
            if (alpha == 0) active = false;
            else {
                const double oldcost = cost;
                evaluate();
                const double gn = warp_sum(grad * grad);
                iter++;
                // MuJoCo factorises H before this test; the direction is unused when the test ends the loop,
                // so the factorisation is skipped then (same qacc, one block-arrow LDL^T less per step)
                if (scale * (oldcost - cost) < SOLVER_TOL || scale * sqrt(gn) < SOLVER_TOL) active = false;
            }
        }
        BLOCK_SYNC();
        if (active) solve_direction();
    }
    info.iters = iter;
    const bool badacc = __any_sync(FULL, bad_value(x));   // mj_checkAcc -> mj_resetData
    if (badacc) info.reset = 1;
    BLOCK_SYNC();
    // ---- mj_Euler with implicit joint damping
    arrow_copy_w(S.H, S.M, T);
    if (p >= 6 && p < NP && !dummy) {
        const double b = TIMESTEP * dof_damping(p);
        if (b != 0) { if (p == 6) S.H.R[tri(6, 6)] += b; else S.H.W[w][tri(l, l)] += b; }
    }
    __syncwarp();
    arrow_factor_w(S.H, T);
    S.X[T] = qfs + fp;
    __syncwarp();
    arrow_solve_w(S.H, S.X, T);
    const double vnew = v + TIMESTEP * S.X[T];
    __syncwarp();
    if (badacc) {
        if (live) {
            for (int i = T; i < NQ; i += 32) qpos_g[i] = (i == 1) ? 2.0 : ((i == 3 || i == 11 || i == 18 || i == 24 || i == 30) ? 1.0 : 0.0);
            if (d >= 0) { qvel_g[d] = 0; warm_g[d] = 0; }
        }
        return;
    }
    if (d >= 0 && live) { warm_g[d] = x; qvel_g[d] = vnew; }
    S.X[T] = vnew;
    __syncwarp();
    if (!live) return;
    // ---- mj_integratePos with the new velocity
    if (T < 3) qpos_g[T] = S.q[T] + TIMESTEP * S.X[T];
    else if (T == 3) {
        double qq[4] = {S.q[3], S.q[4], S.q[5], S.q[6]}, om[3] = {S.X[3], S.X[4], S.X[5]};
        quat_integrate(qq, om, TIMESTEP);
        for (int a = 0; a < 4; a++) qpos_g[3 + a] = qq[a];
    } else if (T == 4) qpos_g[7] = S.q[7] + TIMESTEP * S.X[6];
    else if (T >= 8 && T < 12) {
        const int ww = T - 8, qa0 = chain_q(ww), base = NR + NC * ww;
        qpos_g[qa0] = S.q[qa0] + TIMESTEP * S.X[base];
        if (front(ww)) { qpos_g[qa0 + 1] = S.q[qa0 + 1] + TIMESTEP * S.X[base + 1]; qpos_g[qa0 + 2] = S.q[qa0 + 2] + TIMESTEP * S.X[base + 2]; }
        else qpos_g[qa0 + 1] = S.q[qa0 + 1] + TIMESTEP * S.X[base + 2];
    } else if (T >= 12 && T < 16) {
        const int ww = T - 12, qb = chain_q(ww) + (front(ww) ? 3 : 2), base = NR + NC * ww + 3;
        double qq[4] = {S.q[qb], S.q[qb + 1], S.q[qb + 2], S.q[qb + 3]}, om[3] = {S.X[base], S.X[base + 1], S.X[base + 2]};
        quat_integrate(qq, om, TIMESTEP);
        for (int a = 0; a < 4; a++) qpos_g[qb + a] = qq[a];
    }
    __syncwarp();
#undef BLOCK_SYNC
}

#endif  // __CUDACC__

}  // namespace mushr
}  // namespace ftgp
