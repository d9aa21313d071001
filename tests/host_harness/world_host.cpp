// TEST-ONLY: compiles the product's world step (ft_grandprix_b200/csrc/mushr_world.cuh: cars that touch each other, one
// Newton problem per world solved from the per-car block-arrow factors by the Woodbury identity) for the HOST, so that
// it can be compared with the oracle's dense world solver (oracle/step.c fto_world_step) on the CPU build box.
#include "../../ft_grandprix_b200/csrc/mushr_consts.h"
#include "../../ft_grandprix_b200/csrc/mushr_step_quad.cuh"
#include "../../ft_grandprix_b200/csrc/mushr_world.cuh"
#include <memory>
using namespace ftgp::mushr;

static ModelConsts g_mc;
static bool g_ready = false;
static QHfWalls g_walls = {{nullptr, nullptr, 0, 0, 1.0, 1.0}, false, false};
extern "C" void hw_set_walls(const uint16_t* index, const uint32_t* chunks, int hc, int vc, double size_x, double size_y) {
    g_walls.on = index != nullptr;
    g_walls.hv.index = index; g_walls.hv.chunks = chunks; g_walls.hv.hc = hc; g_walls.hv.vc = vc; g_walls.hv.size_x = size_x; g_walls.hv.size_y = size_y;
}
// qpos [n][34], qvel / warm [n][29], ctrl [n][2], shadowed [n] or NULL; info: iters, car-car contacts, reset, rows of G at the last direction
extern "C" int hw_world_step(double* qpos, double* qvel, double* warm, const double* ctrl, int n, const unsigned char* shadowed, int* info4) {
    if (!g_ready) { g_mc = model_constants(); g_ready = true; }
    if (n < 1 || n > WMAXCARS) return -1;
    std::unique_ptr<WorldWork> W(new WorldWork());
    std::unique_ptr<Kin> kin(new Kin());
    double *q[WMAXCARS], *v[WMAXCARS], *w[WMAXCARS]; const double* u[WMAXCARS];
    QHfWalls walls[WMAXCARS]; bool sh[WMAXCARS];
    QHfWalls none = g_walls; none.on = false;
    for (int c = 0; c < n; c++) {
        q[c] = qpos + 34 * c; v[c] = qvel + 29 * c; w[c] = warm + 29 * c; u[c] = ctrl + 2 * c;
        sh[c] = shadowed && shadowed[c];
        walls[c] = sh[c] ? none : g_walls;
    }
    WorldInfo wi;
    world_step(SeqComm(), g_mc, n, q, v, w, u, walls, sh, *W, *kin, wi);
    if (info4) { info4[0] = wi.iters; info4[1] = wi.ncc; info4[2] = wi.reset; info4[3] = W->nrows; }
    return 0;
}
extern "C" int hw_world_has_contact(const double* qpos, int n, const unsigned char* shadowed) {
    const double* q[WMAXCARS]; bool sh[WMAXCARS];
    for (int c = 0; c < n; c++) { q[c] = qpos + 34 * c; sh[c] = shadowed && shadowed[c]; }
    return world_has_contact(n, q, sh) ? 1 : 0;
}

// TEST SUPPORT: the two root-only solves of the coupled-world direction against the full block-arrow solve.
// dense: [NP x NP] SPD matrix with the block-arrow pattern (root 7x7, four 6x6 chain blocks, borders chain x root), row-major.
// out[0] = max |S6[i][j] - (A^-1 e_i)[j]| over i < 6, j < NR;  out[1] = max |arrow_solve_root_rhs(b) - arrow_solve((b, 0))|.
extern "C" void hw_arrow_root_check(const double* dense, const double* br, double* out2) {
    std::unique_ptr<Arrow> A(new Arrow());
    for (int i = 0; i < NR; i++) for (int j = 0; j <= i; j++) A->R[tri(i, j)] = dense[i * NP + j];
    for (int w = 0; w < 4; w++) {
        for (int l = 0; l < NC; l++) for (int k = 0; k <= l; k++) A->W[w][tri(l, k)] = dense[(NR + NC * w + l) * NP + NR + NC * w + k];
        for (int l = 0; l < NC; l++) for (int j = 0; j < NR; j++) A->B[w][l][j] = dense[(NR + NC * w + l) * NP + j];
    }
    arrow_factor(*A);
    double S6[6][NR], worst_s = 0, worst_r = 0;
    arrow_root_inverse6(*A, S6);
    for (int i = 0; i < 6; i++) {
        double e[NP];
        for (int p = 0; p < NP; p++) e[p] = p == i ? 1.0 : 0.0;
        arrow_solve(*A, e);
        for (int j = 0; j < NR; j++) worst_s = fmax(worst_s, fabs(S6[i][j] - e[j]));
    }
    double x[NP], y[NP];
    for (int p = 0; p < NP; p++) x[p] = y[p] = p < NR ? br[p] : 0.0;
    arrow_solve_root_rhs(*A, x);
    arrow_solve(*A, y);
    for (int p = 0; p < NP; p++) worst_r = fmax(worst_r, fabs(x[p] - y[p]));
    out2[0] = worst_s; out2[1] = worst_r;
}
