// TEST-ONLY: compiles the product's kernel source (ft_grandprix_b200/csrc/mushr_step.cuh) for the HOST
// with g++ so that its arithmetic can be compared with the oracle on the CPU build box, where no GPU
// exists.  Nothing in the product links this file; on the GPU box the same source runs as the CUDA kernel.
#include "../../ft_grandprix_b200/csrc/mushr_consts.h"
using namespace ftgp::mushr;
static ModelConsts g_mc;
static bool g_ready = false;
extern "C" {
void hh_constants(double* dofinv31, double* wheelinv4, double* chassis_mean2) {
    if (!g_ready) { g_mc = model_constants(); g_ready = true; }
    for (int p = 0; p < NP; p++) dofinv31[p] = g_mc.dof_invweight0[p];
    for (int w = 0; w < 4; w++) wheelinv4[w] = g_mc.wheel_invweight0[w];
    chassis_mean2[0] = g_mc.chassis_invweight0; chassis_mean2[1] = g_mc.meaninertia;
}
int hh_step(double* qpos, double* qvel, double* warm, const double* ctrl, long n, int* info4) {
    if (!g_ready) { g_mc = model_constants(); g_ready = true; }
    for (long i = 0; i < n; i++) {
        StepInfo si;
        step_car(g_mc, qpos + i * NQ, qvel + i * NV, warm + i * NV, ctrl + 2 * i, NoWalls(), si);
        if (info4) { info4[4 * i] = si.iters; info4[4 * i + 1] = si.ncon_wheel; info4[4 * i + 2] = si.ncon_wall; info4[4 * i + 3] = si.reset; }
    }
    return 0;
}
}
