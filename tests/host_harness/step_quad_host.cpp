// TEST-ONLY: compiles the product's quad-per-car kernel source (ft_grandprix_b200/csrc/mushr_step_quad.cuh) for
// the HOST.  The four lanes of a quad are four OS threads; the quad's shuffles become exchanges through a small
// buffer between two barriers, summed in the same order as the device butterfly ((l + l^1) + (l^2 + l^3)), so the
// arithmetic is the device's up to FMA contraction.  Nothing in the product links this file.
#include "../../ft_grandprix_b200/csrc/mushr_consts.h"
#include "../../ft_grandprix_b200/csrc/mushr_step_quad.cuh"
#include <atomic>
#include <thread>
#include <vector>
using namespace ftgp::mushr;

struct SpinBarrier {
    std::atomic<int> count{0}, gen{0};
    void wait() {
        const int g = gen.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) == 3) { count.store(0, std::memory_order_relaxed); gen.store(g + 1, std::memory_order_release); }
        else { int spins = 0; while (gen.load(std::memory_order_acquire) == g) if (++spins > 2000) std::this_thread::yield(); }
    }
};
struct QuadHostShared {
    SpinBarrier bar;
    double slot[4]; unsigned bits[4];
    double priv[QP_N * 4]; double shr[QC_N]; double ktab[QK_N];
};
struct QuadHost : QuadMem<4, 1> {
    QuadHostShared* s;
    double sum(double v) const {
        s->slot[w] = v; s->bar.wait();
        const double a = v + s->slot[w ^ 1], b = s->slot[w ^ 2] + s->slot[w ^ 3];
        s->bar.wait();
        return a + b;
    }
    unsigned ballot(bool p) const {
        s->bits[w] = p ? 1u : 0u; s->bar.wait();
        const unsigned m = s->bits[0] | s->bits[1] << 1 | s->bits[2] << 2 | s->bits[3] << 3;
        s->bar.wait();
        return m;
    }
    bool any(bool p) const { return ballot(p) != 0; }
    // the host "warp" is the quad; `ghost` extra rounds emulate a quad that has to keep pace with slower cars of its
    // warp / CTA (its stores are switched off: the results must not change)
    mutable int ghost_w = 0, ghost_c = 0;
    bool wany(bool p) const { const bool a = any(p); if (!a && ghost_w > 0) { ghost_w--; return true; } return a; }
    bool cany(bool p) const { const bool a = any(p); if (!a && ghost_c > 0) { ghost_c--; return true; } return a; }
    void sync() const { s->bar.wait(); }
};

static ModelConsts g_mc;
static bool g_ready = false;
extern "C" int hq_step_ghost(double* qpos, double* qvel, double* warm, const double* ctrl, long n, int nsteps, int* info4, int ghost_w, int ghost_c) {
    if (!g_ready) { g_mc = model_constants(); g_ready = true; }
    QuadHostShared sh;
    for (int g = 0; g < 25; g++) quad_const_entry(g_mc, g, sh.ktab);
    std::vector<std::thread> th;
    for (int w = 0; w < 4; w++)
        th.emplace_back([&, w]() {
            QuadHost q; q.s = &sh; q.w = w; q.priv = sh.priv + w; q.shr = sh.shr; q.ktab = sh.ktab;
            for (long i = 0; i < n; i++)
                for (int k = 0; k < nsteps; k++) {
                    StepInfo si;
                    q.ghost_w = ghost_w; q.ghost_c = ghost_c;
                    step_car_quad(q, g_mc, qpos + i * NQ, qvel + i * NV, warm + i * NV, ctrl + 2 * i, QNoWalls(), true, si);
                    q.sync();
                    if (info4 && w == 0) { info4[4 * i] = si.iters; info4[4 * i + 1] = si.ncon_wheel; info4[4 * i + 2] = si.ncon_wall; info4[4 * i + 3] = si.reset; }
                }
        });
    for (auto& t : th) t.join();
    return 0;
}
extern "C" int hq_step(double* qpos, double* qvel, double* warm, const double* ctrl, long n, int nsteps, int* info4) {
    return hq_step_ghost(qpos, qvel, warm, ctrl, n, nsteps, info4, 0, 0);
}
