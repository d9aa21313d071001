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
// walls: one compiled track of the product's geometry blob (ftgp_geom_blob / ftgp_blob_track_view), or none
static QHfWalls g_walls = {{nullptr, nullptr, 0, 0, 1.0, 1.0}, false, false};
extern "C" void hq_set_walls(const uint16_t* index, const uint32_t* chunks, int hc, int vc, double size_x, double size_y) {
    g_walls.on = index != nullptr;
    g_walls.hv.index = index; g_walls.hv.chunks = chunks; g_walls.hv.hc = hc; g_walls.hv.vc = vc; g_walls.hv.size_x = size_x; g_walls.hv.size_y = size_y;
}
extern "C" void hq_set_bubble_wrap(int on) { g_walls.bubble = on != 0; }
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
                    step_car_quad(q, g_mc, qpos + i * NQ, qvel + i * NV, warm + i * NV, ctrl + 2 * i, g_walls, true, si, QStage{0, false, nullptr});
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

// staged solve: at most k1 Newton rounds, then (suspended cars) k2 more, then to convergence -- the records live on the host
extern "C" int hq_step_staged(double* qpos, double* qvel, double* warm, const double* ctrl, long n, int* info4, int k1, int k2, int* nsuspended) {
    if (!g_ready) { g_mc = model_constants(); g_ready = true; }
    QuadHostShared sh;
    for (int g = 0; g < 25; g++) quad_const_entry(g_mc, g, sh.ktab);
    std::vector<double> rec(QREC_DOUBLES);
    std::vector<std::thread> th;
    int nsus[4] = {0, 0, 0, 0};
    for (int w = 0; w < 4; w++)
        th.emplace_back([&, w]() {
            QuadHost q; q.s = &sh; q.w = w; q.priv = sh.priv + w; q.shr = sh.shr; q.ktab = sh.ktab;
            for (long i = 0; i < n; i++) {
                StepInfo si;
                const int budget[3] = {k1, k2, 0};
                bool sus = false;
                for (int stage = 0; stage < 3; stage++) {
                    sus = step_car_quad(q, g_mc, qpos + i * NQ, qvel + i * NV, warm + i * NV, ctrl + 2 * i, g_walls, true, si,
                                        QStage{budget[stage], stage > 0, rec.data()});
                    q.sync();
                    if (!sus) break;
                    nsus[w]++;
                    // scribble over the shared slots: the next stage must rely on the record alone
                    for (int k = w; k < QP_N * 4; k += 4) sh.priv[k] = -7e77;
                    if (w == 0) for (int k = 0; k < QC_N; k++) sh.shr[k] = 3e99;
                    q.sync();
                }
                if (info4 && w == 0) { info4[4 * i] = si.iters; info4[4 * i + 1] = si.ncon_wheel; info4[4 * i + 2] = si.ncon_wall; info4[4 * i + 3] = si.reset; }
            }
        });
    for (auto& t2 : th) t2.join();
    if (nsuspended) *nsuspended = nsus[0];
    return 0;
}

// ---- a whole "warp" (and CTA) of NQ quads on the host: EVERY collective is a barrier over all 4 NQ threads, exactly as
// the device's full-mask shuffles / ballots / __syncthreads_or need every lane to arrive.  Control flow that is not
// uniform across the quads of a warp (a quad that returns before a trailing collective, a loop one quad leaves early)
// therefore hangs here, on the CPU build box, instead of hanging a GPU.
struct WarpBarrier {
    int nthreads; std::atomic<int> count{0}, gen{0};
    void wait() {
        const int g = gen.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) == nthreads - 1) { count.store(0, std::memory_order_relaxed); gen.store(g + 1, std::memory_order_release); }
        else { int spins = 0; while (gen.load(std::memory_order_acquire) == g) if (++spins > 200) std::this_thread::yield(); }
    }
};
struct WarpHostShared {
    WarpBarrier bar; int nq;
    std::vector<double> slot; std::vector<unsigned> bits;
    std::vector<double> priv, shr; double ktab[QK_N];
};
template <int NQ_>
struct WarpQuadHost {
    static constexpr int PS = 4 * NQ_, CS = NQ_;
    WarpHostShared* s; int w, quad, tid;
    double& P(int i) const { return s->priv[(size_t)i * PS + tid]; }
    double& C(int i) const { return s->shr[(size_t)i * CS + quad]; }
    double K(int i) const { return s->ktab[i]; }
    int lane() const { return w; }
    double sum(double v) const {
        s->slot[tid] = v; s->bar.wait();
        const int b = 4 * quad;
        const double a = v + s->slot[b + (w ^ 1)], c = s->slot[b + (w ^ 2)] + s->slot[b + (w ^ 3)];
        s->bar.wait();
        return a + c;
    }
    unsigned ballot(bool p) const {
        s->bits[tid] = p ? 1u : 0u; s->bar.wait();
        const int b = 4 * quad;
        const unsigned m = s->bits[b] | s->bits[b + 1] << 1 | s->bits[b + 2] << 2 | s->bits[b + 3] << 3;
        s->bar.wait();
        return m;
    }
    bool any(bool p) const { return ballot(p) != 0; }
    bool wany(bool p) const {
        s->bits[tid] = p ? 1u : 0u; s->bar.wait();
        unsigned m = 0; for (int i = 0; i < 4 * NQ_; i++) m |= s->bits[i];
        s->bar.wait();
        return m != 0;
    }
    bool cany(bool p) const { return wany(p); }
    void sync() const { s->bar.wait(); }
};

// NQ cars stepped together as one warp / CTA with a staged solve: k1 Newton rounds, then the suspended cars (alone or
// with finished neighbours idling as dead quads) to convergence.  Returns the number of suspensions.
static unsigned g_shadow_mask = 0;      // bit q: quad q of the host "warp" steps a shadowed car (no walls)
extern "C" void hq_set_shadow_mask(unsigned m) { g_shadow_mask = m; }
template <int NQ_>
static int warp_step(double* qpos, double* qvel, double* warm, const double* ctrl, int* info4, int k1) {
    WarpHostShared sh; sh.bar.nthreads = 4 * NQ_; sh.nq = NQ_;
    sh.slot.assign(4 * NQ_, 0); sh.bits.assign(4 * NQ_, 0); sh.priv.assign((size_t)QP_N * 4 * NQ_, 0); sh.shr.assign((size_t)QC_N * NQ_, 0);
    for (int g = 0; g < 25; g++) quad_const_entry(g_mc, g, sh.ktab);
    std::vector<double> recs((size_t)QREC_DOUBLES * NQ_);
    std::vector<int> suspended(NQ_, 0);
    std::vector<std::thread> th;
    for (int tid = 0; tid < 4 * NQ_; tid++)
        th.emplace_back([&, tid]() {
            WarpQuadHost<NQ_> q; q.s = &sh; q.tid = tid; q.w = tid & 3; q.quad = tid >> 2;
            const int i = q.quad;
            StepInfo si;
            QHfWalls g_walls = ::g_walls;                                  // (shadows the global: this quad's walls)
            if (g_shadow_mask >> i & 1u) g_walls.on = false;
            bool sus = step_car_quad(q, g_mc, qpos + i * NQ, qvel + i * NV, warm + i * NV, ctrl + 2 * i, g_walls, true, si,
                                     QStage{k1, false, recs.data() + (size_t)i * QREC_DOUBLES});
            q.sync();
            if (q.w == 0) suspended[i] = sus;
            q.sync();
            // second launch: every quad takes part (the finished ones as dead quads, like the padding of a resume batch)
            bool anysus = false; for (int c = 0; c < NQ_; c++) anysus |= suspended[c] != 0;
            if (anysus) {
                StepInfo s2;
                const bool live = suspended[i] != 0;
                step_car_quad(q, g_mc, qpos + i * NQ, qvel + i * NV, warm + i * NV, ctrl + 2 * i, g_walls, live, s2,
                              QStage{0, true, recs.data() + (size_t)i * QREC_DOUBLES});
                if (live) si = s2;
                q.sync();
            }
            if (info4 && q.w == 0) { info4[4 * i] = si.iters; info4[4 * i + 1] = si.ncon_wheel; info4[4 * i + 2] = si.ncon_wall; info4[4 * i + 3] = si.reset; }
        });
    for (auto& t2 : th) t2.join();
    int ns = 0; for (int c = 0; c < NQ_; c++) ns += suspended[c];
    return ns;
}
extern "C" int hq_step_warp4(double* qpos, double* qvel, double* warm, const double* ctrl, int* info4, int k1) {
    if (!g_ready) { g_mc = model_constants(); g_ready = true; }
    return warp_step<4>(qpos, qvel, warm, ctrl, info4, k1);
}
