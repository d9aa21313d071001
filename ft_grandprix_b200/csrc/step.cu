// step.cu -- vehicle-step kernel and the fused per-tick entry point.
//
// ftgp_step replaces mujoco.mj_step(model, data) (ft_grandprix/custom.py:1425) for a fleet of
// independent cars of template/mushr.em.xml; the arithmetic is the block-arrow Newton solver specialised
// for the car's fixed topology (csrc/mushr_step.cuh), mapped four lanes per car, one per wheel chain
// (csrc/mushr_step_quad.cuh), in fp64; cars never interact on this path, so nothing crosses a quad.
// ftgp_tick runs whole iterations of physics_thread (custom.py:1337-1426) without leaving the
// device: lap update -> built-in driver on last tick's ranges -> ctrl -> rangefinders from the
// pre-step pose -> mj_step (the reference's one-tick sensor lag is kept).
#include "common.h"
#include "mushr_consts.h"
#include "mushr_step_quad.cuh"
#include <cstdlib>
#include <mutex>
#include <utility>

namespace ftgp {
using namespace mushr;

__constant__ ModelConsts c_model;
static bool g_model_ready[16] = {false};
static ModelConsts g_model_host;
static bool g_model_host_ready = false;

static int ensure_model(int device) {
    if (!g_model_host_ready) { g_model_host = model_constants(); g_model_host_ready = true; }
    if (device >= 0 && device < 16 && g_model_ready[device]) return FTGP_OK;
    FTGP_CUDA(cudaMemcpyToSymbol(c_model, &g_model_host, sizeof(ModelConsts)));
    if (device >= 0 && device < 16) g_model_ready[device] = true;
    return FTGP_OK;
}

// chassis-vs-wall contacts (this framework's definition, identical to oracle/step.c wall_contacts()):
// each chassis hull vertex below the hfield surface gives one condim-3 contact against the surface
// triangle's plane.  Reads the compiled track from global memory (L2-resident, ~50 KB).
typedef QWallHit WallHit;     // { dist, nrm[3], t1[3], t2[3], pnt[3] }

__device__ bool wall_probe(const uint32_t* blob, const TrackHeader* th, const double* R1, const double* p1, int v, WallHit& h) {
    const double hull[MUSHR_CHASSIS_NHULL][3] = MUSHR_CHASSIS_HULL;
    const uint16_t* index = reinterpret_cast<const uint16_t*>(blob + th->index_off);
    const uint32_t* chunks = blob + th->chunks_off;
    double p[3];
    mat_vec3(p, R1, hull[v]);
    for (int a = 0; a < 3; a++) p[a] += p1[a];
    const int i = (int)floor(p[0] / th->dsize_x + 0.5), j = (int)floor(-p[1] / th->dsize_y + 0.5);
    if (i < 0 || i >= th->hc || j < 0 || j >= th->vc) return false;
    const uint32_t cid = index[(th->vc - 1 - j) * th->hc + i];
    if (cid == EMPTY_CHUNK) return false;
    const uint32_t* m = chunks + cid * CHUNK_WORDS;
    const int ncol = m[13] & 0xFF, nrow = (m[13] >> 8) & 0xFF;
    const double sx = 0.5 * th->dsize_x, sy = 0.5 * th->dsize_y;
    const double dx = 2 * sx / (ncol - 1), dy = 2 * sy / (nrow - 1);
    const double u = (p[0] - th->dsize_x * i + sx) / dx, vv = (p[1] + th->dsize_y * j + sy) / dy;
    int cc = (int)floor(u), rr = (int)floor(vv);
    cc = cc < 0 ? 0 : (cc > ncol - 2 ? ncol - 2 : cc); rr = rr < 0 ? 0 : (rr > nrow - 2 ? nrow - 2 : rr);
    const double fu = u - cc, fv = vv - rr;
    auto bit = [&](int r_, int c_) { int b = r_ * ncol + c_; return (double)((m[b >> 5] >> (b & 31)) & 1u) * 0.3; };
    const double z00 = bit(rr, cc), z10 = bit(rr, cc + 1), z01 = bit(rr + 1, cc), z11 = bit(rr + 1, cc + 1);
    double gx, gy, z;
    if (fv <= fu) { gx = (z10 - z00) / dx; gy = (z11 - z10) / dy; z = z00 + (z10 - z00) * fu + (z11 - z10) * fv; }
    else { gx = (z11 - z01) / dx; gy = (z01 - z00) / dy; z = z00 + (z11 - z01) * fu + (z01 - z00) * fv; }
    const double nn = sqrt(gx * gx + gy * gy + 1);
    h.nrm[0] = -gx / nn; h.nrm[1] = -gy / nn; h.nrm[2] = 1 / nn;
    const double hh = -0.1 + z;
    if (hh <= -0.1 + 1e-12 && h.nrm[2] > 0.999999) return false;
    h.dist = (p[2] - hh) * h.nrm[2];
    if (h.dist >= 0) return false;
    // frame (mju_makeFrame)
    h.t1[0] = h.t1[1] = h.t1[2] = 0;
    if (h.nrm[1] < 0.5 && h.nrm[1] > -0.5) h.t1[1] = 1; else h.t1[2] = 1;
    const double d = dot3(h.nrm, h.t1);
    for (int a = 0; a < 3; a++) h.t1[a] -= d * h.nrm[a];
    const double tn = sqrt(dot3(h.t1, h.t1));
    for (int a = 0; a < 3; a++) h.t1[a] /= tn;
    cross3(h.t2, h.nrm, h.t1);
    for (int a = 0; a < 3; a++) h.pnt[a] = p[a] - h.nrm[a] * h.dist * 0.5;
    return true;
}

struct Walls {                                  // thread-per-car flavour
    const uint32_t* blob; const TrackHeader* th;
    __device__ void operator()(const ModelConsts& mc, const Kin& k, Rows& r) const {
        if (!blob) return;
        for (int v = 0; v < MUSHR_CHASSIS_NHULL && r.ncon < MAXCON; v++) {
            WallHit h;
            if (!wall_probe(blob, th, k.R1, k.p1, v, h)) continue;
            Contact& c = r.con[r.ncon++];
            c.dist = h.dist; c.mu = 1.0; c.dmin = 0.9; c.wheel = -1; c.tran = mc.chassis_invweight0;
            double off[3];
            for (int a = 0; a < 3; a++) off[a] = h.pnt[a] - k.com[a];
            for (int col = 0; col < 9; col++) {
                double jp[3] = {0, 0, 0};
                if (col < 6) { cross3(jp, k.cdof[col], off); for (int a = 0; a < 3; a++) jp[a] += k.cdof[col][3 + a]; }
                c.J[0][col] = dot3(h.nrm, jp); c.J[1][col] = dot3(h.t1, jp); c.J[2][col] = dot3(h.t2, jp);
            }
        }
    }
};

__global__ void __launch_bounds__(64)
step_kernel(const uint32_t* __restrict__ blob, double* __restrict__ qpos, double* __restrict__ qvel,
            double* __restrict__ warm, const double* __restrict__ ctrl, const int32_t* __restrict__ track_id,
            const int32_t* __restrict__ lap, int64_t ncars, int nsteps, int32_t* __restrict__ status) {
    const int64_t car = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (car >= ncars) return;
    double q[NQ], v[NV], w[NV], u[2];
    for (int i = 0; i < NQ; i++) q[i] = qpos[car * NQ + i];
    for (int i = 0; i < NV; i++) { v[i] = qvel[car * NV + i]; w[i] = warm[car * NV + i]; }
    u[0] = ctrl[2 * car]; u[1] = ctrl[2 * car + 1];
    Walls walls{nullptr, nullptr};
    // a finished ("shadowed") car no longer collides with walls: conaffinity 0 / contype 2 (custom.py:1455-1464)
    const bool shadowed = lap && lap[car * FTGP_LAP_FIELDS + FTGP_LAP_FINISHED];
    if (blob && !shadowed) {
        const GeomHeader* gh = reinterpret_cast<const GeomHeader*>(blob);
        int tid = track_id ? track_id[car] : 0;
        if (tid < 0 || tid >= gh->ntracks) tid = 0;
        walls.blob = blob; walls.th = reinterpret_cast<const TrackHeader*>(blob + gh->track_off[tid]);
    }
    int st = 0;
    for (int s = 0; s < nsteps; s++) {
        StepInfo info;
        step_car(c_model, q, v, w, u, walls, info);
        st = (info.iters & 0xFF) | (info.reset ? 0x100 : (st & 0x100)) | ((info.ncon_wall & 0xFF) << 16) | ((info.ncon_wheel & 0xF) << 24);
    }
    for (int i = 0; i < NQ; i++) qpos[car * NQ + i] = q[i];
    for (int i = 0; i < NV; i++) { qvel[car * NV + i] = v[i]; warm[car * NV + i] = w[i]; }
    if (status) status[car] = st;
}

struct WallsQuad {                              // quad-per-car flavour: probe of one hull vertex
    const uint32_t* blob; const TrackHeader* th;
    __device__ __forceinline__ bool enabled() const { return blob != nullptr; }
    __device__ __forceinline__ bool operator()(const double* R1, const double* p1, int v, QWallHit& h) const { return wall_probe(blob, th, R1, p1, v, h); }
};

// Quad-per-car: four lanes (one per wheel chain) advance one car, 8 cars per warp; see mushr_step_quad.cuh.
// Shared memory: [slot][thread] for the lane-private slots, [slot][car] for the per-car slots, then the table of
// friction-loss row constants.  1 072 B per lane: one 216-thread CTA (54 cars) per SM.
template <int NT> constexpr size_t quad_smem_bytes() { return (size_t)(NT * QP_N + NT / 4 * QC_N + QK_N + 1) * sizeof(double); }

template <int NT, bool LOCK>
__global__ void __launch_bounds__(NT, 1)
step_quad_kernel(const uint32_t* __restrict__ blob, double* __restrict__ qpos, double* __restrict__ qvel,
                 double* __restrict__ warm, const double* __restrict__ ctrl, const int32_t* __restrict__ track_id,
                 const int32_t* __restrict__ lap, const int32_t* __restrict__ perm, int64_t ncars, int nsteps,
                 int32_t* __restrict__ status, double* __restrict__ recs, int32_t* __restrict__ list_out,
                 int32_t* __restrict__ count_out, int max_rounds) {
    const int tid = threadIdx.x, cib = tid >> 2;
    constexpr int KO = NT * QP_N + NT / 4 * QC_N;
    for (int g = tid; g < 25; g += NT) quad_const_entry(c_model, g, quad_sm + KO);
    __syncthreads();
    int64_t car = (int64_t)blockIdx.x * (NT / 4) + cib;
    const bool live = car < ncars;
    if (!live) car = ncars - 1;                     // padding quad: same collectives, no stores
    if (perm) car = perm[car];                      // cars grouped by their last Newton iteration count
    QuadDev<NT, NT / 4, LOCK> qd;
    qd.w = tid & 3; qd.po = tid; qd.co = NT * QP_N + cib; qd.ko = KO; qd.qs = tid & 28;
    WallsQuad walls{nullptr, nullptr};
    const bool shadowed = lap && lap[car * FTGP_LAP_FIELDS + FTGP_LAP_FINISHED];
    if (blob && !shadowed) {
        const GeomHeader* gh = reinterpret_cast<const GeomHeader*>(blob);
        int tid_ = track_id ? track_id[car] : 0;
        if (tid_ < 0 || tid_ >= gh->ntracks) tid_ = 0;
        walls.blob = blob; walls.th = reinterpret_cast<const TrackHeader*>(blob + gh->track_off[tid_]);
    }
    int st = 0;
    for (int s = 0; s < nsteps; s++) {
        StepInfo info;
        // staged solve (recs != NULL, one step per launch): a car that is not done after max_rounds Newton rounds of its
        // CTA is parked in its record and listed for the continuation kernel below
        const QStage stage{recs ? max_rounds : 0, false, recs ? recs + car * QREC_DOUBLES : nullptr};
        const bool suspended = step_car_quad(qd, c_model, qpos + car * NQ, qvel + car * NV, warm + car * NV, ctrl + 2 * car, walls, live, info, stage);
        if (suspended) {
            if (live && qd.w == 0) list_out[atomicAdd(count_out, 1)] = (int32_t)car;
            return;
        }
        st = (info.iters & 0xFF) | (info.reset ? 0x100 : (st & 0x100)) | ((info.ncon_wall & 0xFF) << 16) | ((info.ncon_wheel & 0xF) << 24);
    }
    if (status && qd.w == 0 && live) status[car] = st;
}

// Continuation of the staged solve: a persistent grid (one CTA per SM) packs the suspended cars of the whole fleet,
// 54 at a time, restores their solver state from the records and goes on for max_rounds more Newton rounds
// (<= 0: to convergence); cars that are still not done are listed again.
template <int NT>
__global__ void __launch_bounds__(NT, 1)
step_quad_resume_kernel(double* __restrict__ qpos, double* __restrict__ qvel, double* __restrict__ warm,
                        const double* __restrict__ ctrl, int32_t* __restrict__ status, double* __restrict__ recs,
                        const int32_t* __restrict__ list_in, const int32_t* __restrict__ count_in,
                        int32_t* __restrict__ list_out, int32_t* __restrict__ count_out, int max_rounds) {
    const int tid = threadIdx.x, cib = tid >> 2;
    constexpr int KO = NT * QP_N + NT / 4 * QC_N, CARS = NT / 4;
    for (int g = tid; g < 25; g += NT) quad_const_entry(c_model, g, quad_sm + KO);
    __syncthreads();
    const int n = *count_in;
    QuadDev<NT, NT / 4, true> qd;
    qd.w = tid & 3; qd.po = tid; qd.co = NT * QP_N + cib; qd.ko = KO; qd.qs = tid & 28;
    const WallsQuad walls{nullptr, nullptr};             // (the position stage, which probes the walls, is behind us)
    for (int base = blockIdx.x * CARS; base < n; base += gridDim.x * CARS) {
        const int e = base + cib;
        const bool live = e < n;
        const int64_t car = list_in[live ? e : n - 1];
        StepInfo info;
        const QStage stage{max_rounds, true, recs + car * QREC_DOUBLES};
        const bool suspended = step_car_quad(qd, c_model, qpos + car * NQ, qvel + car * NV, warm + car * NV, ctrl + 2 * car, walls, live, info, stage);
        if (live && qd.w == 0) {
            if (suspended) list_out[atomicAdd(count_out, 1)] = (int32_t)car;
            else if (status) status[car] = (info.iters & 0xFF) | (info.reset ? 0x100 : 0) | ((info.ncon_wall & 0xFF) << 16) | ((info.ncon_wheel & 0xF) << 24);
        }
        __syncthreads();                                 // the next batch reuses the shared-memory slots
    }
}

// The quad-per-car kernel runs the cars of a CTA in lock-step, so a CTA takes as many Newton rounds as its slowest
// car.  The iteration count is strongly correlated from one step to the next (measured: mean 2, max over 8 random
// cars 3.6), so cars are grouped by (last iteration count, in wall contact or not) with a counting sort.
constexpr int NBIN = 16;
__device__ __forceinline__ int order_bin(int st) { return min(st & 0xFF, 7) + (((st >> 16) & 0xFF) ? 8 : 0); }
__global__ void order_hist_kernel(const int32_t* __restrict__ status, int64_t ncars, int32_t* __restrict__ hist) {
    __shared__ int h[NBIN];
    if (threadIdx.x < NBIN) h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ncars) atomicAdd(&h[order_bin(status[i])], 1);
    __syncthreads();
    if (threadIdx.x < NBIN && h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}
__global__ void order_scatter_kernel(const int32_t* __restrict__ status, int64_t ncars, const int32_t* __restrict__ hist,
                                     int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
    __shared__ int h[NBIN], base[NBIN];
    if (threadIdx.x < NBIN) h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int bin = 0, rank = 0;
    if (i < ncars) { bin = order_bin(status[i]); rank = atomicAdd(&h[bin], 1); }
    __syncthreads();
    if (threadIdx.x < NBIN) {
        int start = 0;
        for (int b = 0; b < (int)threadIdx.x; b++) start += hist[b];
        base[threadIdx.x] = start + (h[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], h[threadIdx.x]) : 0);
    }
    __syncthreads();
    if (i < ncars) perm[base[bin] + rank] = (int32_t)i;
}
// one scratch per (device, stream): fleets stepped concurrently on different streams must not share it
struct OrderScratch {
    int dev = -1; cudaStream_t stream = nullptr; int32_t* perm = nullptr; int32_t* counters = nullptr; int64_t cap = 0; uint64_t used = 0;
    double* recs = nullptr; int32_t* lists = nullptr; int32_t* stage_counts = nullptr; int64_t stage_cap = 0;      // staged solve
};
static OrderScratch g_order[64];
static std::mutex g_order_mutex;
static uint64_t g_order_clock = 0;
static OrderScratch* order_scratch(int dev, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(g_order_mutex);
    OrderScratch* lru = &g_order[0];
    for (auto& o : g_order) {
        if (o.dev == dev && o.stream == stream) { o.used = ++g_order_clock; return &o; }
        if (o.used < lru->used) lru = &o;
    }
    // not found: take a free slot, else recycle the least recently used one (cudaFree waits for work that uses it)
    if (lru->dev >= 0) {
        int cur = 0;
        cudaGetDevice(&cur); cudaSetDevice(lru->dev);
        if (lru->perm) cudaFree(lru->perm);
        if (lru->counters) cudaFree(lru->counters);
        if (lru->recs) cudaFree(lru->recs);
        if (lru->lists) cudaFree(lru->lists);
        if (lru->stage_counts) cudaFree(lru->stage_counts);
        cudaSetDevice(cur);
    }
    *lru = OrderScratch();
    lru->dev = dev; lru->stream = stream; lru->used = ++g_order_clock;
    return lru;
}

// cars grouped by (last Newton iteration count, wall contact) for the kernels that run several cars in lock-step
static int order_cars(const int32_t* status, int64_t ncars, int dev, cudaStream_t stream, const int32_t** perm) {
    *perm = nullptr;
    static int use_order = -1;
    if (use_order < 0) { const char* e = getenv("FTGP_STEP_ORDER"); use_order = (e && e[0] == '0') ? 0 : 1; }
    if (!use_order || !status || ncars < 1024 || ncars >= (int64_t)1 << 31 || dev >= 16) return FTGP_OK;
    OrderScratch* op = order_scratch(dev, stream);
    if (!op) return FTGP_OK;
    OrderScratch& o = *op;
    if (o.cap < ncars) {
        if (o.perm) cudaFree(o.perm);
        if (!o.counters) FTGP_CUDA(cudaMalloc(&o.counters, 2 * NBIN * sizeof(int32_t)));
        o.perm = nullptr; o.cap = 0;
        FTGP_CUDA(cudaMalloc(&o.perm, ncars * sizeof(int32_t)));
        o.cap = ncars;
    }
    FTGP_CUDA(cudaMemsetAsync(o.counters, 0, 2 * NBIN * sizeof(int32_t), stream));
    const unsigned nb = (unsigned)((ncars + 255) / 256);
    order_hist_kernel<<<nb, 256, 0, stream>>>(status, ncars, o.counters);
    order_scatter_kernel<<<nb, 256, 0, stream>>>(status, ncars, o.counters, o.counters + NBIN, o.perm);
    count_launch(2);
    *perm = o.perm;
    return FTGP_OK;
}

int launch_step(const ftgp_geom* g, double* qpos, double* qvel, double* warm, const double* ctrl,
                const int32_t* track_id, const int32_t* lap, int64_t ncars, int nsteps, int32_t* status,
                cudaStream_t stream) {
    int dev = 0;
    FTGP_CUDA(cudaGetDevice(&dev));
    int rc = ensure_model(dev); if (rc) return rc;
    // Two implementations of the same arithmetic (both parity-tested), FTGP_STEP_IMPL = quad (default) | thread:
    //   quad   four lanes per car, one per wheel chain (mushr_step_quad.cuh): 1.47 ms per 65,536 cars
    //   thread one thread per car (mushr_step.cuh): 16.7 KB local frame per thread, DRAM-latency bound, 5.3 ms; kept as the
    //          A/B reference (it is also the source the CPU tests compile for the host)
    // (a third mapping, one warp per car with the state in shared memory, measured 5.2 ms and was removed: DESIGN.md 3.2)
    static int impl = -1;
    if (impl < 0) { const char* e = getenv("FTGP_STEP_IMPL"); impl = (e && e[0] == 't') ? 1 : 2; }
    const uint32_t* blob = g ? g->d_blob : nullptr;
    if (impl == 2) {
        static int qt = 0, lock = 1;
        if (!qt) {
            const char* e = getenv("FTGP_STEP_QT"); qt = e ? atoi(e) : 216; if (qt != 128 && qt != 192 && qt != 216) qt = 216;
            const char* l = getenv("FTGP_STEP_LOCK"); lock = (l && l[0] == '0') ? 0 : 1;
        }
        const int32_t* perm = nullptr;
        if ((rc = order_cars(status, ncars, dev, stream, &perm))) return rc;
        // Staged solve (default for one step of a big fleet): K1 Newton rounds in the first launch, then the cars that are
        // not done yet are packed into fresh CTAs: (optionally K2 more rounds, then) to convergence.  FTGP_STEP_K1=0
        // switches it off.  Measured at 65,536 cars: 1.39 ms unstaged, 1.32 ms with K1 = 2 (the records cost 5.9 KB per car;
        // if they cannot be allocated the step runs unstaged).
        static int k1 = -1, k2 = 1;
        if (k1 < 0) {
            const char* e = getenv("FTGP_STEP_K1"); k1 = e ? atoi(e) : 2;
            const char* f = getenv("FTGP_STEP_K2"); k2 = f ? atoi(f) : 0;
        }
        double* recs = nullptr; int32_t* lists = nullptr; int32_t* counts = nullptr;
        if (k1 > 0 && nsteps == 1 && ncars >= 4096 && ncars < (int64_t)1 << 31 && qt == 216) {
            OrderScratch* o = order_scratch(dev, stream);
            if (o->stage_cap < ncars) {
                if (o->recs) cudaFree(o->recs);
                if (o->lists) cudaFree(o->lists);
                o->recs = nullptr; o->lists = nullptr; o->stage_cap = 0;
                if (!o->stage_counts) FTGP_CUDA(cudaMalloc(&o->stage_counts, 4 * sizeof(int32_t)));
                if (cudaMalloc(&o->recs, (size_t)ncars * QREC_DOUBLES * sizeof(double)) == cudaSuccess &&
                    cudaMalloc(&o->lists, (size_t)ncars * 2 * sizeof(int32_t)) == cudaSuccess) o->stage_cap = ncars;
                else {                                    // no room for the records: run unstaged
                    cudaGetLastError();
                    if (o->recs) cudaFree(o->recs);
                    o->recs = nullptr; o->lists = nullptr;
                }
            }
            if (o->stage_cap >= ncars) {
                recs = o->recs; lists = o->lists; counts = o->stage_counts;
                FTGP_CUDA(cudaMemsetAsync(counts, 0, 4 * sizeof(int32_t), stream));
            }
        }
        auto launch = [&](auto kern, size_t smem) -> int {
            FTGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            FTGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            kern<<<(unsigned)((ncars + qt / 4 - 1) / (qt / 4)), qt, smem, stream>>>(
                blob, qpos, qvel, warm, ctrl, track_id, lap, perm, ncars, nsteps, status, recs, lists, counts, k1);
            return FTGP_OK;
        };
        int rc2;
        if (qt == 128) rc2 = launch(step_quad_kernel<128, true>, quad_smem_bytes<128>());
        else if (qt == 216) rc2 = launch(step_quad_kernel<216, true>, quad_smem_bytes<216>());   // 54 cars: 6 warps + 24 lanes
        else rc2 = lock ? launch(step_quad_kernel<192, true>, quad_smem_bytes<192>()) : launch(step_quad_kernel<192, false>, quad_smem_bytes<192>());
        if (rc2) return rc2;
        if (recs) {
            static int nsm[16] = {0};
            if (dev < 16 && !nsm[dev]) FTGP_CUDA(cudaDeviceGetAttribute(&nsm[dev], cudaDevAttrMultiProcessorCount, dev));
            const int grid = dev < 16 ? nsm[dev] : 148;
            const size_t smem = quad_smem_bytes<216>();
            FTGP_CUDA(cudaFuncSetAttribute(step_quad_resume_kernel<216>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            FTGP_CUDA(cudaFuncSetAttribute(step_quad_resume_kernel<216>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            int32_t* la = lists; int32_t* lb = lists + ncars;
            int ci = 0;
            if (k2 > 0) {
                step_quad_resume_kernel<216><<<grid, 216, smem, stream>>>(qpos, qvel, warm, ctrl, status, recs, la, counts + ci, lb, counts + ci + 1, k2);
                count_launch();
                std::swap(la, lb); ci++;
            }
            step_quad_resume_kernel<216><<<grid, 216, smem, stream>>>(qpos, qvel, warm, ctrl, status, recs, la, counts + ci, lb, counts + ci + 1, 0);
            count_launch();
        }
    } else if (impl == 1) {
        static int threads = 0;
        if (!threads) { const char* e = getenv("FTGP_STEP_BLOCK"); threads = e ? atoi(e) : 64; if (threads < 32 || threads > 64) threads = 64; }
        step_kernel<<<(unsigned)((ncars + threads - 1) / threads), threads, 0, stream>>>(
            blob, qpos, qvel, warm, ctrl, track_id, lap, ncars, nsteps, status);
    }
    count_launch();
    FTGP_CUDA(cudaGetLastError());
    return FTGP_OK;
}

int launch_lidar(const ftgp_geom* g, const double* qpos, int64_t stride, const int32_t* track_id,
                 const uint8_t* visible, const int32_t* lap, int64_t ncars, int cpw, float* ranges, float* min_range,
                 cudaStream_t stream);
int launch_drivers(const float* ranges, const int32_t* kind, int default_kind, const int32_t* lap, double* ctrl,
                   int64_t ncars, cudaStream_t stream);
int launch_lap(const ftgp_geom* g, const double* qpos, int64_t stride, const int32_t* track_id, int32_t* lap,
               int32_t* times, int32_t* winners, const int32_t* status, int64_t ncars, int cpw, int32_t steps,
               int32_t lap_target, cudaStream_t stream);

}  // namespace ftgp
using namespace ftgp;

extern "C" int ftgp_step(const ftgp_geom* g, double* qpos, double* qvel, double* warm, const double* ctrl,
                         const int32_t* track_id, int64_t ncars, int nsteps, int32_t* status, void* stream) {
    if (!qpos || !qvel || !warm || !ctrl || ncars < 0 || nsteps < 0) { set_error("ftgp_step: bad argument"); return FTGP_ERR_ARG; }
    if (ncars == 0 || nsteps == 0) return FTGP_OK;
    if (g) FTGP_CUDA(cudaSetDevice(g->device));
    return launch_step(g, qpos, qvel, warm, ctrl, track_id, nullptr, ncars, nsteps, status, (cudaStream_t)stream);
}

extern "C" int ftgp_tick(const ftgp_tick_args* a, int nticks, void* stream) {
    if (!a || !a->geom || !a->qpos || !a->qvel || !a->warm || !a->ctrl || !a->ranges || !a->lap || !a->times ||
        a->ncars < 0 || a->cars_per_world < 1 || nticks < 0) { set_error("ftgp_tick: bad argument"); return FTGP_ERR_ARG; }
    if (a->ncars == 0) return FTGP_OK;
    FTGP_CUDA(cudaSetDevice(a->geom->device));
    cudaStream_t s = (cudaStream_t)stream;
    for (int t = 0; t < nticks; t++) {
        int rc;
        // custom.py:1340-1372 progress + lap logic from the current pose
        if ((rc = launch_lap(a->geom, a->qpos, FTGP_NQ, a->track_id, a->lap, a->times, a->winners, a->status, a->ncars,
                             a->cars_per_world, a->steps + t, a->lap_target, s))) return rc;
        // custom.py:1395-1423 driver on the ranges of the previous mj_step, control write
        if ((rc = launch_drivers(a->ranges, a->driver_kind, a->default_driver, a->lap, a->ctrl, a->ncars, s))) return rc;
        // custom.py:1425 mj_step: rangefinders are evaluated from the pre-step pose, then the state advances
        if ((rc = launch_lidar(a->geom, a->qpos, FTGP_NQ, a->track_id, nullptr, a->lap, a->ncars, a->cars_per_world, a->ranges,
                               nullptr, s))) return rc;
        if ((rc = launch_step(a->geom, a->qpos, a->qvel, a->warm, a->ctrl, a->track_id, a->lap, a->ncars, 1, a->status, s))) return rc;
    }
    return FTGP_OK;
}
