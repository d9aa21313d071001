// step.cu -- vehicle-step kernel and the fused per-tick entry point.
//
// ftgp_step replaces mujoco.mj_step(model, data) (ft_grandprix/custom.py:1425) for a fleet of
// independent cars of template/mushr.em.xml; the arithmetic is csrc/mushr_step.cuh (block-arrow
// Newton solver specialised for the car's fixed topology).  One thread advances one car in fp64;
// cars never interact on this path, so there is no inter-thread communication at all.
// ftgp_tick runs whole iterations of physics_thread (custom.py:1337-1426) without leaving the
// device: lap update -> built-in driver on last tick's ranges -> ctrl -> rangefinders from the
// pre-step pose -> mj_step (the reference's one-tick sensor lag is kept).
#include "common.h"
#include "mushr_consts.h"

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
struct Walls {
    const uint32_t* blob; const TrackHeader* th;
    __device__ void operator()(const ModelConsts& mc, const Kin& k, Rows& r) const {
        if (!blob) return;
        const double hull[MUSHR_CHASSIS_NHULL][3] = MUSHR_CHASSIS_HULL;
        const uint16_t* index = reinterpret_cast<const uint16_t*>(blob + th->index_off);
        const uint32_t* chunks = blob + th->chunks_off;
        for (int v = 0; v < MUSHR_CHASSIS_NHULL && r.ncon < MAXCON; v++) {
            double p[3];
            mat_vec3(p, k.R1, hull[v]);
            for (int a = 0; a < 3; a++) p[a] += k.p1[a];
            const int i = (int)floor(p[0] / th->dsize_x + 0.5), j = (int)floor(-p[1] / th->dsize_y + 0.5);
            if (i < 0 || i >= th->hc || j < 0 || j >= th->vc) continue;
            const uint32_t cid = index[(th->vc - 1 - j) * th->hc + i];
            if (cid == EMPTY_CHUNK) continue;
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
            double nrm[3] = {-gx / nn, -gy / nn, 1 / nn};
            const double h = -0.1 + z;
            if (h <= -0.1 + 1e-12 && nrm[2] > 0.999999) continue;
            const double dist = (p[2] - h) * nrm[2];
            if (dist >= 0) continue;
            Contact& c = r.con[r.ncon++];
            c.dist = dist; c.mu = 1.0; c.dmin = 0.9; c.wheel = -1; c.tran = mc.chassis_invweight0;
            // frame (mju_makeFrame)
            double t1[3] = {0, 0, 0};
            if (nrm[1] < 0.5 && nrm[1] > -0.5) t1[1] = 1; else t1[2] = 1;
            const double d = dot3(nrm, t1);
            for (int a = 0; a < 3; a++) t1[a] -= d * nrm[a];
            const double tn = sqrt(dot3(t1, t1));
            for (int a = 0; a < 3; a++) t1[a] /= tn;
            double t2[3];
            cross3(t2, nrm, t1);
            double off[3];
            for (int a = 0; a < 3; a++) off[a] = p[a] - nrm[a] * dist * 0.5 - k.com[a];
            for (int col = 0; col < 9; col++) {
                double jp[3] = {0, 0, 0};
                if (col < 6) { cross3(jp, k.cdof[col], off); for (int a = 0; a < 3; a++) jp[a] += k.cdof[col][3 + a]; }
                c.J[0][col] = dot3(nrm, jp); c.J[1][col] = dot3(t1, jp); c.J[2][col] = dot3(t2, jp);
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

int launch_step(const ftgp_geom* g, double* qpos, double* qvel, double* warm, const double* ctrl,
                const int32_t* track_id, const int32_t* lap, int64_t ncars, int nsteps, int32_t* status,
                cudaStream_t stream) {
    int dev = 0;
    FTGP_CUDA(cudaGetDevice(&dev));
    int rc = ensure_model(dev); if (rc) return rc;
    const int threads = 64;
    step_kernel<<<(unsigned)((ncars + threads - 1) / threads), threads, 0, stream>>>(
        g ? g->d_blob : nullptr, qpos, qvel, warm, ctrl, track_id, lap, ncars, nsteps, status);
    count_launch();
    FTGP_CUDA(cudaGetLastError());
    return FTGP_OK;
}

int launch_lidar(const ftgp_geom* g, const double* qpos, int64_t stride, const int32_t* track_id,
                 const uint8_t* visible, int64_t ncars, int cpw, float* ranges, float* min_range, cudaStream_t stream);
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
    if (a->cars_per_world != 1) { set_error("ftgp_tick: multi-car worlds need car-car contacts (not built yet)"); return FTGP_ERR_UNSUPPORTED; }
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
        if ((rc = launch_lidar(a->geom, a->qpos, FTGP_NQ, a->track_id, nullptr, a->ncars, a->cars_per_world, a->ranges,
                               nullptr, s))) return rc;
        if ((rc = launch_step(a->geom, a->qpos, a->qvel, a->warm, a->ctrl, a->track_id, a->lap, a->ncars, 1, a->status, s))) return rc;
    }
    return FTGP_OK;
}
