// fleet.cu -- the per-car, per-tick kernels around the two heavy ones (lidar.cu, step.cu):
//
//   reset_kernel    mj_resetData + position_vehicles      ft_grandprix/custom.py:1092,1232-1245
//   drivers_kernel  the bundled disparity-extender drivers ft_grandprix/nidc.py:116-131,
//                                                          ft_grandprix/fast.py:118-139,
//                                                          ft_grandprix/lobotomy.py:1-3
//                   + the control write                    ft_grandprix/custom.py:1418-1423
//   lap_kernel      progress / lap state machine           ft_grandprix/custom.py:1340-1372
//
// Layout: one warp per car for the driver (lane l owns beams l, l+32, l+64 of the 68-beam
// front window, so the ranges row is read with coalesced 128-byte loads and the sequential
// "extend disparities" loop runs as warp-uniform control flow with shuffles); one thread per
// world for the lap logic (cars of one world are ranked in car order, exactly as the
// reference's `for vehicle_state in self.vehicle_states` loop does).
#include <math.h>
#include "common.h"

namespace ftgp {

// ------------------------------------------------------------------ reset
__global__ void reset_kernel(double* __restrict__ qpos, double* __restrict__ qvel, double* __restrict__ warm,
                             double* __restrict__ ctrl, const double* __restrict__ xy,
                             const double* __restrict__ yaw, int64_t ncars) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncars) return;
    double* q = qpos + i * FTGP_NQ;
    // qpos0: free joint at the body pos (0, 2, 0) with identity quat, hinges/slides 0, balls identity
    // (mushr.em.xml:96); position_vehicles then overwrites x, y and the quaternion (custom.py:1244-1245)
    for (int k = 0; k < FTGP_NQ; k++) q[k] = 0.0;
    q[0] = xy[2 * i]; q[1] = xy[2 * i + 1]; q[2] = 0.0;
    // euler_to_quaternion([yaw, 0, 0]) (custom.py:81-87): cr = cp = 1, sr = sp = 0
    const double h = yaw[i] * 0.5;
    q[3] = cos(h); q[4] = 0.0; q[5] = 0.0; q[6] = sin(h);
    q[11] = 1.0; q[18] = 1.0; q[24] = 1.0; q[30] = 1.0;          // ball joints of the four softener bodies
    for (int k = 0; k < FTGP_NV; k++) { qvel[i * FTGP_NV + k] = 0.0; warm[i * FTGP_NV + k] = 0.0; }
    if (ctrl) { ctrl[2 * i] = 0.0; ctrl[2 * i + 1] = 0.0; }
}

// ------------------------------------------------------------------ option naive_flatten (custom.py:981,1338-1339)
// qpos[3:7] = euler_to_quaternion([quaternion_to_angle(*qpos[3:7]), 0, 0]) (custom.py:62-87): keep the yaw, drop pitch / roll
__global__ void flatten_kernel(double* __restrict__ qpos, int64_t stride, int64_t ncars) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncars) return;
    double* q = qpos + i * stride + 3;
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double yaw = atan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z));
    q[0] = cos(yaw / 2); q[1] = 0.0; q[2] = 0.0; q[3] = sin(yaw / 2);
}
int launch_flatten(double* qpos, int64_t stride, int64_t ncars, cudaStream_t stream) {
    flatten_kernel<<<(unsigned)((ncars + 127) / 128), 128, 0, stream>>>(qpos, stride, ncars);
    count_launch();
    FTGP_CUDA(cudaGetLastError());
    return FTGP_OK;
}

// ------------------------------------------------------------------ drivers
constexpr int NPROC = 68;             // 90 - 2 * int(90 / 8)          nidc.py:17-18
constexpr int EIGHTH = 11;

__device__ __forceinline__ double pick(double e0, double e1, double e2, int idx) {
    // value of element idx of the warp-distributed array (uniform idx)
    int slot = idx >> 5;
    double v = slot == 0 ? e0 : (slot == 1 ? e1 : e2);
    return __shfl_sync(0xffffffffu, v, idx & 31);
}

__global__ void __launch_bounds__(256)
drivers_kernel(const float* __restrict__ ranges, const int32_t* __restrict__ kind, int default_kind,
               const int32_t* __restrict__ lap, double* __restrict__ ctrl, int64_t ncars, int32_t* __restrict__ steps_dev) {
    const int lane = threadIdx.x & 31;
    const int64_t car = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    // graph-replayed tick: the lap kernel (earlier on this stream) has read the tick counter; nothing else reads it this tick
    if (steps_dev && blockIdx.x == 0 && threadIdx.x == 0) *steps_dev += 1;
    if (car >= ncars) return;
    int k = kind ? kind[car] : default_kind;
    if (lap && lap[car * FTGP_LAP_FIELDS + FTGP_LAP_FINISHED]) k = FTGP_DRIVER_LOBOTOMY;   // shadow(): custom.py:1437
    double* out = ctrl + 2 * car;
    if (k == FTGP_DRIVER_LOBOTOMY) { if (lane == 0) { out[0] = 0.0; out[1] = 0.0; } return; }
    const float* row = ranges + car * FTGP_NBEAMS;
    const double PI = 3.141592653589793;
    const double rpp = (2 * PI) / FTGP_NBEAMS;                  // nidc.py:121
    // proc = ranges[11:79]
    double e0 = (double)row[EIGHTH + lane];
    double e1 = (double)row[EIGHTH + 32 + lane];
    double e2 = lane < NPROC - 64 ? (double)row[EIGHTH + 64 + lane] : 0.0;
    const float r0 = row[0];
    // disparities from the unmodified array (nidc.py:123-124): |proc[i] - proc[i-1]| > 0.6, i >= 1
    double p0 = __shfl_up_sync(0xffffffffu, e0, 1);
    double l0 = __shfl_sync(0xffffffffu, e0, 31), l1 = __shfl_sync(0xffffffffu, e1, 31);
    double p1 = __shfl_up_sync(0xffffffffu, e1, 1); if (lane == 0) p1 = l0;
    double p2 = __shfl_up_sync(0xffffffffu, e2, 1); if (lane == 0) p2 = l1;
    unsigned m0 = __ballot_sync(0xffffffffu, lane > 0 && fabs(e0 - p0) > 0.6);
    unsigned m1 = __ballot_sync(0xffffffffu, fabs(e1 - p1) > 0.6);
    unsigned m2 = __ballot_sync(0xffffffffu, lane < NPROC - 64 && fabs(e2 - p2) > 0.6);
    const double car_width = k == FTGP_DRIVER_NIDC ? 0.12 : 0.06;     // nidc.py:5, fast.py:4
    const double width = (car_width / 2) * (1 + 300. / 100);          // nidc.py:93
    bool raised = false;
    for (int seg = 0; seg < 3 && !raised; seg++) {
        unsigned m = seg == 0 ? m0 : (seg == 1 ? m1 : m2);
        while (m) {
            int i = 32 * seg + __ffs(m) - 1; m &= m - 1;
            int first = i - 1;
            double a = pick(e0, e1, e2, first), b = pick(e0, e1, e2, first + 1);
            // np.argmin / np.argmax over two elements: first NaN wins, ties -> index 0
            int amin = isnan(a) ? 0 : (isnan(b) ? 1 : (b < a ? 1 : 0));
            int amax = isnan(a) ? 0 : (isnan(b) ? 1 : (b > a ? 1 : 0));
            int close = first + amin, far = first + amax;
            double dist = amin ? b : a;
            double angle = 2 * atan(width / (2 * dist));               // nidc.py:57
            double np_ = ceil(angle / rpp);                            // nidc.py:58
            if (isnan(np_)) { raised = true; break; }                  // int(nan) -> ValueError -> ctrl kept
            long num = (long)np_;
            // cover_points (nidc.py:71-84): indices close+1 .. close+num (right) or close-1 .. close-num (left)
            long lo, hi;
            if (close < far) { lo = close + 1; hi = close + num; } else { lo = close - num; hi = close - 1; }
            int j0 = lane, j1 = lane + 32, j2 = lane + 64;
            if (j0 >= lo && j0 <= hi && e0 > dist) e0 = dist;
            if (j1 >= lo && j1 <= hi && e1 > dist) e1 = dist;
            if (j2 >= lo && j2 <= hi && j2 < NPROC && e2 > dist) e2 = dist;
        }
    }
    if (raised) return;                                               // custom.py:1409-1411
    // np.argmax: first maximum, first NaN wins
    double bv = e0; int bi = lane;
    bool bn = isnan(e0);
    if (!bn) {
        if (isnan(e1)) { bv = e1; bi = lane + 32; bn = true; }
        else if (e1 > bv) { bv = e1; bi = lane + 32; }
    }
    if (!bn && lane < NPROC - 64) {
        if (isnan(e2)) { bv = e2; bi = lane + 64; bn = true; }
        else if (e2 > bv) { bv = e2; bi = lane + 64; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, bv, off);
        int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        bool on = isnan(ov);
        bool take;
        if (bn || on) take = on && (!bn || oi < bi);
        else take = ov > bv || (ov == bv && oi < bi);
        if (take) { bv = ov; bi = oi; bn = on; }
    }
    if (lane == 0) {
        double ang = ((double)bi - (NPROC / 2.0)) * rpp;              // nidc.py:112
        const double lim = 90.0 * (PI / 180.0);
        double st = ang < -lim ? -lim : (ang > lim ? lim : ang);
        double sp;
        if (k == FTGP_DRIVER_NIDC) sp = 0.5 * 5 * (1 - fabs(st) / (1.57 * 2));      // nidc.py:130
        else if (fabs(st) < 0.1 && (double)r0 > 0.5) sp = 7;                          // fast.py:135-136
        else { double s = 0.5 * 5 * (1 - fabs(st) / PI); sp = s < 2 ? s : 2; }      // fast.py:138
        out[0] = sp; out[1] = st;                                      // custom.py:1422-1423
    }
}

// ------------------------------------------------------------------ lap logic
__device__ __forceinline__ int pymod(int a, int b) { int r = a % b; return r < 0 ? r + b : r; }

__global__ void lap_kernel(const uint32_t* __restrict__ blob, const double* __restrict__ qpos, int64_t stride,
                           const int32_t* __restrict__ track_id, int32_t* __restrict__ lap,
                           int32_t* __restrict__ times, int32_t* __restrict__ winners,
                           const int32_t* __restrict__ status, int64_t ncars, int cpw, int32_t steps,
                           int32_t lap_target, const int32_t* __restrict__ steps_dev) {
    if (steps_dev) steps = *steps_dev;              // graph-replayed tick: self.steps lives on the device
    const int64_t world = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t first = world * cpw;
    if (first >= ncars) return;
    const GeomHeader* gh = reinterpret_cast<const GeomHeader*>(blob);
    int32_t nwin = winners ? winners[world] : 0;
    for (int64_t car = first; car < first + cpw && car < ncars; car++) {
        int tid = track_id ? track_id[car] : 0;
        if (tid < 0 || tid >= gh->ntracks) tid = 0;
        const TrackHeader* th = reinterpret_cast<const TrackHeader*>(blob + gh->track_off[tid]);
        int32_t* s = lap + car * FTGP_LAP_FIELDS;
        if (status && ((status[car] >> 16) & 0xFF)) s[FTGP_LAP_CONTACT_TICKS] += 1;
        if (th->path_off < 0) continue;
        const double* path = reinterpret_cast<const double*>(blob + th->path_off);
        const double x = qpos[car * stride], y = qpos[car * stride + 1];
        int closest = 0; double best = 0;
        for (int k = 0; k < FTGP_NPATH; k++) {                        // custom.py:1341-1342
            double dx = path[2 * k] - x, dy = path[2 * k + 1] - y;
            double d = dx * dx + dy * dy;
            if (k == 0 || d < best) { best = d; closest = k; }
        }
        const int off = best > 1;                                     // custom.py:1343-1344
        s[FTGP_LAP_OFF_TRACK] = off;
        if (off) { s[FTGP_LAP_OFFTRACK_TICKS] += 1; continue; }
        const int completion = pymod(closest - s[FTGP_LAP_OFFSET], 100);
        const int draw = completion - s[FTGP_LAP_COMPLETION];
        const int delta = pymod(draw + 50, 100) - 50;
        s[FTGP_LAP_DELTA] = delta;
        if (abs(draw) > 90) {
            const int lap_steps = steps - s[FTGP_LAP_START];
            if (delta < 0) {
                s[FTGP_LAP_LAPS] -= 1; s[FTGP_LAP_GOOD_START] = 0;
                if (s[FTGP_LAP_NTIMES] != 0) s[FTGP_LAP_NTIMES] -= 1;
            } else if (delta > 0) {
                if (s[FTGP_LAP_GOOD_START]) {
                    int n = s[FTGP_LAP_NTIMES];
                    if (n < FTGP_MAX_LAPTIMES) times[car * FTGP_MAX_LAPTIMES + n] = lap_steps;
                    s[FTGP_LAP_NTIMES] = n + 1;
                    s[FTGP_LAP_START] = steps;
                }
                s[FTGP_LAP_LAPS] += 1; s[FTGP_LAP_GOOD_START] = 1;
            }
        }
        if (s[FTGP_LAP_LAPS] >= lap_target) {                         // custom.py:1367-1371
            if (s[FTGP_LAP_RANK] == 0) { nwin += 1; s[FTGP_LAP_RANK] = nwin; }
            s[FTGP_LAP_FINISHED] = 1;
        }
        s[FTGP_LAP_COMPLETION] = completion;                          // custom.py:1372
    }
    if (winners) winners[world] = nwin;
}

int launch_drivers(const float* ranges, const int32_t* kind, int default_kind, const int32_t* lap, double* ctrl,
                   int64_t ncars, cudaStream_t stream, int32_t* steps_dev) {
    const int threads = 256;
    int64_t blocks = (ncars * 32 + threads - 1) / threads;
    drivers_kernel<<<(unsigned)blocks, threads, 0, stream>>>(ranges, kind, default_kind, lap, ctrl, ncars, steps_dev);
    count_launch();
    FTGP_CUDA(cudaGetLastError());
    return FTGP_OK;
}

int launch_lap(const ftgp_geom* g, const double* qpos, int64_t stride, const int32_t* track_id, int32_t* lap,
               int32_t* times, int32_t* winners, const int32_t* status, int64_t ncars, int cpw, int32_t steps,
               int32_t lap_target, cudaStream_t stream, const int32_t* steps_dev) {
    int64_t nworlds = (ncars + cpw - 1) / cpw;
    const int threads = 128;
    lap_kernel<<<(unsigned)((nworlds + threads - 1) / threads), threads, 0, stream>>>(
        g->d_blob, qpos, stride, track_id, lap, times, winners, status, ncars, cpw, steps, lap_target, steps_dev);
    count_launch();
    FTGP_CUDA(cudaGetLastError());
    return FTGP_OK;
}

}  // namespace ftgp
using namespace ftgp;

extern "C" int ftgp_reset(double* qpos, double* qvel, double* warm, double* ctrl, const double* xy,
                          const double* yaw, int64_t ncars, void* stream) {
    if (!qpos || !qvel || !warm || !xy || !yaw || ncars < 0) { set_error("ftgp_reset: bad argument"); return FTGP_ERR_ARG; }
    if (ncars == 0) return FTGP_OK;
    reset_kernel<<<(unsigned)((ncars + 127) / 128), 128, 0, (cudaStream_t)stream>>>(qpos, qvel, warm, ctrl, xy, yaw, ncars);
    count_launch();
    FTGP_CUDA(cudaGetLastError());
    return FTGP_OK;
}

extern "C" int ftgp_naive_flatten(double* qpos, int64_t qpos_stride, int64_t ncars, void* stream) {
    if (!qpos || ncars < 0 || qpos_stride < 7) { set_error("ftgp_naive_flatten: bad argument"); return FTGP_ERR_ARG; }
    if (ncars == 0) return FTGP_OK;
    return launch_flatten(qpos, qpos_stride, ncars, (cudaStream_t)stream);
}

extern "C" int ftgp_drivers(const float* ranges, const int32_t* kind, int default_kind, const int32_t* lap,
                            double* ctrl, int64_t ncars, void* stream) {
    if (!ranges || !ctrl || ncars < 0 || default_kind < 0 || default_kind > 2) { set_error("ftgp_drivers: bad argument"); return FTGP_ERR_ARG; }
    if (ncars == 0) return FTGP_OK;
    return launch_drivers(ranges, kind, default_kind, lap, ctrl, ncars, (cudaStream_t)stream, nullptr);
}

extern "C" int ftgp_lap_update(const ftgp_geom* g, const double* qpos, int64_t qpos_stride,
                               const int32_t* track_id, int32_t* lap, int32_t* times, int32_t* winners,
                               const int32_t* status, int64_t ncars, int cars_per_world, int32_t steps,
                               int32_t lap_target, void* stream) {
    if (!g || !qpos || !lap || !times || ncars < 0 || cars_per_world < 1 || qpos_stride < 2) {
        set_error("ftgp_lap_update: bad argument"); return FTGP_ERR_ARG;
    }
    if (ncars == 0) return FTGP_OK;
    FTGP_CUDA(cudaSetDevice(g->device));
    return launch_lap(g, qpos, qpos_stride, track_id, lap, times, winners, status, ncars, cars_per_world, steps,
                      lap_target, (cudaStream_t)stream, nullptr);
}
