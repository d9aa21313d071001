// lidar.cu -- the rangefinder kernel.
//
// Replaces `data.sensordata[vehicle_state.sensors]` (ft_grandprix/custom.py:1395), i.e. the
// 90 `rangefinder` sensors per car of template/mushr.em.xml:112-117,204-206, which MuJoCo
// evaluates with mj_ray against the hfield walls (mushr.em.xml:92), the ground plane
// (mushr.em.xml:94) and, in multi-car worlds, the other cars' lidar cylinders
// (mushr.em.xml:108).  Semantics restated in SURVEY.md B.10.
//
// Design (B200):
//   * the whole compiled track (chunk index grid + 400-bit vertex masks, ~40-60 KB) is staged
//     once per CTA into shared memory by a persistent grid of 148 x k CTAs;
//   * one warp per car, lane l traces beams l, l+32, l+64 (< 90): within one round the 32
//     lanes hold 32 angularly adjacent beams (4 deg apart) and the 32 results are written as
//     one coalesced 128-byte store;
//   * two-level traversal: Amanatides-Woo DDA over the 0.5 m chunk grid (empty chunks cost
//     one shared-memory load), then a major-axis sweep over the 19x19 cells of a non-empty
//     chunk where only cells with at least one wall vertex (4 mask bits) are intersected;
//   * a cell is two height-field triangles split along (c,r)-(c+1,r+1) exactly as mj_rayHfield
//     does; the chunk's vertical side faces count as hits below the boundary elevation profile;
//   * pose math (quaternion -> ray origin/direction -> grid coordinates) is done in fp64, the
//     traversal in fp32 on cell-relative coordinates (error budget ~4e-6 m << 1e-4 m);
//   * a warp-level min-reduction yields the per-car closest obstacle (min_range).
#include <math.h>
#include "common.h"

namespace ftgp {

__constant__ double c_beam_sc[FTGP_NBEAMS][2];   // (sin b, cos b), b = radians(4j - 90)
static bool g_beam_ready[16] = {false};

constexpr float HF_RANGE = 0.3f;      // hfield elevation range: border_height + affordance (mushr.em.xml:16,22,55)
constexpr float HF_Z0 = -0.1f;        // hfield geom z (mushr.em.xml:92)
constexpr float PLANE_Z = 0.01f;      // ground plane (mushr.em.xml:94)
constexpr float PLANE_HALF = 300.0f;
constexpr float CELL_EPS = 1e-4f;     // inside-test tolerance in cell units (~2.6 um)
constexpr float BIG = 3.0e38f;

struct Ray {
    int ix0, iy0;        // chunk cell containing the origin (may be outside the grid)
    float fx0, fy0;      // fractional position inside that cell, [0,1)
    float dgx, dgy;      // chunk-grid units per metre of ray parameter
    float lz, dz;        // height above the hfield base plane (z + 0.1), vertical direction
};

__device__ __forceinline__ uint32_t mask_bits2(const uint32_t* m, int bit) {
    // two consecutive vertex bits starting at `bit` (may straddle a word)
    int w = bit >> 5, s = bit & 31;
    uint32_t lo = m[w], hi = m[min(w + 1, 12)];
    return __funnelshift_r(lo, hi, s) & 3u;
}
__device__ __forceinline__ uint32_t mask_bit(const uint32_t* m, int bit) {
    return (m[bit >> 5] >> (bit & 31)) & 1u;
}

// Intersect the ray (cell-local start (u,v), height z, per-metre increments du,dv,dz) with
// the two triangles of one cell.  bits: b0 = (c,r), b1 = (c+1,r), b2 = (c,r+1), b3 = (c+1,r+1).
// Returns the smallest parameter s >= smin at which it hits, or BIG.  `unsure` is set when an
// inside test landed within CELL_EPS of a triangle edge (or s within 1e-5 of smin): there the
// fp32 decision is not trustworthy (a ray grazing a ridge by microns) and the caller re-runs
// the cell in fp64 (cell_hit_exact).
__device__ __forceinline__ float cell_hit(uint32_t bits, float u, float v, float z,
                                          float du, float dv, float dz, float smin, bool& unsure) {
    float h00 = (bits & 1u) ? HF_RANGE : 0.f, h10 = (bits & 2u) ? HF_RANGE : 0.f;
    float h01 = (bits & 4u) ? HF_RANGE : 0.f, h11 = (bits & 8u) ? HF_RANGE : 0.f;
    float best = BIG;
    {   // triangle {(c,r), (c+1,r+1), (c+1,r)}: v <= u
        float a = h10 - h00, b = h11 - h10;
        float den = dz - a * du - b * dv;
        float num = z - h00 - a * u - b * v;
        if (fabsf(den) > 1e-15f) {
            float s = -num / den;
            float uu = u + s * du, vv = v + s * dv;
            float m = fminf(fminf(vv, uu - vv), 1.f - uu);          // > 0 inside
            if (s >= smin - 1e-5f && m >= -CELL_EPS) {
                if (m <= CELL_EPS || s < smin + 1e-5f) unsure = true; else best = s;
            }
        }
    }
    {   // triangle {(c,r), (c+1,r+1), (c,r+1)}: v >= u
        float a = h11 - h01, b = h01 - h00;
        float den = dz - a * du - b * dv;
        float num = z - h00 - a * u - b * v;
        if (fabsf(den) > 1e-15f) {
            float s = -num / den;
            float uu = u + s * du, vv = v + s * dv;
            float m = fminf(fminf(uu, vv - uu), 1.f - vv);
            if (s >= smin - 1e-5f && m >= -CELL_EPS) {
                if (m <= CELL_EPS || s < smin + 1e-5f) unsure = true; else if (s < best) best = s;
            }
        }
    }
    return best;
}

// fp64 re-evaluation of one cell for a grazing ray: the ray is rebuilt from the car pose exactly as
// the CPU restatement does, and the two triangles are tested without tolerance (hit point inside
// the closed triangle, s >= 0).  Returns the absolute ray parameter or BIG.  Rare path (~1e-4 of cells).
__device__ __noinline__ float cell_hit_exact(const TrackHeader* th, const double* __restrict__ pose7, int beam,
                                             int ix, int iy, int ncol, int nrow, int cc, int rr, uint32_t bits) {
    double q[7];
#pragma unroll
    for (int k = 0; k < 7; k++) q[k] = pose7[k];
    double n = sqrt(q[3] * q[3] + q[4] * q[4] + q[5] * q[5] + q[6] * q[6]);
    double w = q[3], x = q[4], y = q[5], z = q[6];
    if (n < 1e-15) { w = 1; x = y = z = 0; } else { w /= n; x /= n; y /= n; z /= n; }
    const double R0 = 1 - 2 * (y * y + z * z), R1 = 2 * (x * y - w * z), R2 = 2 * (x * z + w * y);
    const double R3 = 2 * (x * y + w * z), R4 = 1 - 2 * (x * x + z * z), R5 = 2 * (y * z - w * x);
    const double R6 = 2 * (x * z - w * y), R7 = 2 * (y * z + w * x), R8 = 1 - 2 * (x * x + y * y);
    const double sb = c_beam_sc[beam][0], cb = c_beam_sc[beam][1];
    const double rx = -0.0525, rz = 0.065, lr = 0.030;
    const double dx = sb * R0 - cb * R1, dy = sb * R3 - cb * R4, dz = sb * R6 - cb * R7;
    const double lx = rx - lr * sb, ly = lr * cb;
    const double ox = q[0] + R0 * lx + R1 * ly + R2 * rz, oy = q[1] + R3 * lx + R4 * ly + R5 * rz,
                 oz = q[2] + R6 * lx + R7 * ly + R8 * rz;
    // chunk (ix, iy) is centred at (size_x * ix, size_y * (iy - (vc-1))); cell pitch = size / (n - 1)
    const double cw = th->dsize_x / (ncol - 1), chh = th->dsize_y / (nrow - 1);
    const double u = (ox - th->dsize_x * ix + 0.5 * th->dsize_x) / cw - cc;
    const double v = (oy - th->dsize_y * (iy - (th->vc - 1)) + 0.5 * th->dsize_y) / chh - rr;
    const double du = dx / cw, dv = dy / chh, zz = oz + 0.1;      // hfield geom z = -0.1 (mushr.em.xml:92)
    const double H = 0.2 + 0.1;                                      // border_height + affordance (mushr.em.xml:55)
    const double h00 = (bits & 1u) ? H : 0., h10 = (bits & 2u) ? H : 0., h01 = (bits & 4u) ? H : 0., h11 = (bits & 8u) ? H : 0.;
    double best = 1e300;
    {
        double a = h10 - h00, b = h11 - h10, den = dz - a * du - b * dv, num = zz - h00 - a * u - b * v;
        if (fabs(den) > 1e-15) {
            double s = -num / den, uu = u + s * du, vv = v + s * dv;
            if (s >= 0 && vv >= 0 && vv <= uu && uu <= 1) best = s;
        }
    }
    {
        double a = h11 - h01, b = h01 - h00, den = dz - a * du - b * dv, num = zz - h00 - a * u - b * v;
        if (fabs(den) > 1e-15) {
            double s = -num / den, uu = u + s * du, vv = v + s * dv;
            if (s >= 0 && s < best && uu >= 0 && uu <= vv && vv <= 1) best = s;
        }
    }
    return best < 1e300 ? (float)best : BIG;
}

// Vertical side face of a chunk's top box, crossed at parameter t.  axis 0: face normal to x
// (profile runs along rows), axis 1: normal to y.  far_side: the +x / +y face.
__device__ __forceinline__ bool face_hit(const uint32_t* m, int ncol, int nrow, int axis, bool far_side,
                                         float along, float zt) {
    if (zt < 0.f || zt > 1.f) return false;
    int n = axis == 0 ? nrow : ncol;
    if (along < -CELL_EPS || along > (float)(n - 1) + CELL_EPS) return false;
    float y0 = fminf(fmaxf(floorf(along), 0.f), (float)(n - 2));
    int k = (int)y0;
    uint32_t z0, z1;
    if (axis == 0) {
        int col = far_side ? ncol - 1 : 0;
        z0 = mask_bit(m, k * ncol + col); z1 = mask_bit(m, (k + 1) * ncol + col);
    } else {
        int row = far_side ? nrow - 1 : 0;
        z0 = mask_bit(m, row * ncol + k); z1 = mask_bit(m, row * ncol + k + 1);
    }
    float prof = (float)z0 * (y0 + 1.f - along) + (float)z1 * (along - y0);
    return zt < prof;
}

// All hfield geoms + ground plane for one ray.  Returns distance or -1.
__device__ float trace_walls(const uint32_t* __restrict__ sm, const TrackHeader* th, const Ray& ry,
                             const double* __restrict__ pose7, int beam) {
    float best = BIG;
    // ground plane: hit only from the front side, inside the 300 m rendered square
    if (ry.dz < -1e-15f) {
        float tp = -(ry.lz + HF_Z0 - PLANE_Z) / ry.dz;
        if (tp >= 0.f) {
            float px = ((float)ry.ix0 + ry.fx0 - 0.5f + tp * ry.dgx) * th->size_x;
            float py = ((float)(ry.iy0 - (th->vc - 1)) + ry.fy0 - 0.5f + tp * ry.dgy) * th->size_y;
            if (fabsf(px) <= PLANE_HALF && fabsf(py) <= PLANE_HALF) best = tp;
        }
    }
    // parameter interval in which the ray is inside the walls' height slab [0, 0.3]
    float tz0 = 0.f, tz1 = BIG;
    if (fabsf(ry.dz) > 1e-15f) {
        float a = (0.f - ry.lz) / ry.dz, b = (HF_RANGE - ry.lz) / ry.dz;
        tz0 = fmaxf(0.f, fminf(a, b)); tz1 = fmaxf(a, b);
    } else if (ry.lz < 0.f || ry.lz > HF_RANGE) tz1 = -1.f;
    float tend = fminf(tz1, best);
    if (tend < tz0) return best < BIG ? best : -1.f;

    // clip against the chunk grid [0,hc] x [0,vc]
    const int hc = th->hc, vc = th->vc;
    float gx = (float)ry.ix0 + ry.fx0, gy = (float)ry.iy0 + ry.fy0;
    float tg0 = 0.f, tg1 = tend;
    int enter_axis = -1;
    const float inv_dgx = fabsf(ry.dgx) > 1e-20f ? 1.f / ry.dgx : 0.f;
    const float inv_dgy = fabsf(ry.dgy) > 1e-20f ? 1.f / ry.dgy : 0.f;
    if (inv_dgx != 0.f) {
        float a = (0.f - gx) * inv_dgx, b = ((float)hc - gx) * inv_dgx;
        float lo = fminf(a, b), hi = fmaxf(a, b);
        if (lo > tg0) { tg0 = lo; enter_axis = 0; }
        tg1 = fminf(tg1, hi);
    } else if (gx < 0.f || gx >= (float)hc) tg1 = -1.f;
    if (inv_dgy != 0.f) {
        float a = (0.f - gy) * inv_dgy, b = ((float)vc - gy) * inv_dgy;
        float lo = fminf(a, b), hi = fmaxf(a, b);
        if (lo > tg0) { tg0 = lo; enter_axis = 1; }
        tg1 = fminf(tg1, hi);
    } else if (gy < 0.f || gy >= (float)vc) tg1 = -1.f;
    if (tg1 < tg0) return best < BIG ? best : -1.f;
    float ts = tg0;
    if (tz0 > ts) { ts = tz0; enter_axis = -1; }   // dropping into the slab from above: no side face
    if (ts > tend) return best < BIG ? best : -1.f;

    const int stepx = ry.dgx >= 0.f ? 1 : -1, stepy = ry.dgy >= 0.f ? 1 : -1;
    // chunk containing the start point (relative to the origin cell to keep fp32 exact)
    float relx = ry.fx0 + ts * ry.dgx, rely = ry.fy0 + ts * ry.dgy;
    int ix = ry.ix0 + (int)floorf(relx), iy = ry.iy0 + (int)floorf(rely);
    if (enter_axis == 0) ix = stepx > 0 ? 0 : hc - 1;
    if (enter_axis == 1) iy = stepy > 0 ? 0 : vc - 1;
    ix = min(max(ix, 0), hc - 1); iy = min(max(iy, 0), vc - 1);

    const uint16_t* index = reinterpret_cast<const uint16_t*>(sm + th->index_off);
    const uint32_t* chunks = sm + th->chunks_off;
    float t0 = ts;
    int entry_axis = enter_axis;
    for (int guard = 0; guard < 512; guard++) {
        // exit of this chunk
        float tmx = inv_dgx != 0.f ? ((float)(ix - ry.ix0 + (stepx > 0 ? 1 : 0)) - ry.fx0) * inv_dgx : BIG;
        float tmy = inv_dgy != 0.f ? ((float)(iy - ry.iy0 + (stepy > 0 ? 1 : 0)) - ry.fy0) * inv_dgy : BIG;
        int exit_axis = tmx <= tmy ? 0 : 1;
        float t1 = fminf(tmx, tmy);
        uint32_t cid = index[iy * hc + ix];
        if (cid != EMPTY_CHUNK) {
            const uint32_t* m = chunks + cid * CHUNK_WORDS;
            const int ncol = m[13] & 0xFF, nrow = (m[13] >> 8) & 0xFF;
            const float nx = (float)(ncol - 1), ny = (float)(nrow - 1);
            // fine-cell coordinates inside this chunk: f(t) = fo + t * df
            const float fxo = ((float)(ry.ix0 - ix) + ry.fx0) * nx, dfx = ry.dgx * nx;
            const float fyo = ((float)(ry.iy0 - iy) + ry.fy0) * ny, dfy = ry.dgy * ny;
            const float inv_h = 1.f / HF_RANGE;
            // entry side face
            if (entry_axis >= 0 && t0 <= tend) {
                bool far_side = entry_axis == 0 ? stepx < 0 : stepy < 0;
                float along = entry_axis == 0 ? fyo + t0 * dfy : fxo + t0 * dfx;
                if (face_hit(m, ncol, nrow, entry_axis, far_side, along, (ry.lz + t0 * ry.dz) * inv_h))
                    return fminf(best, t0);
            }
            float ta = fmaxf(t0, tz0), tb = fminf(t1, tend);
            if (ta <= tb) {
                // re-origin at ta
                const float xa = fxo + ta * dfx, ya = fyo + ta * dfy, za = ry.lz + ta * ry.dz;
                const float span = tb - ta;
                const bool floor_reach = fminf(za, za + span * ry.dz) <= 0.f;
                const bool major_x = fabsf(dfx) >= fabsf(dfy);
                const float dM = major_x ? dfx : dfy, dm = major_x ? dfy : dfx;
                const float Ma = major_x ? xa : ya, ma = major_x ? ya : xa;
                const int nM = major_x ? ncol - 1 : nrow - 1, nm = major_x ? nrow - 1 : ncol - 1;
                const int sg = dM >= 0.f ? 1 : -1;
                const float inv_dM = fabsf(dM) > 1e-20f ? 1.f / dM : 0.f;
                float Mb = Ma + span * dM;
                int c = (int)floorf(Ma - (float)sg * 1e-3f), cend = (int)floorf(Mb + (float)sg * 1e-3f);
                c = min(max(c, 0), nM - 1); cend = min(max(cend, 0), nM - 1);
                float found = BIG;
                for (;;) {
                    // parameter interval (relative to ta) spent in major-cell c
                    float s0 = 0.f, s1 = span;
                    if (inv_dM != 0.f) {
                        float e0 = ((float)(c + (sg > 0 ? 0 : 1)) - Ma) * inv_dM;
                        float e1 = ((float)(c + (sg > 0 ? 1 : 0)) - Ma) * inv_dM;
                        s0 = fmaxf(s0, e0); s1 = fminf(s1, e1);
                    }
                    float m0 = ma + s0 * dm, m1 = ma + s1 * dm;
                    int rlo = (int)floorf(fminf(m0, m1) - 1e-3f), rhi = (int)floorf(fmaxf(m0, m1) + 1e-3f);
                    rlo = min(max(rlo, 0), nm - 1); rhi = min(max(rhi, 0), nm - 1);
                    for (int r = rlo; r <= rhi; r++) {
                        int cc = major_x ? c : r, rr = major_x ? r : c;
                        uint32_t bits = mask_bits2(m, rr * ncol + cc) | (mask_bits2(m, (rr + 1) * ncol + cc) << 2);
                        if (bits == 0u && !floor_reach) continue;
                        bool unsure = false;
                        float s = cell_hit(bits, xa - (float)cc, ya - (float)rr, za, dfx, dfy, ry.dz, -ta, unsure);
                        if (unsure) {
                            float sx = cell_hit_exact(th, pose7, beam, ix, iy, ncol, nrow, cc, rr, bits);
                            s = sx < BIG ? sx - ta : BIG;
                        }
                        if (s <= span + 1e-4f) found = fminf(found, s);
                    }
                    if (found < BIG || c == cend) break;
                    c += sg;
                }
                if (found < BIG) return fminf(best, fmaxf(ta + found, 0.f));
            }
            // exit side face
            if (t1 <= tend) {
                bool far_side = exit_axis == 0 ? stepx > 0 : stepy > 0;
                float along = exit_axis == 0 ? fyo + t1 * dfy : fxo + t1 * dfx;
                if (face_hit(m, ncol, nrow, exit_axis, far_side, along, (ry.lz + t1 * ry.dz) * inv_h))
                    return fminf(best, t1);
            }
        }
        if (t1 >= tend) break;
        if (exit_axis == 0) ix += stepx; else iy += stepy;
        if (ix < 0 || ix >= hc || iy < 0 || iy >= vc) break;
        t0 = t1; entry_axis = exit_axis;
    }
    return best < BIG ? best : -1.f;
}

// other car's lidar cylinder (mushr.em.xml:108): r = 0.03, half height 0.015, centred at
// (rx, 0, rz - lh/2) in that car's frame.  p/v: ray in the cylinder frame.  fp64: the target is 3 cm
// wide, so b^2 - a c cancels badly in fp32 for grazing rays (mju_rayGeom's cylinder branch restated).
__device__ __forceinline__ double ray_cylinder(double px, double py, double pz, double vx, double vy, double vz,
                                               double r, double hh) {
    double best = -1.0;
    const double a = vx * vx + vy * vy, b = px * vx + py * vy, c = px * px + py * py - r * r;
    if (a > 1e-15) {
        double det = b * b - a * c;
        if (det >= 1e-15) {
            det = sqrt(det);
            const double x0 = (-b - det) / a, x1 = (-b + det) / a;
            const double s = x0 >= 0 ? x0 : (x1 >= 0 ? x1 : -1.0);
            if (s >= 0 && fabs(pz + s * vz) <= hh) best = s;
        }
    }
    if (fabs(vz) > 1e-15) {
#pragma unroll
        for (int side = -1; side <= 1; side += 2) {
            const double s = ((double)side * hh - pz) / vz;
            if (s >= 0) {
                const double x = px + s * vx, y = py + s * vy;
                if (x * x + y * y <= r * r && (best < 0 || s < best)) best = s;
            }
        }
    }
    return best;
}

struct Pose { double p[3], R[9]; };

__device__ __forceinline__ void load_pose(const double* __restrict__ qpos, int64_t stride, int64_t car,
                                          int lane, Pose& P) {
    double v = lane < 7 ? qpos[car * stride + lane] : 0.0;
    double q[7];
#pragma unroll
    for (int k = 0; k < 7; k++) q[k] = __shfl_sync(0xffffffffu, v, k);
    double n = sqrt(q[3] * q[3] + q[4] * q[4] + q[5] * q[5] + q[6] * q[6]);
    double w = q[3], x = q[4], y = q[5], z = q[6];
    if (n < 1e-15) { w = 1; x = y = z = 0; } else { w /= n; x /= n; y /= n; z /= n; }
    P.p[0] = q[0]; P.p[1] = q[1]; P.p[2] = q[2];
    P.R[0] = 1 - 2 * (y * y + z * z); P.R[1] = 2 * (x * y - w * z); P.R[2] = 2 * (x * z + w * y);
    P.R[3] = 2 * (x * y + w * z); P.R[4] = 1 - 2 * (x * x + z * z); P.R[5] = 2 * (y * z - w * x);
    P.R[6] = 2 * (x * z - w * y); P.R[7] = 2 * (y * z + w * x); P.R[8] = 1 - 2 * (x * x + y * y);
}

__global__ void __launch_bounds__(512)
lidar_kernel(const uint32_t* __restrict__ blob, int lidar_words, const double* __restrict__ qpos,
             int64_t stride, const int32_t* __restrict__ track_id, const uint8_t* __restrict__ visible,
             int64_t ncars, int cpw, float* __restrict__ ranges, float* __restrict__ min_range) {
    extern __shared__ __align__(16) uint32_t sm[];
    {
        const uint4* src = reinterpret_cast<const uint4*>(blob);
        uint4* dst = reinterpret_cast<uint4*>(sm);
        int n4 = (lidar_words + 3) >> 2;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const GeomHeader* gh = reinterpret_cast<const GeomHeader*>(sm);
    const int lane = threadIdx.x & 31;
    const int64_t warps_per_block = blockDim.x >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * warps_per_block;
    const double rx = -0.0525, rz = 0.065, lr = 0.030;   // mushr.em.xml:101-103

    for (int64_t car = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); car < ncars; car += nwarps) {
        int tid = track_id ? track_id[car] : 0;
        if (tid < 0 || tid >= gh->ntracks) tid = 0;
        const TrackHeader* th = reinterpret_cast<const TrackHeader*>(sm + gh->track_off[tid]);
        Pose P;
        load_pose(qpos, stride, car, lane, P);
        const double inv_sx = 1.0 / th->dsize_x, inv_sy = 1.0 / th->dsize_y;
        float rng[3];
        double dwx[3], dwy[3], dwz[3], owx[3], owy[3], owz[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int j = lane + 32 * k;
            rng[k] = -1.f;
            if (j < FTGP_NBEAMS) {
                const double sb = c_beam_sc[j][0], cb = c_beam_sc[j][1];
                // site +Z axis in the car frame = (sin b, -cos b, 0); origin = lidar axis - lr * dir
                dwx[k] = sb * P.R[0] - cb * P.R[1]; dwy[k] = sb * P.R[3] - cb * P.R[4]; dwz[k] = sb * P.R[6] - cb * P.R[7];
                const double lx = rx - lr * sb, ly = lr * cb;
                owx[k] = P.p[0] + P.R[0] * lx + P.R[1] * ly + P.R[2] * rz;
                owy[k] = P.p[1] + P.R[3] * lx + P.R[4] * ly + P.R[5] * rz;
                owz[k] = P.p[2] + P.R[6] * lx + P.R[7] * ly + P.R[8] * rz;
                const double gxd = owx[k] * inv_sx + 0.5, gyd = owy[k] * inv_sy + 0.5 + (double)(th->vc - 1);
                const double fgx = floor(gxd), fgy = floor(gyd);
                Ray ry;
                ry.ix0 = (int)fgx; ry.iy0 = (int)fgy;
                ry.fx0 = (float)(gxd - fgx); ry.fy0 = (float)(gyd - fgy);
                ry.dgx = (float)(dwx[k] * inv_sx); ry.dgy = (float)(dwy[k] * inv_sy);
                ry.lz = (float)(owz[k] - (double)HF_Z0); ry.dz = (float)dwz[k];
                rng[k] = trace_walls(sm, th, ry, qpos + car * stride, j);
            }
        }
        if (cpw > 1) {
            const int64_t w0 = (car / cpw) * cpw;
            for (int64_t oc = w0; oc < w0 + cpw && oc < ncars; oc++) {
                if (oc == car) continue;                       // bodyexclude: own root body
                if (visible && !visible[oc]) continue;         // shadowed car: alpha 0 material
                Pose Q;
                load_pose(qpos, stride, oc, lane, Q);
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    if (lane + 32 * k >= FTGP_NBEAMS) continue;
                    // ray in the other car's frame, relative to its cylinder centre
                    const double ex = owx[k] - Q.p[0], ey = owy[k] - Q.p[1], ez = owz[k] - Q.p[2];
                    const double px = Q.R[0] * ex + Q.R[3] * ey + Q.R[6] * ez - rx;
                    const double py = Q.R[1] * ex + Q.R[4] * ey + Q.R[7] * ez;
                    const double pz = Q.R[2] * ex + Q.R[5] * ey + Q.R[8] * ez - (rz - 0.015 / 2);
                    const double vx = Q.R[0] * dwx[k] + Q.R[3] * dwy[k] + Q.R[6] * dwz[k];
                    const double vy = Q.R[1] * dwx[k] + Q.R[4] * dwy[k] + Q.R[7] * dwz[k];
                    const double vz = Q.R[2] * dwx[k] + Q.R[5] * dwy[k] + Q.R[8] * dwz[k];
                    const double s = ray_cylinder(px, py, pz, vx, vy, vz, 0.03, 0.015);
                    if (s >= 0 && (rng[k] < 0.f || (float)s < rng[k])) rng[k] = (float)s;
                }
            }
        }
        float* out = ranges + car * FTGP_NBEAMS;
        float mn = BIG;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int j = lane + 32 * k;
            if (j < FTGP_NBEAMS) {
                out[j] = rng[k];
                if (rng[k] >= 0.f) mn = fminf(mn, rng[k]);
            }
        }
        if (min_range) {
            // warp-level min-reduction (non-negative floats order like their bit patterns)
            unsigned r = __reduce_min_sync(0xffffffffu, __float_as_uint(mn));
            if (lane == 0) min_range[car] = r == __float_as_uint(BIG) ? INFINITY : __uint_as_float(r);
        }
    }
}

static int ensure_beams(int device) {
    if (device >= 0 && device < 16 && g_beam_ready[device]) return FTGP_OK;
    double h[FTGP_NBEAMS][2];
    const double PI = 3.14159265358979323846;
    for (int j = 0; j < FTGP_NBEAMS; j++) {
        double b = (360.0 / FTGP_NBEAMS * j - 90.0) * PI / 180.0;   // radians(inter_ray_angle*j - 90), mushr.em.xml:21,114
        h[j][0] = sin(b); h[j][1] = cos(b);
    }
    FTGP_CUDA(cudaMemcpyToSymbol(c_beam_sc, h, sizeof h));
    if (device >= 0 && device < 16) g_beam_ready[device] = true;
    return FTGP_OK;
}

int launch_lidar(const ftgp_geom* g, const double* qpos, int64_t stride, const int32_t* track_id,
                 const uint8_t* visible, int64_t ncars, int cpw, float* ranges, float* min_range,
                 cudaStream_t stream) {
    GeomHeader gh; memcpy(&gh, g->h_blob.data(), sizeof gh);
    size_t smem = (size_t)((gh.lidar_words + 3) / 4) * 16;
    static int sm_count[16] = {0};
    int dev = g->device;
    if (dev < 16 && sm_count[dev] == 0) {
        FTGP_CUDA(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
        FTGP_CUDA(cudaFuncSetAttribute(lidar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    if (smem > 227 * 1024) { set_error("geometry blob (%zu B) exceeds shared memory", smem); return FTGP_ERR_UNSUPPORTED; }
    int rc = ensure_beams(dev); if (rc) return rc;
    const int threads = 512;
    int per_sm = (int)std::min<size_t>(4, (227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    int64_t need = (ncars + (threads / 32) - 1) / (threads / 32);
    int nsm = dev < 16 ? sm_count[dev] : 148;
    int grid = (int)std::min<int64_t>(need, (int64_t)nsm * per_sm);
    if (grid < 1) return FTGP_OK;
    lidar_kernel<<<grid, threads, smem, stream>>>(g->d_blob, gh.lidar_words, qpos, stride, track_id, visible,
                                                  ncars, cpw, ranges, min_range);
    count_launch();
    FTGP_CUDA(cudaGetLastError());
    return FTGP_OK;
}

}  // namespace ftgp
using namespace ftgp;

extern "C" int ftgp_lidar(const ftgp_geom* g, const double* qpos, int64_t qpos_stride,
                          const int32_t* track_id, const uint8_t* visible, int64_t ncars,
                          int cars_per_world, float* ranges, float* min_range, void* stream) {
    if (!g || !qpos || !ranges || ncars < 0 || qpos_stride < 7 || cars_per_world < 1) {
        set_error("ftgp_lidar: bad argument"); return FTGP_ERR_ARG;
    }
    if (ncars == 0) return FTGP_OK;
    FTGP_CUDA(cudaSetDevice(g->device));
    return launch_lidar(g, qpos, qpos_stride, track_id, visible, ncars, cars_per_world, ranges, min_range,
                        (cudaStream_t)stream);
}

extern "C" int ftgp_lidar_host(const ftgp_geom* g, const double* qpos, int64_t qpos_stride,
                               const int32_t* track_id, int64_t ncars, float* ranges) {
    if (!g || !qpos || !ranges || ncars < 0 || qpos_stride < 7) { set_error("ftgp_lidar_host: bad argument"); return FTGP_ERR_ARG; }
    if (ncars == 0) return FTGP_OK;
    FTGP_CUDA(cudaSetDevice(g->device));
    ftgp_geom* gm = const_cast<ftgp_geom*>(g);
    size_t b_pose = (size_t)ncars * 7 * sizeof(double), b_tid = track_id ? (size_t)ncars * 4 : 0;
    size_t b_rng = (size_t)ncars * FTGP_NBEAMS * sizeof(float);
    size_t o_tid = (b_pose + 255) & ~(size_t)255, o_rng = (o_tid + b_tid + 255) & ~(size_t)255;
    size_t need = o_rng + b_rng;
    if (need > gm->scratch_bytes) {
        if (gm->d_scratch) cudaFree(gm->d_scratch);
        gm->d_scratch = nullptr; gm->scratch_bytes = 0;
        FTGP_CUDA(cudaMalloc(&gm->d_scratch, need));
        gm->scratch_bytes = need;
    }
    char* base = (char*)gm->d_scratch;
    cudaStream_t s = gm->host_stream;
    // only the free-joint pose (7 doubles per car) crosses the bus
    FTGP_CUDA(cudaMemcpy2DAsync(base, 7 * sizeof(double), qpos, qpos_stride * sizeof(double), 7 * sizeof(double),
                                ncars, cudaMemcpyHostToDevice, s));
    if (track_id) FTGP_CUDA(cudaMemcpyAsync(base + o_tid, track_id, b_tid, cudaMemcpyHostToDevice, s));
    int rc = launch_lidar(g, (const double*)base, 7, track_id ? (const int32_t*)(base + o_tid) : nullptr, nullptr,
                          ncars, 1, (float*)(base + o_rng), nullptr, s);
    if (rc) return rc;
    FTGP_CUDA(cudaMemcpyAsync(ranges, base + o_rng, b_rng, cudaMemcpyDeviceToHost, s));
    FTGP_CUDA(cudaStreamSynchronize(s));
    return FTGP_OK;
}
