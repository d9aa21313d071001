// hfield_contact.cuh -- contacts of the car's geoms with the height-field walls, read straight from the compiled
// geometry blob (chunk index grid + 400-bit vertex masks, common.h).  Replaces what MuJoCo's collision stage does for
// the pairs (chassis mesh | lidar cylinder | wheel ellipsoid) x hfield (contype 1 vs conaffinity 1:
// template/mushr.em.xml:69,92,108,119) inside mujoco.mj_step (ft_grandprix/custom.py:1425).
//
// MuJoCo runs MPR / GJK against per-triangle prisms; that contact set cannot be reproduced bit for bit, so the
// framework defines its own rules (the CPU oracle states the same rules in oracle/step.c from
// independent code, and tests/test_contacts_cpu.py checks them against the explicit triangle mesh):
//   rule V  a probe point (chassis hull vertex) below the surface triangle it projects into -> contact with that plane
//   rule S  a smooth convex geom: for every surface triangle under its bounding square that has a raised vertex, the
//           support point in the direction opposite to the triangle's normal; if it projects into the triangle and lies
//           below its plane it is a candidate; the deepest candidate is the geom's one wall contact.
// __host__ __device__: the CPU tests compile this file with g++ against the host copy of the blob.
#pragma once
#include "mushr_step.cuh"

// the probes are rarely executed (a car near a wall) and long: kept out of line on the device so that the step kernel's
// hot path does not carry four inlined copies of them
#if defined(__CUDACC__)
#define FT_HDNI __host__ __device__ __noinline__
#else
#define FT_HDNI inline
#endif

namespace ftgp {
namespace mushr {

struct QWallHit { double dist, nrm[3], t1[3], t2[3], pnt[3]; };      // pnt = contact position (point - n dist / 2)

constexpr double HF_ELEV = 0.3;           // border_height + affordance (mushr.em.xml:16,22,55)
constexpr double HF_BASE_Z = -0.1;        // hfield geom z (mushr.em.xml:92)
constexpr int HF_CHUNK_WORDS = 28;        // == CHUNK_WORDS (common.h)
constexpr unsigned HF_EMPTY = 0xFFFFu;

// one compiled track as the contact code sees it
struct HfView {
    const uint16_t* index;                // [vc][hc], row gy = vc - 1 - j
    const uint32_t* chunks;               // [nchunks][HF_CHUNK_WORDS]
    int hc, vc;
    double size_x, size_y;                // chunk pitch (m)
};

FT_HD void hf_frame(QWallHit& h, const double* point) {               // mju_makeFrame + contact position
    h.t1[0] = h.t1[1] = h.t1[2] = 0;
    if (h.nrm[1] < 0.5 && h.nrm[1] > -0.5) h.t1[1] = 1; else h.t1[2] = 1;
    const double d = dot3(h.nrm, h.t1);
    for (int a = 0; a < 3; a++) h.t1[a] -= d * h.nrm[a];
    const double tn = sqrt(dot3(h.t1, h.t1));
    for (int a = 0; a < 3; a++) h.t1[a] /= tn;
    cross3(h.t2, h.nrm, h.t1);
    for (int a = 0; a < 3; a++) h.pnt[a] = point[a] - h.nrm[a] * h.dist * 0.5;
}

FT_HD double hf_bit(const uint32_t* m, int ncol, int r, int c) {      // vertex elevation above the hfield base
    const int b = r * ncol + c;
    return (double)((m[b >> 5] >> (b & 31)) & 1u) * HF_ELEV;
}

// n (<= 21) consecutive vertex bits of hfield row r starting at column c (two mask words, one shift)
FT_HD uint32_t hf_row_bits(const uint32_t* m, int ncol, int r, int c, int n) {
    const int b = r * ncol + c, w = b >> 5, sh = b & 31;
    const uint64_t two = ((uint64_t)m[w + 1 < 13 ? w + 1 : 12] << 32) | m[w];
    return (uint32_t)(two >> sh) & ((1u << n) - 1u);
}

// Can anything within `radius` of (x, y) touch a wall?  Corner k (0..3) of the square selects one of the (at most four)
// chunks under it; true if that chunk has raised vertices whose cells (one cell of slope around them) reach the square.
// The four corners together are a conservative gate for every probe below.
FT_HD bool hf_near_corner(const HfView& hv, double x, double y, double radius, int k) {
    const double xs = (k & 1) ? x + radius : x - radius, ys = (k & 2) ? y + radius : y - radius;
    const int i = (int)floor(xs / hv.size_x + 0.5), j = (int)floor(-ys / hv.size_y + 0.5);
    if (i < 0 || i >= hv.hc || j < 0 || j >= hv.vc) return false;
    const unsigned cid = hv.index[(hv.vc - 1 - j) * hv.hc + i];
    if (cid == HF_EMPTY) return false;
    const uint32_t* m = hv.chunks + cid * HF_CHUNK_WORDS;
    const int ncol = m[13] & 0xFF, nrow = (m[13] >> 8) & 0xFF;
    const uint32_t bb = m[14];
    const int cmin = bb & 0xFF, cmax = (bb >> 8) & 0xFF, rmin = (bb >> 16) & 0xFF, rmax = bb >> 24;
    if (cmin > cmax) return false;
    const double dx = hv.size_x / (ncol - 1), dy = hv.size_y / (nrow - 1);
    const double x0 = hv.size_x * i - 0.5 * hv.size_x, y0 = -hv.size_y * j - 0.5 * hv.size_y;
    if (!(x + radius >= x0 + (cmin - 1) * dx && x - radius <= x0 + (cmax + 1) * dx &&
          y + radius >= y0 + (rmin - 1) * dy && y - radius <= y0 + (rmax + 1) * dy)) return false;
    // the bounding box of a diagonal stroke is the whole chunk: look at the vertex bits themselves -- is there a raised
    // vertex within one cell of the square?  (one two-word extraction per vertex row)
    int c0 = (int)floor((x - radius - x0) / dx) - 1, c1 = (int)floor((x + radius - x0) / dx) + 2;
    int r0 = (int)floor((y - radius - y0) / dy) - 1, r1 = (int)floor((y + radius - y0) / dy) + 2;
    c0 = c0 < cmin ? cmin : c0; c1 = c1 > cmax ? cmax : c1;
    r0 = r0 < rmin ? rmin : r0; r1 = r1 > rmax ? rmax : r1;
    if (c0 > c1) return false;
    for (int r = r0; r <= r1; r++)
        if (hf_row_bits(m, ncol, r, c0, c1 - c0 + 1)) return true;
    return false;
}

// rule V: world point p against the surface triangle under it
FT_HDNI bool hf_vertex_probe(const HfView& hv, const double* p, QWallHit& h) {
    const int i = (int)floor(p[0] / hv.size_x + 0.5), j = (int)floor(-p[1] / hv.size_y + 0.5);
    if (i < 0 || i >= hv.hc || j < 0 || j >= hv.vc) return false;
    const unsigned cid = hv.index[(hv.vc - 1 - j) * hv.hc + i];
    if (cid == HF_EMPTY) return false;
    const uint32_t* m = hv.chunks + cid * HF_CHUNK_WORDS;
    const int ncol = m[13] & 0xFF, nrow = (m[13] >> 8) & 0xFF;
    const double sx = 0.5 * hv.size_x, sy = 0.5 * hv.size_y;
    const double dx = 2 * sx / (ncol - 1), dy = 2 * sy / (nrow - 1);
    const double u = (p[0] - hv.size_x * i + sx) / dx, vv = (p[1] + hv.size_y * j + sy) / dy;
    int cc = (int)floor(u), rr = (int)floor(vv);
    cc = cc < 0 ? 0 : (cc > ncol - 2 ? ncol - 2 : cc); rr = rr < 0 ? 0 : (rr > nrow - 2 ? nrow - 2 : rr);
    const double fu = u - cc, fv = vv - rr;
    const uint32_t la = hf_row_bits(m, ncol, rr, cc, 2), lb = hf_row_bits(m, ncol, rr + 1, cc, 2);
    if ((la | lb) == 0) return false;                                        // flat floor cell: below the ground plane
    const double z00 = (la & 1u) * HF_ELEV, z10 = (la >> 1) * HF_ELEV, z01 = (lb & 1u) * HF_ELEV, z11 = (lb >> 1) * HF_ELEV;
    double gx, gy, z;
    if (fv <= fu) { gx = (z10 - z00) / dx; gy = (z11 - z10) / dy; z = z00 + (z10 - z00) * fu + (z11 - z10) * fv; }
    else { gx = (z11 - z01) / dx; gy = (z01 - z00) / dy; z = z00 + (z11 - z01) * fu + (z01 - z00) * fv; }
    const double nn = sqrt(gx * gx + gy * gy + 1);
    h.nrm[0] = -gx / nn; h.nrm[1] = -gy / nn; h.nrm[2] = 1 / nn;
    const double hh = HF_BASE_Z + z;
    if (hh <= HF_BASE_Z + 1e-12 && h.nrm[2] > 0.999999) return false;        // flat floor cell: below the ground plane
    h.dist = (p[2] - hh) * h.nrm[2];
    if (h.dist >= 0) return false;
    hf_frame(h, p);
    return true;
}

// support point (world) of an ellipsoid with semi-axes size[0..2] / a cylinder with radius size[0], half height size[1]
// about its local z, in the world direction dir
enum { HF_ELLIPSOID = 0, HF_CYLINDER = 1, HF_SPHERE = 2 };       // sphere: radius size[0]
FT_HD void hf_support(int kind, const double* size, const double* pos, const double* R, const double* dir, double* out) {
    const double d0 = R[0] * dir[0] + R[3] * dir[1] + R[6] * dir[2], d1 = R[1] * dir[0] + R[4] * dir[1] + R[7] * dir[2],
                 d2 = R[2] * dir[0] + R[5] * dir[1] + R[8] * dir[2];
    double s[3];
    if (kind == HF_ELLIPSOID) {
        s[0] = size[0] * d0; s[1] = size[1] * d1; s[2] = size[2] * d2;
        const double n = sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
        s[0] = size[0] * s[0] / n; s[1] = size[1] * s[1] / n; s[2] = size[2] * s[2] / n;
    } else if (kind == HF_SPHERE) {
        const double n = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        s[0] = size[0] * d0 / n; s[1] = size[0] * d1 / n; s[2] = size[0] * d2 / n;
    } else {
        const double hn = sqrt(d0 * d0 + d1 * d1);
        s[0] = hn > MINVAL ? size[0] * d0 / hn : 0.0; s[1] = hn > MINVAL ? size[0] * d1 / hn : 0.0;
        s[2] = d2 >= 0 ? size[1] : -size[1];
    }
    mat_vec3(out, R, s);
    for (int a = 0; a < 3; a++) out[a] += pos[a];
}

// rule S: the geom's one wall contact (deepest candidate), or false
FT_HDNI bool hf_convex(const HfView& hv, int kind, const double* size, double bound, const double* pos, const double* R, QWallHit& h) {
    bool found = false;
    double best = 0, bp[3] = {0, 0, 0};
    const int i0 = (int)floor((pos[0] - bound) / hv.size_x + 0.5), i1 = (int)floor((pos[0] + bound) / hv.size_x + 0.5);
    const int j0 = (int)floor(-(pos[1] + bound) / hv.size_y + 0.5), j1 = (int)floor(-(pos[1] - bound) / hv.size_y + 0.5);
    for (int i = i0; i <= i1; i++)
        for (int j = j0; j <= j1; j++) {
            if (i < 0 || i >= hv.hc || j < 0 || j >= hv.vc) continue;
            const unsigned cid = hv.index[(hv.vc - 1 - j) * hv.hc + i];
            if (cid == HF_EMPTY) continue;
            const uint32_t* m = hv.chunks + cid * HF_CHUNK_WORDS;
            const int ncol = m[13] & 0xFF, nrow = (m[13] >> 8) & 0xFF;
            const uint32_t bb = m[14];                               // bounding box of the raised vertices
            const int cmin = bb & 0xFF, cmax = (bb >> 8) & 0xFF, rmin = (bb >> 16) & 0xFF, rmax = bb >> 24;
            if (cmin > cmax) continue;
            const double dx = hv.size_x / (ncol - 1), dy = hv.size_y / (nrow - 1);
            const double x0 = hv.size_x * i - 0.5 * hv.size_x, y0 = -hv.size_y * j - 0.5 * hv.size_y;
            int c0 = (int)floor((pos[0] - bound - x0) / dx), c1 = (int)floor((pos[0] + bound - x0) / dx);
            int r0 = (int)floor((pos[1] - bound - y0) / dy), r1 = (int)floor((pos[1] + bound - y0) / dy);
            // only cells with a raised vertex carry a triangle that can be touched: cells cmin-1 .. cmax, rmin-1 .. rmax
            c0 = c0 < cmin - 1 ? cmin - 1 : c0; c0 = c0 < 0 ? 0 : c0;
            r0 = r0 < rmin - 1 ? rmin - 1 : r0; r0 = r0 < 0 ? 0 : r0;
            c1 = c1 > cmax ? cmax : c1; c1 = c1 > ncol - 2 ? ncol - 2 : c1;
            r1 = r1 > rmax ? rmax : r1; r1 = r1 > nrow - 2 ? nrow - 2 : r1;
            if (c0 > c1) continue;
            for (int rr = r0; rr <= r1; rr++) {
                // vertex bits of rows rr and rr + 1 over columns c0 .. c1 + 1: two loads per row instead of four per cell
                const uint32_t la = hf_row_bits(m, ncol, rr, c0, c1 - c0 + 2), lb = hf_row_bits(m, ncol, rr + 1, c0, c1 - c0 + 2);
                if ((la | lb) == 0) continue;
                for (int cc = c0; cc <= c1; cc++) {
                    const uint32_t qa = (la >> (cc - c0)) & 3u, qb = (lb >> (cc - c0)) & 3u;
                    if ((qa | qb) == 0) continue;
                    const double z00 = (qa & 1u) * HF_ELEV, z10 = (qa >> 1) * HF_ELEV, z01 = (qb & 1u) * HF_ELEV, z11 = (qb >> 1) * HF_ELEV;
                    for (int tri = 0; tri < 2; tri++) {
                        double gx, gy;
                        if (tri == 0) { if (z00 + z10 + z11 == 0) continue; gx = (z10 - z00) / dx; gy = (z11 - z10) / dy; }
                        else { if (z00 + z01 + z11 == 0) continue; gx = (z11 - z01) / dx; gy = (z01 - z00) / dy; }
                        const double nn = sqrt(gx * gx + gy * gy + 1);
                        const double dir[3] = {gx / nn, gy / nn, -1 / nn};
                        double sp[3];
                        hf_support(kind, size, pos, R, dir, sp);
                        const double fu = (sp[0] - x0) / dx - cc, fv = (sp[1] - y0) / dy - rr;
                        const bool inside = tri == 0 ? (fv >= 0 && fv <= fu && fu <= 1) : (fu >= 0 && fu <= fv && fv <= 1);
                        if (!inside) continue;
                        const double zs = HF_BASE_Z + z00 + gx * fu * dx + gy * fv * dy;
                        const double dist = (sp[2] - zs) / nn;
                        if (dist >= 0 || (found && dist >= best)) continue;
                        found = true; best = dist;
                        h.nrm[0] = -dir[0]; h.nrm[1] = -dir[1]; h.nrm[2] = -dir[2]; h.dist = dist;
                        bp[0] = sp[0]; bp[1] = sp[1]; bp[2] = sp[2];
                    }
                }
            }
        }
    if (found) hf_frame(h, bp);
    return found;
}

// ground plane z = PLANE_Z (mushr.em.xml:94), normal +z: contact of a point / support point below it
FT_HD bool ground_probe(const double* p, QWallHit& h) {
    if (!(p[2] - PLANE_Z < 0)) return false;
    h.nrm[0] = 0; h.nrm[1] = 0; h.nrm[2] = 1; h.dist = p[2] - PLANE_Z;
    hf_frame(h, p);
    return true;
}

}  // namespace mushr
}  // namespace ftgp
