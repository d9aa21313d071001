/*
 * oracle/ray.c -- TEST INFRASTRUCTURE (see ftgp_oracle.h header).  "parity unpinned".
 *
 * Restates mujoco's mj_ray as used by the 90 rangefinder sensors of
 * template/mushr.em.xml:112-117,204-206 (read at ft_grandprix/custom.py:1395):
 * ray vs every lidar-visible geom not on the car's own root body, smallest
 * non-negative distance, -1 when nothing is hit, no cutoff.
 * Geoms: the hfield walls (mushr.em.xml:92), the ground plane (mushr.em.xml:94),
 * and in multi-car worlds the other cars' lidar cylinder (mushr.em.xml:108) and
 * wheel ellipsoids (mushr.em.xml:69,132).
 */
#include "oracle_internal.h"
#include "mushr_mesh.h"
#include <math.h>
#include <string.h>

static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(double* r, const double* a, const double* b) {
    r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0];
}
static void normalize3(double* a) {
    double n = sqrt(dot3(a, a));
    if (n < FTO_MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; } else { a[0] /= n; a[1] /= n; a[2] /= n; }
}

/* ray vs axis-aligned box centred at the local origin; all[6] = per-face solutions or -1
 * (faces: -x,+x,-y,+y,-z,+z).  Returns nearest non-negative solution or -1. */
static double ray_box(const double* size, const double* lpnt, const double* lvec, double* all) {
    static const int IFACE[3][2] = {{1, 2}, {0, 2}, {0, 1}};
    double x = -1;
    if (all) for (int i = 0; i < 6; i++) all[i] = -1;
    for (int i = 0; i < 3; i++) {
        if (fabs(lvec[i]) > FTO_MINVAL) {
            for (int side = -1; side <= 1; side += 2) {
                double sol = (side * size[i] - lpnt[i]) / lvec[i];
                if (sol >= 0) {
                    int id0 = IFACE[i][0], id1 = IFACE[i][1];
                    double p0 = lpnt[id0] + sol * lvec[id0], p1 = lpnt[id1] + sol * lvec[id1];
                    if (fabs(p0) <= size[id0] && fabs(p1) <= size[id1]) {
                        if (x < 0 || sol < x) x = sol;
                        if (all) all[2 * i + (side + 1) / 2] = sol;
                    }
                }
            }
        }
    }
    return x;
}

static double ray_triangle(double v[3][3], const double* lpnt, const double* lvec,
                           const double* b0, const double* b1) {
    double dif[3][3], planar[3][2];
    for (int i = 0; i < 3; i++) {
        for (int k = 0; k < 3; k++) dif[i][k] = v[i][k] - lpnt[k];
        planar[i][0] = dot3(b0, dif[i]);
        planar[i][1] = dot3(b1, dif[i]);
    }
    if ((planar[0][0] > 0 && planar[1][0] > 0 && planar[2][0] > 0) ||
        (planar[0][0] < 0 && planar[1][0] < 0 && planar[2][0] < 0) ||
        (planar[0][1] > 0 && planar[1][1] > 0 && planar[2][1] > 0) ||
        (planar[0][1] < 0 && planar[1][1] < 0 && planar[2][1] < 0))
        return -1;
    double A[4] = {planar[0][0] - planar[2][0], planar[1][0] - planar[2][0],
                   planar[0][1] - planar[2][1], planar[1][1] - planar[2][1]};
    double b[2] = {-planar[2][0], -planar[2][1]};
    double det = A[0] * A[3] - A[1] * A[2];
    if (fabs(det) < FTO_MINVAL) return -1;
    double t0 = (A[3] * b[0] - A[1] * b[1]) / det;
    double t1 = (-A[2] * b[0] + A[0] * b[1]) / det;
    if (t0 < 0 || t1 < 0 || t0 + t1 > 1) return -1;
    double e0[3], e1[3], e2[3], nrm[3];
    for (int k = 0; k < 3; k++) { e0[k] = v[0][k] - v[2][k]; e1[k] = v[1][k] - v[2][k]; e2[k] = lpnt[k] - v[2][k]; }
    cross3(nrm, e0, e1);
    double denom = dot3(lvec, nrm);
    if (fabs(denom) < FTO_MINVAL) return -1;
    return -dot3(e2, nrm) / denom;
}

/* mj_rayHfield restated; hfield frames are axis aligned here. */
double fto_ray_hfield(const fto_chunk* c, const double* pnt, const double* vec) {
    const double* size = c->size;
    int nrow = c->nrow, ncol = c->ncol;
    const float* data = c->data;
    /* base box: z in [-size[3], 0]; top box: z in [0, size[2]] (hfield local frame) */
    double base_size[3] = {size[0], size[1], size[3] * 0.5};
    double top_size[3] = {size[0], size[1], size[2] * 0.5};
    double lb[3] = {pnt[0] - c->pos[0], pnt[1] - c->pos[1], pnt[2] - (c->pos[2] - size[3] * 0.5)};
    double x = ray_box(base_size, lb, vec, 0);
    double lt[3] = {pnt[0] - c->pos[0], pnt[1] - c->pos[1], pnt[2] - (c->pos[2] + size[2] * 0.5)};
    double all[6];
    double top_intersect = ray_box(top_size, lt, vec, all);
    if (top_intersect < 0) return x;
    double lpnt[3] = {pnt[0] - c->pos[0], pnt[1] - c->pos[1], pnt[2] - c->pos[2]};
    const double* lvec = vec;
    /* basis of the plane normal to the ray */
    double b0[3] = {1, 1, 1}, b1[3];
    if (fabs(lvec[0]) >= fabs(lvec[1]) && fabs(lvec[0]) >= fabs(lvec[2])) b0[0] = 0;
    else if (fabs(lvec[1]) >= fabs(lvec[2])) b0[1] = 0;
    else b0[2] = 0;
    double s = -dot3(lvec, b0) / dot3(lvec, lvec);
    for (int k = 0; k < 3; k++) b1[k] = b0[k] + s * lvec[k];
    normalize3(b1);
    cross3(b0, b1, lvec);
    normalize3(b0);
    /* ray segment inside the top box */
    double seg[2] = {0, top_intersect};
    for (int i = 0; i < 6; i++)
        if (all[i] > seg[1]) { seg[0] = top_intersect; seg[1] = all[i]; }
    double dx = (2.0 * size[0]) / (ncol - 1), dy = (2.0 * size[1]) / (nrow - 1);
    double SX[2], SY[2];
    for (int i = 0; i < 2; i++) {
        SX[i] = (lpnt[0] + seg[i] * lvec[0] + size[0]) / dx;
        SY[i] = (lpnt[1] + seg[i] * lvec[1] + size[1]) / dy;
    }
    int cmin = (int)floor(fmin(SX[0], SX[1])) - 1; if (cmin < 0) cmin = 0;
    int cmax = (int)ceil(fmax(SX[0], SX[1])) + 1; if (cmax > ncol - 1) cmax = ncol - 1;
    int rmin = (int)floor(fmin(SY[0], SY[1])) - 1; if (rmin < 0) rmin = 0;
    int rmax = (int)ceil(fmax(SY[0], SY[1])) + 1; if (rmax > nrow - 1) rmax = nrow - 1;
    for (int r = rmin; r < rmax; r++) {
        for (int cc = cmin; cc < cmax; cc++) {
            double va[3][3] = {
                {dx * cc - size[0], dy * r - size[1], data[r * ncol + cc] * size[2]},
                {dx * (cc + 1) - size[0], dy * (r + 1) - size[1], data[(r + 1) * ncol + (cc + 1)] * size[2]},
                {dx * (cc + 1) - size[0], dy * r - size[1], data[r * ncol + (cc + 1)] * size[2]}};
            double sol = ray_triangle(va, lpnt, lvec, b0, b1);
            if (sol >= 0 && (x < 0 || sol < x)) x = sol;
            double vb[3][3] = {
                {dx * cc - size[0], dy * r - size[1], data[r * ncol + cc] * size[2]},
                {dx * (cc + 1) - size[0], dy * (r + 1) - size[1], data[(r + 1) * ncol + (cc + 1)] * size[2]},
                {dx * cc - size[0], dy * (r + 1) - size[1], data[(r + 1) * ncol + cc] * size[2]}};
            sol = ray_triangle(vb, lpnt, lvec, b0, b1);
            if (sol >= 0 && (x < 0 || sol < x)) x = sol;
        }
    }
    /* vertical sides of the top box: solid below the boundary elevation profile */
    for (int i = 0; i < 4; i++) {
        if (all[i] >= 0 && (all[i] < x || x < 0)) {
            double z = (lpnt[2] + all[i] * lvec[2]) / size[2];
            double y, y0, z0, z1;
            if (i < 2) {
                y = (lpnt[1] + all[i] * lvec[1] + size[1]) / dy;
                y0 = fmax(0, fmin(nrow - 2, floor(y)));
                z0 = data[(int)lround(y0) * ncol + (i == 1 ? ncol - 1 : 0)];
                z1 = data[(int)lround(y0 + 1) * ncol + (i == 1 ? ncol - 1 : 0)];
            } else {
                y = (lpnt[0] + all[i] * lvec[0] + size[0]) / dx;
                y0 = fmax(0, fmin(ncol - 2, floor(y)));
                z0 = data[(int)lround(y0) + (i == 3 ? (nrow - 1) * ncol : 0)];
                z1 = data[(int)lround(y0 + 1) + (i == 3 ? (nrow - 1) * ncol : 0)];
            }
            if (z < z0 * (y0 + 1 - y) + z1 * (y - y0)) x = all[i];
        }
    }
    return x;
}

/* ground plane, mushr.em.xml:94: pos (0,0,0.01), size 300 300; front side only */
static double ray_plane(const double* pnt, const double* vec) {
    double lz = pnt[2] - 0.01;
    if (vec[2] > -FTO_MINVAL) return -1;
    double x = -lz / vec[2];
    if (x < 0) return -1;
    double p0 = pnt[0] + x * vec[0], p1 = pnt[1] + x * vec[1];
    if (fabs(p0) <= 300.0 && fabs(p1) <= 300.0) return x;
    return -1;
}

double fto_ray(const fto_track* t, const double* pnt, const double* vec) {
    double best = ray_plane(pnt, vec);
    if (!t) return best;
    /* every hfield geom, with a conservative cull: the horizontal footprint of the ray
     * must come within the chunk's half-diagonal (MuJoCo culls by bounding sphere). */
    double hx = t->size_x / 2, hy = t->size_y / 2;
    double vn2 = vec[0] * vec[0] + vec[1] * vec[1];
    for (int k = 0; k < t->nchunks; k++) {
        const fto_chunk* c = &t->chunks[k];
        double ox = c->pos[0] - pnt[0], oy = c->pos[1] - pnt[1];
        double rad = hx + hy;
        if (vn2 > 1e-30) {
            double along = (ox * vec[0] + oy * vec[1]);
            double perp2 = ox * ox + oy * oy - along * along / vn2;
            if (perp2 > rad * rad) continue;
            if (along < 0 && ox * ox + oy * oy > rad * rad) continue;
        } else if (fabs(ox) > hx || fabs(oy) > hy) continue;
        double x = fto_ray_hfield(c, pnt, vec);
        if (x >= 0 && (best < 0 || x < best)) best = x;
    }
    return best;
}

/* ------------------------------------------------------------ rangefinder sites */
static void quat_rot(const double* q, const double* v, double* r) {
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
    double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                   2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                   2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
    for (int i = 0; i < 3; i++) r[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
}

/* site j of mushr.em.xml:112-117: pos (rx + lr sin(theta), lr cos(theta), rz),
 * theta = -radians(4j-90); +Z axis = (sin b, -cos b, 0), b = radians(4j-90). */
static void beam_local(int j, double* o, double* d) {
    const double rx = -0.0525, rz = 0.065, lr = 0.030;
    const double PI = 3.14159265358979323846;
    double b = (360.0 / FTO_NBEAMS * j - 90.0) * PI / 180.0;   /* radians(inter_ray_angle*j - 90) */
    double theta = -b;
    o[0] = rx + lr * sin(theta); o[1] = lr * cos(theta); o[2] = rz;
    d[0] = sin(b); d[1] = -cos(b); d[2] = 0;
}

/* --- other cars' geoms (multi-car worlds) --- */
static double ray_quadric_cyl(const double* lp, const double* lv, double r, double hh) {
    /* cylinder along local z, radius r, half height hh (mju_rayGeom mjGEOM_CYLINDER) */
    double best = -1;
    double a = lv[0] * lv[0] + lv[1] * lv[1], b = lp[0] * lv[0] + lp[1] * lv[1], c = lp[0] * lp[0] + lp[1] * lp[1] - r * r;
    if (a > FTO_MINVAL) {
        double det = b * b - a * c;
        if (det >= FTO_MINVAL) {
            det = sqrt(det);
            double x0 = (-b - det) / a, x1 = (-b + det) / a;
            double sol = x0 >= 0 ? x0 : (x1 >= 0 ? x1 : -1);      /* ray_quad: first non-negative root */
            if (sol >= 0 && fabs(lp[2] + sol * lv[2]) <= hh) best = sol;
        }
    }
    if (fabs(lv[2]) > FTO_MINVAL)
        for (int side = -1; side <= 1; side += 2) {
            double x = (side * hh - lp[2]) / lv[2];
            if (x >= 0) {
                double p0 = lp[0] + x * lv[0], p1 = lp[1] + x * lv[1];
                if (p0 * p0 + p1 * p1 <= r * r && (best < 0 || x < best)) best = x;
            }
        }
    return best;
}

/* ellipsoid with semi-axes size[], ray in the geom frame (mju_rayGeom mjGEOM_ELLIPSOID): quadratic in scaled space */
static double ray_ellipsoid(const double* lp, const double* lv, const double* size) {
    double a = 0, b = 0, c = -1;
    for (int k = 0; k < 3; k++) {
        const double s = 1.0 / (size[k] * size[k]);
        a += s * lv[k] * lv[k]; b += s * lp[k] * lv[k]; c += s * lp[k] * lp[k];
    }
    if (a < FTO_MINVAL) return -1;
    double det = b * b - a * c;
    if (det < FTO_MINVAL) return -1;
    det = sqrt(det);
    const double x0 = (-b - det) / a, x1 = (-b + det) / a;
    return x0 >= 0 ? x0 : (x1 >= 0 ? x1 : -1);
}

/* one mesh triangle (mj_rayMesh -> ray_triangle): nearest non-negative hit or -1 */
static double ray_tri(const double* v0, const double* v1, const double* v2, const double* p, const double* d) {
    const double e1[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]}, e2[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
    const double h[3] = {d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0]};
    const double det = e1[0] * h[0] + e1[1] * h[1] + e1[2] * h[2];
    if (fabs(det) < 1e-300) return -1;
    const double tv[3] = {p[0] - v0[0], p[1] - v0[1], p[2] - v0[2]};
    const double u = (tv[0] * h[0] + tv[1] * h[1] + tv[2] * h[2]) / det;
    if (u < 0 || u > 1) return -1;
    const double q[3] = {tv[1] * e1[2] - tv[2] * e1[1], tv[2] * e1[0] - tv[0] * e1[2], tv[0] * e1[1] - tv[1] * e1[0]};
    const double v = (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]) / det;
    if (v < 0 || u + v > 1) return -1;
    const double x = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) / det;
    return x >= 0 ? x : -1;
}

/* Rangefinders of car `self` in an N-car world: mj_ray with bodyexclude = own root body against the walls, the ground and
 * every lidar-visible geom of the OTHER cars: the lidar cylinder (mushr.em.xml:108), the chassis mesh's own
 * triangles (:119) and the four wheel ellipsoids (:69; semi-axes 0.03, 0.01, 0.03 -- invariant under the throttle
 * rotation about y, so only suspension travel and steering angle matter).  qpos_all: [ncars][stride]; with stride >= 34
 * the wheel poses come from the joint values, with a bare pose (stride 7) the joints sit at qpos0.  (The car's OWN wheels
 * are not excluded by mj_ray but lie below every beam, A.1; they are not tested.) */
void fto_lidar_scan_world_stride(const fto_track* t, const double* qpos_all, int stride, int ncars, int self,
                                 const uint8_t* visible, double* out) {
    static const double tri[MUSHR_CHASSIS_NTRI][9] = MUSHR_CHASSIS_TRI;
    const double wsize[3] = {0.03, 0.01, 0.03};
    const int qsusp[4] = {8, 15, 22, 28};
    const double wpos[4][3] = {{0.06925, 0.0575, 0.0244}, {0.06925, -0.0575, 0.0244}, {-0.079, 0.0575, 0.0244}, {-0.079, -0.0575, 0.0244}};
    const double* pose = qpos_all + (size_t)self * stride;
    for (int j = 0; j < FTO_NBEAMS; j++) {
        double o[3], d[3], ow[3], dw[3];
        beam_local(j, o, d);
        quat_rot(pose + 3, o, ow);
        quat_rot(pose + 3, d, dw);
        for (int k = 0; k < 3; k++) ow[k] += pose[k];
        double best = fto_ray(t, ow, dw);
        for (int c = 0; c < ncars; c++) {
            if (c == self || (visible && !visible[c])) continue;
            const double* pc = qpos_all + (size_t)c * stride;
            double qi[4] = {pc[3], -pc[4], -pc[5], -pc[6]};
            double rel[3] = {ow[0] - pc[0], ow[1] - pc[1], ow[2] - pc[2]}, lp[3], lv[3];
            quat_rot(qi, rel, lp); quat_rot(qi, dw, lv);                  /* the ray in the other car's frame */
            {                                                             /* lidar cylinder: pos (rx, 0, rz - lh/2), r 0.03, hh 0.015 */
                double cp[3] = {lp[0] + 0.0525, lp[1], lp[2] - (0.065 - 0.015 / 2)};
                double x = ray_quadric_cyl(cp, lv, 0.03, 0.015);
                if (x >= 0 && (best < 0 || x < best)) best = x;
            }
            for (int f = 0; f < MUSHR_CHASSIS_NTRI; f++) {                /* chassis mesh */
                double x = ray_tri(tri[f], tri[f] + 3, tri[f] + 6, lp, lv);
                if (x >= 0 && (best < 0 || x < best)) best = x;
            }
            for (int w = 0; w < 4; w++) {                                 /* wheel ellipsoids */
                const double susp = stride >= FTO_NQ ? pc[qsusp[w]] : 0.0, steer = (stride >= FTO_NQ && w < 2) ? pc[qsusp[w] + 1] : 0.0;
                const double cx = wpos[w][0], cy = wpos[w][1], cz = wpos[w][2] + susp, cs = cos(steer), sn = sin(steer);
                const double rp[3] = {lp[0] - cx, lp[1] - cy, lp[2] - cz};
                const double wp[3] = {cs * rp[0] + sn * rp[1], -sn * rp[0] + cs * rp[1], rp[2]};      /* Rz(steer)^T */
                const double wv[3] = {cs * lv[0] + sn * lv[1], -sn * lv[0] + cs * lv[1], lv[2]};
                double x = ray_ellipsoid(wp, wv, wsize);
                if (x >= 0 && (best < 0 || x < best)) best = x;
            }
        }
        out[j] = best;
    }
}

void fto_lidar_scan_world(const fto_track* t, const double* qpos_all, int ncars, int self,
                          const uint8_t* visible, double* out) {
    fto_lidar_scan_world_stride(t, qpos_all, 7, ncars, self, visible, out);
}

void fto_lidar_scan(const fto_track* t, const double* pose, double* out) {
    fto_lidar_scan_world(t, pose, 1, 0, 0, out);
}
