/*
 * oracle/track.c -- TEST INFRASTRUCTURE (see ftgp_oracle.h header).
 *
 * Restates the reference's track compiler for the hot path:
 *   ft_grandprix/chunk.py:39-64   threshold + 20x20 chunk scan (column-major i,j)
 *   template/mushr.em.xml:19-20   size_x = 20*scale/horizontal_chunks, size_y likewise
 *   template/mushr.em.xml:55      hfield size = (size_x/2, size_y/2, 0.3, 1e-4)
 *   template/mushr.em.xml:92      hfield geom pos = (size_x*i, -size_y*j, -0.1)
 * and MuJoCo's PNG->hfield load (rows flipped so the image top is +y; elevation
 * normalised (v - min)/(max - min), a constant image becoming all zeros).
 * Centreline: ft_grandprix/curve.py:6-18 with svg.path 6.3 Path.point()/length()
 * semantics, scaled as in ft_grandprix/custom.py:1184-1186.
 */
#include "oracle_internal.h"
#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

fto_track* fto_track_create(const uint8_t* wall, int w, int h, double scale, int chunk_px) {
    if (!wall || w <= 0 || h <= 0 || chunk_px < 2) return NULL;
    fto_track* t = (fto_track*)calloc(1, sizeof(fto_track));
    t->w = w; t->h = h; t->chunk_px = chunk_px; t->scale = scale;
    t->hc = (w + chunk_px - 1) / chunk_px;           /* chunk.py:45 ceil */
    t->vc = (h + chunk_px - 1) / chunk_px;           /* chunk.py:46 */
    t->size_x = 20.0 * scale / t->hc;                /* mushr.em.xml:17,19 */
    t->size_y = 20.0 * scale / t->vc;                /* mushr.em.xml:17,20 */
    t->chunks = (fto_chunk*)calloc((size_t)t->hc * t->vc, sizeof(fto_chunk));
    t->index = (int*)malloc(sizeof(int) * (size_t)t->hc * t->vc);
    for (int k = 0; k < t->hc * t->vc; k++) t->index[k] = -1;
    int n = 0;
    for (int i = 0; i < t->hc; i++) {                /* chunk.py:49 outer loop over columns */
        for (int j = 0; j < t->vc; j++) {            /* chunk.py:50 */
            int x0 = i * chunk_px, y0 = j * chunk_px;
            int x1 = x0 + chunk_px < w ? x0 + chunk_px : w;
            int y1 = y0 + chunk_px < h ? y0 + chunk_px : h;
            int ncol = x1 - x0, nrow = y1 - y0, any = 0, all = 1;
            for (int y = y0; y < y1; y++)
                for (int x = x0; x < x1; x++) {
                    if (wall[(size_t)y * w + x]) any = 1; else all = 0;
                }
            if (!any) continue;                      /* chunk.py:59 np_cropped.sum() > 0 */
            fto_chunk* c = &t->chunks[n];
            c->i = i; c->j = j; c->ncol = ncol; c->nrow = nrow;
            c->data = (float*)calloc((size_t)ncol * nrow, sizeof(float));
            /* MuJoCo hfield from PNG: row r of the hfield is image row (nrow-1-r);
             * normalise: constant image -> zeros */
            if (!all)
                for (int r = 0; r < nrow; r++)
                    for (int cc = 0; cc < ncol; cc++)
                        c->data[r * ncol + cc] =
                            wall[(size_t)(y0 + (nrow - 1 - r)) * w + x0 + cc] ? 1.0f : 0.0f;
            c->pos[0] = t->size_x * i;               /* mushr.em.xml:92 */
            c->pos[1] = -t->size_y * j;
            c->pos[2] = -0.1;
            c->size[0] = t->size_x / 2; c->size[1] = t->size_y / 2;   /* mushr.em.xml:55 */
            c->size[2] = 0.2 + 0.1; c->size[3] = 0.0001;
            t->index[i * t->vc + j] = n;
            n++;
        }
    }
    t->nchunks = n;
    return t;
}

void fto_track_destroy(fto_track* t) {
    if (!t) return;
    for (int k = 0; k < t->nchunks; k++) free(t->chunks[k].data);
    free(t->chunks); free(t->index); free(t);
}
int fto_track_nchunks(const fto_track* t) { return t->nchunks; }
void fto_track_dims(const fto_track* t, int* hc, int* vc, double* sx, double* sy) {
    *hc = t->hc; *vc = t->vc; *sx = t->size_x; *sy = t->size_y;
}
void fto_track_chunks(const fto_track* t, int32_t* out) {
    for (int k = 0; k < t->nchunks; k++) { out[2 * k] = t->chunks[k].i; out[2 * k + 1] = t->chunks[k].j; }
}
void fto_track_chunk_data(const fto_track* t, int k, float* out, int* nrow, int* ncol) {
    const fto_chunk* c = &t->chunks[k];
    *nrow = c->nrow; *ncol = c->ncol;
    memcpy(out, c->data, sizeof(float) * (size_t)c->nrow * c->ncol);
}

/* ------------------------------------------------------------------ centreline
 * svg.path 6.3: parse_path -> [Move, CubicBezier|Line ..., Close]; Path.point(pos).
 */
typedef struct { int kind; /*0 move,1 line/close,2 cubic*/ double p[8]; double len; } seg_t;

static void seg_point(const seg_t* s, double u, double* x, double* y) {
    if (s->kind == 0) { *x = s->p[0]; *y = s->p[1]; return; }
    if (s->kind == 1) {                              /* Linear.point: start + (end-start)*pos */
        *x = s->p[0] + (s->p[2] - s->p[0]) * u; *y = s->p[1] + (s->p[3] - s->p[1]) * u; return;
    }
    /* CubicBezier.point */
    double a = (1 - u) * (1 - u) * (1 - u), b = 3 * (1 - u) * (1 - u) * u,
           c = 3 * (1 - u) * u * u, d = u * u * u;
    *x = a * s->p[0] + b * s->p[2] + c * s->p[4] + d * s->p[6];
    *y = a * s->p[1] + b * s->p[3] + c * s->p[5] + d * s->p[7];
}

/* svg.path segment_length(): recursive midpoint subdivision, error 1e-12, min_depth 5 */
static double seg_len_rec(const seg_t* s, double a, double b, double ax, double ay,
                          double bx, double by, int depth) {
    double m = (a + b) / 2, mx, my;
    seg_point(s, m, &mx, &my);
    double length = hypot(bx - ax, by - ay);
    double l2 = hypot(mx - ax, my - ay) + hypot(bx - mx, by - my);
    if ((l2 - length > 1e-12 || depth < 5) && depth < 40) {
        depth++;
        return seg_len_rec(s, a, m, ax, ay, mx, my, depth) + seg_len_rec(s, m, b, mx, my, bx, by, depth);
    }
    return l2;
}
static double seg_length(const seg_t* s) {
    if (s->kind == 0) return 0.0;
    if (s->kind == 1) return hypot(s->p[2] - s->p[0], s->p[3] - s->p[1]);
    double ax, ay, bx, by;
    seg_point(s, 0, &ax, &ay); seg_point(s, 1, &bx, &by);
    return seg_len_rec(s, 0, 1, ax, ay, bx, by, 0);
}

static int read_num(const char** pp, double* v) {
    const char* p = *pp;
    while (*p && (isspace((unsigned char)*p) || *p == ',')) p++;
    if (!*p) return 0;
    char* e;
    double x = strtod(p, &e);
    if (e == p) return 0;
    *v = x; *pp = e; return 1;
}

int fto_centreline(const char* d, int npoints, int img_w, int img_h, int chunk_w, int chunk_h,
                   double scale, double* out) {
    int cap = 64, n = 0;
    seg_t* segs = (seg_t*)malloc(sizeof(seg_t) * cap);
    double cx = 0, cy = 0, sx = 0, sy = 0;
    const char* p = d;
    char cmd = 0;
    int ok = 1;
    while (ok) {
        while (*p && (isspace((unsigned char)*p) || *p == ',')) p++;
        if (!*p) break;
        if (isalpha((unsigned char)*p)) { cmd = *p++; if (cmd != 'z' && cmd != 'Z') continue; }
        if (n + 2 > cap) { cap *= 2; segs = (seg_t*)realloc(segs, sizeof(seg_t) * cap); }
        seg_t s; memset(&s, 0, sizeof s);
        int rel = islower((unsigned char)cmd);
        double v[6];
        switch (cmd) {
        case 'm': case 'M':
            if (!read_num(&p, &v[0]) || !read_num(&p, &v[1])) { ok = 0; break; }
            if (rel) { v[0] += cx; v[1] += cy; }
            cx = sx = v[0]; cy = sy = v[1];
            s.kind = 0; s.p[0] = cx; s.p[1] = cy; segs[n++] = s;
            cmd = rel ? 'l' : 'L';                    /* implicit lineto for further pairs */
            break;
        case 'l': case 'L':
            if (!read_num(&p, &v[0]) || !read_num(&p, &v[1])) { ok = 0; break; }
            if (rel) { v[0] += cx; v[1] += cy; }
            s.kind = 1; s.p[0] = cx; s.p[1] = cy; s.p[2] = v[0]; s.p[3] = v[1]; segs[n++] = s;
            cx = v[0]; cy = v[1];
            break;
        case 'c': case 'C':
            for (int k = 0; k < 6; k++) if (!read_num(&p, &v[k])) { ok = 0; break; }
            if (!ok) break;
            if (rel) for (int k = 0; k < 6; k += 2) { v[k] += cx; v[k + 1] += cy; }
            s.kind = 2; s.p[0] = cx; s.p[1] = cy;
            for (int k = 0; k < 6; k++) s.p[2 + k] = v[k];
            segs[n++] = s; cx = v[4]; cy = v[5];
            break;
        case 'z': case 'Z':
            s.kind = 1; s.p[0] = cx; s.p[1] = cy; s.p[2] = sx; s.p[3] = sy; segs[n++] = s;
            cx = sx; cy = sy; cmd = 0;
            break;
        default: ok = 0;
        }
        if (cmd == 0) { /* after close, expect a new command or end */ }
    }
    if (n == 0) { free(segs); return 1; }
    /* Path._calc_lengths */
    double total = 0;
    for (int k = 0; k < n; k++) { segs[k].len = seg_length(&segs[k]); total += segs[k].len; }
    double* frac = (double*)malloc(sizeof(double) * n);
    double f = 0;
    for (int k = 0; k < n; k++) { f += total > 0 ? segs[k].len / total : 0; frac[k] = f; }
    frac[n - 1] = 1.0;
    for (int q = 0; q < npoints; q++) {
        double pos = (double)q / (double)npoints, x, y;   /* curve.py:16 p.point(i/points) */
        if (pos == 0.0 || total == 0.0) seg_point(&segs[0], 0.0, &x, &y);
        else {
            int i = 0;                                /* bisect_right(fractions, pos) */
            while (i < n && frac[i] <= pos) i++;
            if (i >= n) i = n - 1;
            double u = i == 0 ? pos / frac[0] : (pos - frac[i - 1]) / (frac[i] - frac[i - 1]);
            seg_point(&segs[i], u, &x, &y);
        }
        out[2 * q] = x / img_w * chunk_w * scale;     /* custom.py:1185 */
        out[2 * q + 1] = -y / img_h * chunk_h * scale; /* custom.py:1186 */
    }
    free(frac); free(segs);
    return ok ? 0 : 2;
}
