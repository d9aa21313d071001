/* oracle/oracle_internal.h -- TEST INFRASTRUCTURE (see ftgp_oracle.h). */
#ifndef FTGP_ORACLE_INTERNAL_H
#define FTGP_ORACLE_INTERNAL_H
#include "ftgp_oracle.h"

#define FTO_MINVAL 1e-15 /* mjMINVAL */

typedef struct {
    int i, j, ncol, nrow;
    float* data;      /* [nrow][ncol], row 0 = -y edge (image bottom row) */
    double pos[3];    /* geom pos */
    double size[4];   /* hfield size: half x, half y, elevation range, base */
} fto_chunk;

struct fto_track {
    int w, h, chunk_px, hc, vc, nchunks;
    double scale, size_x, size_y;
    fto_chunk* chunks;
    int* index;       /* [hc*vc] i*vc + j -> chunk id or -1 */
};

/* ray helpers shared by ray.c and (for wall contacts) step.c */
double fto_ray_hfield(const fto_chunk* c, const double* pnt, const double* vec);
double fto_hfield_height(const fto_track* t, double x, double y); /* world z of wall surface, -0.1 if none */

#endif
