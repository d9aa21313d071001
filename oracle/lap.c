/*
 * oracle/lap.c -- TEST INFRASTRUCTURE (see ftgp_oracle.h header).
 * Restates the progress / lap state machine of ft_grandprix/custom.py:1340-1372.
 */
#include "oracle_internal.h"

static int pymod(int a, int b) { int r = a % b; return r < 0 ? r + b : r; }

void fto_lap_update(fto_lap_state* s, int32_t* times, int max_times, const double* path,
                    const double* xy, int32_t steps, int32_t lap_target, int32_t* nwinners) {
    /* custom.py:1341-1344 */
    int closest = 0;
    double best = 0;
    for (int k = 0; k < FTO_NPATH; k++) {
        double dx = path[2 * k] - xy[0], dy = path[2 * k + 1] - xy[1];
        double d = dx * dx + dy * dy;
        if (k == 0 || d < best) { best = d; closest = k; }      /* argmin: first minimum */
    }
    s->off_track = best > 1;
    if (s->off_track) return;
    int completion = pymod(closest - s->offset, 100);            /* custom.py:1346 */
    int delta = completion - s->completion;                       /* custom.py:1347 */
    s->delta = pymod(completion - s->completion + 50, 100) - 50;  /* custom.py:1348 */
    if ((delta < 0 ? -delta : delta) > 90) {                      /* custom.py:1350 */
        int lap_steps = steps - s->start;                         /* custom.py:1351 (x timestep on host) */
        if (s->delta < 0) {                                       /* custom.py:1352-1356 */
            s->laps -= 1; s->good_start = 0;
            if (s->ntimes != 0) s->ntimes -= 1;
        } else if (s->delta > 0) {                                /* custom.py:1357-1366 */
            if (s->good_start) {
                if (s->ntimes < max_times) times[s->ntimes] = lap_steps;
                s->ntimes += 1;
                s->start = steps;
            }
            s->laps += 1; s->good_start = 1;
        }
    }
    if (s->laps >= lap_target) {                                  /* custom.py:1367-1371 */
        if (s->rank == 0) { *nwinners += 1; s->rank = *nwinners; }
        s->finished = 1;
    }
    s->completion = completion;                                   /* custom.py:1372 */
}
