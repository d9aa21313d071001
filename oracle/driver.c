/*
 * oracle/driver.c -- TEST INFRASTRUCTURE (see ftgp_oracle.h header).
 *
 * Restates the bundled "disparity extender" drivers with numpy's float64 semantics:
 *   ft_grandprix/nidc.py:12-131  (kind 0, CAR_WIDTH 0.12)
 *   ft_grandprix/fast.py:12-139  (kind 1, CAR_WIDTH 0.06, straight-line boost)
 *   ft_grandprix/lobotomy.py:1-3 (kind 2)
 * Pinned against tests/golden/drivers.npz (made by calling the reference classes).
 */
#include "oracle_internal.h"
#include <math.h>

#define MAXB 512

static int np_argmax(const double* a, int n) {      /* first max; first NaN wins */
    int best = 0;
    if (isnan(a[0])) return 0;
    for (int i = 1; i < n; i++) {
        if (isnan(a[i])) return i;
        if (a[i] > a[best]) best = i;
    }
    return best;
}

int fto_driver(int kind, const double* ranges, int n, double* speed, double* steer) {
    if (kind == 2) { *speed = 0; *steer = 0; return 0; }
    if (n > MAXB || n < 8) return 1;
    const double PI = 3.141592653589793;
    double rpp = (2 * PI) / n;                       /* nidc.py:121 */
    int eighth = (int)(n / 8.0);                     /* nidc.py:17 */
    int m = n - 2 * eighth;
    double proc[MAXB];
    for (int i = 0; i < m; i++) proc[i] = ranges[eighth + i];   /* nidc.py:18 (copy) */
    int disp[MAXB], nd = 0;
    for (int i = 1; i < m; i++)                      /* nidc.py:26-40, threshold 0.6 */
        if (fabs(proc[i] - proc[i - 1]) > 0.6) disp[nd++] = i;
    double car_width = kind == 0 ? 0.12 : 0.06;      /* nidc.py:5, fast.py:4 */
    double width = (car_width / 2) * (1 + 300. / 100);           /* nidc.py:93 */
    for (int k = 0; k < nd; k++) {                   /* nidc.py:94-104 */
        int first = disp[k] - 1;
        double p0 = proc[first], p1 = proc[first + 1];
        /* np.argmin / np.argmax over 2 elements: first NaN wins, ties -> index 0 */
        int amin = isnan(p0) ? 0 : (isnan(p1) ? 1 : (p1 < p0 ? 1 : 0));
        int amax = isnan(p0) ? 0 : (isnan(p1) ? 1 : (p1 > p0 ? 1 : 0));
        int close = first + amin, far = first + amax;
        double dist = proc[close];
        double angle = 2 * atan(width / (2 * dist)); /* nidc.py:57; x/0 -> inf as in numpy */
        double np_ = ceil(angle / rpp);              /* nidc.py:58 */
        if (isnan(np_)) return 1;                    /* int(nan) raises ValueError */
        long num = (long)np_;
        double nd_ = proc[close];
        if (close < far) {                           /* nidc.py:71-77 cover right */
            for (long i = 0; i < num; i++) {
                long idx = close + 1 + i;
                if (idx >= m) break;
                if (proc[idx] > nd_) proc[idx] = nd_;
            }
        } else {                                     /* nidc.py:78-83 cover left */
            for (long i = 0; i < num; i++) {
                long idx = close - 1 - i;
                if (idx < 0) break;
                if (proc[idx] > nd_) proc[idx] = nd_;
            }
        }
    }
    int best = np_argmax(proc, m);
    double ang = (best - (m / 2.0)) * rpp;           /* nidc.py:112 */
    double lim = 90.0 * (PI / 180.0);                /* np.radians(90) */
    double st = ang < -lim ? -lim : (ang > lim ? lim : ang);
    if (kind == 0) {
        *speed = 0.5 * 5 * (1 - fabs(st) / (1.57 * 2));          /* nidc.py:130 */
    } else {
        if (fabs(st) < 0.1 && ranges[0] > 0.5) *speed = 7;       /* fast.py:135-136 */
        else { double s = 0.5 * 5 * (1 - fabs(st) / PI); *speed = s < 2 ? s : 2; }  /* fast.py:138 */
    }
    *steer = st;
    return 0;
}
