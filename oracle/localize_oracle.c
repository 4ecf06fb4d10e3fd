/*
 * oracle/localize_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's spot identification and ROI
 * extraction (picasso/localize.py @ 96e0da51): _local_maxima :97-134,
 * _gradient_at :153-181, _net_gradient :202-244, identify_in_image :247-292,
 * _cut_spots_numba :917-931, _to_photons :1101-1112.
 * Pinned by tests/test_oracle_golden_localize.py against golden vectors made
 * from the real reference (tools/gen_golden.py identify).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* localize.py:97-134.  A pixel is a maximum iff np.argmax of its box x box
 * window (first maximum in row-major order) is the centre.  Scan range
 * i in [h, Y-h-1), j in [h, X-h-1).  Output in np.where (row-major) order.
 * Returns the number of maxima (writes at most `cap`). */
long long orc_local_maxima(const float *frame, int Y, int X, int box, long long *ys,
                           long long *xs, long long cap) {
    int h = box / 2;
    long long n = 0;
    for (int i = h; i < Y - (h + 1); i++)
        for (int j = h; j < X - (h + 1); j++) {
            int best = 0;
            float bv = frame[(i - h) * X + (j - h)];
            int flat = 0;
            for (int a = i - h; a <= i + h; a++)
                for (int b = j - h; b <= j + h; b++, flat++) {
                    float v = frame[a * X + b];
                    if (v > bv) { bv = v; best = flat; }
                }
            if (best / box == h && best % box == h) {
                if (n < cap) { ys[n] = i; xs[n] = j; }
                n++;
            }
        }
    return n;
}

/* numba wraps negative indices: frame[-1] is the last row / column */
static inline float at_wrap(const float *frame, int Y, int X, int y, int x) {
    if (y < 0) y += Y;
    if (x < 0) x += X;
    return frame[y * X + x];
}

/* localize.py:202-244 (+ :279-286 unit vectors).  f32 accumulation, row-major. */
void orc_net_gradient(const float *frame, int Y, int X, const long long *ys, const long long *xs,
                      long long n, int box, float *ng) {
    int h = box / 2;
    float *ux = malloc(sizeof(float) * box * box), *uy = malloc(sizeof(float) * box * box);
    for (int r = 0; r < box; r++)
        for (int c = 0; c < box; c++) {
            ux[r * box + c] = (float)(h - c);
            uy[r * box + c] = (float)(h - r);
        }
    for (int q = 0; q < box * box; q++) {
        float un = sqrtf(ux[q] * ux[q] + uy[q] * uy[q]);
        ux[q] = ux[q] / un;     /* centre: 0/0 = NaN, never read */
        uy[q] = uy[q] / un;
    }
    for (long long i = 0; i < n; i++) {
        int yi = (int)ys[i], xi = (int)xs[i];
        float acc = 0.0f;
        int ki = 0;
        for (int k = yi - h; k <= yi + h; k++, ki++) {
            int li = 0;
            for (int m = xi - h; m <= xi + h; m++, li++) {
                if (k == yi && m == xi) continue;
                float gy = at_wrap(frame, Y, X, k + 1, m) - at_wrap(frame, Y, X, k - 1, m);
                float gx = at_wrap(frame, Y, X, k, m + 1) - at_wrap(frame, Y, X, k, m - 1);
                float t = gy * uy[ki * box + li] + gx * ux[ki * box + li];
                acc = acc + t;
            }
        }
        ng[i] = acc;
    }
    free(ux); free(uy);
}

/* localize.py:247-292 identify_in_image: maxima + net gradient + (ng > min_ng).
 * Returns the number kept (writes at most cap). */
long long orc_identify_in_image(const float *image, int Y, int X, double minimum_ng, int box,
                                long long *ys, long long *xs, float *ngs, long long cap) {
    long long maxcand = (long long)Y * X;
    long long *cy = malloc(sizeof(long long) * maxcand), *cx = malloc(sizeof(long long) * maxcand);
    long long nc = orc_local_maxima(image, Y, X, box, cy, cx, maxcand);
    float *ng = malloc(sizeof(float) * (nc > 0 ? nc : 1));
    orc_net_gradient(image, Y, X, cy, cx, nc, box, ng);
    long long n = 0;
    for (long long i = 0; i < nc; i++)
        if ((double)ng[i] > minimum_ng) {
            if (n < cap) { ys[n] = cy[i]; xs[n] = cx[i]; ngs[n] = ng[i]; }
            n++;
        }
    free(cy); free(cx); free(ng);
    return n;
}

/* identify over a uint16 movie, frame by frame (localize.py:295-337, 340-421,
 * 604-636): optional roi (y0,x0,y1,x1) slices the frame first and offsets the
 * coordinates; frame_bounds = inclusive [lo, hi] test on the frame number.
 * Output arrays sized `cap`; returns total count. */
long long orc_identify_movie_u16(const uint16_t *movie, long long F, int Y, int X,
                                 double minimum_ng, int box, const int *roi /*nullable*/,
                                 long long fb_lo, long long fb_hi, long long *frames,
                                 long long *xs, long long *ys, float *ngs, long long cap) {
    int y0 = 0, x0 = 0, y1 = Y, x1 = X;
    if (roi) {
        y0 = roi[0]; x0 = roi[1]; y1 = roi[2]; x1 = roi[3];
        /* python slicing clamps */
        if (y0 < 0) y0 += Y; if (x0 < 0) x0 += X; if (y1 < 0) y1 += Y; if (x1 < 0) x1 += X;
        if (y0 < 0) y0 = 0; if (x0 < 0) x0 = 0; if (y1 > Y) y1 = Y; if (x1 > X) x1 = X;
        if (y1 < y0) y1 = y0; if (x1 < x0) x1 = x0;
    }
    int Ys = y1 - y0, Xs = x1 - x0;
    float *img = malloc(sizeof(float) * (size_t)(Ys > 0 ? Ys : 1) * (Xs > 0 ? Xs : 1));
    long long n = 0;
    for (long long f = 0; f < F; f++) {
        if (f < fb_lo || f > fb_hi) continue;
        const uint16_t *fr = movie + (size_t)f * Y * X;
        for (int a = 0; a < Ys; a++)
            for (int b = 0; b < Xs; b++) img[a * Xs + b] = (float)fr[(a + y0) * X + (b + x0)];
        long long room = cap - n > 0 ? cap - n : 0;
        long long k = orc_identify_in_image(img, Ys, Xs, minimum_ng, box, ys + (room ? n : 0),
                                            xs + (room ? n : 0), ngs + (room ? n : 0), room);
        long long w = k < room ? k : room;
        for (long long q = 0; q < w; q++) {
            frames[n + q] = f;
            ys[n + q] += y0;
            xs[n + q] += x0;
        }
        n += k;
    }
    free(img);
    return n;
}

/* localize.py:917-931 _cut_spots_numba + :1101-1112 _to_photons (f32:
 * (s - baseline) * sensitivity / gain with weak Python scalars). */
void orc_get_spots_u16(const uint16_t *movie, int Y, int X, const long long *frames,
                       const long long *xs, const long long *ys, long long n, int box,
                       float baseline, float sensitivity, float gain, float *spots) {
    int r = box / 2;
    for (long long i = 0; i < n; i++) {
        const uint16_t *fr = movie + (size_t)frames[i] * Y * X;
        for (int a = 0; a < box; a++)
            for (int b = 0; b < box; b++) {
                float s = (float)fr[(ys[i] - r + a) * X + (xs[i] - r + b)];
                float v = s - baseline;
                v = v * sensitivity;
                v = v / gain;
                spots[(i * box + a) * box + b] = v;
            }
    }
}
