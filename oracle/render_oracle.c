/*
 * oracle/render_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's unrotated renderers
 * (picasso/render.py @ 96e0da51): _render_setup :177-232, _fill :451-467,
 * _draw_gaussian_loc :494-540, _fill_gaussian :543-575, _render_hist :798-853,
 * _render_gaussian :1020-1112, _render_gaussian_iso :1148-1216.
 * Pinned by tests/test_oracle_golden_render.py against images rendered by the
 * real reference (tools/gen_golden.py render).
 *
 * Precision model (numba typing, verified): x, y are float32 arrays; after
 * `oversampling * (x - x_min)` they are float64; sx, sy stay float32
 * (float32(oversampling) * max(lp, float32(min_blur_width)), numpy weak-scalar
 * rules); window bounds are np.int32 truncations; the 1-D kernels are evaluated
 * in float64 and stored float32; the image accumulates in float32 in
 * localisation order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int trunc_i32(double v) { return (int)v; }   /* np.int32(f64): toward zero */

/* mode: 0 = no blur (histogram), 1 = "gaussian", 2 = "gaussian_iso".
 * image must hold n_pixel_y * n_pixel_x floats and is zeroed here.
 * Returns n = number of localisations in view. */
long long orc_render(const float *x, const float *y, const float *lpx, const float *lpy,
                     long long n_locs, double oversampling, double y_min, double x_min,
                     double y_max, double x_max, double min_blur_width, int mode, float *image,
                     int n_pixel_y, int n_pixel_x) {
    memset(image, 0, sizeof(float) * (size_t)n_pixel_y * n_pixel_x);
    long long n = 0;
    const float osf = (float)oversampling;
    const float mbw = (float)min_blur_width;
    for (long long k = 0; k < n_locs; k++) {
        double xv = (double)x[k], yv = (double)y[k];
        if (!(xv > x_min && yv > y_min && xv < x_max && yv < y_max)) continue;
        n++;
        double x_ = oversampling * (xv - x_min);
        double y_ = oversampling * (yv - y_min);
        if (mode == 0) {
            int i = trunc_i32(x_), j = trunc_i32(y_);
            if (j >= 0 && j < n_pixel_y && i >= 0 && i < n_pixel_x)
                image[(size_t)j * n_pixel_x + i] += 1.0f;
            continue;
        }
        float bw = osf * (lpx[k] > mbw ? lpx[k] : mbw);   /* np.maximum, f32 */
        float bh = osf * (lpy[k] > mbw ? lpy[k] : mbw);
        float sx_, sy_;
        if (mode == 2) { sy_ = (bh + bw) / 2.0f; sx_ = sy_; }
        else { sx_ = bw; sy_ = bh; }
        double max_y_off = 3.0 * (double)sy_;
        int i_min = trunc_i32(y_ - max_y_off);
        if (i_min < 0) i_min = 0;
        int i_max = trunc_i32(y_ + max_y_off + 1.0);
        if (i_max > n_pixel_y) i_max = n_pixel_y;
        double max_x_off = 3.0 * (double)sx_;
        int j_min = trunc_i32(x_ - max_x_off);
        if (j_min < 0) j_min = 0;
        int j_max = trunc_i32(x_ + max_x_off) + 1;
        if (j_max > n_pixel_x) j_max = n_pixel_x;
        int nx = j_max - j_min, ny = i_max - i_min;
        if (nx <= 0 || ny <= 0) continue;
        double inv_2sx2 = 1.0 / (2.0 * (double)sx_ * (double)sx_);
        double inv_2sy2 = 1.0 / (2.0 * (double)sy_ * (double)sy_);
        double norm = 1.0 / (2.0 * M_PI * (double)sx_ * (double)sy_);
        float *gx = malloc(sizeof(float) * nx), *gy = malloc(sizeof(float) * ny);
        for (int jj = 0; jj < nx; jj++) {
            double dx = (double)(j_min + jj) + 0.5 - x_;
            gx[jj] = (float)exp(-dx * dx * inv_2sx2);
        }
        for (int ii = 0; ii < ny; ii++) {
            double dy = (double)(i_min + ii) + 0.5 - y_;
            gy[ii] = (float)(norm * exp(-dy * dy * inv_2sy2));
        }
        for (int ii = 0; ii < ny; ii++) {
            float *row = image + (size_t)(i_min + ii) * n_pixel_x;
            for (int jj = 0; jj < nx; jj++) row[j_min + jj] += gy[ii] * gx[jj];
        }
        free(gx); free(gy);
    }
    return n;
}
