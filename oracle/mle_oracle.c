/*
 * oracle/mle_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's per-spot 2-D Gaussian maximum
 * likelihood fit (picasso/gaussmle.py, jungmannlab/picasso @ 96e0da51,
 * v0.10.3).  It is the CPU checker for the CUDA path in picasso_b200/csrc and
 * the `cpu_baseline` / `--impl reference` arm of bench.py.  Only tests/,
 * __graft_entry__.smoke() and bench.py may load it.
 *
 * Parity status: PINNED.  tools/gen_golden.py imports the real reference in
 * the build container and stores its outputs under tests/golden/;
 * tests/test_oracle_golden.py checks this file against them (bit-identical
 * thetas / iterations on the build container; CRLB within LAPACK-vs-Jacobi
 * rounding).
 *
 * Precision model (numba typing of the reference, verified by experiment,
 * see DESIGN.md "precision model"):
 *   - theta, max_step, dudt, d2udt2, numerator, denominator are float32
 *     arrays; every store rounds to f32.
 *   - `int64 - float32`, `float64_const / float32`, f64*f32 promote to f64.
 *   - `float32 ** int` stays float32 (binary powering in f32), so sigma**2,
 *     sigma**3, sigma**5 and dudt**2 are rounded to f32.
 *   - math.erf / np.exp / np.log are glibc's libm (numba externs).
 * Build with -ffp-contract=off (numba's x86 code has no fused multiply-add).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_BOX 31

/* ---- float32 ** int, binary powering as numba/LLVM emit it -------------- */
static float powi_f32(float a, int n) {
    float r = 1.0f;
    int first = 1;
    while (n) {
        if (n & 1) { r = first ? a : r * a; first = 0; }
        n >>= 1;
        if (n) a = a * a;
    }
    return r;
}
static double powi_f64(double a, int n) {
    double r = 1.0;
    int first = 1;
    while (n) {
        if (n & 1) { r = first ? a : r * a; first = 0; }
        n >>= 1;
        if (n) a = a * a;
    }
    return r;
}

/* gaussmle.py:28-48 _sum_and_center_of_mass */
static void sum_and_com(const float *spot, int size, double *sum, double *yc, double *xc) {
    double y = 0.0, x = 0.0, s = 0.0;
    for (int i = 0; i < size; i++)
        for (int j = 0; j < size; j++) {
            double v = (double)spot[i * size + j];
            y += v * (double)i;
            x += v * (double)j;
            s += v;
        }
    if (s <= 0.0) {
        *sum = 0.01; *yc = (size - 1) / 2.0; *xc = (size - 1) / 2.0;
        return;
    }
    *sum = s; *yc = y / s; *xc = x / s;
}

/* gaussmle.py:61-91 _mean_filter + np.min (gaussmle.py:135) */
static float mean_filter_min(const float *spot, int size) {
    float best = 0.0f;
    for (int k = 0; k < size; k++)
        for (int l = 0; l < size; l++) {
            int min_m = k - 1 < 0 ? 0 : k - 1;
            int max_m = k + 2 > size ? size : k + 2;
            int min_n = l - 1 < 0 ? 0 : l - 1;
            int max_n = l + 2 > size ? size : l + 2;
            int N = (max_m - min_m) * (max_n - min_n);
            double Nsum = 0.0;
            for (int m = min_m; m < max_m; m++)
                for (int n = min_n; n < max_n; n++)
                    Nsum += (double)spot[m * size + n];
            float f = (float)(Nsum / (double)N);
            if ((k == 0 && l == 0) || f < best) best = f;   /* np.min; NaN-free inputs */
        }
    return best;
}

/* gaussmle.py:94-124 _initial_sigmas on (spot - bg) (f32 array - f32 scalar) */
static void initial_sigmas(const float *spot, float bg, int size, double *sy_out, double *sx_out,
                           int *status) {
    int h = size / 2;
    double sdy = 0.0, sdx = 0.0, sumy = 0.0, sumx = 0.0;
    for (int i = 0; i < size; i++) {
        double d2 = (double)((i - h) * (i - h));
        float vy = spot[i * size + h] - bg;    /* f32 subtraction */
        float vx = spot[h * size + i] - bg;
        sdy += (double)vy * d2;
        sdx += (double)vx * d2;
        sumy += (double)vy;
        sumx += (double)vx;
    }
    /* The reference raises ZeroDivisionError here when a sum is exactly 0
     * (numba python error model).  We emit the 0.01 fallback the code
     * intends (gaussmle.py:116-123) and flag it. */
    double sy, sx;
    if (sumy == 0.0) { sy = NAN; *status |= 1; } else sy = sqrt(sdy / sumy);
    if (sumx == 0.0) { sx = NAN; *status |= 1; } else sx = sqrt(sdx / sumx);
    if (!isfinite(sy)) sy = 0.01;
    if (!isfinite(sx)) sx = 0.01;
    if (sx == 0) sx = 0.01;
    if (sy == 0) sy = 0.01;
    *sy_out = sy; *sx_out = sx;
}

/* gaussmle.py:127-139 _initial_parameters -> (x, y, photons, bg, sx, sy) */
static void initial_parameters(const float *spot, int size, double *x, double *y, double *photons,
                               float *bg, double *sx, double *sy, int *status) {
    double sum;
    sum_and_com(spot, size, &sum, y, x);
    *bg = mean_filter_min(spot, size);
    double ph = sum - (double)(size * size) * (double)(*bg);
    *photons = ph > 1.0 ? ph : 1.0;        /* np.maximum(1.0, photons) */
    initial_sigmas(spot, *bg, size, sy, sx, status);
}

/* gaussmle.py:268-280 _gaussian_integral(x:int64, mu:f32, sigma:f32) -> f64 */
static double gaussian_integral(int x, float mu, float sigma) {
    double sq_norm = 0.70710678118654757 / (double)sigma;
    double d = (double)x - (double)mu;
    return 0.5 * (erf((d + 0.5) * sq_norm) - erf((d - 0.5) * sq_norm));
}

/* gaussmle.py:283-303 _derivative_gaussian_integral */
static void deriv_gaussian_integral(int x, float mu, float sigma, float photons, double PSFy,
                                    double *dudt, double *d2udt2) {
    double d = (double)x - (double)mu;
    double s = (double)sigma;
    double ta = (d + 0.5) / s, tb = (d - 0.5) / s;
    double a = exp(-0.5 * (ta * ta));
    double b = exp(-0.5 * (tb * tb));
    const double sq2pi = sqrt(2.0 * M_PI);
    *dudt = (double)photons * PSFy * (b - a) / (sq2pi * s);
    float s3 = powi_f32(sigma, 3);                    /* sigma**3 stays f32 */
    *d2udt2 = (double)photons * ((d - 0.5) * b - (d + 0.5) * a) * PSFy / (sq2pi * (double)s3);
}

/* gaussmle.py:306-316 _G(n, m, x, mu, sigma_x) */
static double G(int n, int m, int x, float mu, float sigma_x) {
    double a_minus = (double)x - (double)mu - 0.5;
    double a_plus = (double)x - (double)mu + 0.5;
    double two_s2 = 2.0 * (double)powi_f32(sigma_x, 2);    /* 2 * f32(sigma**2) */
    double exp_minus = exp(-(a_minus * a_minus) / two_s2);
    double exp_plus = exp(-(a_plus * a_plus) / two_s2);
    const double sq2pi = sqrt(2.0 * M_PI);
    return (powi_f64(a_minus, m) * exp_minus - powi_f64(a_plus, m) * exp_plus) /
           ((double)powi_f32(sigma_x, n) * sq2pi);
}

/* gaussmle.py:319-336 _derivative_gaussian_integral_sigma */
static void deriv_gaussian_integral_sigma(int x, float mu, float sigma_x, float photons,
                                          double PSFy, double *dudt, double *d2udt2) {
    *dudt = (double)photons * PSFy * G(2, 1, x, mu, sigma_x);
    *d2udt2 = (double)photons * PSFy *
              (G(5, 3, x, mu, sigma_x) - 2.0 * G(3, 1, x, mu, sigma_x));
}

/* gaussmle.py:339-383 _derivative_gaussian_integral_iso_sigma (incl. the
 * operator-precedence quirk at :380-382: photons multiplies only term 1) */
static void deriv_gaussian_integral_iso_sigma(int x, int y, float mu, float nu, float sigma,
                                              float photons, double PSFx, double PSFy,
                                              double *dudt, double *d2udt2) {
    const double sq2 = sqrt(2.0), sqpi = sqrt(M_PI);
    double s = (double)sigma;
    double a_plus = ((double)x - (double)mu + 0.5) / (sq2 * s);
    double a_minus = ((double)x - (double)mu - 0.5) / (sq2 * s);
    double b_plus = ((double)y - (double)nu + 0.5) / (sq2 * s);
    double b_minus = ((double)y - (double)nu - 0.5) / (sq2 * s);

    double Fx = a_minus * exp(-(a_minus * a_minus)) - a_plus * exp(-(a_plus * a_plus));
    double Fy = b_minus * exp(-(b_minus * b_minus)) - b_plus * exp(-(b_plus * b_plus));
    double dPSFxdt = Fx / (sqpi * s);
    double dPSFydt = Fy / (sqpi * s);

    double dFxdt = (a_plus * exp(-(a_plus * a_plus)) * (1.0 - 2.0 * (a_plus * a_plus)) -
                    a_minus * exp(-(a_minus * a_minus)) * (1.0 - 2.0 * (a_minus * a_minus))) / s;
    double dFydy = (b_plus * exp(-(b_plus * b_plus)) * (1.0 - 2.0 * (b_plus * b_plus)) -
                    b_minus * exp(-(b_minus * b_minus)) * (1.0 - 2.0 * (b_minus * b_minus))) / s;
    float s2 = powi_f32(sigma, 2);            /* sigma**2   : f32 */
    float sinv = 1.0f / sigma;                /* sigma**(-1): f32 */
    double d2PSFxdt2 = (1.0 / sqpi) * ((-Fx / (double)s2) + (double)sinv * dFxdt);
    double d2PSFydt2 = (1.0 / sqpi) * ((-Fy / (double)s2) + (double)sinv * dFydy);

    *dudt = (double)photons * (PSFy * dPSFxdt + PSFx * dPSFydt);
    *d2udt2 = (double)photons * PSFy * d2PSFxdt2 + 2.0 * dPSFxdt * dPSFydt + PSFx * d2PSFydt2;
}

/* ---- 6x6 (or 5x5) Moore-Penrose pseudo-inverse diagonal -----------------
 * Reference: np.linalg.pinv(M) (LAPACK gesdd, rcond = 1e-15) then the
 * diagonal (gaussmle.py:737-742, 950-954).  M is symmetric PSD, so the SVD
 * equals the eigendecomposition; we use cyclic Jacobi in f64. */
static void pinv_diag_sym(const double *Min, int n, double *diag) {
    double A[36], V[36];
    for (int i = 0; i < n * n; i++) A[i] = Min[i];
    int finite = 1;
    for (int i = 0; i < n * n; i++) if (!isfinite(A[i])) finite = 0;
    if (!finite) { for (int i = 0; i < n; i++) diag[i] = NAN; return; }
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = (i == j);
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0, dsum = 0.0;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                if (i != j) off += A[i * n + j] * A[i * n + j];
                else dsum += A[i * n + j] * A[i * n + j];
            }
        if (off <= 1e-60 || off <= 1e-34 * dsum) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double apq = A[p * n + q];
                if (apq == 0.0) continue;
                double app = A[p * n + p], aqq = A[q * n + q];
                double theta = (aqq - app) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                if (!isfinite(theta)) t = 0.0;
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) {       /* A <- A J */
                    double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {       /* A <- J^T A */
                    double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    double smax = 0.0;
    for (int i = 0; i < n; i++) if (fabs(A[i * n + i]) > smax) smax = fabs(A[i * n + i]);
    double cutoff = 1e-15 * smax;          /* numpy pinv default rcond */
    for (int i = 0; i < n; i++) {
        double acc = 0.0;
        for (int k = 0; k < n; k++) {
            double lam = A[k * n + k];
            /* singular value = |lam|; pinv term = v v^T / lam for |lam| > cutoff */
            if (fabs(lam) > cutoff) acc += V[i * n + k] * V[i * n + k] / lam;
        }
        diag[i] = acc;
    }
}

/* gaussmle.py:887-954 _mlefit_sigmaxy_crlb */
static void crlb_sigmaxy(const float *theta, const float *spot, int size, float *crlb,
                         float *loglik) {
    const int np_ = 6;
    float dudt[6];
    double M[36];
    double ll = 0.0;
    memset(M, 0, sizeof M);
    for (int ii = 0; ii < size; ii++)
        for (int jj = 0; jj < size; jj++) {
            double PSFx = gaussian_integral(ii, theta[0], theta[4]);
            double PSFy = gaussian_integral(jj, theta[1], theta[5]);
            double d, d2;
            deriv_gaussian_integral(ii, theta[0], theta[4], theta[2], PSFy, &d, &d2); dudt[0] = (float)d;
            deriv_gaussian_integral(jj, theta[1], theta[5], theta[2], PSFx, &d, &d2); dudt[1] = (float)d;
            deriv_gaussian_integral_sigma(ii, theta[0], theta[4], theta[2], PSFy, &d, &d2); dudt[4] = (float)d;
            deriv_gaussian_integral_sigma(jj, theta[1], theta[5], theta[2], PSFx, &d, &d2); dudt[5] = (float)d;
            dudt[2] = (float)(PSFx * PSFy);
            dudt[3] = 1.0f;
            double model = (double)theta[2] * PSFx * PSFy + (double)theta[3];
            for (int kk = 0; kk < np_; kk++)
                for (int l = kk; l < np_; l++) {
                    /* dudt[ll]*dudt[kk] is an f32*f32 product (f32), then / f64 model */
                    M[kk * np_ + l] += (double)(dudt[l] * dudt[kk]) / model;
                    M[l * np_ + kk] = M[kk * np_ + l];
                }
            if (model > 0) {
                float data = spot[jj * size + ii];
                if (data > 0)
                    /* np.log(f32) is logf; data*np.log(data) is an f32 product */
                    ll += (double)data * log(model) - model - (double)(data * logf(data)) + (double)data;
                else
                    ll += -model;
            }
        }
    *loglik = (float)ll;
    double dg[6];
    pinv_diag_sym(M, np_, dg);
    for (int k = 0; k < np_; k++) crlb[k] = (float)dg[k];
}

/* gaussmle.py:673-742 _mlefit_sigma_crlb */
static void crlb_sigma(const float *theta, const float *spot, int size, float *crlb,
                       float *loglik) {
    const int np_ = 5;
    float dudt[5];
    double M[25];
    double ll = 0.0;
    memset(M, 0, sizeof M);
    for (int ii = 0; ii < size; ii++)
        for (int jj = 0; jj < size; jj++) {
            double PSFx = gaussian_integral(ii, theta[0], theta[4]);
            double PSFy = gaussian_integral(jj, theta[1], theta[4]);
            double d, d2;
            deriv_gaussian_integral(ii, theta[0], theta[4], theta[2], PSFy, &d, &d2); dudt[0] = (float)d;
            deriv_gaussian_integral(jj, theta[1], theta[4], theta[2], PSFx, &d, &d2); dudt[1] = (float)d;
            deriv_gaussian_integral_iso_sigma(ii, jj, theta[0], theta[1], theta[4], theta[2], PSFx,
                                              PSFy, &d, &d2);
            dudt[4] = (float)d;
            dudt[2] = (float)(PSFx * PSFy);
            dudt[3] = 1.0f;
            double model = (double)theta[2] * PSFx * PSFy + (double)theta[3];
            for (int kk = 0; kk < np_; kk++)
                for (int l = kk; l < np_; l++) {
                    M[kk * np_ + l] += (double)(dudt[l] * dudt[kk]) / model;
                    M[l * np_ + kk] = M[kk * np_ + l];
                }
            if (model > 0) {
                float data = spot[jj * size + ii];
                if (data > 0)
                    /* np.log(f32) is logf; data*np.log(data) is an f32 product */
                    ll += (double)data * log(model) - model - (double)(data * logf(data)) + (double)data;
                else
                    ll += -model;
            }
        }
    *loglik = (float)ll;
    double dg[5];
    pinv_diag_sym(M, np_, dg);
    for (int k = 0; k < np_; k++) crlb[k] = (float)dg[k];
    crlb[5] = crlb[4];
}

static float signf_np(float v) { return v > 0 ? 1.0f : (v < 0 ? -1.0f : (v == 0 ? 0.0f : v)); }

/* gaussmle.py:745-857 _mlefit_sigmaxy (+ :860-884 _update_theta_sigmaxy) */
int orc_mle_sigmaxy_one(const float *spot, int size, double eps, int max_it, float *theta_out,
                        float *crlb_out, float *loglik_out, int *iter_out) {
    const int np_ = 6;
    int status = 0;
    float theta[6], max_step[6], dudt[6], d2udt2[6], num[6], den[6];
    {
        double x, y, ph, sx, sy; float bg;
        initial_parameters(spot, size, &x, &y, &ph, &bg, &sx, &sy, &status);
        theta[0] = (float)x; theta[1] = (float)y; theta[2] = (float)ph; theta[3] = bg;
        theta[4] = (float)sx; theta[5] = (float)sy;
    }
    max_step[0] = max_step[1] = theta[4];
    max_step[2] = (float)(0.1 * (double)theta[2]);
    max_step[3] = (float)(0.1 * (double)theta[3]);
    max_step[4] = (float)(0.2 * (double)theta[4]);
    max_step[5] = (float)(0.2 * (double)theta[5]);

    float old_x = theta[0], old_y = theta[1], old_sx = theta[4], old_sy = theta[5];
    int kk = 0;
    while (kk < max_it) {
        kk++;
        for (int l = 0; l < np_; l++) num[l] = den[l] = 0.0f;
        for (int ii = 0; ii < size; ii++)
            for (int jj = 0; jj < size; jj++) {
                double PSFx = gaussian_integral(ii, theta[0], theta[4]);
                double PSFy = gaussian_integral(jj, theta[1], theta[5]);
                double d, d2;
                deriv_gaussian_integral(ii, theta[0], theta[4], theta[2], PSFy, &d, &d2);
                dudt[0] = (float)d; d2udt2[0] = (float)d2;
                deriv_gaussian_integral(jj, theta[1], theta[5], theta[2], PSFx, &d, &d2);
                dudt[1] = (float)d; d2udt2[1] = (float)d2;
                dudt[2] = (float)(PSFx * PSFy); d2udt2[2] = 0.0f;
                dudt[3] = 1.0f; d2udt2[3] = 0.0f;
                deriv_gaussian_integral_sigma(ii, theta[0], theta[4], theta[2], PSFy, &d, &d2);
                dudt[4] = (float)d; d2udt2[4] = (float)d2;
                deriv_gaussian_integral_sigma(jj, theta[1], theta[5], theta[2], PSFx, &d, &d2);
                dudt[5] = (float)d; d2udt2[5] = (float)d2;

                double model = (double)theta[2] * PSFx * PSFy + (double)theta[3];
                double cf = 0.0, df = 0.0;
                double data = (double)spot[jj * size + ii];
                if (model > 10e-3) {
                    cf = data / model - 1.0;
                    df = data / (model * model);
                }
                cf = cf < 10e4 ? cf : 10e4;        /* np.minimum */
                df = df < 10e4 ? df : 10e4;
                for (int l = 0; l < np_; l++) {
                    num[l] = (float)((double)num[l] + cf * (double)dudt[l]);
                    float du2 = dudt[l] * dudt[l];                      /* f32 ** 2 */
                    den[l] = (float)((double)den[l] + (cf * (double)d2udt2[l] - df * (double)du2));
                }
            }
        /* _update_theta_sigmaxy */
        for (int l = 0; l < np_; l++) {
            if (den[l] == 0.0f) {
                theta[l] = theta[l] - signf_np(num[l]) * max_step[l];
            } else {
                float q = num[l] / den[l];
                float lo = -max_step[l];
                float u = q > lo ? q : lo;               /* np.maximum (NaN-propagating) */
                if (q != q) u = q;
                float v = u < max_step[l] ? u : max_step[l];
                if (u != u) v = u;
                theta[l] = theta[l] - v;
            }
        }
        /* np.maximum(f32, f64 const) -> f64 -> stored f32 */
        if (!(theta[2] != theta[2])) theta[2] = theta[2] > 1.0f ? theta[2] : 1.0f;
        if (!(theta[3] != theta[3])) theta[3] = (float)((double)theta[3] > 0.01 ? (double)theta[3] : 0.01);
        if (!(theta[4] != theta[4])) theta[4] = (float)((double)theta[4] > 0.01 ? (double)theta[4] : 0.01);
        if (!(theta[5] != theta[5])) theta[5] = (float)((double)theta[5] > 0.01 ? (double)theta[5] : 0.01);

        if ((double)fabsf(old_x - theta[0]) < eps && (double)fabsf(old_y - theta[1]) < eps &&
            (double)fabsf(old_sx - theta[4]) < eps && (double)fabsf(old_sy - theta[5]) < eps)
            break;
        old_x = theta[0]; old_y = theta[1]; old_sx = theta[4]; old_sy = theta[5];
    }
    for (int l = 0; l < 6; l++) theta_out[l] = theta[l];
    *iter_out = kk;
    crlb_sigmaxy(theta, spot, size, crlb_out, loglik_out);
    return status;
}

/* gaussmle.py:533-670 _mlefit_sigma (+ _update_theta_sigma) */
int orc_mle_sigma_one(const float *spot, int size, double eps, int max_it, float *theta_out,
                      float *crlb_out, float *loglik_out, int *iter_out) {
    const int np_ = 5;
    int status = 0;
    float theta[5], max_step[5], dudt[5], d2udt2[5], num[5], den[5];
    {
        double x, y, ph, sx, sy; float bg;
        initial_parameters(spot, size, &x, &y, &ph, &bg, &sx, &sy, &status);
        theta[0] = (float)x; theta[1] = (float)y; theta[2] = (float)ph; theta[3] = bg;
        theta[4] = (float)((sx + sy) / 2.0);
    }
    max_step[0] = max_step[1] = theta[4];
    max_step[2] = (float)(0.1 * (double)theta[2]);
    max_step[3] = (float)(0.1 * (double)theta[3]);
    max_step[4] = (float)(0.2 * (double)theta[4]);

    float old_x = theta[0], old_y = theta[1];
    int kk = 0;
    while (kk < max_it) {
        kk++;
        for (int l = 0; l < np_; l++) num[l] = den[l] = 0.0f;
        for (int ii = 0; ii < size; ii++)
            for (int jj = 0; jj < size; jj++) {
                double PSFx = gaussian_integral(ii, theta[0], theta[4]);
                double PSFy = gaussian_integral(jj, theta[1], theta[4]);
                double d, d2;
                deriv_gaussian_integral(ii, theta[0], theta[4], theta[2], PSFy, &d, &d2);
                dudt[0] = (float)d; d2udt2[0] = (float)d2;
                deriv_gaussian_integral(jj, theta[1], theta[4], theta[2], PSFx, &d, &d2);
                dudt[1] = (float)d; d2udt2[1] = (float)d2;
                dudt[2] = (float)(PSFx * PSFy); d2udt2[2] = 0.0f;
                dudt[3] = 1.0f; d2udt2[3] = 0.0f;
                deriv_gaussian_integral_iso_sigma(ii, jj, theta[0], theta[1], theta[4], theta[2],
                                                  PSFx, PSFy, &d, &d2);
                dudt[4] = (float)d; d2udt2[4] = (float)d2;

                double model = (double)theta[2] * PSFx * PSFy + (double)theta[3];
                double cf = 0.0, df = 0.0;
                double data = (double)spot[jj * size + ii];
                if (model > 10e-3) {
                    cf = data / model - 1.0;
                    df = data / (model * model);
                }
                cf = cf < 10e4 ? cf : 10e4;
                df = df < 10e4 ? df : 10e4;
                for (int l = 0; l < np_; l++) {
                    num[l] = (float)((double)num[l] + cf * (double)dudt[l]);
                    float du2 = dudt[l] * dudt[l];
                    den[l] = (float)((double)den[l] + (cf * (double)d2udt2[l] - df * (double)du2));
                }
            }
        /* _update_theta_sigma: den==0 -> step = sign(num*max_step) i.e. +-1 */
        for (int l = 0; l < np_; l++) {
            float upd;
            if (den[l] == 0.0f) {
                upd = signf_np(num[l] * max_step[l]);
            } else {
                float q = num[l] / den[l];
                float lo = -max_step[l];
                float u = q > lo ? q : lo;
                if (q != q) u = q;
                upd = u < max_step[l] ? u : max_step[l];
                if (u != u) upd = u;
            }
            theta[l] = theta[l] - upd;
        }
        if (!(theta[2] != theta[2])) theta[2] = theta[2] > 1.0f ? theta[2] : 1.0f;
        if (!(theta[3] != theta[3])) theta[3] = (float)((double)theta[3] > 0.01 ? (double)theta[3] : 0.01);
        if (!(theta[4] != theta[4])) theta[4] = (float)((double)theta[4] > 0.01 ? (double)theta[4] : 0.01);
        if (!(theta[4] != theta[4])) theta[4] = theta[4] < (float)size ? theta[4] : (float)size;

        if ((double)fabsf(old_x - theta[0]) < eps && (double)fabsf(old_y - theta[1]) < eps) break;
        old_x = theta[0]; old_y = theta[1];
    }
    for (int l = 0; l < 5; l++) theta_out[l] = theta[l];
    theta_out[5] = theta[4];
    *iter_out = kk;
    crlb_sigma(theta, spot, size, crlb_out, loglik_out);
    return status;
}

/* gaussmle.py:409-475 gaussmle(): loop over spots (no Python callback here).
 * method: 0 = "sigma", 1 = "sigmaxy".  Spots [begin, end).  Thread-safe. */
int orc_mle_fit_range(const float *spots, long long begin, long long end, int size, double eps,
                      int max_it, int method, float *thetas, float *crlbs, float *logliks,
                      int *iterations, int *status) {
    if (size < 1 || size > ORC_MAX_BOX) return -1;
    if (method != 0 && method != 1) return -2;
    for (long long i = begin; i < end; i++) {
        const float *spot = spots + i * size * size;
        int st;
        if (method == 1)
            st = orc_mle_sigmaxy_one(spot, size, eps, max_it, thetas + 6 * i, crlbs + 6 * i,
                                     logliks + i, iterations + i);
        else
            st = orc_mle_sigma_one(spot, size, eps, max_it, thetas + 6 * i, crlbs + 6 * i,
                                   logliks + i, iterations + i);
        if (status) status[i] = st;
    }
    return 0;
}

/* test hooks: expose the pieces the golden tests pin individually */
void orc_mle_initial_theta(const float *spot, int size, float *theta6) {
    double x, y, ph, sx, sy; float bg; int st = 0;
    initial_parameters(spot, size, &x, &y, &ph, &bg, &sx, &sy, &st);
    theta6[0] = (float)x; theta6[1] = (float)y; theta6[2] = (float)ph; theta6[3] = bg;
    theta6[4] = (float)sx; theta6[5] = (float)sy;
}

/* ---- multi-threaded driver (pthreads) for bench.py's CPU arm ------------
 * The reference's production path is gaussmle_async: a thread pool pulling
 * spot indices (gaussmle.py:478-530).  Here: static contiguous chunks. */
#include <pthread.h>
typedef struct {
    const float *spots; long long begin, end; int size; double eps; int max_it, method;
    float *thetas, *crlbs, *logliks; int *iterations, *status;
} orc_mle_job;

static void *orc_mle_worker(void *p) {
    orc_mle_job *j = (orc_mle_job *)p;
    orc_mle_fit_range(j->spots, j->begin, j->end, j->size, j->eps, j->max_it, j->method,
                      j->thetas, j->crlbs, j->logliks, j->iterations, j->status);
    return NULL;
}

int orc_mle_fit_mt(const float *spots, long long n, int size, double eps, int max_it, int method,
                   float *thetas, float *crlbs, float *logliks, int *iterations, int *status,
                   int nthreads) {
    if (size < 1 || size > ORC_MAX_BOX) return -1;
    if (method != 0 && method != 1) return -2;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    orc_mle_job jobs[256];
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (orc_mle_job){spots, n * t / nthreads, n * (t + 1) / nthreads, size, eps, max_it,
                                method, thetas, crlbs, logliks, iterations, status};
        pthread_create(&th[t], NULL, orc_mle_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}
