/* oracle/zfit_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU restatement, never on the product path).
 *
 * Astigmatic z fit of picasso.zfit._fit_z (reference picasso/zfit.py:327-367): for every
 * localization, scipy.optimize.minimize_scalar(_fit_z_target, bounds=[-1000, 1000],
 * args=(sx, sy, cx, cy)) -- with bounds and no method scipy runs its bounded Brent minimiser
 * (scipy 1.18.1, scipy/optimize/_optimize.py::_minimize_scalar_bounded, xatol = 1e-5,
 * maxiter = 500), restated here statement by statement -- on the target of zfit.py:255-291:
 *     (sx**0.5 - wx(z)**0.5)**2 + (sy**0.5 - wy(z)**0.5)**2,  wx, wy degree-6 polynomials.
 * Types follow the reference: _fit_z_target is numba-compiled with the signature
 * (float64, float32, float32, float64[:], float64[:]) -- sx, sy are promoted to float64 and
 * `** 0.5` is LLVM's pow(x, 0.5) -> sqrt(x) (verified: every target value, z and nfev of the
 * golden set is reproduced bit for bit); the results are stored into float32 arrays
 * (np.zeros_like(locs["x"])).
 * Pinned by tests/golden/zfit.npz (raw_z / raw_fun / raw_nfev from the real reference).
 */
#include <math.h>
#include <stddef.h>

#ifndef ZROOT
#define ZROOT(v) sqrt(v)
#endif
static double zfit_target(double z, float sx, float sy, const double* cx, const double* cy) {
    const double z2 = z * z, z3 = z * z2, z4 = z * z3, z5 = z * z4, z6 = z * z5;
    const double wx = cx[0] * z6 + cx[1] * z5 + cx[2] * z4 + cx[3] * z3 + cx[4] * z2 + cx[5] * z + cx[6];
    const double wy = cy[0] * z6 + cy[1] * z5 + cy[2] * z4 + cy[3] * z3 + cy[4] * z2 + cy[5] * z + cy[6];
    const double dx = ZROOT((double)sx) - ZROOT(wx), dy = ZROOT((double)sy) - ZROOT(wy);
    return dx * dx + dy * dy;
}

static double np_sign(double v) { return v != v ? v : (v > 0.0) - (v < 0.0); }

/* _minimize_scalar_bounded; returns nfev */
static int fminbound(float sx, float sy, const double* cx, const double* cy, double x1, double x2,
                     double xatol, int maxfun, double* xout, double* fout) {
    const double sqrt_eps = sqrt(2.2e-16);
    const double golden_mean = 0.5 * (3.0 - sqrt(5.0));
    double a = x1, b = x2;
    double fulc = a + golden_mean * (b - a);
    double nfc = fulc, xf = fulc;
    double rat = 0.0, e = 0.0;
    double x = xf;
    double fx = zfit_target(x, sx, sy, cx, cy);
    int num = 1;
    double fu = INFINITY;
    double ffulc = fx, fnfc = fx;
    double xm = 0.5 * (a + b);
    double tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
    double tol2 = 2.0 * tol1;
    while (fabs(xf - xm) > (tol2 - 0.5 * (b - a))) {
        int golden = 1;
        if (fabs(e) > tol1) {
            golden = 0;
            double r = (xf - nfc) * (fx - ffulc);
            double q = (xf - fulc) * (fx - fnfc);
            double p = (xf - fulc) * q - (xf - nfc) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) p = -p;
            q = fabs(q);
            r = e;
            e = rat;
            if (fabs(p) < fabs(0.5 * q * r) && p > q * (a - xf) && p < q * (b - xf)) {
                rat = (p + 0.0) / q;
                x = xf + rat;
                if ((x - a) < tol2 || (b - x) < tol2) {
                    const double si = np_sign(xm - xf) + ((xm - xf) == 0);
                    rat = tol1 * si;
                }
            } else {
                golden = 1;
            }
        }
        if (golden) {
            e = (xf >= xm) ? a - xf : b - xf;
            rat = golden_mean * e;
        }
        const double si = np_sign(rat) + (rat == 0);
        const double ar = fabs(rat);
        /* np.maximum propagates NaN */
        const double step = (ar != ar || tol1 != tol1) ? ar + tol1 : (ar > tol1 ? ar : tol1);
        x = xf + si * step;
        fu = zfit_target(x, sx, sy, cx, cy);
        num++;
        if (fu <= fx) {
            if (x >= xf) a = xf; else b = xf;
            fulc = nfc; ffulc = fnfc;
            nfc = xf; fnfc = fx;
            xf = x; fx = fu;
        } else {
            if (x < xf) a = x; else b = x;
            if (fu <= fnfc || nfc == xf) {
                fulc = nfc; ffulc = fnfc;
                nfc = x; fnfc = fu;
            } else if (fu <= ffulc || fulc == xf || fulc == nfc) {
                fulc = x; ffulc = fu;
            }
        }
        xm = 0.5 * (a + b);
        tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
        tol2 = 2.0 * tol1;
        if (num >= maxfun) break;
    }
    *xout = xf;
    *fout = fx;
    return num;
}

/* z[i] = float32(result.x), sq[i] = float32(result.fun) (zfit.py:338-355); raw_* (nullable) keep
 * the float64 minimiser outputs and the number of function evaluations. */
int orc_zfit(long long n, const float* sx, const float* sy, const double* cx, const double* cy,
             float* z, float* sq, double* raw_z, double* raw_fun, int* raw_nfev) {
    for (long long i = 0; i < n; i++) {
        double x, f;
        const int nf = fminbound(sx[i], sy[i], cx, cy, -1000.0, 1000.0, 1e-5, 500, &x, &f);
        z[i] = (float)x;
        sq[i] = (float)f;
        if (raw_z) raw_z[i] = x;
        if (raw_fun) raw_fun[i] = f;
        if (raw_nfev) raw_nfev[i] = nf;
    }
    return 0;
}
