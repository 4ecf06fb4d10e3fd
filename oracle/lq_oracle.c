/*
 * oracle/lq_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's least-squares Gaussian spot fit
 * (picasso/gausslq.py:33-289 @ 96e0da51): start values (_initial_parameters
 * :95-112), the float32-rounded point-sampled model / residuals
 * (_gaussian :33-39, _outer :151-164, _compute_residuals :187-203) and the
 * optimiser the reference calls, scipy.optimize.leastsq(ftol=1e-2, xtol=1e-2)
 * (gausslq.py:240-242).
 *
 * THIRD-PARTY ARITHMETIC: scipy.optimize.leastsq -> MINPACK `lmdif`
 * (reference pins scipy>=1.15.3,<2, pyproject.toml:36; this container: scipy
 * 1.18.1).  MINPACK is not part of /root/reference; its published algorithm
 * (More, Garbow, Hillstrom, "User Guide for MINPACK-1", ANL-80-74: lmdif,
 * fdjac2, qrfac, lmpar, qrsolv, enorm) is restated below.  scipy's wrapper
 * settings: gtol=0, maxfev=200*(n+1), factor=100, mode 1 (diag from column
 * norms), epsfcn = finfo(float32).eps because the residual callback returns
 * float32 (_minpack_py.py:426-434).
 *
 * Parity status: PINNED against golden vectors generated from the real
 * reference (tools/gen_golden.py lq; tests/test_oracle_golden_lq.py).
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define LQ_N 6
#define LQ_MAXBOX 31
#define LQ_MAXM (LQ_MAXBOX * LQ_MAXBOX)

static const double EPSMCH = 2.220446049250313e-16;   /* dpmpar(1) */
static const double DWARF = 2.2250738585072014e-308;  /* dpmpar(2) */

/* MINPACK enorm: scaled Euclidean norm with under/overflow guards */
static double enorm(int n, const double *x) {
    const double rdwarf = 3.834e-20, rgiant = 1.304e19;
    double s1 = 0, s2 = 0, s3 = 0, x1max = 0, x3max = 0;
    double agiant = rgiant / (double)n;
    for (int i = 0; i < n; i++) {
        double xabs = fabs(x[i]);
        if (xabs > rdwarf && xabs < agiant) {
            s2 += xabs * xabs;
        } else if (xabs <= rdwarf) {
            if (xabs > x3max) {
                double r = x3max / xabs;
                s3 = 1.0 + s3 * (r * r);
                x3max = xabs;
            } else if (xabs != 0.0) {
                double r = xabs / x3max;
                s3 += r * r;
            }
        } else {
            if (xabs > x1max) {
                double r = x1max / xabs;
                s1 = 1.0 + s1 * (r * r);
                x1max = xabs;
            } else {
                double r = xabs / x1max;
                s1 += r * r;
            }
        }
    }
    if (s1 != 0.0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
    if (s2 != 0.0) {
        if (s2 >= x3max) return sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)));
        return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrt(s3);
}

/* ---- the reference's residual callback (gausslq.py:33-39,151-203) ---------
 * theta is float64 (MINPACK's x); model_x/model_y/model/residuals are float32
 * buffers, so every evaluation is rounded to f32. */
typedef struct {
    const float *spot;
    int size;
} lq_ctx;

static void residuals(const lq_ctx *c, const double *theta, double *fvec) {
    int size = c->size, h = size / 2;
    float mx[LQ_MAXBOX], my[LQ_MAXBOX];
    double normx = 0.3989422804014327 / theta[4];
    double normy = 0.3989422804014327 / theta[5];
    for (int i = 0; i < size; i++) {
        double grid = (double)(float)(i - h);
        double tx = (grid - theta[0]) / theta[4];
        double ty = (grid - theta[1]) / theta[5];
        mx[i] = (float)(normx * exp(-0.5 * (tx * tx)));
        my[i] = (float)(normy * exp(-0.5 * (ty * ty)));
    }
    for (int i = 0; i < size; i++)
        for (int j = 0; j < size; j++) {
            float model = (float)(theta[2] * (double)my[i] * (double)mx[j] + theta[3]);
            float r = c->spot[i * size + j] - model;      /* f32 - f32 */
            fvec[i * size + j] = (double)r;
        }
}

/* fjac is stored column-major: fjac[i + m*j] (as in MINPACK) */
static void fdjac2(const lq_ctx *c, int m, int n, double *x, const double *fvec, double *fjac,
                   double epsfcn, double *wa) {
    double eps = sqrt(epsfcn > EPSMCH ? epsfcn : EPSMCH);
    for (int j = 0; j < n; j++) {
        double temp = x[j];
        double h = eps * fabs(temp);
        if (h == 0.0) h = eps;
        x[j] = temp + h;
        residuals(c, x, wa);
        x[j] = temp;
        for (int i = 0; i < m; i++) fjac[i + m * j] = (wa[i] - fvec[i]) / h;
    }
}

static void qrfac(int m, int n, double *a, int *ipvt, double *rdiag, double *acnorm, double *wa) {
    for (int j = 0; j < n; j++) {
        acnorm[j] = enorm(m, a + m * j);
        rdiag[j] = acnorm[j];
        wa[j] = rdiag[j];
        ipvt[j] = j;
    }
    int minmn = m < n ? m : n;
    for (int j = 0; j < minmn; j++) {
        int kmax = j;
        for (int k = j; k < n; k++)
            if (rdiag[k] > rdiag[kmax]) kmax = k;
        if (kmax != j) {
            for (int i = 0; i < m; i++) {
                double t = a[i + m * j];
                a[i + m * j] = a[i + m * kmax];
                a[i + m * kmax] = t;
            }
            rdiag[kmax] = rdiag[j];
            wa[kmax] = wa[j];
            int k = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = k;
        }
        double ajnorm = enorm(m - j, a + j + m * j);
        if (ajnorm != 0.0) {
            if (a[j + m * j] < 0.0) ajnorm = -ajnorm;
            for (int i = j; i < m; i++) a[i + m * j] /= ajnorm;
            a[j + m * j] += 1.0;
            for (int k = j + 1; k < n; k++) {
                double sum = 0.0;
                for (int i = j; i < m; i++) sum += a[i + m * j] * a[i + m * k];
                double temp = sum / a[j + m * j];
                for (int i = j; i < m; i++) a[i + m * k] -= temp * a[i + m * j];
                if (rdiag[k] != 0.0) {
                    temp = a[j + m * k] / rdiag[k];
                    double d = 1.0 - temp * temp;
                    rdiag[k] *= sqrt(d > 0.0 ? d : 0.0);
                    double q = rdiag[k] / wa[k];
                    if (0.05 * (q * q) <= EPSMCH) {
                        rdiag[k] = enorm(m - j - 1, a + (j + 1) + m * k);
                        wa[k] = rdiag[k];
                    }
                }
            }
        }
        rdiag[j] = -ajnorm;
    }
}

/* r is n x n inside fjac (leading dimension ldr = m), column-major */
static void qrsolv(int n, double *r, int ldr, const int *ipvt, const double *diag,
                   const double *qtb, double *x, double *sdiag, double *wa) {
    for (int j = 0; j < n; j++) {
        for (int i = j; i < n; i++) r[i + ldr * j] = r[j + ldr * i];
        x[j] = r[j + ldr * j];
        wa[j] = qtb[j];
    }
    for (int j = 0; j < n; j++) {
        int l = ipvt[j];
        if (diag[l] != 0.0) {
            for (int k = j; k < n; k++) sdiag[k] = 0.0;
            sdiag[j] = diag[l];
            double qtbpj = 0.0;
            for (int k = j; k < n; k++) {
                if (sdiag[k] == 0.0) continue;
                double cs, sn;
                if (fabs(r[k + ldr * k]) < fabs(sdiag[k])) {
                    double cotan = r[k + ldr * k] / sdiag[k];
                    sn = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
                    cs = sn * cotan;
                } else {
                    double tn = sdiag[k] / r[k + ldr * k];
                    cs = 0.5 / sqrt(0.25 + 0.25 * (tn * tn));
                    sn = cs * tn;
                }
                r[k + ldr * k] = cs * r[k + ldr * k] + sn * sdiag[k];
                double temp = cs * wa[k] + sn * qtbpj;
                qtbpj = -sn * wa[k] + cs * qtbpj;
                wa[k] = temp;
                for (int i = k + 1; i < n; i++) {
                    temp = cs * r[i + ldr * k] + sn * sdiag[i];
                    sdiag[i] = -sn * r[i + ldr * k] + cs * sdiag[i];
                    r[i + ldr * k] = temp;
                }
            }
        }
        sdiag[j] = r[j + ldr * j];
        r[j + ldr * j] = x[j];
    }
    int nsing = n;
    for (int j = 0; j < n; j++) {
        if (sdiag[j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa[j] = 0.0;
    }
    for (int j = nsing - 1; j >= 0; j--) {
        double sum = 0.0;
        for (int i = j + 1; i < nsing; i++) sum += r[i + ldr * j] * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
    for (int j = 0; j < n; j++) x[ipvt[j]] = wa[j];
}

static void lmpar(int n, double *r, int ldr, const int *ipvt, const double *diag,
                  const double *qtb, double delta, double *par, double *x, double *sdiag,
                  double *wa1, double *wa2) {
    int nsing = n;
    for (int j = 0; j < n; j++) {
        wa1[j] = qtb[j];
        if (r[j + ldr * j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa1[j] = 0.0;
    }
    for (int j = nsing - 1; j >= 0; j--) {
        wa1[j] /= r[j + ldr * j];
        double temp = wa1[j];
        for (int i = 0; i < j; i++) wa1[i] -= r[i + ldr * j] * temp;
    }
    for (int j = 0; j < n; j++) x[ipvt[j]] = wa1[j];

    int iter = 0;
    for (int j = 0; j < n; j++) wa2[j] = diag[j] * x[j];
    double dxnorm = enorm(n, wa2);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) { *par = 0.0; return; }

    double parl = 0.0;
    if (nsing >= n) {
        for (int j = 0; j < n; j++) {
            int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
        }
        for (int j = 0; j < n; j++) {
            double sum = 0.0;
            for (int i = 0; i < j; i++) sum += r[i + ldr * j] * wa1[i];
            wa1[j] = (wa1[j] - sum) / r[j + ldr * j];
        }
        double temp = enorm(n, wa1);
        parl = ((fp / delta) / temp) / temp;
    }
    for (int j = 0; j < n; j++) {
        double sum = 0.0;
        for (int i = 0; i <= j; i++) sum += r[i + ldr * j] * qtb[i];
        wa1[j] = sum / diag[ipvt[j]];
    }
    double gnorm = enorm(n, wa1);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = DWARF / (delta < 0.1 ? delta : 0.1);
    if (*par < parl) *par = parl;
    if (*par > paru) *par = paru;
    if (*par == 0.0) *par = gnorm / dxnorm;

    for (;;) {
        iter++;
        if (*par == 0.0) { double t = 0.001 * paru; *par = DWARF > t ? DWARF : t; }
        double temp = sqrt(*par);
        for (int j = 0; j < n; j++) wa1[j] = temp * diag[j];
        qrsolv(n, r, ldr, ipvt, wa1, qtb, x, sdiag, wa2);
        for (int j = 0; j < n; j++) wa2[j] = diag[j] * x[j];
        dxnorm = enorm(n, wa2);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10)
            break;
        for (int j = 0; j < n; j++) {
            int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
        }
        for (int j = 0; j < n; j++) {
            wa1[j] /= sdiag[j];
            double t = wa1[j];
            for (int i = j + 1; i < n; i++) wa1[i] -= r[i + ldr * j] * t;
        }
        temp = enorm(n, wa1);
        double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0 && *par > parl) parl = *par;
        if (fp < 0.0 && *par < paru) paru = *par;
        double np_ = *par + parc;
        *par = parl > np_ ? parl : np_;
    }
    if (iter == 0) *par = 0.0;
}

/* MINPACK lmdif with scipy.optimize.leastsq's settings; returns info, nfev */
static int lmdif(const lq_ctx *c, int m, double *x, double ftol, double xtol, double gtol,
                 int maxfev, double epsfcn, double factor, int *nfev_out) {
    enum { n = LQ_N };
    static __thread double fvec[LQ_MAXM], fjac[LQ_MAXM * LQ_N], wa4[LQ_MAXM];
    double diag[n], qtf[n], wa1[n], wa2[n], wa3[n];
    int ipvt[n];
    int info = 0, nfev = 0, iter = 1;
    double par = 0.0, delta = 0.0, xnorm = 0.0, fnorm, gnorm = 0.0;

    residuals(c, x, fvec);
    nfev = 1;
    fnorm = enorm(m, fvec);

    for (;;) {
        fdjac2(c, m, n, x, fvec, fjac, epsfcn, wa4);
        nfev += n;
        qrfac(m, n, fjac, ipvt, wa1, wa2, wa3);
        if (iter == 1) {
            for (int j = 0; j < n; j++) {
                diag[j] = wa2[j];
                if (wa2[j] == 0.0) diag[j] = 1.0;
            }
            for (int j = 0; j < n; j++) wa3[j] = diag[j] * x[j];
            xnorm = enorm(n, wa3);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
        for (int i = 0; i < m; i++) wa4[i] = fvec[i];
        for (int j = 0; j < n; j++) {
            if (fjac[j + m * j] != 0.0) {
                double sum = 0.0;
                for (int i = j; i < m; i++) sum += fjac[i + m * j] * wa4[i];
                double temp = -sum / fjac[j + m * j];
                for (int i = j; i < m; i++) wa4[i] += fjac[i + m * j] * temp;
            }
            fjac[j + m * j] = wa1[j];
            qtf[j] = wa4[j];
        }
        gnorm = 0.0;
        if (fnorm != 0.0) {
            for (int j = 0; j < n; j++) {
                int l = ipvt[j];
                if (wa2[l] != 0.0) {
                    double sum = 0.0;
                    for (int i = 0; i <= j; i++) sum += fjac[i + m * j] * (qtf[i] / fnorm);
                    double g = fabs(sum / wa2[l]);
                    if (g > gnorm) gnorm = g;
                }
            }
        }
        if (gnorm <= gtol) { info = 4; break; }
        for (int j = 0; j < n; j++)
            if (wa2[j] > diag[j]) diag[j] = wa2[j];

        double ratio = 0.0;
        do {
            lmpar(n, fjac, m, ipvt, diag, qtf, delta, &par, wa1, wa2, wa3, wa4);
            for (int j = 0; j < n; j++) {
                wa1[j] = -wa1[j];
                wa2[j] = x[j] + wa1[j];
                wa3[j] = diag[j] * wa1[j];
            }
            double pnorm = enorm(n, wa3);
            if (iter == 1 && pnorm < delta) delta = pnorm;
            residuals(c, wa2, wa4);
            nfev++;
            double fnorm1 = enorm(m, wa4);
            double actred = -1.0;
            if (0.1 * fnorm1 < fnorm) { double q = fnorm1 / fnorm; actred = 1.0 - q * q; }
            for (int j = 0; j < n; j++) {
                wa3[j] = 0.0;
                double temp = wa1[ipvt[j]];
                for (int i = 0; i <= j; i++) wa3[i] += fjac[i + m * j] * temp;
            }
            double temp1 = enorm(n, wa3) / fnorm;
            double temp2 = (sqrt(par) * pnorm) / fnorm;
            double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
            double dirder = -(temp1 * temp1 + temp2 * temp2);
            ratio = 0.0;
            if (prered != 0.0) ratio = actred / prered;
            if (ratio <= 0.25) {
                double temp;
                if (actred >= 0.0) temp = 0.5;
                else temp = 0.5 * dirder / (dirder + 0.5 * actred);
                if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
                double pn = pnorm / 0.1;
                delta = temp * (delta < pn ? delta : pn);
                par = par / temp;
            } else if (par == 0.0 || ratio >= 0.75) {
                delta = pnorm / 0.5;
                par = 0.5 * par;
            }
            if (ratio >= 1e-4) {
                for (int j = 0; j < n; j++) {
                    x[j] = wa2[j];
                    wa2[j] = diag[j] * x[j];
                }
                for (int i = 0; i < m; i++) fvec[i] = wa4[i];
                xnorm = enorm(n, wa2);
                fnorm = fnorm1;
                iter++;
            }
            if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0) info = 1;
            if (delta <= xtol * xnorm) info = 2;
            if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0 && info == 2) info = 3;
            if (info != 0) goto done;
            if (nfev >= maxfev) info = 5;
            if (fabs(actred) <= EPSMCH && prered <= EPSMCH && 0.5 * ratio <= 1.0) info = 6;
            if (delta <= EPSMCH * xnorm) info = 7;
            if (gnorm <= EPSMCH) info = 8;
            if (info != 0) goto done;
        } while (ratio < 1e-4);
    }
done:
    *nfev_out = nfev;
    return info;
}

/* gausslq.py:51-112 _initial_parameters -> theta0 float32[6] (x,y relative to centre) */
void orc_lq_initial_parameters(const float *spot, int size, float *theta) {
    int h = size / 2;
    float mn = spot[0];
    for (int i = 1; i < size * size; i++)
        if (spot[i] < mn) mn = spot[i];
    theta[3] = mn;
    double y = 0.0, x = 0.0, s = 0.0;
    for (int i = 0; i < size; i++)
        for (int j = 0; j < size; j++) {
            float v = spot[i * size + j] - mn;            /* f32 array - f32 scalar */
            y += (double)v * (double)i;
            x += (double)v * (double)j;
            s += (double)v;
        }
    double sum;
    if (s <= 0.0) { sum = 0.01; y = (size - 1) / 2.0; x = (size - 1) / 2.0; }
    else { sum = s; y /= s; x /= s; }
    theta[1] = (float)y;
    theta[0] = (float)x;
    theta[2] = (float)(sum > 1.0 ? sum : 1.0);
    double sdy = 0.0, sdx = 0.0;
    for (int i = 0; i < size; i++)
        for (int j = 0; j < size; j++) {
            float v = spot[i * size + j] - mn;
            double dy = (double)i - (double)theta[1];
            double dx = (double)j - (double)theta[0];
            sdy += (double)v * (dy * dy);
            sdx += (double)v * (dx * dx);
        }
    theta[5] = (float)sqrt(sdy / sum);
    theta[4] = (float)sqrt(sdx / sum);
    theta[0] = (float)((double)theta[0] - (double)h);
    theta[1] = (float)((double)theta[1] - (double)h);
}

/* gausslq.py:206-244 fit_spot: returns MINPACK info; theta_out float32[6] */
int orc_lq_fit_one(const float *spot, int size, float *theta_out, int *nfev_out) {
    float t0[6];
    orc_lq_initial_parameters(spot, size, t0);
    double x[6];
    for (int k = 0; k < 6; k++) x[k] = (double)t0[k];
    lq_ctx c = {spot, size};
    int nfev = 0;
    int info = lmdif(&c, size * size, x, 1e-2, 1e-2, 0.0, 200 * (LQ_N + 1),
                     (double)FLT_EPSILON, 100.0, &nfev);
    for (int k = 0; k < 6; k++) theta_out[k] = (float)x[k];
    if (nfev_out) *nfev_out = nfev;
    return info;
}

typedef struct {
    const float *spots; long long begin, end; int size; float *thetas; int *infos; int *nfevs;
} lq_job;

static void *lq_worker(void *p) {
    lq_job *j = (lq_job *)p;
    for (long long i = j->begin; i < j->end; i++) {
        int nfev;
        int info = orc_lq_fit_one(j->spots + i * j->size * j->size, j->size, j->thetas + 6 * i, &nfev);
        if (j->infos) j->infos[i] = info;
        if (j->nfevs) j->nfevs[i] = nfev;
    }
    return NULL;
}

/* gausslq.py:247-289 fit_spots (+ the process pool of :292-343 as threads) */
int orc_lq_fit_mt(const float *spots, long long n, int size, float *thetas, int *infos, int *nfevs,
                  int nthreads) {
    if (size < 1 || size > LQ_MAXBOX) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    lq_job jobs[256];
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (lq_job){spots, n * t / nthreads, n * (t + 1) / nthreads, size, thetas, infos, nfevs};
        pthread_create(&th[t], NULL, lq_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}
