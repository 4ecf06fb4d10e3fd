// picasso_b200/csrc/lq_core.cuh
//
// Per-spot arithmetic of the least-squares fit (lq_fit.cu): MINPACK-1 `lmdif` (More / Garbow /
// Hillstrom, ANL-80-74: lmdif, fdjac2, qrfac, lmpar, qrsolv, enorm) driving picasso's
// float32-rounded, point-sampled Gaussian model (reference picasso/gausslq.py:33-39, 51-112,
// 151-244; scipy.optimize.leastsq(ftol=xtol=1e-2) with epsfcn = eps_f32).  Scalar code that one
// thread runs for one spot, written once for the device (nvcc, sm_100a) and for the host (g++):
// tests/host_sim compiles the same functions on the CPU so the optimiser's trajectory is checked
// against the oracle / the reference's golden vectors without a GPU.  The host build is test
// scaffolding only -- the product has no CPU path.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define LQ_FN __device__
#define LQ_FN_NOINLINE __device__ __noinline__
#define LQ_FSUB(a, b) __fsub_rn((a), (b))
#else
#define LQ_FN static inline
#define LQ_FN_NOINLINE static
#define LQ_FSUB(a, b) ((float)((float)(a) - (float)(b)))
#endif

namespace lq {
constexpr int kN = 6;
constexpr double kEpsMch = 2.220446049250313e-16;
constexpr double kDwarf = 2.2250738585072014e-308;

LQ_FN double sq(double v) { return v * v; }

// MINPACK enorm (three accumulators guard against under/overflow)
LQ_FN_NOINLINE double lq_enorm(int n, const double* x) {
    const double rdwarf = 3.834e-20, rgiant = 1.304e19;
    double s1 = 0, s2 = 0, s3 = 0, x1max = 0, x3max = 0;
    const double agiant = rgiant / (double)n;
    for (int i = 0; i < n; i++) {
        const double xabs = fabs(x[i]);
        if (xabs > rdwarf && xabs < agiant) {
            s2 += xabs * xabs;
        } else if (xabs <= rdwarf) {
            if (xabs > x3max) { s3 = 1.0 + s3 * sq(x3max / xabs); x3max = xabs; }
            else if (xabs != 0.0) s3 += sq(xabs / x3max);
        } else {
            if (xabs > x1max) { s1 = 1.0 + s1 * sq(x1max / xabs); x1max = xabs; }
            else s1 += sq(xabs / x1max);
        }
    }
    if (s1 != 0.0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
    if (s2 != 0.0) {
        if (s2 >= x3max) return sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)));
        return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrt(s3);
}

// residuals = spot - f32(model), model = f32(N * gy[i] * gx[j] + bg) with
// gx, gy float32 point-sampled normalised Gaussians (gausslq.py:33-39,151-203)
template <int BOX>
LQ_FN void lq_residuals(const float* spot /*smem, stride 1*/, const double* th, double* fvec) {
    constexpr int H = BOX / 2;
    float mx[BOX], my[BOX];
    const double nx = 0.3989422804014327 / th[4];
    const double ny = 0.3989422804014327 / th[5];
#pragma unroll
    for (int i = 0; i < BOX; i++) {
        const double grid = (double)(i - H);
        const double tx = (grid - th[0]) / th[4];
        const double ty = (grid - th[1]) / th[5];
        mx[i] = (float)(nx * exp(-0.5 * (tx * tx)));
        my[i] = (float)(ny * exp(-0.5 * (ty * ty)));
    }
#pragma unroll
    for (int i = 0; i < BOX; i++) {
        const double nmy = th[2] * (double)my[i];
#pragma unroll
        for (int j = 0; j < BOX; j++) {
            const float model = (float)(nmy * (double)mx[j] + th[3]);
            fvec[i * BOX + j] = (double)LQ_FSUB(spot[i * BOX + j], model);
        }
    }
}

template <int M>
LQ_FN void lq_qrfac(double* a /*M x 6 col-major*/, int* ipvt, double* rdiag, double* acnorm,
                         double* wa) {
    for (int j = 0; j < kN; j++) {
        acnorm[j] = lq_enorm(M, a + M * j);
        rdiag[j] = acnorm[j];
        wa[j] = rdiag[j];
        ipvt[j] = j;
    }
    for (int j = 0; j < kN; j++) {
        int kmax = j;
        for (int k = j; k < kN; k++)
            if (rdiag[k] > rdiag[kmax]) kmax = k;
        if (kmax != j) {
            for (int i = 0; i < M; i++) {
                const double t = a[i + M * j];
                a[i + M * j] = a[i + M * kmax];
                a[i + M * kmax] = t;
            }
            rdiag[kmax] = rdiag[j];
            wa[kmax] = wa[j];
            const int k = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = k;
        }
        double ajnorm = lq_enorm(M - j, a + j + M * j);
        if (ajnorm != 0.0) {
            if (a[j + M * j] < 0.0) ajnorm = -ajnorm;
            for (int i = j; i < M; i++) a[i + M * j] /= ajnorm;
            a[j + M * j] += 1.0;
            for (int k = j + 1; k < kN; k++) {
                double sum = 0.0;
                for (int i = j; i < M; i++) sum += a[i + M * j] * a[i + M * k];
                double temp = sum / a[j + M * j];
                for (int i = j; i < M; i++) a[i + M * k] -= temp * a[i + M * j];
                if (rdiag[k] != 0.0) {
                    temp = a[j + M * k] / rdiag[k];
                    rdiag[k] *= sqrt(fmax(0.0, 1.0 - temp * temp));
                    if (0.05 * sq(rdiag[k] / wa[k]) <= kEpsMch) {
                        rdiag[k] = lq_enorm(M - j - 1, a + (j + 1) + M * k);
                        wa[k] = rdiag[k];
                    }
                }
            }
        }
        rdiag[j] = -ajnorm;
    }
}

LQ_FN void lq_qrsolv(double* r, int ldr, const int* ipvt, const double* diag,
                          const double* qtb, double* x, double* sdiag, double* wa) {
#pragma unroll 1
    for (int j = 0; j < kN; j++) {
#pragma unroll 1
        for (int i = j; i < kN; i++) r[i + ldr * j] = r[j + ldr * i];
        x[j] = r[j + ldr * j];
        wa[j] = qtb[j];
    }
#pragma unroll 1
    for (int j = 0; j < kN; j++) {
        const int l = ipvt[j];
        if (diag[l] != 0.0) {
#pragma unroll 1
            for (int k = j; k < kN; k++) sdiag[k] = 0.0;
            sdiag[j] = diag[l];
            double qtbpj = 0.0;
#pragma unroll 1
            for (int k = j; k < kN; k++) {
                if (sdiag[k] == 0.0) continue;
                double cs, sn;
                if (fabs(r[k + ldr * k]) < fabs(sdiag[k])) {
                    const double cotan = r[k + ldr * k] / sdiag[k];
                    sn = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
                    cs = sn * cotan;
                } else {
                    const double tn = sdiag[k] / r[k + ldr * k];
                    cs = 0.5 / sqrt(0.25 + 0.25 * (tn * tn));
                    sn = cs * tn;
                }
                r[k + ldr * k] = cs * r[k + ldr * k] + sn * sdiag[k];
                double temp = cs * wa[k] + sn * qtbpj;
                qtbpj = -sn * wa[k] + cs * qtbpj;
                wa[k] = temp;
#pragma unroll 1
                for (int i = k + 1; i < kN; i++) {
                    temp = cs * r[i + ldr * k] + sn * sdiag[i];
                    sdiag[i] = -sn * r[i + ldr * k] + cs * sdiag[i];
                    r[i + ldr * k] = temp;
                }
            }
        }
        sdiag[j] = r[j + ldr * j];
        r[j + ldr * j] = x[j];
    }
    int nsing = kN;
#pragma unroll 1
    for (int j = 0; j < kN; j++) {
        if (sdiag[j] == 0.0 && nsing == kN) nsing = j;
        if (nsing < kN) wa[j] = 0.0;
    }
#pragma unroll 1
    for (int j = nsing - 1; j >= 0; j--) {
        double sum = 0.0;
#pragma unroll 1
        for (int i = j + 1; i < nsing; i++) sum += r[i + ldr * j] * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
#pragma unroll 1
    for (int j = 0; j < kN; j++) x[ipvt[j]] = wa[j];
}

LQ_FN void lq_lmpar(double* r, int ldr, const int* ipvt, const double* diag,
                         const double* qtb, double delta, double& par, double* x, double* sdiag,
                         double* wa1, double* wa2) {
    int nsing = kN;
#pragma unroll 1
    for (int j = 0; j < kN; j++) {
        wa1[j] = qtb[j];
        if (r[j + ldr * j] == 0.0 && nsing == kN) nsing = j;
        if (nsing < kN) wa1[j] = 0.0;
    }
#pragma unroll 1
    for (int j = nsing - 1; j >= 0; j--) {
        wa1[j] /= r[j + ldr * j];
        const double temp = wa1[j];
#pragma unroll 1
        for (int i = 0; i < j; i++) wa1[i] -= r[i + ldr * j] * temp;
    }
#pragma unroll 1
    for (int j = 0; j < kN; j++) x[ipvt[j]] = wa1[j];
    int iter = 0;
#pragma unroll 1
    for (int j = 0; j < kN; j++) wa2[j] = diag[j] * x[j];
    double dxnorm = lq_enorm(kN, wa2);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) { par = 0.0; return; }
    double parl = 0.0;
    if (nsing >= kN) {
#pragma unroll 1
        for (int j = 0; j < kN; j++) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
#pragma unroll 1
        for (int j = 0; j < kN; j++) {
            double sum = 0.0;
#pragma unroll 1
            for (int i = 0; i < j; i++) sum += r[i + ldr * j] * wa1[i];
            wa1[j] = (wa1[j] - sum) / r[j + ldr * j];
        }
        const double temp = lq_enorm(kN, wa1);
        parl = ((fp / delta) / temp) / temp;
    }
#pragma unroll 1
    for (int j = 0; j < kN; j++) {
        double sum = 0.0;
#pragma unroll 1
        for (int i = 0; i <= j; i++) sum += r[i + ldr * j] * qtb[i];
        wa1[j] = sum / diag[ipvt[j]];
    }
    const double gnorm = lq_enorm(kN, wa1);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = kDwarf / fmin(delta, 0.1);
    par = fmin(fmax(par, parl), paru);
    if (par == 0.0) par = gnorm / dxnorm;
#pragma unroll 1
    for (;;) {
        iter++;
        if (par == 0.0) par = fmax(kDwarf, 0.001 * paru);
        double temp = sqrt(par);
#pragma unroll 1
        for (int j = 0; j < kN; j++) wa1[j] = temp * diag[j];
        lq_qrsolv(r, ldr, ipvt, wa1, qtb, x, sdiag, wa2);
#pragma unroll 1
        for (int j = 0; j < kN; j++) wa2[j] = diag[j] * x[j];
        dxnorm = lq_enorm(kN, wa2);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
#pragma unroll 1
        for (int j = 0; j < kN; j++) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
#pragma unroll 1
        for (int j = 0; j < kN; j++) {
            wa1[j] /= sdiag[j];
            const double t = wa1[j];
#pragma unroll 1
            for (int i = j + 1; i < kN; i++) wa1[i] -= r[i + ldr * j] * t;
        }
        temp = lq_enorm(kN, wa1);
        const double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0) parl = fmax(parl, par);
        if (fp < 0.0) paru = fmin(paru, par);
        par = fmax(parl, par + parc);
    }
}

// ---- MINPACK-order variant: m x 6 forward-difference Jacobian, Householder qrfac ----
template <int BOX>
LQ_FN void fit_spot_qr(const float* spot, double* x, int* info_out, int* nfev_out) {
    constexpr int M = BOX * BOX;
    constexpr int H = BOX / 2;
    // ---- start values (gausslq.py:51-112) ----
    {
        float mn = spot[0];
        for (int i = 1; i < M; i++) mn = fminf(mn, spot[i]);
        double y = 0.0, xs = 0.0, s = 0.0;
        for (int i = 0; i < BOX; i++)
            for (int j = 0; j < BOX; j++) {
                const double v = (double)LQ_FSUB(spot[i * BOX + j], mn);
                y += v * (double)i;
                xs += v * (double)j;
                s += v;
            }
        double sum;
        if (s <= 0.0) { sum = 0.01; y = (BOX - 1) / 2.0; xs = (BOX - 1) / 2.0; }
        else { sum = s; y /= s; xs /= s; }
        const float ty = (float)y, tx = (float)xs;
        double sdy = 0.0, sdx = 0.0;
        for (int i = 0; i < BOX; i++)
            for (int j = 0; j < BOX; j++) {
                const double v = (double)LQ_FSUB(spot[i * BOX + j], mn);
                sdy += v * sq((double)i - (double)ty);
                sdx += v * sq((double)j - (double)tx);
            }
        x[0] = (double)(float)((double)tx - (double)H);
        x[1] = (double)(float)((double)ty - (double)H);
        x[2] = (double)(float)fmax(sum, 1.0);
        x[3] = (double)mn;
        x[4] = (double)(float)sqrt(sdx / sum);
        x[5] = (double)(float)sqrt(sdy / sum);
    }

    // ---- lmdif (ftol = xtol = 1e-2, gtol = 0, maxfev = 1400, factor = 100,
    //      epsfcn = eps_f32, mode 1) ----
    const double ftol = 1e-2, xtol = 1e-2, gtol = 0.0, factor = 100.0;
    const int maxfev = 200 * (kN + 1);
    const double fd_eps = sqrt(fmax((double)1.1920928955078125e-07, kEpsMch));
    double fvec[M], fjac[M * kN], wa4[M];
    double diag[kN], qtf[kN], wa1[kN], wa2[kN], wa3[kN];
    int ipvt[kN];
    int info = 0, nfev, iter = 1;
    double par = 0.0, delta = 0.0, xnorm = 0.0, gnorm = 0.0;

    lq_residuals<BOX>(spot, x, fvec);
    nfev = 1;
    double fnorm = lq_enorm(M, fvec);
    bool finished = false;
    while (!finished) {
        // forward-difference Jacobian (fdjac2)
        for (int j = 0; j < kN; j++) {
            const double temp = x[j];
            double h = fd_eps * fabs(temp);
            if (h == 0.0) h = fd_eps;
            x[j] = temp + h;
            lq_residuals<BOX>(spot, x, wa4);
            x[j] = temp;
            for (int i = 0; i < M; i++) fjac[i + M * j] = (wa4[i] - fvec[i]) / h;
        }
        nfev += kN;
        lq_qrfac<M>(fjac, ipvt, wa1, wa2, wa3);
        if (iter == 1) {
            for (int j = 0; j < kN; j++) diag[j] = (wa2[j] == 0.0) ? 1.0 : wa2[j];
            for (int j = 0; j < kN; j++) wa3[j] = diag[j] * x[j];
            xnorm = lq_enorm(kN, wa3);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
        for (int i = 0; i < M; i++) wa4[i] = fvec[i];
        for (int j = 0; j < kN; j++) {
            if (fjac[j + M * j] != 0.0) {
                double sum = 0.0;
                for (int i = j; i < M; i++) sum += fjac[i + M * j] * wa4[i];
                const double temp = -sum / fjac[j + M * j];
                for (int i = j; i < M; i++) wa4[i] += fjac[i + M * j] * temp;
            }
            fjac[j + M * j] = wa1[j];
            qtf[j] = wa4[j];
        }
        gnorm = 0.0;
        if (fnorm != 0.0) {
            for (int j = 0; j < kN; j++) {
                const int l = ipvt[j];
                if (wa2[l] != 0.0) {
                    double sum = 0.0;
                    for (int i = 0; i <= j; i++) sum += fjac[i + M * j] * (qtf[i] / fnorm);
                    gnorm = fmax(gnorm, fabs(sum / wa2[l]));
                }
            }
        }
        if (gnorm <= gtol) { info = 4; break; }
        for (int j = 0; j < kN; j++) diag[j] = fmax(diag[j], wa2[j]);

        double ratio = 0.0;
        do {
            lq_lmpar(fjac, M, ipvt, diag, qtf, delta, par, wa1, wa2, wa3, wa4);
            for (int j = 0; j < kN; j++) {
                wa1[j] = -wa1[j];
                wa2[j] = x[j] + wa1[j];
                wa3[j] = diag[j] * wa1[j];
            }
            const double pnorm = lq_enorm(kN, wa3);
            if (iter == 1) delta = fmin(delta, pnorm);
            lq_residuals<BOX>(spot, wa2, wa4);
            nfev++;
            const double fnorm1 = lq_enorm(M, wa4);
            double actred = -1.0;
            if (0.1 * fnorm1 < fnorm) actred = 1.0 - sq(fnorm1 / fnorm);
            for (int j = 0; j < kN; j++) {
                wa3[j] = 0.0;
                const double temp = wa1[ipvt[j]];
                for (int i = 0; i <= j; i++) wa3[i] += fjac[i + M * j] * temp;
            }
            const double temp1 = lq_enorm(kN, wa3) / fnorm;
            const double temp2 = (sqrt(par) * pnorm) / fnorm;
            const double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
            const double dirder = -(temp1 * temp1 + temp2 * temp2);
            ratio = (prered != 0.0) ? actred / prered : 0.0;
            if (ratio <= 0.25) {
                double temp = (actred >= 0.0) ? 0.5 : 0.5 * dirder / (dirder + 0.5 * actred);
                if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
                delta = temp * fmin(delta, pnorm / 0.1);
                par = par / temp;
            } else if (par == 0.0 || ratio >= 0.75) {
                delta = pnorm / 0.5;
                par = 0.5 * par;
            }
            if (ratio >= 1e-4) {
                for (int j = 0; j < kN; j++) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
                for (int i = 0; i < M; i++) fvec[i] = wa4[i];
                xnorm = lq_enorm(kN, wa2);
                fnorm = fnorm1;
                iter++;
            }
            const bool small = fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0;
            if (small) info = 1;
            if (delta <= xtol * xnorm) info = 2;
            if (small && info == 2) info = 3;
            if (info != 0) { finished = true; break; }
            if (nfev >= maxfev) info = 5;
            if (fabs(actred) <= kEpsMch && prered <= kEpsMch && 0.5 * ratio <= 1.0) info = 6;
            if (delta <= kEpsMch * xnorm) info = 7;
            if (gnorm <= kEpsMch) info = 8;
            if (info != 0) { finished = true; break; }
        } while (ratio < 1e-4);
    }
    *info_out = info;
    *nfev_out = nfev;
}

// ============================================================================
// Register-resident variant (default).  MINPACK's lmdif needs the m x 6 Jacobian only
// through R (pivoted QR), Q^T f (first 6 components) and the column norms.  In exact
// arithmetic R is the pivoted Cholesky factor of J^T J and Q^T f = R^-T P^T J^T f (up to
// row signs, which cancel in every later use: R^T qtf, R x = qtf, norms).  So the kernel
// streams over the pixels once per LM iteration, evaluating the forward differences on the
// fly and accumulating J^T J (21), J^T f (6) and ||f||^2 -- no m x 6 array, no local-memory
// traffic (the QR variant above keeps 3 KB per thread in local memory and runs at 5 % issue
// utilisation).  cond(J)^2 * eps ~ 1e-8 relative on the LM step: far below what can flip
// one of lmdif's coarse decisions (ftol = xtol = 1e-2, gain-ratio thresholds), so the
// optimiser trajectory -- number of function evaluations -- is the same as scipy's
// (tests/test_lq_gpu.py).  Everything after the factorisation (lmpar, qrsolv, step
// acceptance, convergence tests) is the MINPACK logic unchanged.

// float32 point-sampled normalised Gaussian vector (gausslq.py:33-39)
// lq_enorm, lq_gauss_vec and lq_resnorm are kept out of line and the small linear-algebra loops
// rolled: inlined and unrolled at every call site (10 Gaussian vectors x 7 float64 divisions +
// exponentials, 8 enorm copies with three division branches x 6, 21-fold unrolled Givens /
// back-substitution steps) the kernel was 330 KB of SASS and stalled on instruction fetch
// (ncu: no_instruction 4.0 stall cycles per issued instruction, issue slots 21 % busy).
template <int BOX>
LQ_FN_NOINLINE void lq_gauss_vec(double mu, double sigma, float* out) {
    constexpr int H = BOX / 2;
    const double nrm = 0.3989422804014327 / sigma;
#pragma unroll
    for (int i = 0; i < BOX; i++) {
        const double t = ((double)(i - H) - mu) / sigma;
        out[i] = (float)(nrm * exp(-0.5 * (t * t)));
    }
}

// ||spot - f32(model)|| with MINPACK's enorm summation (mid-range branch) in pixel order
template <int BOX>
LQ_FN_NOINLINE double lq_resnorm(const float* spot, const double* x) {
    float mx[BOX], my[BOX];
    lq_gauss_vec<BOX>(x[0], x[4], mx);
    lq_gauss_vec<BOX>(x[1], x[5], my);
    double s2 = 0.0;
#pragma unroll 1
    for (int i = 0; i < BOX; i++) {
        const double nmy = x[2] * (double)my[i];
#pragma unroll
        for (int j = 0; j < BOX; j++) {
            const float model = (float)(nmy * (double)mx[j] + x[3]);
            const double r = (double)LQ_FSUB(spot[i * BOX + j], model);
            s2 = fma(r, r, s2);
        }
    }
    return sqrt(s2);
}

template <int BOX>
LQ_FN void fit_spot_ne(const float* spot, double* x, int* info_out, int* nfev_out) {
    constexpr int M = BOX * BOX; (void)M;
    constexpr int H = BOX / 2;
    // ---- start values (gausslq.py:51-112) ----
    {
        float mn = spot[0];
        for (int i = 1; i < M; i++) mn = fminf(mn, spot[i]);
        double y = 0.0, xs = 0.0, s = 0.0;
        for (int i = 0; i < BOX; i++)
            for (int j = 0; j < BOX; j++) {
                const double v = (double)LQ_FSUB(spot[i * BOX + j], mn);
                y += v * (double)i;
                xs += v * (double)j;
                s += v;
            }
        double sum;
        if (s <= 0.0) { sum = 0.01; y = (BOX - 1) / 2.0; xs = (BOX - 1) / 2.0; }
        else { sum = s; y /= s; xs /= s; }
        const float ty = (float)y, tx = (float)xs;
        double sdy = 0.0, sdx = 0.0;
        for (int i = 0; i < BOX; i++)
            for (int j = 0; j < BOX; j++) {
                const double v = (double)LQ_FSUB(spot[i * BOX + j], mn);
                sdy += v * sq((double)i - (double)ty);
                sdx += v * sq((double)j - (double)tx);
            }
        x[0] = (double)(float)((double)tx - (double)H);
        x[1] = (double)(float)((double)ty - (double)H);
        x[2] = (double)(float)fmax(sum, 1.0);
        x[3] = (double)mn;
        x[4] = (double)(float)sqrt(sdx / sum);
        x[5] = (double)(float)sqrt(sdy / sum);
    }

    const double ftol = 1e-2, xtol = 1e-2, gtol = 0.0, factor = 100.0;
    const int maxfev = 200 * (kN + 1);
    const double fd_eps = sqrt(fmax((double)1.1920928955078125e-07, kEpsMch));
    double r[kN * kN];                       // R (upper) / S (lower) work matrix, ld = 6
    double diag[kN], qtf[kN], wa1[kN], wa2[kN], wa3[kN], wa4[kN], acn[kN];
    int ipvt[kN];
    int info = 0, nfev, iter = 1;
    double par = 0.0, delta = 0.0, xnorm = 0.0, gnorm = 0.0;

    double fnorm = lq_resnorm<BOX>(spot, x);
    nfev = 1;
    bool finished = false;
    while (!finished) {
        // ---- one pass over the pixels: forward differences -> J^T J, J^T f ----
        double A[21], g[kN];
#pragma unroll
        for (int q = 0; q < 21; q++) A[q] = 0.0;
#pragma unroll
        for (int q = 0; q < kN; q++) g[q] = 0.0;
        {
            double h[kN];
#pragma unroll
            for (int j = 0; j < kN; j++) {
                h[j] = fd_eps * fabs(x[j]);
                if (h[j] == 0.0) h[j] = fd_eps;
            }
            // x[j] + h is formed exactly as fdjac2 does; the divisor is h itself
            const double x0p = x[0] + h[0], x1p = x[1] + h[1], x2p = x[2] + h[2];
            const double x3p = x[3] + h[3], x4p = x[4] + h[4], x5p = x[5] + h[5];
            float mx[BOX], my[BOX], mx0[BOX], mx4[BOX], my1[BOX], my5[BOX];
            lq_gauss_vec<BOX>(x[0], x[4], mx);
            lq_gauss_vec<BOX>(x[1], x[5], my);
            lq_gauss_vec<BOX>(x0p, x[4], mx0);
            lq_gauss_vec<BOX>(x[0], x4p, mx4);
            lq_gauss_vec<BOX>(x1p, x[5], my1);
            lq_gauss_vec<BOX>(x[1], x5p, my5);
#pragma unroll 1
            for (int i = 0; i < BOX; i++) {
                const double nmy = x[2] * (double)my[i];
                const double nmy1 = x[2] * (double)my1[i];
                const double nmy5 = x[2] * (double)my5[i];
                const double pmy = x2p * (double)my[i];
#pragma unroll 1
                for (int j = 0; j < BOX; j++) {
                    const float sp = spot[i * BOX + j];
                    const double mxj = (double)mx[j];
                    const double r0 = (double)LQ_FSUB(sp, (float)(nmy * mxj + x[3]));
                    double J[kN];
                    J[0] = ((double)LQ_FSUB(sp, (float)(nmy * (double)mx0[j] + x[3])) - r0) / h[0];
                    J[1] = ((double)LQ_FSUB(sp, (float)(nmy1 * mxj + x[3])) - r0) / h[1];
                    J[2] = ((double)LQ_FSUB(sp, (float)(pmy * mxj + x[3])) - r0) / h[2];
                    J[3] = ((double)LQ_FSUB(sp, (float)(nmy * mxj + x3p)) - r0) / h[3];
                    J[4] = ((double)LQ_FSUB(sp, (float)(nmy * (double)mx4[j] + x[3])) - r0) / h[4];
                    J[5] = ((double)LQ_FSUB(sp, (float)(nmy5 * mxj + x[3])) - r0) / h[5];
                    int q = 0;
#pragma unroll
                    for (int a_ = 0; a_ < kN; a_++) {
                        g[a_] = fma(J[a_], r0, g[a_]);
#pragma unroll
                        for (int b_ = a_; b_ < kN; b_++) { A[q] = fma(J[a_], J[b_], A[q]); q++; }
                    }
                }
            }
        }
        nfev += kN;
        auto Aat = [&](int a_, int b_) -> double {     // symmetric access to packed A
            const int lo = a_ < b_ ? a_ : b_, hi = a_ < b_ ? b_ : a_;
            return A[lo * kN - (lo * (lo - 1)) / 2 + (hi - lo)];
        };
        // column norms of J (qrfac's acnorm) and pivoted Cholesky = R of the pivoted QR
#pragma unroll
        for (int j = 0; j < kN; j++) { acn[j] = sqrt(Aat(j, j)); ipvt[j] = j; }
        {
            double rem[kN];                  // remaining squared column norms
#pragma unroll
            for (int j = 0; j < kN; j++) rem[j] = Aat(j, j);
#pragma unroll 1
            for (int q = 0; q < kN * kN; q++) r[q] = 0.0;
#pragma unroll 1
            for (int j = 0; j < kN; j++) {
                int kmax = j;
#pragma unroll 1
                for (int k = j; k < kN; k++)
                    if (rem[k] > rem[kmax]) kmax = k;
                if (kmax != j) {
                    // swap pivot columns j <-> kmax: permutation, remaining norms, computed rows
                    const int t = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = t;
                    const double tr = rem[j]; rem[j] = rem[kmax]; rem[kmax] = tr;
#pragma unroll 1
                    for (int l = 0; l < j; l++) {
                        const double tv = r[l + kN * j]; r[l + kN * j] = r[l + kN * kmax];
                        r[l + kN * kmax] = tv;
                    }
                }
                const double d = rem[j];
                const double rjj = d > 0.0 ? sqrt(d) : 0.0;
                r[j + kN * j] = rjj;
#pragma unroll 1
                for (int k = j + 1; k < kN; k++) {
                    double v = 0.0;
                    if (rjj > 0.0) {
                        v = Aat(ipvt[j], ipvt[k]);
#pragma unroll 1
                        for (int l = 0; l < j; l++) v -= r[l + kN * j] * r[l + kN * k];
                        v /= rjj;
                    }
                    r[j + kN * k] = v;
                    rem[k] -= v * v;
                }
            }
        }
        // qtf = R^-T P^T (J^T f)
#pragma unroll 1
        for (int j = 0; j < kN; j++) {
            double v = g[ipvt[j]];
#pragma unroll 1
            for (int l = 0; l < j; l++) v -= r[l + kN * j] * qtf[l];
            qtf[j] = (r[j + kN * j] != 0.0) ? v / r[j + kN * j] : 0.0;
        }
        if (iter == 1) {
#pragma unroll 1
            for (int j = 0; j < kN; j++) diag[j] = (acn[j] == 0.0) ? 1.0 : acn[j];
#pragma unroll 1
            for (int j = 0; j < kN; j++) wa3[j] = diag[j] * x[j];
            xnorm = lq_enorm(kN, wa3);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
        gnorm = 0.0;
        if (fnorm != 0.0) {
#pragma unroll 1
            for (int j = 0; j < kN; j++) {
                const int l = ipvt[j];
                if (acn[l] != 0.0) {
                    double sum = 0.0;
#pragma unroll 1
                    for (int i = 0; i <= j; i++) sum += r[i + kN * j] * (qtf[i] / fnorm);
                    gnorm = fmax(gnorm, fabs(sum / acn[l]));
                }
            }
        }
        if (gnorm <= gtol) { info = 4; break; }
#pragma unroll 1
        for (int j = 0; j < kN; j++) diag[j] = fmax(diag[j], acn[j]);

        double ratio = 0.0;
        do {
            lq_lmpar(r, kN, ipvt, diag, qtf, delta, par, wa1, wa2, wa3, wa4);
#pragma unroll 1
            for (int j = 0; j < kN; j++) {
                wa1[j] = -wa1[j];
                wa2[j] = x[j] + wa1[j];
                wa3[j] = diag[j] * wa1[j];
            }
            const double pnorm = lq_enorm(kN, wa3);
            if (iter == 1) delta = fmin(delta, pnorm);
            const double fnorm1 = lq_resnorm<BOX>(spot, wa2);
            nfev++;
            double actred = -1.0;
            if (0.1 * fnorm1 < fnorm) actred = 1.0 - sq(fnorm1 / fnorm);
#pragma unroll 1
            for (int j = 0; j < kN; j++) {
                wa3[j] = 0.0;
                const double temp = wa1[ipvt[j]];
#pragma unroll 1
                for (int i = 0; i <= j; i++) wa3[i] += r[i + kN * j] * temp;
            }
            const double temp1 = lq_enorm(kN, wa3) / fnorm;
            const double temp2 = (sqrt(par) * pnorm) / fnorm;
            const double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
            const double dirder = -(temp1 * temp1 + temp2 * temp2);
            ratio = (prered != 0.0) ? actred / prered : 0.0;
            if (ratio <= 0.25) {
                double temp = (actred >= 0.0) ? 0.5 : 0.5 * dirder / (dirder + 0.5 * actred);
                if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
                delta = temp * fmin(delta, pnorm / 0.1);
                par = par / temp;
            } else if (par == 0.0 || ratio >= 0.75) {
                delta = pnorm / 0.5;
                par = 0.5 * par;
            }
            if (ratio >= 1e-4) {
#pragma unroll 1
                for (int j = 0; j < kN; j++) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
                xnorm = lq_enorm(kN, wa2);
                fnorm = fnorm1;
                iter++;
            }
            const bool small = fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0;
            if (small) info = 1;
            if (delta <= xtol * xnorm) info = 2;
            if (small && info == 2) info = 3;
            if (info != 0) { finished = true; break; }
            if (nfev >= maxfev) info = 5;
            if (fabs(actred) <= kEpsMch && prered <= kEpsMch && 0.5 * ratio <= 1.0) info = 6;
            if (delta <= kEpsMch * xnorm) info = 7;
            if (gnorm <= kEpsMch) info = 8;
            if (info != 0) { finished = true; break; }
        } while (ratio < 1e-4);
    }
    *info_out = info;
    *nfev_out = nfev;
}

}  // namespace lq
