// picasso_b200/csrc/zfit.cu -- astigmatic z fit, one thread per localization (sm_100a).
//
// Replaces the per-localization loop of picasso.zfit._fit_z (picasso/zfit.py:327-383):
//   scipy.optimize.minimize_scalar(_fit_z_target, bounds=[-1000, 1000], args=(sx, sy, cx, cy))
// = scipy's bounded Brent minimiser (_minimize_scalar_bounded, xatol 1e-5, maxiter 500) on the
// numba-compiled target (zfit.py:255-291, float64 with sx, sy promoted from float32), plus the
// column arithmetic that follows it: z * magnification, sqrt(fun) and the axial localization
// precision _axial_localization_precision_astig (zfit.py:805-890) with gausslq / gaussmle
// sigma_uncertainty (gausslq.py:592-633, gaussmle.py:1040-1074), which the reference evaluates
// on float32 pandas columns.
//
// This file is compiled with -fmad=false: the minimiser follows the reference's float64
// trajectory operation by operation (no fused multiply-adds, IEEE sqrt / division), so z, fun
// and the number of function evaluations are bit-identical to the reference.
#include <atomic>
#include <math.h>

#include "pb_common.cuh"
#include "../../include/picasso_b200.h"

extern std::atomic<long long> g_pb_launches;

namespace {

struct ZfitArgs {
    long long n;
    const float* sx;
    const float* sy;
    const float* photons;
    const float* bg;
    const float* sx_unc;
    const float* sy_unc;
    double cx[7], cy[7];
    float cxf[7], cyf[7];
    float mag, pixelsize;
    int method;          // 0 gausslq, 1 gaussmle (sigma_uncertainty), 2 gaussmle with sx_unc / sy_unc
    float* z;
    float* d_zcalib;
    float* lpz;
    int* nfev;
};

__device__ __forceinline__ double zfit_target(double z, double sx, double sy, const double* cx,
                                              const double* cy) {
    const double z2 = z * z, z3 = z * z2, z4 = z * z3, z5 = z * z4, z6 = z * z5;
    const double wx = cx[0] * z6 + cx[1] * z5 + cx[2] * z4 + cx[3] * z3 + cx[4] * z2 + cx[5] * z + cx[6];
    const double wy = cy[0] * z6 + cy[1] * z5 + cy[2] * z4 + cy[3] * z3 + cy[4] * z2 + cy[5] * z + cy[6];
    const double dx = sqrt(sx) - sqrt(wx), dy = sqrt(sy) - sqrt(wy);
    return dx * dx + dy * dy;
}

__device__ __forceinline__ double np_sign(double v) { return v != v ? v : (double)((v > 0.0) - (v < 0.0)); }

// gausslq.sigma_uncertainty, float32, numpy evaluation order
__device__ __forceinline__ float sigma_unc_lq(float s, float so, float ph, float bg) {
    const float c12 = (float)(1.0 / 12.0);
    const float sa2 = s * s + c12;
    const float sa = sqrtf(sa2);
    const float sao = sqrtf(so * so + c12);
    const float t = ((((float)(64.0 * 3.141592653589793) * sa) * sao) * bg) / (3.0f * ph);
    const float var = ((sa2 * sa2) / ph) * ((float)(512.0 / 81.0) + t);
    return sqrtf(var / (4.0f * (s * s)));
}
// gaussmle.sigma_uncertainty
__device__ __forceinline__ float sigma_unc_mle(float s, float so, float ph, float bg) {
    const float sa2 = s * s + (float)(1.0 / 12.0);
    const float tau = (((float)(2.0 * 3.141592653589793) * sa2) * bg) / ph;
    const float var = ((s * s) / (4.0f * ph)) * ((1.0f + 8.0f * tau) + sqrtf((8.0f * tau) / (1.0f + 2.0f * tau)));
    return sqrtf(var);
}

__global__ void __launch_bounds__(128) zfit_kernel(const ZfitArgs a) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const float sxf = a.sx[i], syf = a.sy[i];
    const double sx = (double)sxf, sy = (double)syf;
    const double* cx = a.cx;
    const double* cy = a.cy;
    // ---- _minimize_scalar_bounded -------------------------------------------------------
    const double xatol = 1e-5;
    const double sqrt_eps = sqrt(2.2e-16);
    const double golden_mean = 0.5 * (3.0 - sqrt(5.0));
    double lo = -1000.0, hi = 1000.0;
    double fulc = lo + golden_mean * (hi - lo);
    double nfc = fulc, xf = fulc;
    double rat = 0.0, e = 0.0;
    double x = xf;
    double fx = zfit_target(x, sx, sy, cx, cy);
    int num = 1;
    double ffulc = fx, fnfc = fx;
    double xm = 0.5 * (lo + hi);
    double tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
    double tol2 = 2.0 * tol1;
    while (fabs(xf - xm) > (tol2 - 0.5 * (hi - lo))) {
        bool golden = true;
        if (fabs(e) > tol1) {
            golden = false;
            double r = (xf - nfc) * (fx - ffulc);
            double q = (xf - fulc) * (fx - fnfc);
            double p = (xf - fulc) * q - (xf - nfc) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) p = -p;
            q = fabs(q);
            r = e;
            e = rat;
            if (fabs(p) < fabs(0.5 * q * r) && p > q * (lo - xf) && p < q * (hi - xf)) {
                rat = (p + 0.0) / q;
                x = xf + rat;
                if ((x - lo) < tol2 || (hi - x) < tol2) {
                    const double si = np_sign(xm - xf) + (double)((xm - xf) == 0.0);
                    rat = tol1 * si;
                }
            } else {
                golden = true;
            }
        }
        if (golden) {
            e = (xf >= xm) ? lo - xf : hi - xf;
            rat = golden_mean * e;
        }
        const double si = np_sign(rat) + (double)(rat == 0.0);
        const double ar = fabs(rat);
        const double step = (ar != ar || tol1 != tol1) ? ar + tol1 : (ar > tol1 ? ar : tol1);
        x = xf + si * step;
        const double fu = zfit_target(x, sx, sy, cx, cy);
        num++;
        if (fu <= fx) {
            if (x >= xf) lo = xf; else hi = xf;
            fulc = nfc; ffulc = fnfc;
            nfc = xf; fnfc = fx;
            xf = x; fx = fu;
        } else {
            if (x < xf) lo = x; else hi = x;
            if (fu <= fnfc || nfc == xf) {
                fulc = nfc; ffulc = fnfc;
                nfc = x; fnfc = fu;
            } else if (fu <= ffulc || fulc == xf || fulc == nfc) {
                fulc = x; ffulc = fu;
            }
        }
        xm = 0.5 * (lo + hi);
        tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
        tol2 = 2.0 * tol1;
        if (num >= 500) break;
    }
    // ---- columns (float32 like the reference's pandas arithmetic) -------------------------
    const float zraw = (float)xf;                  // z[i] = result.x into a float32 array
    const float zcol = zraw * a.mag;               // locs["z"] = z * magnification_factor
    a.z[i] = zcol;
    a.d_zcalib[i] = sqrtf((float)fx);              // np.sqrt(square_d_zcalib)
    if (a.nfev) a.nfev[i] = num;
    if (a.lpz) {
        const float ph = a.photons[i], bg = a.bg[i], px = a.pixelsize;
        float se_sx, se_sy;
        if (a.method == 0) {
            se_sx = sigma_unc_lq(sxf, syf, ph, bg) * px;
            se_sy = sigma_unc_lq(syf, sxf, ph, bg) * px;
        } else if (a.method == 1) {
            se_sx = sigma_unc_mle(sxf, syf, ph, bg) * px;
            se_sy = sigma_unc_mle(syf, sxf, ph, bg) * px;
        } else {
            se_sx = a.sx_unc[i] * px;
            se_sy = a.sy_unc[i] * px;
        }
        const float zz = zcol / a.mag;
        const float z2 = zz * zz, z3 = z2 * zz, z4 = z2 * z2, z5 = z4 * zz, z6 = z3 * z3;
        const float* c = a.cxf;
        const float* d = a.cyf;
        const float wx = (c[0] * z6 + c[1] * z5 + c[2] * z4 + c[3] * z3 + c[4] * z2 + c[5] * zz + c[6]) * px;
        const float wy = (d[0] * z6 + d[1] * z5 + d[2] * z4 + d[3] * z3 + d[4] * z2 + d[5] * zz + d[6]) * px;
        const float wxp = (6.0f * c[0] * z5 + 5.0f * c[1] * z4 + 4.0f * c[2] * z3 + 3.0f * c[3] * z2 + 2.0f * c[4] * zz + c[5]) * px;
        const float wyp = (6.0f * d[0] * z5 + 5.0f * d[1] * z4 + 4.0f * d[2] * z3 + 3.0f * d[3] * z2 + 2.0f * d[4] * zz + d[5]) * px;
        const float swx = sqrtf(wx), swy = sqrtf(wy);
        const float swxp = wxp / (2.0f * swx), swyp = wyp / (2.0f * swy);
        const float dwx = (1.0f / (2.0f * sqrtf(sxf * px))) * se_sx;
        const float dwy = (1.0f / (2.0f * sqrtf(syf * px))) * se_sy;
        const float a2 = swxp * swxp, b2 = swyp * swyp, c2 = dwx * dwx, d2 = dwy * dwy;
        const float den = a2 + b2;
        a.lpz[i] = sqrtf((a2 * c2 + b2 * d2) / (den * den)) * a.mag;
    }
}

}  // namespace

extern "C" int pb_zfit_dev(size_t n, const float* d_sx, const float* d_sy, const float* d_photons,
                           const float* d_bg, const float* d_sx_unc, const float* d_sy_unc,
                           const double* cx, const double* cy, double magnification, double pixelsize,
                           int method, float* d_z, float* d_d_zcalib, float* d_lpz, int* d_nfev,
                           void* stream) {
    if (method < 0 || method > 2) { pb_set_error("pb_zfit: method must be 0 (gausslq), 1 (gaussmle) or 2 (gaussmle with sx_unc/sy_unc)"); return PB_ERR_INVALID; }
    if (n == 0) return PB_OK;
    if (!d_sx || !d_sy || !cx || !cy || !d_z || !d_d_zcalib || (d_lpz && (!d_photons || !d_bg)) ||
        (d_lpz && method == 2 && (!d_sx_unc || !d_sy_unc))) {
        pb_set_error("pb_zfit: null pointer");
        return PB_ERR_INVALID;
    }
    ZfitArgs a{};
    a.n = (long long)n; a.sx = d_sx; a.sy = d_sy; a.photons = d_photons; a.bg = d_bg;
    a.sx_unc = d_sx_unc; a.sy_unc = d_sy_unc;
    for (int k = 0; k < 7; k++) { a.cx[k] = cx[k]; a.cy[k] = cy[k]; a.cxf[k] = (float)cx[k]; a.cyf[k] = (float)cy[k]; }
    a.mag = (float)magnification; a.pixelsize = (float)pixelsize; a.method = method;
    a.z = d_z; a.d_zcalib = d_d_zcalib; a.lpz = d_lpz; a.nfev = d_nfev;
    zfit_kernel<<<(unsigned)((n + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    g_pb_launches++;
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_zfit(size_t n, const float* sx, const float* sy, const float* photons, const float* bg,
                       const float* sx_unc, const float* sy_unc, const double* cx, const double* cy,
                       double magnification, double pixelsize, int method, float* z, float* d_zcalib,
                       float* lpz, int* nfev) {
    if (method < 0 || method > 2) { pb_set_error("pb_zfit: method must be 0 (gausslq), 1 (gaussmle) or 2 (gaussmle with sx_unc/sy_unc)"); return PB_ERR_INVALID; }
    if (n == 0) return PB_OK;
    if (!sx || !sy || !cx || !cy || !z || !d_zcalib || (lpz && (!photons || !bg)) ||
        (lpz && method == 2 && (!sx_unc || !sy_unc))) {
        pb_set_error("pb_zfit: null pointer");
        return PB_ERR_INVALID;
    }
    const size_t b = n * 4;
    float* d = nullptr;           // 10 float columns: sx sy photons bg sx_unc sy_unc | z d_zcalib lpz nfev
    PB_CUDA_CHECK(cudaMalloc(&d, 10 * b));
    cudaError_t e = cudaMemcpy(d, sx, b, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + n, sy, b, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && photons) e = cudaMemcpy(d + 2 * n, photons, b, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && bg) e = cudaMemcpy(d + 3 * n, bg, b, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && sx_unc) e = cudaMemcpy(d + 4 * n, sx_unc, b, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && sy_unc) e = cudaMemcpy(d + 5 * n, sy_unc, b, cudaMemcpyHostToDevice);
    int rc = PB_OK;
    if (e == cudaSuccess)
        rc = pb_zfit_dev(n, d, d + n, photons ? d + 2 * n : nullptr, bg ? d + 3 * n : nullptr,
                         sx_unc ? d + 4 * n : nullptr, sy_unc ? d + 5 * n : nullptr, cx, cy, magnification,
                         pixelsize, method, d + 6 * n, d + 7 * n, lpz ? d + 8 * n : nullptr,
                         nfev ? reinterpret_cast<int*>(d + 9 * n) : nullptr, nullptr);
    if (rc == PB_OK && e == cudaSuccess) e = cudaMemcpy(z, d + 6 * n, b, cudaMemcpyDeviceToHost);
    if (rc == PB_OK && e == cudaSuccess) e = cudaMemcpy(d_zcalib, d + 7 * n, b, cudaMemcpyDeviceToHost);
    if (rc == PB_OK && e == cudaSuccess && lpz) e = cudaMemcpy(lpz, d + 8 * n, b, cudaMemcpyDeviceToHost);
    if (rc == PB_OK && e == cudaSuccess && nfev) e = cudaMemcpy(nfev, d + 9 * n, b, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (rc != PB_OK) return rc;
    if (e != cudaSuccess) { pb_set_error("pb_zfit: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    return PB_OK;
}
