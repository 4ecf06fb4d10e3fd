// picasso_b200/csrc/gpufit_core.cuh
//
// Per-spot arithmetic of picasso's Gpufit path (reference picasso/gausslq.py:128-148 start values,
// :346-395 fit_spots_gpufit -> gf.fit(data, None, ModelID.GAUSS_2D_ELLIPTIC, initial_parameters,
// tolerance=1e-2, max_number_iterations=20), then parameters[:, 0] *= 2 pi sx sy).
//
// The vendored Gpufit 1.2.0 is a Windows DLL without source (picasso/ext/pygpufit/Gpufit.dll); what
// is restated here is the PUBLISHED algorithm of that release (Przybylski et al., Sci. Rep. 7, 15722
// (2017); upstream MIT sources Gpufit/cuda_kernels.cu, lm_fit_cuda.cu, models/gauss_2d_elliptic.cuh,
// estimators/lse.cuh @ v1.2.0), all in float32 like the library's default REAL:
//   model    f = p0 exp(-((x - p1)^2 / (2 p3^2) + (y - p2)^2 / (2 p4^2))) + p5, x / y = pixel indices
//   LSE      chi2 = sum (f - d)^2, gradient_k = sum df/dp_k (d - f), hessian_kl = sum df/dp_k df/dp_l
//   LM       lambda0 = 0.001; per iteration: scaling_k = max(scaling_k, H_kk), H_kk += scaling_k lambda,
//            delta = H^-1 g by Gauss-Jordan with partial pivoting, p += delta, re-evaluate;
//            converged when |chi2 - chi2_prev| < tol max(1, chi2) (also on a failed step);
//            chi2 < chi2_prev: lambda *= 0.1, keep; else lambda *= 10, p = p_prev, chi2 = chi2_prev;
//            at most 20 iterations (state 1 = MAX_ITERATION), singular Hessian -> state 2.
// The sums run in pixel order (Gpufit reduces them with a shared-memory tree: its float32 rounding
// differs in the last bits) -- parity with the Gpufit BINARY is unpinned, parity with this
// restatement (oracle/gpufit_oracle.c compiles the same header for the host) is what the tests hold.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define GF_FN __device__ __forceinline__
#else
#define GF_FN static inline
#endif

namespace gpufit {

constexpr int kNP = 6;
enum State { CONVERGED = 0, MAX_ITERATION = 1, SINGULAR_HESSIAN = 2, NEG_CURVATURE_MLE = 3 };

// values, chi-square, gradient and Hessian (upper triangle mirrored) at p
template <int BOX, class Roi>
GF_FN void evaluate(const Roi& roi, const float p[kNP], float* chi2, float g[kNP], float H[kNP][kNP]) {
    float chi = 0.0f;
    for (int k = 0; k < kNP; k++) {
        g[k] = 0.0f;
        for (int l = 0; l < kNP; l++) H[k][l] = 0.0f;
    }
    const float two_sx2 = 2 * p[3] * p[3], two_sy2 = 2 * p[4] * p[4];
    const float sx2 = p[3] * p[3], sy2 = p[4] * p[4];
    const float sx3 = p[3] * p[3] * p[3], sy3 = p[4] * p[4] * p[4];
    for (int iy = 0; iy < BOX; iy++) {
        const float dy = (float)iy - p[2];
        const float argy = dy * dy / two_sy2;
        for (int ix = 0; ix < BOX; ix++) {
            const float dx = (float)ix - p[1];
            const float argx = dx * dx / two_sx2;
            const float ex = expf(-(argx + argy));
            const float value = p[0] * ex + p[5];
            float d[kNP];
            d[0] = ex;
            d[1] = p[0] * ex * dx / sx2;
            d[2] = p[0] * ex * dy / sy2;
            d[3] = p[0] * ex * dx * dx / sx3;
            d[4] = p[0] * ex * dy * dy / sy3;
            d[5] = 1.0f;
            const float data = roi(iy * BOX + ix);
            const float dev = value - data;
            chi += dev * dev;
            const float r = data - value;
            for (int k = 0; k < kNP; k++) {
                g[k] += d[k] * r;
                for (int l = k; l < kNP; l++) H[k][l] += d[k] * d[l];
            }
        }
    }
    for (int k = 0; k < kNP; k++)
        for (int l = 0; l < k; l++) H[k][l] = H[l][k];
    *chi2 = chi;
}

// Gauss-Jordan elimination with partial pivoting on [A | b]; false when singular
GF_FN bool gauss_jordan(float A[kNP][kNP], float b[kNP]) {
    for (int c = 0; c < kNP; c++) {
        int piv = c;
        float best = fabsf(A[c][c]);
        for (int r = c + 1; r < kNP; r++)
            if (fabsf(A[r][c]) > best) { best = fabsf(A[r][c]); piv = r; }
        if (!(best > 0.0f) || !isfinite(best)) return false;
        if (piv != c) {
            for (int k = 0; k < kNP; k++) { const float t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
            const float t = b[c]; b[c] = b[piv]; b[piv] = t;
        }
        const float inv = 1.0f / A[c][c];
        for (int k = 0; k < kNP; k++) A[c][k] *= inv;
        b[c] *= inv;
        for (int r = 0; r < kNP; r++) {
            if (r == c) continue;
            const float f = A[r][c];
            if (f == 0.0f) continue;
            for (int k = 0; k < kNP; k++) A[r][k] -= f * A[c][k];
            b[r] -= f * b[c];
        }
    }
    return true;
}

// start values of the reference (gausslq.py:128-148): [max - min, c, c, max(size/5, 1), ., min]
template <int BOX, class Roi>
GF_FN void initial_parameters(const Roi& roi, float p[kNP]) {
    float mx = roi(0), mn = roi(0);
    for (int i = 1; i < BOX * BOX; i++) {
        const float v = roi(i);
        mx = v > mx ? v : mx;
        mn = v < mn ? v : mn;
    }
    const float c = (float)(BOX / 2.0 - 0.5);
    const float w = (float)((BOX / 5.0) > 1.0 ? (BOX / 5.0) : 1.0);
    p[0] = mx - mn; p[1] = c; p[2] = c; p[3] = w; p[4] = w; p[5] = mn;
}

// The LM loop.  p: start values in, fitted [amplitude, x, y, sx, sy, offset] out.
template <int BOX, class Roi>
GF_FN void fit(const Roi& roi, float p[kNP], float tolerance, int max_iterations, int* state_out,
               float* chi2_out, int* n_iterations_out) {
    float chi, g[kNP], H[kNP][kNP];
    evaluate<BOX>(roi, p, &chi, g, H);
    float prev_chi = chi, lambda = 0.001f;
    float prev_p[kNP], scaling[kNP];
    for (int k = 0; k < kNP; k++) { prev_p[k] = p[k]; scaling[k] = 0.0f; }
    int state = CONVERGED, n_it = 0;
    bool finished = false;
    for (int it = 0; !finished && it < max_iterations; it++) {
        // cuda_modify_step_widths: adaptive (running maximum) Marquardt scaling
        for (int k = 0; k < kNP; k++) {
            scaling[k] = fmaxf(scaling[k], H[k][k]);
            H[k][k] += scaling[k] * lambda;
        }
        float delta[kNP];
        for (int k = 0; k < kNP; k++) delta[k] = g[k];
        if (!gauss_jordan(H, delta)) { state = SINGULAR_HESSIAN; n_it = it + 1; break; }
        // cuda_update_parameters
        for (int k = 0; k < kNP; k++) { prev_p[k] = p[k]; p[k] += delta[k]; }
        evaluate<BOX>(roi, p, &chi, g, H);
        // cuda_check_for_convergence (evaluated on accepted and rejected steps alike)
        const bool found = fabsf(chi - prev_chi) < tolerance * fmaxf(1.0f, chi);
        if (found) finished = true;
        else if (it == max_iterations - 1) { state = MAX_ITERATION; finished = true; }
        if (finished) n_it = it + 1;
        // cuda_prepare_next_iteration
        if (chi < prev_chi) {
            lambda *= 0.1f;
            prev_chi = chi;
        } else {
            lambda *= 10.0f;
            chi = prev_chi;
            for (int k = 0; k < kNP; k++) p[k] = prev_p[k];
            if (!finished) evaluate<BOX>(roi, p, &chi, g, H);   // derivatives at the restored parameters
        }
    }
    *state_out = state;
    *chi2_out = chi;
    *n_iterations_out = n_it;
}

}  // namespace gpufit
