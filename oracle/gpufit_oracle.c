/* oracle/gpufit_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of picasso's Gpufit path (reference picasso/gausslq.py:128-148 start values,
 * :346-395 fit_spots_gpufit = gf.fit(..., GAUSS_2D_ELLIPTIC, tolerance 1e-2, 20 iterations) followed
 * by amplitude * 2 pi sx sy).  The fit itself lives in a third-party binary that is absent from
 * /root/reference as source: Gpufit 1.2.0 (picasso/ext/pygpufit/Gpufit.dll, Windows only).  What is
 * restated is its published Levenberg-Marquardt algorithm (Przybylski et al., Sci. Rep. 7, 15722
 * (2017); upstream v1.2.0 sources cuda_kernels.cu / lm_fit_cuda.cu / models/gauss_2d_elliptic.cuh /
 * estimators/lse.cuh), float32 like the library's default REAL:
 *   model    f = p0 exp(-((x - p1)^2 / (2 p3^2) + (y - p2)^2 / (2 p4^2))) + p5, x / y = pixel indices
 *   LSE      chi2 = sum (f - d)^2, gradient_k = sum df/dp_k (d - f), hessian_kl = sum df/dp_k df/dp_l
 *   LM       lambda0 = 0.001; per iteration: scaling_k = max(scaling_k, H_kk), H_kk += scaling_k lambda
 *            (cuda_modify_step_widths), delta = H^-1 g by Gauss-Jordan with partial pivoting, p += delta,
 *            re-evaluate; converged when |chi2 - chi2_prev| < tol max(1, chi2) (cuda_check_for_convergence,
 *            evaluated on rejected steps too); chi2 < chi2_prev: lambda *= 0.1, keep; else lambda *= 10,
 *            p = p_prev, chi2 = chi2_prev (cuda_prepare_next_iteration); at most max_it iterations.
 * Written independently of the CUDA kernel (picasso_b200/csrc/gpufit_core.cuh states the same
 * algorithm for the device); sums run in pixel order.
 *
 * PARITY STATUS: UNPINNED against the Gpufit binary -- it cannot run here (no Windows, no GPU in the
 * build container) and the reference's only test of this path is skipped
 * (tests/test_gausslq.py:388-398).  Pinned instead by ground truth: the reference's own tolerance
 * regime for the LQ fit (tests/test_gausslq.py:38-50: centred noiseless spot within 1e-3 px) and
 * agreement with the MINPACK path within the LQ tolerance on Poisson spots (tests/). */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>

enum { GF_CONVERGED = 0, GF_MAX_ITERATION = 1, GF_SINGULAR_HESSIAN = 2 };

static void gf_evaluate(const float* spot, int box, const float* p, float* chi2, float* g, float H[6][6]) {
    float chi = 0.0f;
    for (int k = 0; k < 6; k++) { g[k] = 0.0f; for (int l = 0; l < 6; l++) H[k][l] = 0.0f; }
    for (int iy = 0; iy < box; iy++)
        for (int ix = 0; ix < box; ix++) {
            const float dx = (float)ix - p[1], dy = (float)iy - p[2];
            const float argx = dx * dx / (2 * p[3] * p[3]);
            const float argy = dy * dy / (2 * p[4] * p[4]);
            const float ex = expf(-(argx + argy));
            const float value = p[0] * ex + p[5];
            float d[6];
            d[0] = ex;
            d[1] = p[0] * ex * dx / (p[3] * p[3]);
            d[2] = p[0] * ex * dy / (p[4] * p[4]);
            d[3] = p[0] * ex * dx * dx / (p[3] * p[3] * p[3]);
            d[4] = p[0] * ex * dy * dy / (p[4] * p[4] * p[4]);
            d[5] = 1.0f;
            const float data = spot[iy * box + ix];
            const float dev = value - data;
            chi += dev * dev;
            for (int k = 0; k < 6; k++) {
                g[k] += d[k] * (data - value);
                for (int l = k; l < 6; l++) H[k][l] += d[k] * d[l];
            }
        }
    for (int k = 0; k < 6; k++) for (int l = 0; l < k; l++) H[k][l] = H[l][k];
    *chi2 = chi;
}

static int gf_gauss_jordan(float A[6][6], float* b) {
    for (int c = 0; c < 6; c++) {
        int piv = c;
        float best = fabsf(A[c][c]);
        for (int r = c + 1; r < 6; r++) if (fabsf(A[r][c]) > best) { best = fabsf(A[r][c]); piv = r; }
        if (!(best > 0.0f) || !isfinite(best)) return 0;
        if (piv != c) {
            for (int k = 0; k < 6; k++) { float t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
            float t = b[c]; b[c] = b[piv]; b[piv] = t;
        }
        const float inv = 1.0f / A[c][c];
        for (int k = 0; k < 6; k++) A[c][k] *= inv;
        b[c] *= inv;
        for (int r = 0; r < 6; r++) {
            if (r == c) continue;
            const float f = A[r][c];
            if (f == 0.0f) continue;
            for (int k = 0; k < 6; k++) A[r][k] -= f * A[c][k];
            b[r] -= f * b[c];
        }
    }
    return 1;
}

static void gf_fit_spot(const float* spot, int box, float tol, int max_it, float* out, int* state_out,
                        float* chi2_out, int* n_it_out) {
    /* start values, gausslq.py:128-148 */
    float mx = spot[0], mn = spot[0];
    for (int i = 1; i < box * box; i++) { if (spot[i] > mx) mx = spot[i]; if (spot[i] < mn) mn = spot[i]; }
    float p[6], prev_p[6], scaling[6] = {0, 0, 0, 0, 0, 0};
    const float c = (float)(box / 2.0 - 0.5);
    const float w = (float)((box / 5.0) > 1.0 ? (box / 5.0) : 1.0);
    p[0] = mx - mn; p[1] = c; p[2] = c; p[3] = w; p[4] = w; p[5] = mn;
    float chi, g[6], H[6][6];
    gf_evaluate(spot, box, p, &chi, g, H);
    float prev_chi = chi, lambda = 0.001f;
    int state = GF_CONVERGED, n_it = 0, finished = 0;
    for (int k = 0; k < 6; k++) prev_p[k] = p[k];
    for (int it = 0; !finished && it < max_it; it++) {
        float delta[6];
        for (int k = 0; k < 6; k++) {
            scaling[k] = fmaxf(scaling[k], H[k][k]);
            H[k][k] += scaling[k] * lambda;
            delta[k] = g[k];
        }
        if (!gf_gauss_jordan(H, delta)) { state = GF_SINGULAR_HESSIAN; n_it = it + 1; break; }
        for (int k = 0; k < 6; k++) { prev_p[k] = p[k]; p[k] += delta[k]; }
        gf_evaluate(spot, box, p, &chi, g, H);
        if (fabsf(chi - prev_chi) < tol * fmaxf(1.0f, chi)) finished = 1;
        else if (it == max_it - 1) { state = GF_MAX_ITERATION; finished = 1; }
        if (finished) n_it = it + 1;
        if (chi < prev_chi) { lambda *= 0.1f; prev_chi = chi; }
        else {
            lambda *= 10.0f; chi = prev_chi;
            for (int k = 0; k < 6; k++) p[k] = prev_p[k];
            if (!finished) gf_evaluate(spot, box, p, &chi, g, H);
        }
    }
    /* fit_spots_gpufit: parameters[:, 0] *= 2 pi sx sy in float32 numpy arithmetic (gausslq.py:393):
       2.0 * np.pi is a Python float (weak scalar), the right-hand side is evaluated first:
       a * ((f32(2 pi) * sx) * sy) */
    const float twopi = (float)(2.0 * 3.141592653589793);
    out[0] = p[0] * ((twopi * p[3]) * p[4]);
    out[1] = p[1]; out[2] = p[2]; out[3] = p[3]; out[4] = p[4]; out[5] = p[5];
    if (state_out) *state_out = state;
    if (chi2_out) *chi2_out = chi;
    if (n_it_out) *n_it_out = n_it;
}

typedef struct {
    const float* spots; long long first, last; int box; float tol; int max_it;
    float* params; int* states; float* chi2; int* n_it;
} gf_job;

static void* gf_worker(void* arg) {
    gf_job* j = (gf_job*)arg;
    for (long long s = j->first; s < j->last; s++)
        gf_fit_spot(j->spots + s * (long long)j->box * j->box, j->box, j->tol, j->max_it, j->params + s * 6,
                    j->states ? j->states + s : 0, j->chi2 ? j->chi2 + s : 0, j->n_it ? j->n_it + s : 0);
    return 0;
}

/* parameters (n, 6) float32 = [photons = amplitude 2 pi sx sy, x, y, sx, sy, bg] (gausslq.py:393) */
int orc_gpufit_mt(const float* spots, long long n, int box, float tolerance, int max_it, float* params,
                  int* states, float* chi2, int* n_it, int nthreads) {
    if (box < 3 || box > 21 || !(box & 1)) return 1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    gf_job jobs[256];
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (gf_job){spots, n * t / nthreads, n * (t + 1) / nthreads, box, tolerance, max_it, params,
                           states, chi2, n_it};
        pthread_create(&th[t], 0, gf_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], 0);
    return 0;
}
