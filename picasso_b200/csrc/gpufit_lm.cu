// picasso_b200/csrc/gpufit_lm.cu
//
// picasso's "gausslq-gpu" fit on B200: replaces the call into the vendored Gpufit 1.2.0 Windows DLL
// (reference picasso/gausslq.py:346-395 fit_spots_gpufit -> gf.fit(GAUSS_2D_ELLIPTIC, tolerance 1e-2,
// 20 iterations), picasso/ext/pygpufit/gpufit.py:40-61 gpufit_fit) with a thread-per-spot kernel that
// runs Gpufit's published float32 Levenberg-Marquardt (gpufit_core.cuh) -- its own start values
// (gausslq.py:128-148), its own trajectory, its own column layout [photons, x, y, sx, sy, bg] with
// photons = amplitude * 2 pi sx sy (gausslq.py:393).  Gpufit itself spends one CTA per fit and
// tree-reduces chi-square / gradient / Hessian in shared memory; a 7 x 7 ROI is 49 points, so here one
// THREAD owns a fit: no reductions, no synchronisation, 128 ROIs per CTA staged with coalesced loads
// (odd row stride -> conflict-free per-thread reads), everything in registers.
#include <atomic>

#include "gpufit_core.cuh"
#include "pb_common.cuh"
#include "../../include/picasso_b200.h"

extern std::atomic<long long> g_pb_launches;

namespace {

constexpr int kThreads = 128;

struct RoiSmem {
    const float* p;
    __device__ __forceinline__ float operator()(int k) const { return p[k]; }
};

template <int BOX>
__global__ void __launch_bounds__(kThreads) gpufit_lm_kernel(const float* __restrict__ spots, long long n,
                                                             float tolerance, int max_it,
                                                             float* __restrict__ params, int* __restrict__ states,
                                                             float* __restrict__ chi2, int* __restrict__ n_it) {
    constexpr int M = BOX * BOX;
    extern __shared__ float gf_smem[];
    const long long base = (long long)blockIdx.x * kThreads;
    const long long nblk = min((long long)kThreads, n - base);
    for (long long i = threadIdx.x; i < nblk * M; i += kThreads) gf_smem[i] = spots[base * M + i];
    __syncthreads();
    if (threadIdx.x >= nblk) return;
    const RoiSmem roi{gf_smem + threadIdx.x * M};
    const long long s = base + threadIdx.x;
    float p[gpufit::kNP];
    gpufit::initial_parameters<BOX>(roi, p);
    int state, it;
    float chi;
    gpufit::fit<BOX>(roi, p, tolerance, max_it, &state, &chi, &it);
    // parameters[:, 0] *= 2.0 * np.pi * parameters[:, 3] * parameters[:, 4]  (float32, gausslq.py:393)
    const float twopi = (float)(2.0 * 3.141592653589793);
    float* out = params + s * 6;
    out[0] = __fmul_rn(p[0], __fmul_rn(__fmul_rn(twopi, p[3]), p[4]));
    out[1] = p[1]; out[2] = p[2]; out[3] = p[3]; out[4] = p[4]; out[5] = p[5];
    if (states) states[s] = state;
    if (chi2) chi2[s] = chi;
    if (n_it) n_it[s] = it;
}

template <int BOX>
int launch(const float* spots, long long n, float tol, int max_it, float* params, int* states, float* chi2,
           int* n_it, cudaStream_t stream) {
    const int smem = kThreads * BOX * BOX * 4;
    auto kern = gpufit_lm_kernel<BOX>;
    PB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, smem, stream>>>(spots, n, tol, max_it, params,
                                                                               states, chi2, n_it);
    g_pb_launches++;
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

}  // namespace

extern "C" int pb_gpufit_fit_dev(size_t n, int box, const float* d_spots, float tolerance, int max_iterations,
                                 float* d_params, int* d_states, float* d_chi2, int* d_n_iterations,
                                 void* stream) {
    if (n == 0) return PB_OK;
    if (!d_spots || !d_params) { pb_set_error("pb_gpufit_fit_dev: null pointer"); return PB_ERR_INVALID; }
    if (max_iterations < 0) { pb_set_error("pb_gpufit_fit_dev: negative iteration count"); return PB_ERR_INVALID; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long nn = (long long)n;
    switch (box) {
        case 5:  return launch<5>(d_spots, nn, tolerance, max_iterations, d_params, d_states, d_chi2, d_n_iterations, s);
        case 7:  return launch<7>(d_spots, nn, tolerance, max_iterations, d_params, d_states, d_chi2, d_n_iterations, s);
        case 9:  return launch<9>(d_spots, nn, tolerance, max_iterations, d_params, d_states, d_chi2, d_n_iterations, s);
        case 11: return launch<11>(d_spots, nn, tolerance, max_iterations, d_params, d_states, d_chi2, d_n_iterations, s);
        case 13: return launch<13>(d_spots, nn, tolerance, max_iterations, d_params, d_states, d_chi2, d_n_iterations, s);
        case 15: return launch<15>(d_spots, nn, tolerance, max_iterations, d_params, d_states, d_chi2, d_n_iterations, s);
        default:
            pb_set_error("unsupported box size %d for the Gpufit-path fit (odd 5..15)", box);
            return PB_ERR_INVALID;
    }
}
// The host-buffer variant pb_gpufit_fit lives in api.cu (shared chunked H2D -> kernel -> D2H pipeline).
