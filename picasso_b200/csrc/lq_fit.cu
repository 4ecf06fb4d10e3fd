// picasso_b200/csrc/lq_fit.cu
//
// Least-squares 2-D Gaussian spot fit on B200 -- replaces
// picasso.gausslq.fit_spot / fit_spots / fit_spots_parallel
// (reference picasso/gausslq.py:33-343) and stands in for the vendored Gpufit
// DLL (picasso/ext/pygpufit, sm_86 Windows binary, gausslq.py:346-395).
//
// The reference's fit is scipy.optimize.leastsq(ftol=xtol=1e-2) == MINPACK
// `lmdif` driving a float32-rounded, point-sampled Gaussian model with a
// forward-difference Jacobian (step sqrt(eps_f32)*|x|).  With such coarse
// tolerances the result IS the optimiser's trajectory (about 3 LM iterations),
// so the kernel follows the published MINPACK-1 algorithm (More/Garbow/
// Hillstrom, ANL-80-74: lmdif, fdjac2, qrfac, lmpar, qrsolv, enorm) step by
// step in float64, one thread per spot:
//   * a CTA of 128 threads stages its 128 ROIs in shared memory with coalesced
//     loads (row stride box^2 words is odd -> conflict-free per-thread reads);
//   * the per-spot optimiser lives in lq_core.cuh (shared with the host-side test build);
//   * the model / residual values are rounded to float32 exactly where the
//     reference's numpy buffers are (gausslq.py:232-236).
#include <algorithm>
#include <atomic>

#include "pb_common.cuh"
#include "lq_core.cuh"
#include "../../include/picasso_b200.h"

extern std::atomic<long long> g_pb_launches;

namespace {

#ifndef PB_LQ_THREADS
#define PB_LQ_THREADS 128
#endif
constexpr int kThreads = PB_LQ_THREADS;

// VARIANT 0: register-resident normal-equations factorisation, 1: MINPACK-order Householder QR
template <int BOX, int VARIANT>
__global__ void __launch_bounds__(kThreads) lq_fit_kernel(const float* __restrict__ spots,
                                                          long long n, float* __restrict__ thetas,
                                                          int* __restrict__ infos,
                                                          int* __restrict__ nfevs) {
    constexpr int M = BOX * BOX;
    extern __shared__ float lq_smem[];          // kThreads * M floats
    const long long base = (long long)blockIdx.x * kThreads;
    const long long nblk = min((long long)kThreads, n - base);
    for (long long i = threadIdx.x; i < nblk * M; i += kThreads)
        lq_smem[i] = spots[base * M + i];
    __syncthreads();
    if (threadIdx.x >= nblk) return;
    const float* spot = lq_smem + threadIdx.x * M;
    const long long sid = base + threadIdx.x;
    double x[6];
    int info, nfev;
    if (VARIANT == 1) lq::fit_spot_qr<BOX>(spot, x, &info, &nfev);
    else lq::fit_spot_ne<BOX>(spot, x, &info, &nfev);
#pragma unroll
    for (int k = 0; k < 6; k++) thetas[sid * 6 + k] = (float)x[k];
    if (infos) infos[sid] = info;
    if (nfevs) nfevs[sid] = nfev;
}

static int g_lq_variant = 0;

template <int BOX, int VARIANT>
int launch_lq_v(const float* spots, long long n, float* thetas, int* infos, int* nfevs,
                cudaStream_t stream) {
    const int smem = kThreads * BOX * BOX * 4;
    auto kern = lq_fit_kernel<BOX, VARIANT>;
    PB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const long long grid = (n + kThreads - 1) / kThreads;
    kern<<<(unsigned)grid, kThreads, smem, stream>>>(spots, n, thetas, infos, nfevs);
    g_pb_launches++;
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

template <int BOX>
int launch_lq(const float* spots, long long n, float* thetas, int* infos, int* nfevs,
              cudaStream_t stream) {
    if (g_lq_variant == 1) return launch_lq_v<BOX, 1>(spots, n, thetas, infos, nfevs, stream);
    return launch_lq_v<BOX, 0>(spots, n, thetas, infos, nfevs, stream);
}

}  // namespace

// measurement hook: 0 = register-resident factorisation (default), 1 = MINPACK-order QR with the
// m x 6 Jacobian in local memory (A/B parity and speed comparisons, profiles/)
extern "C" int pb_lq_set_impl(int impl) {
    if (impl != 0 && impl != 1) { pb_set_error("pb_lq_set_impl: impl must be 0 or 1"); return PB_ERR_INVALID; }
    g_lq_variant = impl;
    return PB_OK;
}

extern "C" int pb_lq_fit_dev(size_t n, int box, const float* d_spots, float* d_thetas,
                             int* d_infos, int* d_nfevs, void* stream) {
    if (n == 0) return PB_OK;
    if (!d_spots || !d_thetas) { pb_set_error("pb_lq_fit_dev: null pointer"); return PB_ERR_INVALID; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long nn = (long long)n;
    switch (box) {
        case 5:  return launch_lq<5>(d_spots, nn, d_thetas, d_infos, d_nfevs, s);
        case 7:  return launch_lq<7>(d_spots, nn, d_thetas, d_infos, d_nfevs, s);
        case 9:  return launch_lq<9>(d_spots, nn, d_thetas, d_infos, d_nfevs, s);
        case 11: return launch_lq<11>(d_spots, nn, d_thetas, d_infos, d_nfevs, s);
        case 13: return launch_lq<13>(d_spots, nn, d_thetas, d_infos, d_nfevs, s);
        case 15: return launch_lq<15>(d_spots, nn, d_thetas, d_infos, d_nfevs, s);
        default:
            pb_set_error("unsupported box size %d for LQ fit (odd 5..15)", box);
            return PB_ERR_INVALID;
    }
}

// The host-buffer variant pb_lq_fit lives in api.cu (shared chunked H2D -> kernel -> D2H pipeline).
