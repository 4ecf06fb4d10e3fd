// picasso_b200/csrc/pb_common.cuh -- shared device/host helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define PB_OK 0
#define PB_ERR_INVALID 1   // bad argument (box, method, null pointer ...)
#define PB_ERR_CUDA 2      // CUDA runtime error, see pb_last_error()
#define PB_ERR_NOGPU 3     // no sm_100 device visible
#define PB_ERR_CAPACITY 4  // caller-provided output buffer too small
#define PB_ERR_CUFFT 5

// thread-local last-error string (Gpufit convention, ext/pygpufit/gpufit.py:361-366)
void pb_set_error(const char* fmt, ...);

#define PB_CUDA_CHECK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            pb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                      \
            return PB_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

// ---- host <-> device transfers for pageable memory (csrc/transfer.cu) ---------------------
// numpy arrays are pageable: cudaMemcpy then runs at 5-10 GB/s through the driver's staging
// copy.  These helpers stage through pinned double buffers filled / drained by several host
// threads (PB_COPY_THREADS, default min(12, cores)) so that the DMA of one chunk overlaps the host copy of
// the next; pinned memory (pb_host_alloc) is copied directly.
bool pb_host_is_pinned(const void* p);
void pb_parallel_memcpy(void* dst, const void* src, size_t bytes);
// Enqueue host -> device on `s`.  On return `h_src` is no longer referenced when it is pageable;
// pinned sources follow cudaMemcpyAsync semantics (valid until the stream reaches the copy).
int pb_h2d(void* d_dst, const void* h_src, size_t bytes, cudaStream_t s);
// Device -> host after everything enqueued on `s` so far; blocking, `h_dst` is complete on return.
int pb_d2h(void* h_dst, const void* d_src, size_t bytes, cudaStream_t s);

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t pb_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP / SYNCS) ----
__device__ __forceinline__ void pb_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pb_smem_u32(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void pb_mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void pb_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pb_smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void pb_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(pb_smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(pb_smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ bool pb_mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(pb_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void pb_mbar_wait(uint64_t* bar, unsigned parity) {
    while (!pb_mbar_try_wait(bar, parity)) {
    }
}
// generic-proxy accesses -> async-proxy (TMA) ordering on shared memory
__device__ __forceinline__ void pb_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- sub-warp group shuffles (G lanes per group, G in {8,16,32}) ----------
template <int G>
__device__ __forceinline__ double pb_gshfl(double v, int src) {
    return __shfl_sync(0xffffffffu, v, src, G);
}
template <int G>
__device__ __forceinline__ float pb_gshfl(float v, int src) {
    return __shfl_sync(0xffffffffu, v, src, G);
}
template <int G>
__device__ __forceinline__ double pb_gshfl_down1(double v) {
    return __shfl_down_sync(0xffffffffu, v, 1, G);
}
template <int G>
__device__ __forceinline__ double pb_gsum(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, G);
    return v;
}
template <int G>
__device__ __forceinline__ float pb_gmin(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o, G));
    return v;
}

#endif  // __CUDACC__
