// picasso_b200/csrc/mle_tps.cu
//
// Thread-per-spot MLE fit: the fast path of pb_mle_fit[_dev] for box <= 13
// (reference picasso/gaussmle.py:28-168, 533-954; arithmetic in mle_tps_core.cuh).
//
// Three kernels per batch, results handed over through the output arrays themselves:
//
//   tps_init_kernel   thread per spot; 128-spot ROI chunks staged in shared memory by one 1-D
//                     bulk async copy (TMA engine); start values -> thetas, iterations = 0
//   tps_iter_kernel   persistent, warp-autonomous.  Every LANE owns one spot and runs one
//                     Newton iteration per trip; a lane whose spot has converged writes theta /
//                     iterations and takes the next unclaimed spot immediately (block of 32
//                     indices per atomic, ROI copied into the lane's shared-memory slot by
//                     the whole warp, coalesced).  Spots need 3..100 iterations: with
//                     lane-level refill no lane ever waits for a slower neighbour, and there
//                     are no shuffles, no cross-lane reductions and no idle lanes inside an
//                     iteration.
//   tps_crlb_kernel   thread per spot, staged like init; Fisher matrix (float32 pair sums per pixel
//                     row, float64 across rows; all-float64 repeat for near-singular matrices),
//                     Cholesky inverse diagonal (Jacobi pseudo-inverse fallback), log-likelihood with
//                     a table-driven ln(); on multi-GPU runs it can also store the spot's results
//                     through an NVSwitch multicast mapping (fused all-gather, csrc/multicast.cu)
//
// The start-value and CRLB passes are separate kernels so that each runs with all 32 lanes
// busy (inside the iteration kernel they would execute for the ~4 lanes per trip that
// finish or start a spot).  The ROI is therefore read three times (DRAM traffic ~3x the
// algorithmic 252 B/spot, still < 5 % of HBM bandwidth: the fit is compute bound).
//
// Per pixel the Newton sums run in float32 by default (FFMA pipe; the reference
// accumulates them in float32 too) with float64 edge terms, row factors and parameter
// update; PB_MLE_IMPL / pb_mle_set_impl select the all-float64 pixel variant or the
// lane-group kernel of mle_fit.cu.
#include <atomic>
#include <mutex>
#include <stdlib.h>
#include <vector>

#include "mle_tps_core.cuh"
#include "pb_common.cuh"

extern std::atomic<long long> g_pb_launches;

namespace {

constexpr int kThreads = 128;            // 4 warps per CTA in all three kernels
constexpr unsigned kFull = 0xffffffffu;

#ifndef PB_CRLB_MINB
#define PB_CRLB_MINB 4
#endif
#ifndef PB_TPS_MINB
#define PB_TPS_MINB 4
#endif

struct TpsArgs {
    const float* spots;
    long long n;
    double eps;
    int max_it;
    float* thetas;
    float* crlbs;
    float* logliks;
    int* iterations;
    int* status;
    unsigned long long* counter;
    float* mc_block;     // multicast address of this rank's gather block (fused all-gather), or null
    int mc_mode;         // 1: the CRLB kernel stores all 14 words; 2: the iteration kernel stores theta +
                         //    iterations when a lane finishes its spot (spread over the whole kernel), the
                         //    CRLB / logL half is left to the caller (copy through the multicast mapping)
};

// multimem.st: a store to a multicast address, replicated by the NVSwitch into the bound memory of
// every device of the team (SASS: STG.E.STRONG.SYS on a multicast mapping)
__device__ __forceinline__ void mc_st_v2(float* p, float a, float b) {
    asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mc_st_b32(void* p, unsigned v) {
    asm volatile("multimem.st.relaxed.sys.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- accessors -----------------------------------------------------------------
struct RoiSlot {   // one spot's pixels, contiguous floats in shared memory
    const float* p;
    __host__ __device__ __forceinline__ float operator()(int k) const { return p[k]; }
};
// column factors of the x axis for the lane's spot; column stride = 32 lanes (one warp)
struct XfF32 {
    float4* a;   // (PSF, d/dmu, d2/dmu2, d/dsigma)
    float2* b;   // (d2/dsigma2, low word of PSF for the float-float residual)
    __host__ __device__ __forceinline__ void put(int c, const double f[5]) {
        const float px = (float)f[0];
        a[c * 32] = make_float4(px, (float)f[1], (float)f[2], (float)f[3]);
        b[c * 32] = make_float2((float)f[4], (float)(f[0] - (double)px));
    }
    __host__ __device__ __forceinline__ void put_f(int c, double psf, const float d[4]) {
        const float px = (float)psf;
        a[c * 32] = make_float4(px, d[0], d[1], d[2]);
        b[c * 32] = make_float2(d[3], (float)(psf - (double)px));
    }
    __host__ __device__ __forceinline__ void get(int c, float f[6]) const {
        const float4 v = a[c * 32];
        const float2 w = b[c * 32];
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
        f[4] = w.x; f[5] = w.y;
    }
};
struct XfF64 {
    double2* a;   // (PSF, d/dmu)
    double2* b;   // (d2/dmu2, d/dsigma)
    double* c;    // d2/dsigma2
    __host__ __device__ __forceinline__ void put(int col, const double f[5]) {
        a[col * 32] = make_double2(f[0], f[1]);
        b[col * 32] = make_double2(f[2], f[3]);
        c[col * 32] = f[4];
    }
    __host__ __device__ __forceinline__ void get(int col, double f[6]) const {
        const double2 u = a[col * 32], v = b[col * 32];
        f[0] = u.x; f[1] = u.y; f[2] = v.x; f[3] = v.y;
        f[4] = c[col * 32];
        f[5] = 0.0;
    }
};
// CRLB pass: (PSF, d/dmu, d/dsigma) per column; column stride = kThreads
struct Xf3 {
    double* p;
    __host__ __device__ __forceinline__ void put(int c, const double f[5]) {
        p[(c * 3 + 0) * kThreads] = f[0];
        p[(c * 3 + 1) * kThreads] = f[1];
        p[(c * 3 + 2) * kThreads] = f[3];
    }
    __host__ __device__ __forceinline__ void get(int c, double f[3]) const {
        f[0] = p[(c * 3 + 0) * kThreads];
        f[1] = p[(c * 3 + 1) * kThreads];
        f[2] = p[(c * 3 + 2) * kThreads];
    }
};

// fast CRLB pass: PSF as double, (d/dmu, d/dsigma) as one float2 per column; same footprint rule
// (column stride = kThreads) inside the memory of Xf3, which the float64 fallback re-uses
struct Xf3F {
    double* p;
    __host__ __device__ __forceinline__ void put_f(int c, double psf, const float d[4]) {
        p[(c * 2 + 0) * kThreads] = psf;
        reinterpret_cast<float2*>(p)[(c * 2 + 1) * kThreads] = make_float2(d[0], d[2]);
    }
    __host__ __device__ __forceinline__ void get(int c, double& px, float& c1, float& g1) const {
        px = p[(c * 2 + 0) * kThreads];
        const float2 v = reinterpret_cast<const float2*>(p)[(c * 2 + 1) * kThreads];
        c1 = v.x; g1 = v.y;
    }
};
// ln() table staged in shared memory as (rc, lc) pairs
struct LogTabSmem {
    const double2* t;
    __device__ __forceinline__ void get(int i, double& rc, double& lc) const {
        const double2 v = t[i];
        rc = v.x; lc = v.y;
    }
};

// erf() table staged in shared memory: (c0, c1) as one double2 and (c2..c5) as one float4 per interval
// (tps::kErfTabA / kErfTabB; the interval index diverges across lanes, which would serialise
// constant-bank reads)
constexpr int kErfN = 97;
constexpr int kErfBytes = kErfN * 32;
struct ErfTabSmem {
    const double2* a;
    const float4* b;
    __device__ __forceinline__ void get(int k, double& c0, double& c1, float& c2, float& c3, float& c4,
                                        float& c5) const {
        const double2 u = a[k];
        const float4 v = b[k];
        c0 = u.x; c1 = u.y;
        c2 = v.x; c3 = v.y; c4 = v.z; c5 = v.w;
    }
};
__device__ __forceinline__ ErfTabSmem stage_erf_table(unsigned char* dst) {
    double2* a = reinterpret_cast<double2*>(dst);
    float4* b = reinterpret_cast<float4*>(dst + kErfN * 16);
    for (int k = threadIdx.x; k < kErfN; k += blockDim.x) {
        a[k] = make_double2(tps::kErfTabA[k][0], tps::kErfTabA[k][1]);
        b[k] = make_float4(tps::kErfTabB[k][0], tps::kErfTabB[k][1], tps::kErfTabB[k][2], tps::kErfTabB[k][3]);
    }
    return ErfTabSmem{a, b};
}

template <typename T> struct XfSel;
template <> struct XfSel<float> {
    using type = XfF32;
    static constexpr int kBytesPerLaneCol = 24;
};
template <> struct XfSel<double> {
    using type = XfF64;
    static constexpr int kBytesPerLaneCol = 40;
};

// Stage the ROIs of spots [first, first + count) (count <= kThreads) into shared memory as
// they lie in global memory (spot s at s * PIX floats: stride PIX is odd, so thread-per-spot
// access is bank-conflict free).  Full, 16-byte aligned chunks go through the TMA engine.
template <int PIX>
__device__ __forceinline__ void stage_chunk(const float* spots, long long first, int count,
                                            float* sm, uint64_t* bar) {
    const float* src = spots + first * PIX;
    const bool bulk = (count == kThreads) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (bulk) {
        constexpr unsigned kBytes = kThreads * PIX * 4;
        static_assert(kBytes % 16 == 0, "bulk copy size");
        if (threadIdx.x == 0) {
            pb_mbar_init(bar, 1);
            pb_mbar_fence_init();
            pb_mbar_expect_tx(bar, kBytes);
            pb_bulk_g2s(sm, src, kBytes, bar);
        }
        __syncthreads();          // barrier initialised before anyone polls it
        pb_mbar_wait(bar, 0);
    } else {
        for (int i = threadIdx.x; i < count * PIX; i += kThreads) sm[i] = src[i];
        __syncthreads();
    }
}

// ---- start values -----------------------------------------------------------------
template <int BOX, int METHOD>
__global__ void __launch_bounds__(kThreads) tps_init_kernel(const TpsArgs a, long long seg_first,
                                                            long long seg_n) {
    constexpr int PIX = BOX * BOX;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sm = reinterpret_cast<float*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + kThreads * PIX * 4);
    const long long first = seg_first + (long long)blockIdx.x * kThreads;
    const long long rem = seg_first + seg_n - first;
    const int count = rem < kThreads ? (int)rem : kThreads;
    stage_chunk<PIX>(a.spots, first, count, sm, bar);
    if ((int)threadIdx.x < count) {
        const long long s = first + threadIdx.x;
        RoiSlot roi{sm + threadIdx.x * PIX};
        float th[6];
        const int st = tps::initial_theta<BOX, METHOD>(roi, th);
        float2* out = reinterpret_cast<float2*>(a.thetas + s * 6);
        out[0] = make_float2(th[0], th[1]);
        out[1] = make_float2(th[2], th[3]);
        out[2] = make_float2(th[4], th[5]);
        a.iterations[s] = 0;
        if (a.status) a.status[s] = st;
    }
}

// ---- Newton iterations ---------------------------------------------------------------
template <int BOX, typename T>
struct IterSmem {
    static constexpr int PIX = BOX * BOX;
    static constexpr int kRoi = 32 * PIX * 4;                                  // per warp
    static constexpr int kXf = ((32 * BOX * XfSel<T>::kBytesPerLaneCol + 15) / 16) * 16;
    static constexpr int kMs = 32 * 6 * 4;
    static constexpr int kPerWarp = ((kRoi + kXf + kMs + 127) / 128) * 128;
    static constexpr int kWarps = kPerWarp * (kThreads / 32);
    static constexpr int kTotal = kWarps + kErfBytes;                           // + the erf table
};

template <int BOX, int METHOD, typename T>
__global__ void __launch_bounds__(kThreads, PB_TPS_MINB) tps_iter_kernel(const TpsArgs a,
                                                                         long long seg_first,
                                                                         long long seg_n) {
    using SM = IterSmem<BOX, T>;
    constexpr int PIX = SM::PIX;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem_raw + warp * SM::kPerWarp;
    float* roi_w = reinterpret_cast<float*>(wbase);                 // [32 lanes][PIX]
    unsigned char* xf_raw = wbase + SM::kRoi;
    float* ms_s = reinterpret_cast<float*>(wbase + SM::kRoi + SM::kXf);   // [6][32]

    typename XfSel<T>::type xf;
    if constexpr (sizeof(T) == 4) {
        xf.a = reinterpret_cast<float4*>(xf_raw) + lane;
        xf.b = reinterpret_cast<float2*>(xf_raw + 32 * BOX * 16) + lane;
    } else {
        xf.a = reinterpret_cast<double2*>(xf_raw) + lane;
        xf.b = reinterpret_cast<double2*>(xf_raw + 32 * BOX * 16) + lane;
        xf.c = reinterpret_cast<double*>(xf_raw + 32 * BOX * 32) + lane;
    }
    const RoiSlot roi{roi_w + lane * PIX};
    const long long seg_end = seg_first + seg_n;
    const ErfTabSmem etab = stage_erf_table(smem_raw + SM::kWarps);
    __syncthreads();               // the only CTA-wide barrier: the warps are autonomous from here on

    long long idx = -1;            // spot owned by this lane (-1: none)
    // Two claimed blocks of up to 32 spot indices per warp (warp-uniform): `cur` is being handed
    // out, `nxt` was claimed one block ahead and its ROIs / start values prefetched into L2, so
    // the refill loads below do not wait for DRAM.
    long long wnext = 0, wend = 0, nnext = 0, nend = 0;
    bool exhausted = false;        // the claim counter ran past the end (uniform)
    float th[6] = {0.f, 0.f, 1.f, 1.f, 1.f, 1.f};
    int kk = 0;

    auto claim = [&](long long& b, long long& e) {
        b = e = 0;
        if (exhausted) return;
        unsigned long long v = 0;
        if (lane == 0) v = atomicAdd(a.counter, 32ull);
        v = __shfl_sync(kFull, v, 0);
        const long long first = seg_first + (long long)v;
        if (first >= seg_end) { exhausted = true; return; }
        b = first;
        e = first + 32 < seg_end ? first + 32 : seg_end;
        const char* p0 = reinterpret_cast<const char*>(a.spots + b * PIX);
        const long long nbytes = (e - b) * PIX * 4;
        for (long long off = (long long)lane * 128; off < nbytes + 127; off += 32 * 128) {
            const char* q = p0 + (off < nbytes ? off : nbytes - 1);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
        }
        const char* t0 = reinterpret_cast<const char*>(a.thetas + b * 6);
        const long long tbytes = (e - b) * 24;
        if ((long long)lane * 128 < tbytes + 127) {
            const long long off = (long long)lane * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(t0 + (off < tbytes ? off : tbytes - 1)));
        }
    };
    claim(wnext, wend);
    claim(nnext, nend);

    // (the trip bound is a watchdog only: a warp makes ~ spots/warp * iterations / 32 trips)
    for (unsigned trip = 0; trip < (1u << 24); trip++) {
        // ---- lanes without a spot take the next unclaimed ones ------------------------
        const bool idle = idx < 0;
        const unsigned need = __ballot_sync(kFull, idle);
        if (need != 0u && (wnext < wend || nnext < nend)) {
            const int cnt = __popc(need);
            const long long leftc = wend - wnext, leftn = nend - nnext;
            const int rank = __popc(need & ((1u << lane) - 1u));
            if (idle) {
                if (rank < leftc) idx = wnext + rank;
                else if (rank - leftc < leftn) idx = nnext + (rank - leftc);
            }
            if (cnt <= leftc) wnext += cnt;
            else {
                const long long used = cnt - leftc < leftn ? cnt - leftc : leftn;
                wnext = nnext + used;
                wend = nend;
                claim(nnext, nend);
            }
            const bool fresh = idle && idx >= 0;
            // start values of the new spots (per-lane loads, in flight during the ROI copy)
            float2 v0 = make_float2(0.f, 0.f), v1 = v0, v2 = v0;
            if (fresh) {
                const float2* in = reinterpret_cast<const float2*>(a.thetas + idx * 6);
                v0 = in[0]; v1 = in[1]; v2 = in[2];
            }
            // the whole warp copies each new spot's ROI into its lane's slot (coalesced),
            // four spots per round so that their loads are in flight together
            constexpr int NCH = (PIX + 31) / 32;
            unsigned got = __ballot_sync(kFull, fresh);
            while (got) {
                int t[4];
                float v[4][NCH];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    t[u] = got ? __ffs(got) - 1 : -1;
                    got &= got - 1;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (t[u] >= 0) {
                        const long long sidx = __shfl_sync(kFull, idx, t[u]);
                        const float* src = a.spots + sidx * PIX;
#pragma unroll
                        for (int c = 0; c < NCH; c++)
                            v[u][c] = (lane + 32 * c < PIX) ? src[lane + 32 * c] : 0.f;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (t[u] >= 0) {
                        float* dst = roi_w + t[u] * PIX;
#pragma unroll
                        for (int c = 0; c < NCH; c++)
                            if (lane + 32 * c < PIX) dst[lane + 32 * c] = v[u][c];
                    }
                }
            }
            if (fresh) {
                th[0] = v0.x; th[1] = v0.y; th[2] = v1.x; th[3] = v1.y; th[4] = v2.x; th[5] = v2.y;
                float ms[6];
                tps::max_steps(th, ms);
#pragma unroll
                for (int l = 0; l < 6; l++) ms_s[l * 32 + lane] = ms[l];
                kk = 0;
            }
            __syncwarp();
        }
        if (__all_sync(kFull, idx < 0)) break;

        // ---- one Newton iteration on every lane that owns a spot -------------------------
        if (idx >= 0) {
            tps::column_stage<BOX, METHOD, T>(th, xf, etab);
            T num[6], den[6];   // float32 pixel sums also run the row stage in float32
            tps::newton_sums<BOX, METHOD, T, T>(roi, th, xf, num, den, etab);
            float ms[6];
#pragma unroll
            for (int l = 0; l < 6; l++) ms[l] = ms_s[l * 32 + lane];
            const bool conv = tps::update_theta<BOX, METHOD, T>(th, ms, num, den, a.eps);
            kk++;
            if (conv || kk >= a.max_it) {
                float2* out = reinterpret_cast<float2*>(a.thetas + idx * 6);
                out[0] = make_float2(th[0], th[1]);
                out[1] = make_float2(th[2], th[3]);
                out[2] = make_float2(th[4], th[5]);
                a.iterations[idx] = kk;
                if (a.mc_block && a.mc_mode == 2) {
                    float* mt = a.mc_block + idx * 6;
                    mc_st_v2(mt, th[0], th[1]); mc_st_v2(mt + 2, th[2], th[3]); mc_st_v2(mt + 4, th[4], th[5]);
                    mc_st_b32(a.mc_block + 13 * a.n + idx, (unsigned)kk);
                }
                idx = -1;
            }
        }
        __syncwarp();   // slots of finished lanes may be refilled by the warp next trip
    }
}

// ---- CRLB + log-likelihood ---------------------------------------------------------------
template <int BOX>
struct CrlbSmem {
    static constexpr int PIX = BOX * BOX;
    static constexpr int kRoi = ((kThreads * PIX * 4 + 15) / 16) * 16;
    static constexpr int kXf = kThreads * BOX * 3 * 8;
    static constexpr int kTab = 128 * 16;
    static constexpr int kTotal = kRoi + kXf + kTab + 16 + kErfBytes;
};

// FAST: float32 pair sums + table ln() with the all-float64 pass as per-spot fallback
// (mle_tps_core.cuh, crlb_loglik_fast); !FAST: the all-float64 pass (pb_mle_set_impl(1)).
template <int BOX, int METHOD, bool FAST>
__global__ void __launch_bounds__(kThreads, PB_CRLB_MINB) tps_crlb_kernel(const TpsArgs a, long long seg_first,
                                                            long long seg_n) {
    using SM = CrlbSmem<BOX>;
    constexpr int PIX = BOX * BOX;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sm = reinterpret_cast<float*>(smem_raw);
    double* xfp = reinterpret_cast<double*>(smem_raw + SM::kRoi);
    double2* tabp = reinterpret_cast<double2*>(smem_raw + SM::kRoi + SM::kXf);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + SM::kRoi + SM::kXf + SM::kTab);
    const long long first = seg_first + (long long)blockIdx.x * kThreads;
    const long long rem = seg_first + seg_n - first;
    const int count = rem < kThreads ? (int)rem : kThreads;
    ErfTabSmem etab{nullptr, nullptr};
    if (FAST) {
        static_assert(kThreads == 128, "one table entry per thread");
        tabp[threadIdx.x] = make_double2(tps::kLogRc[threadIdx.x], tps::kLogLc[threadIdx.x]);
        etab = stage_erf_table(smem_raw + SM::kRoi + SM::kXf + SM::kTab + 16);
    }
    stage_chunk<PIX>(a.spots, first, count, sm, bar);      // (its __syncthreads publishes the table)
    if ((int)threadIdx.x < count) {
        const long long s = first + threadIdx.x;
        RoiSlot roi{sm + threadIdx.x * PIX};
        Xf3 xf{xfp + threadIdx.x};
        const float2* in = reinterpret_cast<const float2*>(a.thetas + s * 6);
        const float2 v0 = in[0], v1 = in[1], v2 = in[2];
        const float th[6] = {v0.x, v0.y, v1.x, v1.y, v2.x, v2.y};
        float cr[6], ll;
        int st = -1;
        if (FAST) {
            Xf3F xff{xfp + threadIdx.x};
            const LogTabSmem tab{tabp};
            st = tps::crlb_loglik_fast<BOX, METHOD>(roi, th, xff, tab, etab, cr, &ll);
        }
        if (st < 0) st = tps::crlb_loglik<BOX, METHOD>(roi, th, xf, cr, &ll);
        float2* out = reinterpret_cast<float2*>(a.crlbs + s * 6);
        out[0] = make_float2(cr[0], cr[1]);
        out[1] = make_float2(cr[2], cr[3]);
        out[2] = make_float2(cr[4], cr[5]);
        a.logliks[s] = ll;
        if (a.status && st) a.status[s] |= st;
        if (a.mc_block && a.mc_mode == 1) {
            // fused all-gather: this spot's 14 output words go through the multicast mapping into the
            // gather buffers of ALL ranks, block layout [thetas 6n | crlbs 6n | logliks n | iterations n]
            float* mt = a.mc_block + s * 6;
            mc_st_v2(mt, th[0], th[1]); mc_st_v2(mt + 2, th[2], th[3]); mc_st_v2(mt + 4, th[4], th[5]);
            float* mc = a.mc_block + 6 * a.n + s * 6;
            mc_st_v2(mc, cr[0], cr[1]); mc_st_v2(mc + 2, cr[2], cr[3]); mc_st_v2(mc + 4, cr[4], cr[5]);
            mc_st_b32(a.mc_block + 12 * a.n + s, __float_as_uint(ll));
            mc_st_b32(a.mc_block + 13 * a.n + s, (unsigned)a.iterations[s]);
        }
    }
}

// ---- host side -----------------------------------------------------------------------
// A ring of zeroed claim counters per device (launches on different streams may be in flight
// together); each iteration-kernel launch takes the next slot and clears it on its own stream.
constexpr int kCounterSlots = 256;
int next_counter(int dev, cudaStream_t stream, unsigned long long** out) {
    static std::mutex mu;
    static std::vector<unsigned long long*> rings;
    static std::vector<unsigned> cursor;
    std::lock_guard<std::mutex> lk(mu);
    if ((int)rings.size() <= dev) { rings.resize(dev + 1, nullptr); cursor.resize(dev + 1, 0); }
    if (!rings[dev])
        PB_CUDA_CHECK(cudaMalloc(&rings[dev], kCounterSlots * sizeof(unsigned long long)));
    unsigned long long* slot = rings[dev] + (cursor[dev]++ % kCounterSlots);
    PB_CUDA_CHECK(cudaMemsetAsync(slot, 0, sizeof(unsigned long long), stream));
    *out = slot;
    return PB_OK;
}

// The batch can be walked in segments (PB_MLE_SEGMENT_SPOTS) so that the ROIs of a segment are
// still in L2 for the second and third kernel.  Off by default: the iteration kernel is
// persistent with lane-level refill, and its drain phase (lanes idle once the claim counter
// runs out) costs more on short segments than the L2 hits save -- the fit is compute bound
// (measured: 48 MB segments 163 M fits/s, see profiles/).
long long segment_spots(int pix) {
    static long long env = -1;
    if (env < 0) {
        env = 0;
        if (const char* e = getenv("PB_MLE_SEGMENT_SPOTS")) env = atoll(e);
    }
    (void)pix;
    return env > 0 ? env : (1ll << 40);
}

// Optional per-kernel timing (pb_mle_profile / pb_mle_profile_read): CUDA events recorded on
// the launch stream around the three kernels of the most recent call.
struct TpsProfile {
    std::atomic<bool> on{false};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool recorded = false;
};
TpsProfile g_prof;

// Optional phase event (pb_mle_set_phase_event): recorded on the launch stream after the iteration kernel
// of the next thread-per-spot call(s) -- thetas and iterations are final from there on, so a multi-GPU
// caller can start moving that half of the results while the CRLB kernel still runs.
std::atomic<cudaEvent_t> g_phase_event{nullptr};

template <int BOX, int METHOD, typename T>
int launch_tps(const TpsArgs& a0, cudaStream_t stream) {
    constexpr int PIX = BOX * BOX;
    using ISM = IterSmem<BOX, T>;
    using CSM = CrlbSmem<BOX>;
    auto k_init = tps_init_kernel<BOX, METHOD>;
    auto k_iter = tps_iter_kernel<BOX, METHOD, T>;
    auto k_crlb = tps_crlb_kernel<BOX, METHOD, sizeof(T) == 4>;
    constexpr int init_smem = kThreads * PIX * 4 + 16;
    int dev = 0;
    PB_CUDA_CHECK(cudaGetDevice(&dev));
    // per-device launch configuration of this instantiation, set up once
    struct Cfg { int num_sms = 0, per_sm = 0; };
    static std::mutex cfg_mu;
    static std::vector<Cfg> cfgs;
    Cfg cfg;
    {
        std::lock_guard<std::mutex> lk(cfg_mu);
        if ((int)cfgs.size() <= dev) cfgs.resize(dev + 1);
        if (cfgs[dev].per_sm == 0) {
            Cfg c;
            PB_CUDA_CHECK(cudaDeviceGetAttribute(&c.num_sms, cudaDevAttrMultiProcessorCount, dev));
            PB_CUDA_CHECK(cudaFuncSetAttribute(k_init, cudaFuncAttributeMaxDynamicSharedMemorySize, init_smem));
            PB_CUDA_CHECK(cudaFuncSetAttribute(k_iter, cudaFuncAttributeMaxDynamicSharedMemorySize, ISM::kTotal));
            PB_CUDA_CHECK(cudaFuncSetAttribute(k_crlb, cudaFuncAttributeMaxDynamicSharedMemorySize, CSM::kTotal));
            PB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.per_sm, k_iter, kThreads, ISM::kTotal));
            if (c.per_sm < 1) {
                pb_set_error("mle iteration kernel does not fit on an SM (box=%d)", BOX);
                return PB_ERR_CUDA;
            }
            cfgs[dev] = c;
        }
        cfg = cfgs[dev];
    }
    const int num_sms = cfg.num_sms, per_sm = cfg.per_sm;
    const bool prof = g_prof.on.load() && g_prof.ev[0] != nullptr;
    const long long seg = segment_spots(PIX);
    for (long long first = 0; first < a0.n; first += seg) {
        const long long m = a0.n - first < seg ? a0.n - first : seg;
        TpsArgs a = a0;
        const int chunks = (int)((m + kThreads - 1) / kThreads);
        if (a.max_it > 0) {
            int rc = next_counter(dev, stream, &a.counter);
            if (rc != PB_OK) return rc;
        }
        const bool p = prof && first == 0;
        if (p) cudaEventRecord(g_prof.ev[0], stream);
        k_init<<<chunks, kThreads, init_smem, stream>>>(a, first, m);
        g_pb_launches++;
        if (p) cudaEventRecord(g_prof.ev[1], stream);
        if (a.max_it > 0) {
            long long cap = (long long)num_sms * per_sm;
            int grid = (int)(chunks < cap ? chunks : cap);
            k_iter<<<grid, kThreads, ISM::kTotal, stream>>>(a, first, m);
            g_pb_launches++;
        }
        if (p) cudaEventRecord(g_prof.ev[2], stream);
        if (cudaEvent_t pe = g_phase_event.load()) cudaEventRecord(pe, stream);
        k_crlb<<<chunks, kThreads, CSM::kTotal, stream>>>(a, first, m);
        g_pb_launches++;
        if (p) { cudaEventRecord(g_prof.ev[3], stream); g_prof.recorded = true; }
    }
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

template <int METHOD, typename T>
int dispatch_tps(int box, const TpsArgs& a, cudaStream_t stream) {
    switch (box) {
        case 5:  return launch_tps<5, METHOD, T>(a, stream);
        case 7:  return launch_tps<7, METHOD, T>(a, stream);
        case 9:  return launch_tps<9, METHOD, T>(a, stream);
        case 11: return launch_tps<11, METHOD, T>(a, stream);
        case 13: return launch_tps<13, METHOD, T>(a, stream);
        default:
            pb_set_error("thread-per-spot MLE path: unsupported box %d", box);
            return PB_ERR_INVALID;
    }
}

}  // namespace

// Called by pb_mle_fit_dev (mle_fit.cu).  pixel_f32: 1 = float32 per-pixel sums, 0 = float64.
bool pb_mle_tps_supports(int box) { return box >= 5 && box <= 13 && (box & 1); }

int pb_mle_tps_fit(size_t n, int box, const float* d_spots, double eps, int max_it, int method,
                   float* d_thetas, float* d_crlbs, float* d_logliks, int* d_iterations,
                   int* d_status, cudaStream_t stream, int pixel_f32, float* mc_block, int mc_mode) {
    TpsArgs a{d_spots, (long long)n, eps, max_it, d_thetas, d_crlbs, d_logliks, d_iterations,
              d_status, nullptr, mc_block, mc_mode};
    if (method == 1)
        return pixel_f32 ? dispatch_tps<1, float>(box, a, stream) : dispatch_tps<1, double>(box, a, stream);
    return pixel_f32 ? dispatch_tps<0, float>(box, a, stream) : dispatch_tps<0, double>(box, a, stream);
}

// Measurement hooks (include/picasso_b200.h): time the three kernels of the next
// thread-per-spot calls with CUDA events on their launch stream.
extern "C" int pb_mle_profile(int enable) {
    if (enable && g_prof.ev[0] == nullptr)
        for (int i = 0; i < 4; i++) PB_CUDA_CHECK(cudaEventCreate(&g_prof.ev[i]));
    g_prof.on.store(enable != 0);
    return PB_OK;
}
extern "C" int pb_mle_set_phase_event(void* cuda_event) {
    g_phase_event.store(reinterpret_cast<cudaEvent_t>(cuda_event));
    return PB_OK;
}
extern "C" int pb_mle_profile_read(float* ms3) {
    if (!ms3 || !g_prof.recorded) {
        pb_set_error("pb_mle_profile_read: nothing recorded (enable pb_mle_profile and run a fit)");
        return PB_ERR_INVALID;
    }
    PB_CUDA_CHECK(cudaEventSynchronize(g_prof.ev[3]));
    for (int i = 0; i < 3; i++) PB_CUDA_CHECK(cudaEventElapsedTime(&ms3[i], g_prof.ev[i], g_prof.ev[i + 1]));
    return PB_OK;
}
