// picasso_b200/csrc/mle_fit.cu
//
// Per-spot 2-D Gaussian maximum-likelihood fit (Smith et al. 2010, integrated
// pixel model, Poisson likelihood, per-parameter Newton steps, Fisher/CRLB) --
// the B200 replacement for picasso.gaussmle._mlefit_sigmaxy / _mlefit_sigma
// (reference picasso/gaussmle.py:533-954).
//
// Mapping (see DESIGN.md "MLE kernel"):
//   * a cooperative group of G lanes owns one spot (G = 8 for box <= 7, 16 for
//     box <= 15, 32 above) -> 32/G spots per warp in flight; all control flow
//     is warp-uniform (converged groups are predicated off).
//   * the model is separable: lane g evaluates the erf/exp terms of pixel
//     EDGE g on both axes (box+1 edges per axis) -- 2 erf + 2 exp per lane per
//     Newton iteration instead of the reference's 4 erf + 16 exp per PIXEL.
//   * lane j then walks pixel row j (box pixels), accumulating the 2*n_par
//     Newton sums in f64; sums are combined through shared memory and lane l
//     applies the clamped Newton update for parameter l in float32 exactly as
//     the reference does (theta is float32 state between iterations).
//   * ROIs are staged into shared memory 4 spots at a time per warp with 1-D
//     bulk async copies (TMA engine, mbarrier completion), double buffered.
//   * CRLB: Fisher matrix in f64, diagonal of the inverse via a scaled
//     Cholesky; singular / ill-conditioned matrices fall back to a Jacobi
//     eigen pseudo-inverse with numpy's rcond=1e-15 (np.linalg.pinv semantics).
//
// No tensor cores: 2*n_par sums over box^2 pixels is not a dense contraction.
#include <atomic>
#include <mutex>
#include <stdlib.h>
#include <vector>

#include "pb_common.cuh"

extern std::atomic<long long> g_pb_launches;

namespace {

constexpr double kInvSqrt2Pi = 0.3989422804014326779;   // 1/sqrt(2*pi)
constexpr double kInvSqrt2 = 0.70710678118654757;        // gaussmle.py:276
constexpr double kInvSqrtPi = 0.5641895835477562869;     // 1/sqrt(pi)
constexpr int kTileSpots = 4;      // spots per TMA tile (per warp): 16*box^2 bytes
constexpr int kWarpsPerBlock = 4;
constexpr int kRedStride = 23;     // doubles per lane in the reduction scratch (22 + pad)

// tuning knobs (tools/tune_mle.py builds variants with -D overrides)
#ifndef PB_MLE_MINB
#define PB_MLE_MINB 4          // min resident CTAs per SM requested from ptxas
#endif
#ifndef PB_MLE_PIX_UNROLL
#define PB_MLE_PIX_UNROLL 8   // unroll factor of the per-row pixel loops
#endif
#ifndef PB_MLE_LIBM_ERF
#define PB_MLE_LIBM_ERF 0      // 1: CUDA libdevice erf(); 0: erf_from_gauss (shares the exp)
#endif
#ifndef PB_MLE_F32_PIXELS
#define PB_MLE_F32_PIXELS 0    // 1: float32 per-pixel Newton sums (experiment, not the default:
#endif                         //    trades ~0.1 % iteration-count parity for FP32-pipe throughput)
#define PB_STR2(x) #x
#define PB_STR(x) PB_STR2(x)
#define PB_PIX_UNROLL _Pragma(PB_STR(unroll PB_MLE_PIX_UNROLL))

struct MleArgs {
    const float* spots;   // (n, box, box) f32, device
    long long n;
    double eps;
    int max_it;
    float* thetas;        // (n, 6)
    float* crlbs;         // (n, 6)
    float* logliks;       // (n,)
    int* iterations;      // (n,)
    int* status;          // (n,) nullable
    unsigned long long* tile_counter;   // dynamic tile scheduler (zeroed before launch)
};

constexpr int kNFx = 14;          // per-column x-factors kept in shared memory (12 used)

template <int BOX, int G>
struct MleSmem {
    static constexpr int S = 32 / G;
    static constexpr int PIX = BOX * BOX;
    static constexpr int kRoiBytes = 2 * kTileSpots * PIX * 4;          // 2 TMA stages (f32)
    static constexpr int kDataBytes = ((S * PIX * 8 + 15) / 16) * 16;    // current spots as f64
    static constexpr int kFxBytes = ((S * kNFx * BOX * 8 + 15) / 16) * 16;
    static constexpr int kFxfBytes = PB_MLE_F32_PIXELS ? S * 12 * BOX * 4 : 0;
    static constexpr int kRedBytes = 32 * kRedStride * 8;
    static constexpr int kSumBytes = S * 24 * 8;
    static constexpr int kBarBytes = 16;
    static constexpr int kPerWarp =
        ((kRoiBytes + kDataBytes + kFxBytes + kFxfBytes + kRedBytes + kSumBytes + kBarBytes + 127) / 128) * 128;
    static constexpr int kTotal = kPerWarp * kWarpsPerBlock;
};

// 1/x for x in the normal range: MUFU seed (rcp.approx.ftz.f64) refined by one cubic
// and one quadratic Newton step (seed error e -> e^6), i.e. full double precision
// without the slow-path branch that '/' carries for denormals and infinities.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    r = fma(r, fma(e, e, e), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

// ---- Jacobi eigen pseudo-inverse diagonal (rare fallback; np.linalg.pinv) --
template <int NP>
__device__ __noinline__ void pinv_diag_jacobi(const double* Msym /*NP*NP*/, double* diag) {
    double A[NP * NP], V[NP * NP];
    bool finite = true;
    for (int i = 0; i < NP * NP; i++) {
        A[i] = Msym[i];
        finite = finite && isfinite(A[i]);
    }
    if (!finite) {
        for (int i = 0; i < NP; i++) diag[i] = nan("");
        return;
    }
    for (int i = 0; i < NP; i++)
        for (int j = 0; j < NP; j++) V[i * NP + j] = (i == j) ? 1.0 : 0.0;
    double prev_off = INFINITY;      // stop when a sweep no longer reduces the off-diagonal weight (mle_tps_core.cuh)
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0, dsum = 0.0;
        for (int i = 0; i < NP; i++)
            for (int j = 0; j < NP; j++) {
                double v = A[i * NP + j] * A[i * NP + j];
                if (i != j) off += v; else dsum += v;
            }
        if (off <= 1e-60 || off <= 1e-34 * dsum || off >= 0.25 * prev_off) break;
        prev_off = off;
        for (int p = 0; p < NP - 1; p++)
            for (int q = p + 1; q < NP; q++) {
                double apq = A[p * NP + q];
                if (apq == 0.0) continue;
                double th = (A[q * NP + q] - A[p * NP + p]) / (2.0 * apq);
                double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
                if (!isfinite(th)) t = 0.0;
                double c = rsqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < NP; k++) {
                    double akp = A[k * NP + p], akq = A[k * NP + q];
                    A[k * NP + p] = c * akp - s * akq;
                    A[k * NP + q] = s * akp + c * akq;
                }
                for (int k = 0; k < NP; k++) {
                    double apk = A[p * NP + k], aqk = A[q * NP + k];
                    A[p * NP + k] = c * apk - s * aqk;
                    A[q * NP + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < NP; k++) {
                    double vkp = V[k * NP + p], vkq = V[k * NP + q];
                    V[k * NP + p] = c * vkp - s * vkq;
                    V[k * NP + q] = s * vkp + c * vkq;
                }
            }
    }
    double smax = 0.0;
    for (int i = 0; i < NP; i++) smax = fmax(smax, fabs(A[i * NP + i]));
    double cutoff = 1e-15 * smax;
    for (int i = 0; i < NP; i++) {
        double acc = 0.0;
        for (int k = 0; k < NP; k++) {
            double lam = A[k * NP + k];
            if (fabs(lam) > cutoff) acc += V[i * NP + k] * V[i * NP + k] / lam;
        }
        diag[i] = acc;
    }
}

// Diagonal of inv(M) for a symmetric positive-definite NP x NP matrix given as
// its packed upper triangle m[idx(k,l)], k<=l (row-major packed).  Returns
// false when the (diagonally scaled) Cholesky meets a non-positive / tiny
// pivot -> caller falls back to the pseudo-inverse.
template <int NP>
__device__ __forceinline__ bool inv_diag_cholesky(const double* m, double* diag) {
    auto idx = [](int k, int l) { return k * NP - (k * (k - 1)) / 2 + (l - k); };
    double d[NP];
    bool ok = true;
#pragma unroll
    for (int i = 0; i < NP; i++) {
        double mii = m[idx(i, i)];
        ok = ok && (mii > 0.0) && isfinite(mii);
        d[i] = rsqrt(mii);
    }
    if (!ok) return false;
    // scaled matrix C = D M D has unit diagonal; C = L L^T
    double L[NP][NP];
    double rl[NP];   // 1 / L[j][j]
#pragma unroll
    for (int j = 0; j < NP; j++) {
        double s = 1.0;
#pragma unroll
        for (int k = 0; k < j; k++) s -= L[j][k] * L[j][k];
        ok = ok && (s > 1e-12);
        rl[j] = rsqrt(s);
        L[j][j] = s * rl[j];
#pragma unroll
        for (int i = j + 1; i < NP; i++) {
            double c = m[idx(j, i)] * d[i] * d[j];
#pragma unroll
            for (int k = 0; k < j; k++) c -= L[i][k] * L[j][k];
            L[i][j] = c * rl[j];
        }
    }
    if (!ok) return false;
    // X = inv(L) column by column; diag(inv(C))_j = sum_i X_ij^2
#pragma unroll
    for (int j = 0; j < NP; j++) {
        double X[NP];
        double acc = 0.0;
#pragma unroll
        for (int i = j; i < NP; i++) {
            double s = (i == j) ? 1.0 : 0.0;
#pragma unroll
            for (int k = j; k < i; k++) s -= L[i][k] * X[k];
            X[i] = s * rl[i];
            acc += X[i] * X[i];
        }
        diag[j] = acc * d[j] * d[j];
    }
    return true;
}

// erf(z) from the Gaussian term A = exp(-z^2) the kernel computes anyway:
//   erf(z) = sign(z) * (1 - A * P(x)),  t = 1/(1 + |z|/2),  x = (8 t - 5) / 3,
// P = degree-16 fit of erfcx on z in [0, 6] (tools/gen_erf_coeffs.py: max abs error
// 2.2e-15 in float64 Horner form; beyond |z| = 6, A < 3e-16 so the term vanishes).
// One branch-free polynomial instead of libdevice's two-branch erf (both branches execute
// when the 8 edges of a spot straddle |z| ~ 1).
__constant__ double kErfcxPoly[17] = {
    0.3785374169292369,      0.4221875836134109,     0.16940759095488894,
    0.03299342947957943,     -0.0016702444148986467, -0.001665364388974453,
    0.00013358437327083263,  0.00010054202594586505, -2.2700690081718415e-05,
    -4.271484393949245e-06,  2.839667893118267e-06,  -2.8180849640529505e-07,
    -2.0096403079505821e-07, 8.619074557878024e-08,  -3.971810944694673e-09,
    -7.52808888456569e-09,   2.015803325273065e-09,
};

__device__ __forceinline__ double fast_rcp(double x);

__device__ __forceinline__ double erf_from_gauss(double z, double A) {
    double a = fabs(z);
    a = a > 6.0 ? 6.0 : a;
    const double t = fast_rcp(fma(0.5, a, 1.0));
    const double x = fma(t, 2.6666666666666665, -1.6666666666666667);
    double p = kErfcxPoly[16];
#pragma unroll
    for (int k = 15; k >= 0; k--) p = fma(p, x, kErfcxPoly[k]);
    return copysign(fma(-A, p, 1.0), z);
}

template <int BOX, int G, int METHOD>   // METHOD 1 = sigmaxy (6 par), 0 = sigma (5 par)
__global__ void __launch_bounds__(kWarpsPerBlock * 32, PB_MLE_MINB)
mle_fit_kernel(const MleArgs a) {
    using SM = MleSmem<BOX, G>;
    constexpr int S = SM::S;
    constexpr int PIX = SM::PIX;
    constexpr int NP = METHOD == 1 ? 6 : 5;
    constexpr int NSUB = kTileSpots / S;
    constexpr int NFISH = NP * (NP + 1) / 2;
    static_assert(BOX + 1 <= G, "group too small for box");
    // shared-memory x-factor slots (per column i)
    enum { F_NPX = 0, F_PX, F_C1, F_C2, F_G1, F_G2, F_PX2, F_C1SQ, F_G1SQ, F_G1PX, F_C1PX, F_C1G1 };

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int g = lane % G;         // lane within the spot group
    const int grp = lane / G;       // group within the warp
    unsigned char* wbase = smem_raw + warp * SM::kPerWarp;
    float* roi = reinterpret_cast<float*>(wbase);                                   // [2][4*PIX]
    double* dspot = reinterpret_cast<double*>(wbase + SM::kRoiBytes) + grp * PIX;    // [PIX] f64
    double* fx = reinterpret_cast<double*>(wbase + SM::kRoiBytes + SM::kDataBytes) +
                 grp * kNFx * BOX;                                                  // [BOX][14]
    constexpr int kOffF = SM::kRoiBytes + SM::kDataBytes + SM::kFxBytes;
    float* fxf = reinterpret_cast<float*>(wbase + kOffF) + grp * 12 * BOX;          // [BOX][12]
    constexpr int kOffR = kOffF + SM::kFxfBytes;
    double* red = reinterpret_cast<double*>(wbase + kOffR);                         // [32][23]
    double* sums = reinterpret_cast<double*>(wbase + kOffR + SM::kRedBytes) + grp * 24;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wbase + kOffR + SM::kRedBytes + SM::kSumBytes);
    (void)fxf;

    const long long n = a.n;
    const long long ntiles = (n + kTileSpots - 1) / kTileSpots;
    // Dynamic tile scheduler: every warp claims its next 4-spot tile with one atomic (one
    // per ~100 us of work).  CTAs that become resident late -- e.g. because a concurrent NCCL
    // all-gather of the previous step's outputs holds a few SMs -- simply claim fewer tiles,
    // so collectives overlap the fit without stretching its tail.
    auto claim = [&]() -> long long {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(a.tile_counter, 1ull);
        return (long long)__shfl_sync(0xffffffffu, t, 0);
    };
    long long tile = claim();
    long long next_tile = (tile < ntiles) ? claim() : ntiles;
    const bool tma_ok = ((reinterpret_cast<uintptr_t>(a.spots) & 15) == 0);
    constexpr unsigned kTileBytes = kTileSpots * PIX * 4;

    if (lane == 0) {
        pb_mbar_init(&bars[0], 1);
        pb_mbar_init(&bars[1], 1);
        pb_mbar_fence_init();
    }
    __syncwarp();

    auto issue = [&](long long t, int stage) {
        // full tiles via the TMA engine; the ragged tail tile by plain loads
        const long long first = t * kTileSpots;
        const bool full = tma_ok && (first + kTileSpots <= n);
        float* dst = roi + stage * kTileSpots * PIX;
        if (full) {
            if (lane == 0) {
                pb_fence_proxy_async();
                pb_mbar_expect_tx(&bars[stage], kTileBytes);
                pb_bulk_g2s(dst, a.spots + first * PIX, kTileBytes, &bars[stage]);
            }
        } else {
            const long long avail = (n - first < kTileSpots ? n - first : kTileSpots) * PIX;
            for (int i = lane; i < kTileSpots * PIX; i += 32)
                dst[i] = i < avail ? a.spots[first * PIX + i] : 0.0f;
        }
        return full;
    };

    unsigned phase0 = 0, phase1 = 0;
    bool cur_tma = false;
    if (tile < ntiles) cur_tma = issue(tile, 0);
    int stage = 0;

    for (; tile < ntiles; stage ^= 1) {
        // prefetch the next tile into the other stage (it was fully consumed
        // before the __syncwarp at the end of the previous trip)
        bool next_tma = false;
        if (next_tile < ntiles) next_tma = issue(next_tile, stage ^ 1);
        if (cur_tma) {
            if (stage == 0) { pb_mbar_wait(&bars[0], phase0); phase0 ^= 1; }
            else            { pb_mbar_wait(&bars[1], phase1); phase1 ^= 1; }
        }
        __syncwarp();
        const float* tile_roi = roi + stage * kTileSpots * PIX;

#pragma unroll 1
        for (int sub = 0; sub < NSUB; sub++) {
            const int local = sub * S + grp;
            const long long spot_idx = tile * kTileSpots + local;
            const bool valid = spot_idx < n;
            const float* sp = tile_roi + local * PIX;
            int st_flags = 0;

            // the ROI as float64 (converted once; every iteration re-reads it)
            __syncwarp();
            for (int q = g; q < PIX; q += G) dspot[q] = (double)sp[q];

            // ---------------- initial parameters (gaussmle.py:28-168) -----
            float th[6];
            {
                double rs_ = 0.0, rxs = 0.0;
                if (g < BOX) {
#pragma unroll 1
                    for (int i = 0; i < BOX; i++) {
                        double v = (double)sp[g * BOX + i];
                        rs_ += v;
                        rxs += v * (double)i;
                    }
                }
                double sum = pb_gsum<G>(rs_);
                double ysum = pb_gsum<G>(rs_ * (double)g);
                double xsum = pb_gsum<G>(rxs);
                double xc, yc;
                if (sum <= 0.0) { sum = 0.01; yc = (BOX - 1) / 2.0; xc = (BOX - 1) / 2.0; }
                else { yc = ysum / sum; xc = xsum / sum; }
                // 3x3 edge-truncated mean filter, minimum (gaussmle.py:61-91,135)
                float fmin_ = INFINITY;
                if (g < BOX) {
                    const int k = g;
                    const int min_m = k - 1 < 0 ? 0 : k - 1, max_m = k + 2 > BOX ? BOX : k + 2;
#pragma unroll 1
                    for (int l = 0; l < BOX; l++) {
                        const int min_n = l - 1 < 0 ? 0 : l - 1, max_n = l + 2 > BOX ? BOX : l + 2;
                        double ns = 0.0;
                        for (int m = min_m; m < max_m; m++)
                            for (int q = min_n; q < max_n; q++) ns += (double)sp[m * BOX + q];
                        float f = (float)(ns / (double)((max_m - min_m) * (max_n - min_n)));
                        fmin_ = fminf(fmin_, f);
                    }
                }
                const float bg = pb_gmin<G>(fmin_);
                double ph = sum - (double)(BOX * BOX) * (double)bg;
                ph = ph > 1.0 ? ph : 1.0;
                // initial sigmas from the centre row / column of (spot - bg)
                constexpr int H = BOX / 2;
                double sdy = 0.0, sy_ = 0.0, sdx = 0.0, sx_ = 0.0;
                if (g < BOX) {
                    const float vy = sp[g * BOX + H] - bg;
                    const float vx = sp[H * BOX + g] - bg;
                    const double d2 = (double)((g - H) * (g - H));
                    sdy = (double)vy * d2; sy_ = (double)vy;
                    sdx = (double)vx * d2; sx_ = (double)vx;
                }
                sdy = pb_gsum<G>(sdy); sy_ = pb_gsum<G>(sy_);
                sdx = pb_gsum<G>(sdx); sx_ = pb_gsum<G>(sx_);
                double sy0, sx0;
                // the reference raises ZeroDivisionError when a sum is 0; we
                // emit its 0.01 fallback and set status bit 0
                if (sy_ == 0.0) { sy0 = 0.01; st_flags |= 1; } else sy0 = sqrt(sdy / sy_);
                if (sx_ == 0.0) { sx0 = 0.01; st_flags |= 1; } else sx0 = sqrt(sdx / sx_);
                if (!isfinite(sy0) || sy0 == 0.0) sy0 = 0.01;
                if (!isfinite(sx0) || sx0 == 0.0) sx0 = 0.01;
                th[0] = (float)xc; th[1] = (float)yc; th[2] = (float)ph; th[3] = bg;
                if constexpr (METHOD == 1) { th[4] = (float)sx0; th[5] = (float)sy0; }
                else { th[4] = (float)((sx0 + sy0) / 2.0); th[5] = th[4]; }
            }
            // max_step (gaussmle.py:558-561 / 770-773)
            float ms_mine;
            {
                const float ms0 = th[4];
                const float ms2 = (float)(0.1 * (double)th[2]);
                const float ms3 = (float)(0.1 * (double)th[3]);
                const float ms4 = (float)(0.2 * (double)th[4]);
                const float ms5 = (float)(0.2 * (double)th[5]);
                ms_mine = g < 2 ? ms0 : g == 2 ? ms2 : g == 3 ? ms3 : g == 4 ? ms4 : ms5;
            }

            // ---- Newton iterations, then one more pass of the same loop body
            //      in CRLB mode (one instance of the factor code: I-cache) ------
            int kk = 0;
            bool done = !valid;
            bool crlb_phase = false;
#pragma unroll 1
            while (true) {
                const bool active = !done && kk < a.max_it;
                if (!__any_sync(0xffffffffu, active)) crlb_phase = true;

                // ===== 1-D factors: lane g evaluates pixel EDGE g on both axes =====
                double PSFy, cy1, cy2, gy1, gy2;   // y factors of row g (registers)
                {
                    // reciprocals of sigma, f32(sigma^2), f32(sigma^3), f32(sigma^5)
                    // (float32 ** int stays float32 in the reference): one per lane
                    const int ax = (METHOD == 1) ? ((g >> 2) & 1) : 0;
                    const float sig = ax ? th[5] : th[4];
                    const float s2f = sig * sig;
                    const float s3f = sig * s2f;
                    const float s5f = sig * (s2f * s2f);
                    const int kq = g & 3;
                    const double v = kq == 0 ? (double)sig : kq == 1 ? (double)s2f
                                     : kq == 2 ? (double)s3f : (double)s5f;
                    const double r = 1.0 / v;
                    const double rsx = pb_gshfl<G>(r, 0), r2x = pb_gshfl<G>(r, 1);
                    const double r3x = pb_gshfl<G>(r, 2), r5x = pb_gshfl<G>(r, 3);
                    double rsy, r2y, r3y, r5y;
                    if constexpr (METHOD == 1) {
                        rsy = pb_gshfl<G>(r, 4); r2y = pb_gshfl<G>(r, 5);
                        r3y = pb_gshfl<G>(r, 6); r5y = pb_gshfl<G>(r, 7);
                    } else { rsy = rsx; r2y = r2x; r3y = r3x; r5y = r5x; }
                    const double sxd = (double)th[4], syd = (double)th[5];
                    // edge g: minus edge of pixel g == plus edge of pixel g-1 (exact)
                    const double ex = ((double)g - (double)th[0]) - 0.5;
                    const double ey = ((double)g - (double)th[1]) - 0.5;
                    const double tx = ex * rsx, ty = ey * rsy;
                    const double qx = 0.5 * tx * tx, qy = 0.5 * ty * ty;
                    const double Ax = exp(-qx), Ay = exp(-qy);
#if PB_MLE_LIBM_ERF
                    const double Ex = erf(ex * (kInvSqrt2 * rsx));
                    const double Ey = erf(ey * (kInvSqrt2 * rsy));
#else
                    const double Ex = erf_from_gauss(ex * (kInvSqrt2 * rsx), Ax);
                    const double Ey = erf_from_gauss(ey * (kInvSqrt2 * rsy), Ay);
#endif
                    // plus-edge values from the next lane
                    const double Exp_ = pb_gshfl_down1<G>(Ex), Eyp_ = pb_gshfl_down1<G>(Ey);
                    const double Axp = pb_gshfl_down1<G>(Ax), Ayp = pb_gshfl_down1<G>(Ay);
                    const double exp_ = ex + 1.0, eyp_ = ey + 1.0;
                    const double PSFx = 0.5 * (Exp_ - Ex);
                    PSFy = 0.5 * (Eyp_ - Ey);
                    const double cx1 = (Ax - Axp) * rsx * kInvSqrt2Pi;
                    cy1 = (Ay - Ayp) * rsy * kInvSqrt2Pi;
                    const double cx2 = (ex * Ax - exp_ * Axp) * r3x * kInvSqrt2Pi;
                    cy2 = (ey * Ay - eyp_ * Ayp) * r3y * kInvSqrt2Pi;
                    double gx1, gx2;
                    if constexpr (METHOD == 1) {
                        // _G uses exp(-(a^2)/(2*f32(sigma^2))): correct Ax by the
                        // f32 rounding of sigma^2 (gaussmle.py:306-316)
                        const double rhox = fma(sxd * sxd, r2x, -1.0), rhoy = fma(syd * syd, r2y, -1.0);
                        const double zx = -qx * rhox, zy = -qy * rhoy;
                        const double AGx = fma(Ax, fma(0.5 * zx, zx, zx), Ax);
                        const double AGy = fma(Ay, fma(0.5 * zy, zy, zy), Ay);
                        const double AGxp = pb_gshfl_down1<G>(AGx), AGyp = pb_gshfl_down1<G>(AGy);
                        const double w1x = ex * AGx - exp_ * AGxp, w1y = ey * AGy - eyp_ * AGyp;
                        const double w3x = ex * (ex * ex) * AGx - exp_ * (exp_ * exp_) * AGxp;
                        const double w3y = ey * (ey * ey) * AGy - eyp_ * (eyp_ * eyp_) * AGyp;
                        gx1 = w1x * r2x * kInvSqrt2Pi;
                        gy1 = w1y * r2y * kInvSqrt2Pi;
                        gx2 = (w3x * r5x - 2.0 * w1x * r3x) * kInvSqrt2Pi;
                        gy2 = (w3y * r5y - 2.0 * w1y * r3y) * kInvSqrt2Pi;
                    } else {
                        // isotropic sigma (gaussmle.py:339-383): dPSF/dsigma, d2PSF/dsigma2
                        const double am = ex * (rsx * kInvSqrt2), ap = exp_ * (rsx * kInvSqrt2);
                        const double bm = ey * (rsx * kInvSqrt2), bp = eyp_ * (rsx * kInvSqrt2);
                        const double Fx = am * Ax - ap * Axp, Fy = bm * Ay - bp * Ayp;
                        gx1 = Fx * rsx * kInvSqrtPi;      // dPSFxdt
                        gy1 = Fy * rsx * kInvSqrtPi;
                        const double dFx =
                            (ap * Axp * (1.0 - 2.0 * ap * ap) - am * Ax * (1.0 - 2.0 * am * am)) * rsx;
                        const double dFy =
                            (bp * Ayp * (1.0 - 2.0 * bp * bp) - bm * Ay * (1.0 - 2.0 * bm * bm)) * rsx;
                        const double rinvf = (double)(1.0f / th[4]);   // sigma ** (-1) is f32
                        gx2 = kInvSqrtPi * (-Fx * r2x + rinvf * dFx);  // d2PSFxdt2
                        gy2 = kInvSqrtPi * (-Fy * r2x + rinvf * dFy);
                    }
                    __syncwarp();   // previous readers of fx are done
                    if (g < BOX) {
                        double* c = fx + g * kNFx;
                        c[F_NPX] = (double)th[2] * PSFx;
                        c[F_PX] = PSFx;
                        c[F_C1] = cx1;
                        c[F_C2] = cx2;
                        c[F_G1] = gx1;
                        c[F_G2] = gx2;
                        c[F_PX2] = PSFx * PSFx;
                        c[F_C1SQ] = cx1 * cx1;
                        c[F_G1SQ] = gx1 * gx1;
                        c[F_G1PX] = gx1 * PSFx;
                        c[F_C1PX] = cx1 * PSFx;
                        c[F_C1G1] = cx1 * gx1;
#if PB_MLE_F32_PIXELS
                        float* cf32 = fxf + g * 12;
                        cf32[0] = (float)((double)th[2] * PSFx); cf32[1] = (float)PSFx;
                        cf32[2] = (float)cx1; cf32[3] = (float)cx2;
                        cf32[4] = (float)gx1; cf32[5] = (float)gx2;
                        cf32[6] = (float)(PSFx * PSFx); cf32[7] = (float)(cx1 * cx1);
                        cf32[8] = (float)(gx1 * gx1); cf32[9] = (float)(gx1 * PSFx);
#endif
                    }
                    __syncwarp();
                }
                const double N = (double)th[2], bg = (double)th[3];
                const double NPy = N * PSFy;
                double* myred = red + lane * kRedStride;
                const double* gr = red + (grp * G) * kRedStride;

                if (!crlb_phase) {
                    // ===== Newton sums.  Every derivative is (row factor) x (column
                    // factor), so lane j accumulates column-weighted sums of cf and df
                    // over its row and applies the row factors once per row. =====
                    double c0 = 0, cpx = 0, cc1 = 0, cc2 = 0, cg1 = 0, cg2 = 0;
                    double d0 = 0, dpx2 = 0, dc1 = 0, dg1 = 0, dgp = 0;
#if PB_MLE_F32_PIXELS
                    if (g < BOX) {
                        const float* frow = sp + g * BOX;
                        const float4* col = reinterpret_cast<const float4*>(fxf);
                        const float PSFy_f = (float)PSFy, bg_f = (float)bg;
                        float c0f = 0, cpxf = 0, cc1f = 0, cc2f = 0, cg1f = 0, cg2f = 0;
                        float d0f = 0, dpx2f = 0, dc1f = 0, dg1f = 0, dgpf = 0;
                        PB_PIX_UNROLL
                        for (int i = 0; i < BOX; i++, col += 3) {
                            const float4 f0 = col[0];    // N*px, px, c1, c2
                            const float4 f1 = col[1];    // g1, g2, px^2, c1^2
                            const float4 f2 = col[2];    // g1^2, g1*px, -, -
                            const float model = fmaf(f0.x, PSFy_f, bg_f);
                            const bool okm = model > 10e-3f;
                            const float inv = __frcp_rn(model);
                            const float t = frow[i] * inv;
                            float cf = t - 1.0f, df = t * inv;
                            cf = cf > 10e4f ? 10e4f : cf;
                            df = df > 10e4f ? 10e4f : df;
                            cf = okm ? cf : 0.0f;
                            df = okm ? df : 0.0f;
                            c0f += cf;
                            cpxf = fmaf(cf, f0.y, cpxf);
                            cc1f = fmaf(cf, f0.z, cc1f);
                            cc2f = fmaf(cf, f0.w, cc2f);
                            cg1f = fmaf(cf, f1.x, cg1f);
                            cg2f = fmaf(cf, f1.y, cg2f);
                            d0f += df;
                            dpx2f = fmaf(df, f1.z, dpx2f);
                            dc1f = fmaf(df, f1.w, dc1f);
                            dg1f = fmaf(df, f2.x, dg1f);
                            if constexpr (METHOD == 0) dgpf = fmaf(df, f2.y, dgpf);
                        }
                        c0 = c0f; cpx = cpxf; cc1 = cc1f; cc2 = cc2f; cg1 = cg1f; cg2 = cg2f;
                        d0 = d0f; dpx2 = dpx2f; dc1 = dc1f; dg1 = dg1f; dgp = dgpf;
                    }
#else
                    if (g < BOX) {
                        const double* drow = dspot + g * BOX;
                        const double2* col = reinterpret_cast<const double2*>(fx);
                        PB_PIX_UNROLL
                        for (int i = 0; i < BOX; i++, col += kNFx / 2) {
                            const double2 f0 = col[0];   // N*px, px
                            const double2 f1 = col[1];   // c1, c2
                            const double2 f2 = col[2];   // g1, g2
                            const double2 f3 = col[3];   // px^2, c1^2
                            const double2 f4 = col[4];   // g1^2, g1*px
                            const double model = fma(f0.x, PSFy, bg);
                            // branch-free guards (gaussmle.py:829-835): model > 0.01, cf/df <= 1e5
                            // (NaN/inf from a non-positive model are discarded by okm;
                            //  '>' selects instead of fmin keep ptxas from emitting the
                            //  NaN-aware DSETP.MIN sequences)
                            const bool okm = model > 10e-3;
                            const double inv = fast_rcp(model);
                            const double t = drow[i] * inv;
                            double cf = t - 1.0, df = t * inv;
                            cf = cf > 10e4 ? 10e4 : cf;
                            df = df > 10e4 ? 10e4 : df;
                            cf = okm ? cf : 0.0;
                            df = okm ? df : 0.0;
                            c0 += cf;
                            cpx = fma(cf, f0.y, cpx);
                            cc1 = fma(cf, f1.x, cc1);
                            cc2 = fma(cf, f1.y, cc2);
                            cg1 = fma(cf, f2.x, cg1);
                            cg2 = fma(cf, f2.y, cg2);
                            d0 += df;
                            dpx2 = fma(df, f3.x, dpx2);
                            dc1 = fma(df, f3.y, dc1);
                            dg1 = fma(df, f4.x, dg1);
                            if constexpr (METHOD == 0) dgp = fma(df, f4.y, dgp);
                        }
                    }
#endif
                    const double Ncy1 = N * cy1, Ncy2 = N * cy2;
                    myred[0] = NPy * cc1;            myred[6] = NPy * cc2 - NPy * NPy * dc1;
                    myred[1] = Ncy1 * cpx;           myred[7] = Ncy2 * cpx - Ncy1 * Ncy1 * dpx2;
                    myred[2] = PSFy * cpx;           myred[8] = -PSFy * PSFy * dpx2;
                    myred[3] = c0;                   myred[9] = -d0;
                    if constexpr (METHOD == 1) {
                        const double Ngy1 = N * gy1, Ngy2 = N * gy2;
                        myred[4] = NPy * cg1;        myred[10] = NPy * cg2 - NPy * NPy * dg1;
                        myred[5] = Ngy1 * cpx;       myred[11] = Ngy2 * cpx - Ngy1 * Ngy1 * dpx2;
                    } else {
                        // dudt = N*(PSFy*dPx + PSFx*dPy); the reference's d2udt2 has
                        // photons on the first term only (gaussmle.py:380-382)
                        myred[4] = N * (PSFy * cg1 + gy1 * cpx);
                        myred[10] = (NPy * cg2 + 2.0 * gy1 * cg1 + gy2 * cpx) -
                                    N * N * (PSFy * PSFy * dg1 + 2.0 * PSFy * gy1 * dgp +
                                             gy1 * gy1 * dpx2);
                    }
                    __syncwarp();
                    float th_new = 0.0f;
                    if (g < NP) {
                        double sn = 0.0, sd = 0.0;
#pragma unroll 1
                        for (int q = 0; q < BOX; q++) {
                            sn += gr[q * kRedStride + g];
                            sd += gr[q * kRedStride + 6 + g];
                        }
                        // clamped per-parameter Newton step in float32
                        // (gaussmle.py:648-670, 860-884)
                        const float nu = (float)sn, de = (float)sd;
                        float thl = g == 0 ? th[0] : g == 1 ? th[1] : g == 2 ? th[2] : g == 3 ? th[3]
                                    : g == 4 ? th[4] : th[5];
                        float upd;
                        if (de == 0.0f) {
                            if constexpr (METHOD == 1) {
                                const float sg = nu > 0.f ? 1.f : (nu < 0.f ? -1.f : nu);
                                upd = sg * ms_mine;
                            } else {
                                const float pr = nu * ms_mine;
                                upd = pr > 0.f ? 1.f : (pr < 0.f ? -1.f : pr);
                            }
                        } else {
                            upd = fminf(fmaxf(nu / de, -ms_mine), ms_mine);
                        }
                        thl -= upd;
                        if (g == 2) thl = fmaxf(thl, 1.0f);
                        if (g >= 3) thl = fmaxf(thl, 0.01f);
                        if (METHOD == 0 && g == 4) thl = fminf(thl, (float)BOX);
                        th_new = thl;
                    }
                    float tn[6];
#pragma unroll
                    for (int l = 0; l < NP; l++) tn[l] = pb_gshfl<G>(th_new, l);
                    if (METHOD == 0) tn[5] = tn[4];
                    if (active) {
                        kk++;
                        bool conv = ((double)fabsf(th[0] - tn[0]) < a.eps) &&
                                    ((double)fabsf(th[1] - tn[1]) < a.eps);
                        if (METHOD == 1)
                            conv = conv && ((double)fabsf(th[4] - tn[4]) < a.eps) &&
                                   ((double)fabsf(th[5] - tn[5]) < a.eps);
#pragma unroll
                        for (int l = 0; l < 6; l++) th[l] = tn[l];
                        if (conv) done = true;
                    }
                    continue;
                }

                // ===== CRLB + log-likelihood (gaussmle.py:673-742, 887-954) =====
                // column factors b in {c1, px, 1, g1}; acc[pair(b, b')] += bb'/model
                double ac[10];
#pragma unroll
                for (int q = 0; q < 10; q++) ac[q] = 0.0;
                double ll = 0.0;
                if (g < BOX) {
                    const double* drow = dspot + g * BOX;
                    const float* frow = sp + g * BOX;
#pragma unroll 1
                    for (int i = 0; i < BOX; i++) {
                        const double* c = fx + i * kNFx;
                        const double model = fma(c[F_NPX], PSFy, bg);
                        const double w = 1.0 / model;
                        ac[0] = fma(w, c[F_C1SQ], ac[0]);   // c1 c1
                        ac[1] = fma(w, c[F_C1PX], ac[1]);   // c1 px
                        ac[2] = fma(w, c[F_C1], ac[2]);     // c1 1
                        ac[3] = fma(w, c[F_C1G1], ac[3]);   // c1 g1
                        ac[4] = fma(w, c[F_PX2], ac[4]);    // px px
                        ac[5] = fma(w, c[F_PX], ac[5]);     // px 1
                        ac[6] = fma(w, c[F_G1PX], ac[6]);   // px g1
                        ac[7] += w;                         // 1 1
                        ac[8] = fma(w, c[F_G1], ac[8]);     // 1 g1
                        ac[9] = fma(w, c[F_G1SQ], ac[9]);   // g1 g1
                        const float dataf = frow[i];
                        if (model > 0.0) {
                            if (dataf > 0.0f)
                                ll += drow[i] * log(model) - model - (double)(dataf * logf(dataf)) +
                                      drow[i];
                            else
                                ll -= model;
                        }
                    }
                }
                // pair index of column-factor kinds: 0 = c1, 1 = px, 2 = one, 3 = g1
                auto pr = [&](int p, int q) -> double {
                    const int lo = p < q ? p : q, hi = p < q ? q : p;
                    return ac[lo * 4 - (lo * (lo - 1)) / 2 + (hi - lo)];
                };
                double F[NFISH];
                if constexpr (METHOD == 1) {
                    // dudt_k = arow[k] * b_{kind[k]}(i)
                    const double arow[6] = {NPy, N * cy1, PSFy, 1.0, NPy, N * gy1};
                    constexpr int kind[6] = {0, 1, 1, 2, 3, 1};
                    int q = 0;
#pragma unroll
                    for (int k = 0; k < 6; k++)
#pragma unroll
                        for (int l = k; l < 6; l++) F[q++] = arow[k] * arow[l] * pr(kind[k], kind[l]);
                } else {
                    // dudt_4 = N*PSFy*g1(i) + N*gy1*px(i)
                    const double arow[4] = {NPy, N * cy1, PSFy, 1.0};
                    constexpr int kind[4] = {0, 1, 1, 2};
                    const double u = NPy, v = N * gy1;
                    int q = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
#pragma unroll
                        for (int l = k; l < 4; l++) F[q++] = arow[k] * arow[l] * pr(kind[k], kind[l]);
                        F[q++] = arow[k] * (u * pr(kind[k], 3) + v * pr(kind[k], 1));
                    }
                    F[q++] = u * u * pr(3, 3) + 2.0 * u * v * pr(3, 1) + v * v * pr(1, 1);
                }
#pragma unroll
                for (int q = 0; q < NFISH; q++) myred[q] = F[q];
                myred[NFISH] = ll;
                __syncwarp();
                for (int q = g; q <= NFISH; q += G) {
                    double s = 0.0;
#pragma unroll 1
                    for (int r = 0; r < BOX; r++) s += gr[r * kRedStride + q];
                    sums[q] = s;
                }
                __syncwarp();
                double m[NFISH];
#pragma unroll
                for (int q = 0; q < NFISH; q++) m[q] = sums[q];
                const double llsum = sums[NFISH];
                double dg[NP];
                // np.linalg.pinv drops singular values below 1e-15 * max: a tiny
                // diagonal entry means a (numerically) null direction -> fallback
                double dmin = INFINITY, dmax = 0.0;
                {
                    int q = 0;
#pragma unroll
                    for (int k = 0; k < NP; k++) {
                        dmin = fmin(dmin, m[q]); dmax = fmax(dmax, m[q]);
                        q += NP - k;
                    }
                }
                bool ok = (dmin > 1e-13 * dmax) && inv_diag_cholesky<NP>(m, dg);
                if (!ok) {
                    st_flags |= 2;
                    double Mfull[NP * NP];
                    int q = 0;
                    for (int k = 0; k < NP; k++)
                        for (int l = k; l < NP; l++) {
                            Mfull[k * NP + l] = m[q];
                            Mfull[l * NP + k] = m[q];
                            q++;
                        }
                    pinv_diag_jacobi<NP>(Mfull, dg);
                }
                __syncwarp();
                float cr = 0.f, tv = 0.f;
#pragma unroll
                for (int l = 0; l < NP; l++)
                    if (g == l) { cr = (float)dg[l]; tv = th[l]; }
                if (METHOD == 0 && g == 5) { cr = (float)dg[4]; tv = th[4]; }
                const unsigned badmask = __ballot_sync(0xffffffffu, g < 6 && !isfinite(cr));
                const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (grp * G));
                if (badmask & gmask) st_flags |= 4;
                if (valid) {
                    if (g < 6) {
                        a.thetas[spot_idx * 6 + g] = tv;
                        a.crlbs[spot_idx * 6 + g] = cr;
                    }
                    if (g == 6) a.logliks[spot_idx] = (float)llsum;
                    if (g == 7) a.iterations[spot_idx] = kk;
                    if (a.status != nullptr && g == 0) a.status[spot_idx] = st_flags;
                }
                break;
            }   // iteration / CRLB loop
        }   // sub
        __syncwarp();   // everyone is done reading this stage before it is refilled
        cur_tma = next_tma;
        tile = next_tile;
        next_tile = (tile < ntiles) ? claim() : ntiles;
    }
}

// A ring of zero-initialised tile counters per device: launches on different streams may be
// in flight together (the host pipeline keeps three), so each launch takes the next slot and
// clears it on its own stream.
constexpr int kCounterSlots = 64;
int next_tile_counter(int dev, cudaStream_t stream, unsigned long long** out) {
    static std::mutex mu;
    static std::vector<unsigned long long*> rings;
    static std::vector<unsigned> cursor;
    std::lock_guard<std::mutex> lk(mu);
    if ((int)rings.size() <= dev) { rings.resize(dev + 1, nullptr); cursor.resize(dev + 1, 0); }
    if (!rings[dev]) PB_CUDA_CHECK(cudaMalloc(&rings[dev], kCounterSlots * sizeof(unsigned long long)));
    unsigned long long* slot = rings[dev] + (cursor[dev]++ % kCounterSlots);
    PB_CUDA_CHECK(cudaMemsetAsync(slot, 0, sizeof(unsigned long long), stream));
    *out = slot;
    return PB_OK;
}

template <int BOX, int G, int METHOD>
int launch_mle(const MleArgs& a, cudaStream_t stream) {
    using SM = MleSmem<BOX, G>;
    auto kern = mle_fit_kernel<BOX, G, METHOD>;
    int blocks_per_sm = 0, num_sms = 0, dev = 0;
    PB_CUDA_CHECK(cudaGetDevice(&dev));
    PB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SM::kTotal));
    PB_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    PB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern,
                                                                kWarpsPerBlock * 32, SM::kTotal));
    if (blocks_per_sm < 1) {
        pb_set_error("mle kernel does not fit on an SM (box=%d)", BOX);
        return PB_ERR_CUDA;
    }
    MleArgs args = a;
    {
        int rc = next_tile_counter(dev, stream, &args.tile_counter);
        if (rc != PB_OK) return rc;
    }
    const long long ntiles = (a.n + kTileSpots - 1) / kTileSpots;
    long long want = (ntiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
    long long cap = (long long)num_sms * blocks_per_sm;   // persistent: one wave
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, kWarpsPerBlock * 32, SM::kTotal, stream>>>(args);
    g_pb_launches++;
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

template <int METHOD>
int dispatch_box(int box, const MleArgs& a, cudaStream_t stream) {
    switch (box) {
        case 5:  return launch_mle<5, 8, METHOD>(a, stream);
        case 7:  return launch_mle<7, 8, METHOD>(a, stream);
        case 9:  return launch_mle<9, 16, METHOD>(a, stream);
        case 11: return launch_mle<11, 16, METHOD>(a, stream);
        case 13: return launch_mle<13, 16, METHOD>(a, stream);
        case 15: return launch_mle<15, 16, METHOD>(a, stream);
        case 17: return launch_mle<17, 32, METHOD>(a, stream);
        case 19: return launch_mle<19, 32, METHOD>(a, stream);
        case 21: return launch_mle<21, 32, METHOD>(a, stream);
        default:
            pb_set_error("unsupported box size %d (supported: odd 5..21)", box);
            return PB_ERR_INVALID;
    }
}

}  // namespace

// thread-per-spot path (mle_tps.cu)
bool pb_mle_tps_supports(int box);
int pb_mle_tps_fit(size_t n, int box, const float* d_spots, double eps, int max_it, int method,
                   float* d_thetas, float* d_crlbs, float* d_logliks, int* d_iterations,
                   int* d_status, cudaStream_t stream, int pixel_f32, float* mc_block = nullptr,
                   int mc_mode = 1);

// Implementation selector: 0 = lane-group kernel (this file), 1 = thread-per-spot with float64
// per-pixel sums, 2 = thread-per-spot with float32 per-pixel sums (default for box <= 13).
// PB_MLE_IMPL in the environment sets the initial value.
static std::atomic<int> g_mle_impl{-1};
static int mle_impl() {
    int v = g_mle_impl.load();
    if (v < 0) {
        v = 2;
        if (const char* e = getenv("PB_MLE_IMPL")) {
            const int w = atoi(e);
            if (w >= 0 && w <= 2) v = w;
        }
        g_mle_impl.store(v);
    }
    return v;
}
extern "C" int pb_mle_set_impl(int impl) {
    if (impl < 0 || impl > 2) {
        pb_set_error("pb_mle_set_impl: impl must be 0, 1 or 2");
        return PB_ERR_INVALID;
    }
    g_mle_impl.store(impl);
    return PB_OK;
}
extern "C" int pb_mle_get_impl(void) { return mle_impl(); }

// Device-pointer entry point, asynchronous on `stream`.  Declared in
// include/picasso_b200.h.
extern "C" int pb_mle_fit_dev(size_t n, int box, const float* d_spots, double eps, int max_it,
                              int method, float* d_thetas, float* d_crlbs, float* d_logliks,
                              int* d_iterations, int* d_status, void* stream) {
    if (method != 0 && method != 1) {
        pb_set_error("Method not available.");   // gaussmle.py:465
        return PB_ERR_INVALID;
    }
    if (n == 0) return PB_OK;
    if (!d_spots || !d_thetas || !d_crlbs || !d_logliks || !d_iterations) {
        pb_set_error("pb_mle_fit_dev: null device pointer");
        return PB_ERR_INVALID;
    }
    if (max_it < 0) max_it = 0;
    MleArgs a{d_spots, (long long)n, eps, max_it, d_thetas, d_crlbs, d_logliks, d_iterations,
              d_status, nullptr};
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int impl = mle_impl();
    if (impl != 0 && pb_mle_tps_supports(box))
        return pb_mle_tps_fit(n, box, d_spots, eps, max_it, method, d_thetas, d_crlbs, d_logliks,
                              d_iterations, d_status, s, impl == 2);
    return method == 1 ? dispatch_box<1>(box, a, s) : dispatch_box<0>(box, a, s);
}

// Fused fit + all-gather (multi-GPU): as pb_mle_fit_dev, and the kernel that finishes a spot (CRLB /
// log-likelihood pass) also stores the spot's 14 output words through `mc_block`, the NVSwitch
// MULTICAST address of this rank's block [thetas 6n | crlbs 6n | logliks n | iterations n] of the gather
// buffer (csrc/multicast.cu): the switch replicates every store into all ranks' buffers, so the
// all-gather of the results costs one store stream per rank and no copy, kernel or collective after the
// fit.  The data are complete on all ranks once every rank's stream has passed the call and the ranks
// have synchronised (barrier).
// 1 (default): the finishing kernel stores all 14 words of a spot; 2: theta + iterations leave from the
// iteration kernel as lanes finish (spread over the step), CRLB + logL are not stored -- the caller copies
// that half of its block through the mapping (pb_mc_copy_async / a copy-engine copy) behind the next step.
static std::atomic<int> g_gather_mode{1};
extern "C" int pb_mle_fit_gather_mode(int mode) {
    if (mode != 1 && mode != 2) { pb_set_error("pb_mle_fit_gather_mode: 1 or 2"); return PB_ERR_INVALID; }
    g_gather_mode.store(mode);
    return PB_OK;
}

extern "C" int pb_mle_fit_gather_dev(size_t n, int box, const float* d_spots, double eps, int max_it,
                                     int method, float* d_thetas, float* d_crlbs, float* d_logliks,
                                     int* d_iterations, int* d_status, void* mc_block, void* stream) {
    if (method != 0 && method != 1) { pb_set_error("Method not available."); return PB_ERR_INVALID; }
    if (n == 0) return PB_OK;
    if (!d_spots || !d_thetas || !d_crlbs || !d_logliks || !d_iterations || !mc_block) {
        pb_set_error("pb_mle_fit_gather_dev: null device pointer");
        return PB_ERR_INVALID;
    }
    if ((reinterpret_cast<uintptr_t>(mc_block) & 15) || (n & 1)) {
        pb_set_error("pb_mle_fit_gather_dev: the gather block must be 16-byte aligned and n even");
        return PB_ERR_INVALID;
    }
    const int impl = mle_impl();
    if (impl == 0 || !pb_mle_tps_supports(box)) {
        pb_set_error("pb_mle_fit_gather_dev: needs the thread-per-spot kernels (box <= 13, impl 1 or 2)");
        return PB_ERR_INVALID;
    }
    if (max_it < 0) max_it = 0;
    return pb_mle_tps_fit(n, box, d_spots, eps, max_it, method, d_thetas, d_crlbs, d_logliks, d_iterations,
                          d_status, reinterpret_cast<cudaStream_t>(stream), impl == 2, static_cast<float*>(mc_block),
                          g_gather_mode.load());
}
