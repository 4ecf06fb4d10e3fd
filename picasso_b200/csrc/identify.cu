// picasso_b200/csrc/identify.cu
//
// Spot identification and ROI extraction on B200 -- replaces
//   localize._local_maxima + _net_gradient + identify_in_image/identify_in_frame
//       (reference picasso/localize.py:97-337)  -> pb_identify*
//   localize._cut_spots_numba + _to_photons (get_spots)
//       (reference picasso/localize.py:917-931, 1101-1145) -> pb_get_spots*
//
// identify kernel (fused local-max + net-gradient filter), one CTA per
// TY x TX output tile of one frame:
//   * the tile plus a halo is staged in shared memory with 16-byte vector
//     loads (8 uint16 / 4 float per load) when the row pitch allows it;
//   * thread j walks DOWN column j keeping the row-window maxima of the last
//     2h+1 rows in registers (separable form of the reference's b x b argmax:
//     centre > every window pixel before it in row-major order and >= every
//     one after it -- np.argmax returns the first maximum);
//   * a thread that finds a maximum evaluates the net gradient in float32 in
//     the reference's row-major order with unfused IEEE ops, including the
//     reference's negative-index wrap-around at the top/left scan border, and
//     appends (frame, y, x, ng) if ng > minimum_ng.
// HBM-bound by design: each pixel is read once (+ halo re-reads through L2).
#include <algorithm>
#include <cmath>
#include <stdlib.h>
#include <atomic>
#include <numeric>
#include <vector>

#include "pb_common.cuh"
#include "../../include/picasso_b200.h"

extern std::atomic<long long> g_pb_launches;

namespace {

constexpr int kTX = 128;    // output columns per tile == threads per CTA
// output rows per tile: chosen so that kTY + 2h is a multiple of the ring period 2h+1
// (the column walk is unrolled by exactly one period -> compact loop body, I-cache friendly)
__host__ __device__ constexpr int tile_rows(int h) {
    return h == 1 ? 34 : h == 2 ? 36 : h == 3 ? 36 : h == 4 ? 37 : h == 5 ? 34 : h == 6 ? 40 : 31;
}
constexpr int kHP = 8;      // column halo in shared memory (>= h+1, keeps 16 B alignment)
constexpr int kCandCap = 96;   // per-warp list of local maxima awaiting their net gradient

struct IdArgs {
    const void* movie;      // frames [n_frames][Y][X]
    long long n_frames;
    int Y, X;               // full frame shape (row pitch X)
    int y0, x0, Ys, Xs;     // ROI window (image = frame[y0:y0+Ys, x0:x0+Xs])
    long long frame_offset; // added to the emitted frame number
    double min_ng_d;        // threshold, compared in double like the reference (ng > minimum_ng)
    double ng_bound;        // |ng| <= (max - min over the (b+2)^2 neighbourhood) * ng_bound; 0: no pre-filter
    long long* out_frame;
    long long* out_x;
    long long* out_y;
    float* out_ng;
    unsigned long long capacity;
    unsigned long long* counter;   // total found (may exceed capacity)
};

template <typename T> struct PixTraits;
template <> struct PixTraits<unsigned short> {
    using Vec = uint4;              // 8 pixels
    static constexpr int kPerVec = 8;
    __device__ static float to_f32(unsigned short v) { return (float)v; }
    // f32(a) - f32(b) for uint16 a, b: the difference is an exact integer
    __device__ static float diff(unsigned short a, unsigned short b) { return (float)((int)a - (int)b); }
};
template <> struct PixTraits<float> {
    using Vec = float4;             // 4 pixels
    static constexpr int kPerVec = 4;
    __device__ static float to_f32(float v) { return v; }
    __device__ static float diff(float a, float b) { return __fsub_rn(a, b); }
};

template <typename T>
__device__ __forceinline__ T pmax(T a, T b) { return a > b ? a : b; }

// packed helpers: two uint16 pixels per 32-bit word (low half = even column)
// (PTX max.u16x2 is a native packed instruction on sm_90+; the __vmaxu2 / __vcmp*2
//  SIMD-video intrinsics are emulated with long LOP3/SEL sequences)
__device__ __forceinline__ unsigned pk_max(unsigned a, unsigned b) {
    unsigned r;
    asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ unsigned pk_min(unsigned a, unsigned b) {
    unsigned r;
    asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
// pixels (2k+1, 2k+2) from the words holding (2k, 2k+1) and (2k+2, 2k+3)
__device__ __forceinline__ unsigned pk_odd(unsigned w0, unsigned w1) { return __byte_perm(w0, w1, 0x5432); }

template <typename T, int H, bool PACKED>
__global__ void __launch_bounds__(kTX) identify_kernel(const IdArgs a) {
    constexpr int BOX = 2 * H + 1;
    constexpr int kTY = tile_rows(H);
    constexpr int P = 2 * H + 1;            // ring period of the column walk
    static_assert((kTY + 2 * H) % P == 0 && kTY <= 64, "tile height must fit the ring period");
    static_assert(!PACKED || sizeof(T) == 2, "packed walk is for uint16 movies");
    constexpr int ROWS = kTY + 2 * H + 2;   // window halo H + gradient halo 1
    constexpr int TXW = PACKED ? 2 * kTX : kTX;   // output columns per tile
    constexpr int COLS = TXW + 2 * kHP;
    static_assert(H + 1 <= kHP, "halo too small");
    __shared__ __align__(16) T tile[ROWS][COLS];
    __shared__ float ux[BOX * BOX], uy[BOX * BOX];
    __shared__ unsigned short cand_list[kTX / 32][kCandCap];

    const int tid = threadIdx.x;
    const long long f = blockIdx.z;
    const int ty0 = blockIdx.y * kTY;            // first output row (image coords)
    const int tx0 = blockIdx.x * TXW;            // first output column (image coords)
    const T* frame = static_cast<const T*>(a.movie) + (size_t)f * a.Y * a.X;
    // image(r, c) = frame[(y0 + r) * X + (x0 + c)], valid for 0<=r<Ys, 0<=c<Xs

    // unit vectors towards the centre (localize.py:279-286), float32 IEEE ops
    for (int q = tid; q < BOX * BOX; q += kTX) {
        const float vx = (float)(H - q % BOX), vy = (float)(H - q / BOX);
        const float un = __fsqrt_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)));
        ux[q] = __fdiv_rn(vx, un);
        uy[q] = __fdiv_rn(vy, un);
    }

    // ---- stage the tile: rows ty0-H-1 .. ty0+kTY+H, cols tx0-kHP .. tx0+TXW+kHP-1
    {
        using V = typename PixTraits<T>::Vec;
        constexpr int PV = PixTraits<T>::kPerVec;
        const bool vec_ok = ((a.X % PV) == 0) && ((a.x0 % PV) == 0) &&
                            ((reinterpret_cast<uintptr_t>(a.movie) & 15) == 0);
        constexpr int VPR = COLS / PV;   // vectors per tile row
        for (int idx = tid; idx < ROWS * VPR; idx += kTX) {
            const int tr = idx / VPR, tv = idx % VPR;
            const int r = ty0 - H - 1 + tr;
            const int c = tx0 - kHP + tv * PV;
            T* dst = &tile[tr][tv * PV];
            if (r >= 0 && r < a.Ys && c >= 0 && c + PV <= a.Xs && vec_ok) {
                // 16-byte asynchronous global->shared copy (LDGSTS): no register staging,
                // all of a thread's vectors are in flight at once
                const V* src = reinterpret_cast<const V*>(frame + (size_t)(a.y0 + r) * a.X + a.x0 + c);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pb_smem_u32(dst)), "l"(src)
                             : "memory");
            } else {
#pragma unroll
                for (int k = 0; k < PV; k++) {
                    const int cc = c + k;
                    dst[k] = (r >= 0 && r < a.Ys && cc >= 0 && cc < a.Xs)
                                 ? frame[(size_t)(a.y0 + r) * a.X + a.x0 + cc]
                                 : T(0);
                }
            }
        }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // Local maxima (localize.py:97-134) in separable form: the centre must be > every
    // window pixel before it in row-major order and >= every one after it (np.argmax
    // returns the first maximum).  Thread walks DOWN its column(s) keeping, in register
    // rings of period P = 2H+1, the row-window maximum, centre, left-part and right-part
    // maxima of the last P rows (slot = tile row mod P).  Scan range of the reference:
    // i in [H, Ys-H-1), j in [H, Xs-H-1).
    unsigned long long cand0 = 0, cand1 = 0;   // bit u: (ty0+u, column) is a local maximum
    // rows of this tile inside the reference's scan range: 0 <= u < kTY, H <= ty0+u < Ys-H-1
    unsigned long long rowmask;
    {
        const int ulo = max(0, H - ty0), uhi = min(kTY, a.Ys - H - 1 - ty0);
        rowmask = (uhi > ulo) ? ((uhi - ulo >= 64 ? ~0ull : ((1ull << (uhi - ulo)) - 1ull)) << ulo) : 0ull;
    }
    if constexpr (!PACKED) {
        const int j = tx0 + tid;                 // image column of this thread
        const int tc = kHP + tid;                // tile column
        const bool col_ok = (j >= H) && (j < a.Xs - H - 1);
        T rm[P], cc[P], ll[P], rr[P];
#pragma unroll
        for (int k = 0; k < P; k++) { rm[k] = T(0); cc[k] = T(0); ll[k] = T(0); rr[k] = T(0); }
#pragma unroll 1
        for (int t0 = 0; t0 < kTY + 2 * H; t0 += P) {
            unsigned pm = 0;
#pragma unroll
            for (int k = 0; k < P; k++) {
                const int t = t0 + k;            // tile row t+1 <-> image row ty0 - H + t
                const T* row = &tile[t + 1][tc];
                T L = row[-H], R = row[1];
#pragma unroll
                for (int q = 1; q < H; q++) { L = pmax(L, row[-H + q]); R = pmax(R, row[1 + q]); }
                const T Cv = row[0];
                rm[k] = pmax(pmax(L, R), Cv);
                cc[k] = Cv; ll[k] = L; rr[k] = R;
                // centre = row t-H (slot k-H); rows above t-2H..t-H-1, below t-H+1..t
                T above = rm[(k + 1) % P], below = rm[(k + P - H + 1) % P];
#pragma unroll
                for (int q = 1; q < H; q++) {
                    above = pmax(above, rm[(k + 1 + q) % P]);
                    below = pmax(below, rm[(k + P - H + 1 + q) % P]);
                }
                const int cs = (k + P - H) % P;
                const T c0 = cc[cs];
                const bool is_max = (c0 > above) && (c0 > ll[cs]) && (c0 >= rr[cs]) && (c0 >= below);
                pm |= (is_max ? 1u : 0u) << k;   // bit k <-> output row t0 + k - 2H
            }
            // one variable shift per period instead of one per row
            if (t0 >= 2 * H) cand0 |= (unsigned long long)pm << (t0 - 2 * H);
            else cand0 |= (unsigned long long)pm >> (2 * H - t0);
        }
        if (!col_ok) cand0 = 0;
        cand0 &= rowmask;
    } else {
        // two adjacent columns per thread, packed uint16x2 SIMD: 2.5 shared loads per
        // pixel instead of 7 and half the compare/max instructions
        const int j = tx0 + 2 * tid;             // even image column of the pair
        const int wc = (kHP >> 1) + tid;         // word index of the pair in a tile row
        const bool ok0 = (j >= H) && (j < a.Xs - H - 1);
        const bool ok1 = (j + 1 >= H) && (j + 1 < a.Xs - H - 1);
        constexpr int W0 = -((H + 1) / 2);       // first / last word offset needed for
        constexpr int W1 = (H + 2) / 2;          // pixel offsets -H .. H+1
        unsigned rm[P], cc[P], ll[P], rr[P];
#pragma unroll
        for (int k = 0; k < P; k++) { rm[k] = 0; cc[k] = 0; ll[k] = 0; rr[k] = 0; }
#pragma unroll 1
        for (int t0 = 0; t0 < kTY + 2 * H; t0 += P) {
            unsigned pm0 = 0, pm1 = 0;
#pragma unroll
            for (int k = 0; k < P; k++) {
                const int t = t0 + k;
                const unsigned* row = reinterpret_cast<const unsigned*>(&tile[t + 1][0]) + wc;
                unsigned w[W1 - W0 + 1];
#pragma unroll
                for (int q = W0; q <= W1; q++) w[q - W0] = row[q];
                // V(d) = pixels (j+d, j+1+d): aligned word for even d, byte-permute for odd d
                auto V = [&](int d) -> unsigned {
                    if ((d & 1) == 0) return w[d / 2 - W0];
                    const int lo = (d - 1) / 2;   // floor for negative odd d
                    return pk_odd(w[lo - W0], w[lo + 1 - W0]);
                };
                unsigned L = V(-H), R = V(1);
#pragma unroll
                for (int q = 1; q < H; q++) { L = pk_max(L, V(-H + q)); R = pk_max(R, V(1 + q)); }
                const unsigned Cv = V(0);
                rm[k] = pk_max(pk_max(L, R), Cv);
                cc[k] = Cv; ll[k] = L; rr[k] = R;
                unsigned above = rm[(k + 1) % P], below = rm[(k + P - H + 1) % P];
#pragma unroll
                for (int q = 1; q < H; q++) {
                    above = pk_max(above, rm[(k + 1 + q) % P]);
                    below = pk_max(below, rm[(k + P - H + 1 + q) % P]);
                }
                const int cs = (k + P - H) % P;
                const unsigned c0 = cc[cs];
                // strict part: c0 > max(above, left)  <=>  max(c0, m1) != m1   (per half)
                // weak part:   c0 >= max(right, below) <=> max(c0, m2) == c0   (per half)
                const unsigned m1 = pk_max(above, ll[cs]);
                const unsigned m2 = pk_max(rr[cs], below);
                const unsigned sgt = pk_max(c0, m1) ^ m1;      // half != 0: strictly greater
                const unsigned wge = pk_max(c0, m2) ^ c0;      // half == 0: greater or equal
                const unsigned h0 = ((sgt & 0x0000ffffu) != 0) & ((wge & 0x0000ffffu) == 0);
                const unsigned h1 = ((sgt & 0xffff0000u) != 0) & ((wge & 0xffff0000u) == 0);
                pm0 |= h0 << k;                  // bit k <-> output row t0 + k - 2H
                pm1 |= h1 << k;
            }
            if (t0 >= 2 * H) {
                cand0 |= (unsigned long long)pm0 << (t0 - 2 * H);
                cand1 |= (unsigned long long)pm1 << (t0 - 2 * H);
            } else {
                cand0 |= (unsigned long long)pm0 >> (2 * H - t0);
                cand1 |= (unsigned long long)pm1 >> (2 * H - t0);
            }
        }
        if (!ok0) cand0 = 0;
        if (!ok1) cand1 = 0;
        cand0 &= rowmask;
        cand1 &= rowmask;
    }

    // ---- net gradient of the local maxima (localize.py:202-244) -------------------
    // float32, row-major accumulation, unfused IEEE ops, numba's negative-index wrap.
    // Noise alone makes ~1 pixel in BOX^2 a local maximum, unevenly spread over the
    // lanes, so the warp's candidates are first compacted into a shared list and then
    // dealt out one per lane (the per-candidate sum stays sequential -> bit-exact).
    const int lane = tid & 31, wid = tid >> 5;
    const int mine = __popcll(cand0) + __popcll(cand1);
    int prefix = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, prefix, o);
        if (lane >= o) prefix += v;
    }
    const int total = __shfl_sync(0xffffffffu, prefix, 31);
    int wpos = prefix - mine;
    // list entry: (u << 8) | column-in-tile.  If the list overflows, the excess stays in
    // left0/left1 and is handled by its own lane afterwards.
    unsigned long long left0 = 0, left1 = 0;
    const int col0 = PACKED ? 2 * tid : tid;
    while (cand0) {
        const int u = __ffsll((long long)cand0) - 1;
        cand0 &= cand0 - 1;
        if (wpos < kCandCap) cand_list[wid][wpos] = (unsigned short)((u << 8) | col0);
        else left0 |= (1ull << u);
        wpos++;
    }
    while (cand1) {
        const int u = __ffsll((long long)cand1) - 1;
        cand1 &= cand1 - 1;
        if (wpos < kCandCap) cand_list[wid][wpos] = (unsigned short)((u << 8) | (col0 + 1));
        else left1 |= (1ull << u);
        wpos++;
    }
    __syncwarp();
    const int listed = min(total, kCandCap);
    for (int base = 0; base < listed || __any_sync(0xffffffffu, (left0 | left1) != 0); base += 32) {
        int u = -1, col = col0;
        if (base + lane < listed) {
            const unsigned short e = cand_list[wid][base + lane];
            u = e >> 8; col = e & 0xff;
        } else if (base >= listed && left0) {
            u = __ffsll((long long)left0) - 1;
            left0 &= left0 - 1;
        } else if (base >= listed && left1) {
            u = __ffsll((long long)left1) - 1;
            left1 &= left1 - 1;
            col = col0 + 1;
        }
        if (u < 0) continue;
        const int i = ty0 + u;
        const int jc = tx0 + col;                      // image column of the candidate
        const int tcc = kHP + col;                     // its tile column
        const int tr_c = u + H + 1;                    // tile row of the centre
        float acc = 0.0f;
        if (i > H && jc > H) {
            // Pre-filter: every gradient entering the sum is a difference of two pixels of the
            // (b+2) x (b+2) neighbourhood, so |ng| <= (max - min) * sum_q(|ux_q| + |uy_q|).  Noise maxima
            // (range ~20 counts -> bound ~1 300 against a threshold of 5 000) are dropped without the
            // 48-term sequential sum; whatever passes is summed exactly as before, so the detections and
            // their net gradients stay bit-identical.  (A superset of the neighbourhood only loosens the bound.)
            if (a.ng_bound > 0.0) {
                float range;
                if constexpr (PACKED && H <= 6) {
                    const int w0 = (tcc - H - 1) >> 1;          // aligned words covering the columns
                    unsigned mn = 0xffffffffu, mx = 0u;
#pragma unroll 1
                    for (int rr = -H - 1; rr <= H + 1; rr++) {
                        const unsigned* rw = reinterpret_cast<const unsigned*>(&tile[tr_c + rr][0]) + w0;
#pragma unroll
                        for (int w = 0; w < H + 2; w++) {
                            const unsigned v = rw[w];
                            mn = pk_min(mn, v);
                            mx = pk_max(mx, v);
                        }
                    }
                    const unsigned lo = min(mn & 0xffffu, mn >> 16), hi = max(mx & 0xffffu, mx >> 16);
                    range = (float)(hi - lo);
                } else {
                    T mn = tile[tr_c][tcc], mx = mn;
#pragma unroll 1
                    for (int rr = -H - 1; rr <= H + 1; rr++)
                        for (int cc2 = -H - 1; cc2 <= H + 1; cc2++) {
                            const T v = tile[tr_c + rr][tcc + cc2];
                            mn = v < mn ? v : mn;
                            mx = v > mx ? v : mx;
                        }
                    range = PixTraits<T>::diff(mx, mn);
                }
                if ((double)range * a.ng_bound <= a.min_ng_d) continue;
            }
            // interior: every neighbour is in the tile; u16 differences are exact integers
#pragma unroll 1
            for (int kk = -H; kk <= H; kk++) {
                const T* r0 = &tile[tr_c + kk - 1][tcc];
                const T* r1 = &tile[tr_c + kk][tcc];
                const T* r2 = &tile[tr_c + kk + 1][tcc];
#pragma unroll
                for (int mm = -H; mm <= H; mm++) {
                    if (kk == 0 && mm == 0) continue;
                    const float gy = PixTraits<T>::diff(r2[mm], r0[mm]);
                    const float gx = PixTraits<T>::diff(r1[mm + 1], r1[mm - 1]);
                    const int q = (kk + H) * BOX + (mm + H);
                    acc = __fadd_rn(acc, __fadd_rn(__fmul_rn(gy, uy[q]), __fmul_rn(gx, ux[q])));
                }
            }
        } else {
            for (int kk = -H; kk <= H; kk++) {
                for (int mm = -H; mm <= H; mm++) {
                    if (kk == 0 && mm == 0) continue;
                    const int k = i + kk, m = jc + mm;
                    float up, lf;
                    const float dn = PixTraits<T>::to_f32(tile[tr_c + kk + 1][tcc + mm]);
                    const float rt = PixTraits<T>::to_f32(tile[tr_c + kk][tcc + mm + 1]);
                    if (k - 1 >= 0) up = PixTraits<T>::to_f32(tile[tr_c + kk - 1][tcc + mm]);
                    else  // numba negative index: frame[-1] is the LAST image row
                        up = PixTraits<T>::to_f32(frame[(size_t)(a.y0 + a.Ys - 1) * a.X + a.x0 + m]);
                    if (m - 1 >= 0) lf = PixTraits<T>::to_f32(tile[tr_c + kk][tcc + mm - 1]);
                    else
                        lf = PixTraits<T>::to_f32(frame[(size_t)(a.y0 + k) * a.X + a.x0 + a.Xs - 1]);
                    const float gy = __fsub_rn(dn, up);
                    const float gx = __fsub_rn(rt, lf);
                    const int q = (kk + H) * BOX + (mm + H);
                    acc = __fadd_rn(acc, __fadd_rn(__fmul_rn(gy, uy[q]), __fmul_rn(gx, ux[q])));
                }
            }
        }
        if ((double)acc > a.min_ng_d) {
            const unsigned long long slot = atomicAdd(a.counter, 1ull);
            if (slot < a.capacity) {
                a.out_frame[slot] = f + a.frame_offset;
                a.out_y[slot] = i + a.y0;
                a.out_x[slot] = jc + a.x0;
                a.out_ng[slot] = acc;
            }
        }
    }
}

template <typename T>
int launch_identify(const IdArgs& a, int box, cudaStream_t stream) {
    if (a.Ys <= 0 || a.Xs <= 0 || a.n_frames <= 0) return PB_OK;
    constexpr bool PACKED = (sizeof(T) == 2);      // uint16 movies: two pixels per thread
    constexpr int TXW = PACKED ? 2 * kTX : kTX;
    const int ty = tile_rows(box / 2);
    dim3 grid((a.Xs + TXW - 1) / TXW, (a.Ys + ty - 1) / ty, 1);
    // gridDim.z is limited to 65535: loop over frame batches
    const long long zmax = 32768;
    // pre-filter constant: sum over the b*b - 1 unit vectors of |ux| + |uy| (float64, 0.1 % head room for the
    // float32 rounding of the sequential sum); PB_IDENTIFY_PREFILTER=0 disables it for A/B runs
    static const bool prefilter = [] {
        const char* e = getenv("PB_IDENTIFY_PREFILTER");
        return !(e && atoi(e) == 0);
    }();
    double bound = 0.0;
    if (prefilter) {
        const int h = box / 2;
        for (int ky = -h; ky <= h; ky++)
            for (int kx = -h; kx <= h; kx++)
                if (ky || kx) bound += (std::abs((double)ky) + std::abs((double)kx)) / std::sqrt((double)(ky * ky + kx * kx));
        bound *= 1.001;
    }
    for (long long f0 = 0; f0 < a.n_frames; f0 += zmax) {
        IdArgs b = a;
        b.ng_bound = bound;
        const long long nf = std::min(zmax, a.n_frames - f0);
        size_t fsz = (size_t)a.Y * a.X * sizeof(T);
        b.movie = static_cast<const char*>(a.movie) + (size_t)f0 * fsz;
        b.n_frames = nf;
        b.frame_offset = a.frame_offset + f0;
        grid.z = (unsigned)nf;
        switch (box / 2) {
            case 1: identify_kernel<T, 1, PACKED><<<grid, kTX, 0, stream>>>(b); break;
            case 2: identify_kernel<T, 2, PACKED><<<grid, kTX, 0, stream>>>(b); break;
            case 3: identify_kernel<T, 3, PACKED><<<grid, kTX, 0, stream>>>(b); break;
            case 4: identify_kernel<T, 4, PACKED><<<grid, kTX, 0, stream>>>(b); break;
            case 5: identify_kernel<T, 5, PACKED><<<grid, kTX, 0, stream>>>(b); break;
            case 6: identify_kernel<T, 6, PACKED><<<grid, kTX, 0, stream>>>(b); break;
            case 7: identify_kernel<T, 7, PACKED><<<grid, kTX, 0, stream>>>(b); break;
            default:
                pb_set_error("unsupported box size %d for identify (odd 3..15)", box);
                return PB_ERR_INVALID;
        }
        g_pb_launches++;
        PB_CUDA_CHECK(cudaGetLastError());
    }
    return PB_OK;
}

// ---- ROI gather + photon conversion ----------------------------------------
struct CutArgs {
    const void* movie;
    long long n_frames;        // frames present in `movie`
    long long frame_offset;    // movie[0] is frame number frame_offset
    int Y, X, box;
    const long long* frame;
    const long long* x;
    const long long* y;
    long long n;
    float baseline, sensitivity, gain;
    float* spots;
};

template <typename T>
__global__ void cut_spots_kernel(const CutArgs a) {
    const int pix = a.box * a.box;
    const long long total = a.n * pix;
    const int r = a.box / 2;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long id = idx / pix;
        const int p = (int)(idx - id * pix);
        const long long fr = a.frame[id] - a.frame_offset;
        if (fr < 0 || fr >= a.n_frames) continue;     // spot belongs to another chunk
        const int yy = (int)a.y[id] - r + p / a.box;
        const int xx = (int)a.x[id] - r + p % a.box;
        const T* frame = static_cast<const T*>(a.movie) + (size_t)fr * a.Y * a.X;
        // out-of-frame windows cannot come from identify(); clamp defensively
        const int yc = min(max(yy, 0), a.Y - 1), xc = min(max(xx, 0), a.X - 1);
        const float s = PixTraits<T>::to_f32(frame[(size_t)yc * a.X + xc]);
        // (spots - baseline) * sensitivity / gain in float32 (localize.py:1106-1112)
        a.spots[idx] = __fdiv_rn(__fmul_rn(__fsub_rn(s, a.baseline), a.sensitivity), a.gain);
    }
}

}  // namespace

extern "C" int pb_identify_dev(const void* d_movie, int dtype, size_t n_frames, int Y, int X,
                               long long frame_offset, int box, double min_ng, const int* roi,
                               long long* d_frame, long long* d_x, long long* d_y, float* d_ng,
                               size_t capacity, unsigned long long* d_counter, void* stream) {
    if (box < 3 || box > 15 || (box & 1) == 0) {
        pb_set_error("unsupported box size %d for identify (odd 3..15)", box);
        return PB_ERR_INVALID;
    }
    if (dtype != PB_DTYPE_U16 && dtype != PB_DTYPE_F32) {
        pb_set_error("unsupported movie dtype %d (0 = uint16, 1 = float32)", dtype);
        return PB_ERR_INVALID;
    }
    IdArgs a{};
    a.movie = d_movie; a.n_frames = (long long)n_frames; a.Y = Y; a.X = X;
    a.y0 = 0; a.x0 = 0; a.Ys = Y; a.Xs = X;
    if (roi) {   // ((y0, x0), (y1, x1)) with python slice clamping
        int y0 = roi[0], x0 = roi[1], y1 = roi[2], x1 = roi[3];
        if (y0 < 0) y0 += Y; if (x0 < 0) x0 += X; if (y1 < 0) y1 += Y; if (x1 < 0) x1 += X;
        y0 = std::max(0, std::min(y0, Y)); x0 = std::max(0, std::min(x0, X));
        y1 = std::max(y0, std::min(y1, Y)); x1 = std::max(x0, std::min(x1, X));
        a.y0 = y0; a.x0 = x0; a.Ys = y1 - y0; a.Xs = x1 - x0;
    }
    a.frame_offset = frame_offset;
    a.min_ng_d = min_ng;
    a.out_frame = d_frame; a.out_x = d_x; a.out_y = d_y; a.out_ng = d_ng;
    a.capacity = capacity; a.counter = d_counter;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    return dtype == PB_DTYPE_U16 ? launch_identify<unsigned short>(a, box, s)
                                 : launch_identify<float>(a, box, s);
}

extern "C" int pb_get_spots_dev(const void* d_movie, int dtype, size_t n_frames, int Y, int X,
                                long long frame_offset, size_t n, const long long* d_frame,
                                const long long* d_x, const long long* d_y, int box, float baseline,
                                float sensitivity, float gain, float* d_spots, void* stream) {
    if (box < 1 || (box & 1) == 0) { pb_set_error("box must be odd"); return PB_ERR_INVALID; }
    if (dtype != PB_DTYPE_U16 && dtype != PB_DTYPE_F32) {
        pb_set_error("unsupported movie dtype %d", dtype);
        return PB_ERR_INVALID;
    }
    if (n == 0) return PB_OK;
    CutArgs a{d_movie, (long long)n_frames, frame_offset, Y, X, box, d_frame, d_x, d_y,
              (long long)n, baseline, sensitivity, gain, d_spots};
    const long long total = (long long)n * box * box;
    int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == PB_DTYPE_U16) cut_spots_kernel<unsigned short><<<grid, 256, 0, s>>>(a);
    else cut_spots_kernel<float><<<grid, 256, 0, s>>>(a);
    g_pb_launches++;
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

// ---- host-buffer entry points ------------------------------------------------
namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        PB_CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 1));
        return PB_OK;
    }
};
inline size_t dtype_size(int dtype) { return dtype == PB_DTYPE_U16 ? 2 : 4; }
}  // namespace

// Identify over a host movie chunk; results sorted by (frame, y, x) like the
// serial reference (localize.py:604-636).  If more than `capacity` spots are
// found, *n_found holds the required capacity and PB_ERR_CAPACITY is returned.
extern "C" int pb_identify(const void* movie, int dtype, size_t n_frames, int Y, int X,
                           long long frame_offset, int box, double min_ng, const int* roi,
                           long long* frame, long long* x, long long* y, float* ng,
                           size_t capacity, size_t* n_found) {
    if (!n_found) { pb_set_error("pb_identify: n_found is null"); return PB_ERR_INVALID; }
    *n_found = 0;
    if (n_frames == 0) return PB_OK;
    if (!movie) { pb_set_error("pb_identify: null movie"); return PB_ERR_INVALID; }
    const size_t fsz = (size_t)Y * X * dtype_size(dtype);
    // frame chunks of ~128 MB, double buffered H2D
    size_t chunk = std::max<size_t>(1, (128u << 20) / fsz);
    chunk = std::min(chunk, n_frames);
    DevBuf mv[2], dfr, dx, dy, dng, dcnt;
    int rc;
    for (int s = 0; s < 2; s++) if ((rc = mv[s].alloc(chunk * fsz))) return rc;
    const size_t cap = std::max<size_t>(capacity, 1);
    if ((rc = dfr.alloc(cap * 8)) || (rc = dx.alloc(cap * 8)) || (rc = dy.alloc(cap * 8)) ||
        (rc = dng.alloc(cap * 4)) || (rc = dcnt.alloc(8)))
        return rc;
    // streams / events are released on every path (no early return between create and destroy)
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaError_t e = cudaSuccess;
    auto ck = [&](cudaError_t err) { if (err != cudaSuccess && e == cudaSuccess) e = err; return e == cudaSuccess; };
    for (int s = 0; s < 2; s++) {
        ck(cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking));
        ck(cudaEventCreateWithFlags(&ev[s], cudaEventDisableTiming));
    }
    rc = PB_OK;
    if (e == cudaSuccess) {
        ck(cudaMemsetAsync(dcnt.p, 0, 8, st[0]));
        ck(cudaStreamSynchronize(st[0]));
    }
    size_t c = 0;
    for (size_t f0 = 0; f0 < n_frames && rc == PB_OK && e == cudaSuccess; f0 += chunk, c++) {
        const int s = (int)(c & 1);
        const size_t nf = std::min(chunk, n_frames - f0);
        if (c >= 2 && !ck(cudaEventSynchronize(ev[s]))) break;
        if ((rc = pb_h2d(mv[s].p, static_cast<const char*>(movie) + f0 * fsz, nf * fsz, st[s])) != PB_OK) break;
        rc = pb_identify_dev(mv[s].p, dtype, nf, Y, X, frame_offset + (long long)f0, box, min_ng,
                             roi, static_cast<long long*>(dfr.p), static_cast<long long*>(dx.p),
                             static_cast<long long*>(dy.p), static_cast<float*>(dng.p), capacity,
                             static_cast<unsigned long long*>(dcnt.p), st[s]);
        ck(cudaEventRecord(ev[s], st[s]));
    }
    for (int s = 0; s < 2; s++) if (st[s]) ck(cudaStreamSynchronize(st[s]));
    unsigned long long found = 0;
    if (e == cudaSuccess && rc == PB_OK) ck(cudaMemcpy(&found, dcnt.p, 8, cudaMemcpyDeviceToHost));
    for (int s = 0; s < 2; s++) {
        if (st[s]) cudaStreamDestroy(st[s]);
        if (ev[s]) cudaEventDestroy(ev[s]);
    }
    if (rc != PB_OK) return rc;
    if (e != cudaSuccess) { pb_set_error("pb_identify: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    *n_found = (size_t)found;
    if (found > capacity) {
        pb_set_error("pb_identify: found %llu spots, capacity %zu", found, capacity);
        return PB_ERR_CAPACITY;
    }
    if (found == 0) return PB_OK;
    std::vector<long long> hf(found), hx(found), hy(found);
    std::vector<float> hn(found);
    PB_CUDA_CHECK(cudaMemcpy(hf.data(), dfr.p, found * 8, cudaMemcpyDeviceToHost));
    PB_CUDA_CHECK(cudaMemcpy(hx.data(), dx.p, found * 8, cudaMemcpyDeviceToHost));
    PB_CUDA_CHECK(cudaMemcpy(hy.data(), dy.p, found * 8, cudaMemcpyDeviceToHost));
    PB_CUDA_CHECK(cudaMemcpy(hn.data(), dng.p, found * 4, cudaMemcpyDeviceToHost));
    std::vector<size_t> order(found);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](size_t p, size_t q) {
        if (hf[p] != hf[q]) return hf[p] < hf[q];
        if (hy[p] != hy[q]) return hy[p] < hy[q];
        return hx[p] < hx[q];
    });
    for (size_t k = 0; k < found; k++) {
        frame[k] = hf[order[k]]; x[k] = hx[order[k]]; y[k] = hy[order[k]]; ng[k] = hn[order[k]];
    }
    return PB_OK;
}

// get_spots over a host movie chunk: spots whose frame lies in
// [frame_offset, frame_offset + n_frames) are written, others left untouched.
extern "C" int pb_get_spots(const void* movie, int dtype, size_t n_frames, int Y, int X,
                            long long frame_offset, size_t n, const long long* frame,
                            const long long* x, const long long* y, int box, float baseline,
                            float sensitivity, float gain, float* spots) {
    if (n == 0 || n_frames == 0) return PB_OK;
    if (!movie || !frame || !x || !y || !spots) { pb_set_error("pb_get_spots: null pointer"); return PB_ERR_INVALID; }
    const size_t pix = (size_t)box * box;
    // only the spots that fall into this chunk are gathered (ids are normally frame-sorted)
    std::vector<size_t> sel;
    sel.reserve(n);
    for (size_t k = 0; k < n; k++)
        if (frame[k] >= frame_offset && frame[k] < frame_offset + (long long)n_frames) sel.push_back(k);
    if (sel.empty()) return PB_OK;
    const size_t m = sel.size();
    std::vector<long long> hf(m), hx(m), hy(m);
    for (size_t k = 0; k < m; k++) { hf[k] = frame[sel[k]]; hx[k] = x[sel[k]]; hy[k] = y[sel[k]]; }
    const size_t fsz = (size_t)Y * X * dtype_size(dtype);
    DevBuf mv, dfr, dx, dy, dsp;
    int rc;
    // upload only the frame range that is actually referenced
    long long fmin = hf[0], fmax = hf[0];
    for (size_t k = 1; k < m; k++) { fmin = std::min(fmin, hf[k]); fmax = std::max(fmax, hf[k]); }
    const size_t nf = (size_t)(fmax - fmin + 1);
    if ((rc = mv.alloc(nf * fsz)) || (rc = dfr.alloc(m * 8)) || (rc = dx.alloc(m * 8)) ||
        (rc = dy.alloc(m * 8)) || (rc = dsp.alloc(m * pix * 4)))
        return rc;
    if ((rc = pb_h2d(mv.p, static_cast<const char*>(movie) + (size_t)(fmin - frame_offset) * fsz, nf * fsz,
                     nullptr)) != PB_OK)
        return rc;
    PB_CUDA_CHECK(cudaMemcpy(dfr.p, hf.data(), m * 8, cudaMemcpyHostToDevice));
    PB_CUDA_CHECK(cudaMemcpy(dx.p, hx.data(), m * 8, cudaMemcpyHostToDevice));
    PB_CUDA_CHECK(cudaMemcpy(dy.p, hy.data(), m * 8, cudaMemcpyHostToDevice));
    rc = pb_get_spots_dev(mv.p, dtype, nf, Y, X, fmin, m, static_cast<long long*>(dfr.p),
                          static_cast<long long*>(dx.p), static_cast<long long*>(dy.p), box,
                          baseline, sensitivity, gain, static_cast<float*>(dsp.p), nullptr);
    if (rc != PB_OK) return rc;
    std::vector<float> hs(m * pix);
    if ((rc = pb_d2h(hs.data(), dsp.p, m * pix * 4, nullptr)) != PB_OK) return rc;
    for (size_t k = 0; k < m; k++)
        memcpy(spots + sel[k] * pix, hs.data() + k * pix, pix * 4);
    return PB_OK;
}

// Fused pass for the end-to-end localize path: every frame chunk is uploaded ONCE, spots
// are identified and (after the host-side (frame, y, x) sort of the few identifications)
// their ROIs are cut from the still-resident chunk.  Equivalent to pb_identify followed by
// pb_get_spots (localize.py:1787-1811 identify -> fit2D -> get_spots) without the second
// upload of the movie.  `spots` must hold capacity * box * box floats.
extern "C" int pb_identify_get_spots(const void* movie, int dtype, size_t n_frames, int Y, int X,
                                     long long frame_offset, int box, double min_ng,
                                     const int* roi, float baseline, float sensitivity, float gain,
                                     long long* frame, long long* x, long long* y, float* ng,
                                     float* spots, size_t capacity, size_t* n_found) {
    if (!n_found) { pb_set_error("pb_identify_get_spots: n_found is null"); return PB_ERR_INVALID; }
    *n_found = 0;
    if (n_frames == 0) return PB_OK;
    if (!movie) { pb_set_error("pb_identify_get_spots: null movie"); return PB_ERR_INVALID; }
    const size_t fsz = (size_t)Y * X * dtype_size(dtype);
    const size_t pix = (size_t)box * box;
    size_t chunk = std::max<size_t>(1, (128u << 20) / fsz);
    chunk = std::min(chunk, n_frames);
    // per-chunk device capacity for identifications (grown on demand)
    size_t dcap = std::max<size_t>(4096, chunk * 512);
    DevBuf mv[2], dfr, dx, dy, dng, dcnt, dsp;
    int rc;
    for (int s = 0; s < 2; s++) if ((rc = mv[s].alloc(chunk * fsz))) return rc;
    auto alloc_ids = [&](size_t cap) -> int {
        DevBuf* bufs[5] = {&dfr, &dx, &dy, &dng, &dsp};
        for (auto* b : bufs) { if (b->p) cudaFree(b->p); b->p = nullptr; }
        int r;
        if ((r = dfr.alloc(cap * 8)) || (r = dx.alloc(cap * 8)) || (r = dy.alloc(cap * 8)) ||
            (r = dng.alloc(cap * 4)) || (r = dsp.alloc(cap * pix * 4)))
            return r;
        return PB_OK;
    };
    if ((rc = alloc_ids(dcap)) || (rc = dcnt.alloc(8))) return rc;
    // Every CUDA call below is checked; the first failure is latched in (rc, e) and control falls
    // through to the common clean-up that waits for and destroys the streams -- no early return
    // leaks them, and a failed copy can never be mistaken for success through stale buffers.
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaError_t e = cudaSuccess;
    rc = PB_OK;
    auto ck = [&](cudaError_t err) { if (err != cudaSuccess && e == cudaSuccess) e = err; return e == cudaSuccess && rc == PB_OK; };
    for (int s = 0; s < 2; s++) ck(cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking));
    auto upload = [&](size_t c) {
        const size_t f0 = c * chunk;
        if (f0 >= n_frames || rc != PB_OK || e != cudaSuccess) return;
        const size_t nf = std::min(chunk, n_frames - f0);
        rc = pb_h2d(mv[c & 1].p, static_cast<const char*>(movie) + f0 * fsz, nf * fsz, st[c & 1]);
    };
    upload(0);
    size_t total = 0;
    bool overflow = false;
    std::vector<long long> hf, hx, hy;
    std::vector<float> hn;
    std::vector<size_t> order;
    const size_t nchunks = (n_frames + chunk - 1) / chunk;
    for (size_t c = 0; c < nchunks && rc == PB_OK && e == cudaSuccess; c++) {
        const int s = (int)(c & 1);
        const size_t f0 = c * chunk, nf = std::min(chunk, n_frames - f0);
        unsigned long long found = 0;
        for (;;) {   // retry loop if the per-chunk device capacity was too small
            if (!ck(cudaMemsetAsync(dcnt.p, 0, 8, st[s]))) break;
            rc = pb_identify_dev(mv[s].p, dtype, nf, Y, X, frame_offset + (long long)f0, box, min_ng,
                                 roi, static_cast<long long*>(dfr.p), static_cast<long long*>(dx.p),
                                 static_cast<long long*>(dy.p), static_cast<float*>(dng.p), dcap,
                                 static_cast<unsigned long long*>(dcnt.p), st[s]);
            if (rc != PB_OK) break;
            if (!ck(cudaMemcpyAsync(&found, dcnt.p, 8, cudaMemcpyDeviceToHost, st[s]))) break;
            if (!ck(cudaStreamSynchronize(st[s]))) break;
            if (found <= dcap) break;
            dcap = (size_t)found;
            if ((rc = alloc_ids(dcap))) break;
        }
        if (rc != PB_OK || e != cudaSuccess) break;
        upload(c + 1);   // next chunk streams in while this one is post-processed
        if (rc != PB_OK) break;
        if (total + found > capacity) overflow = true;
        if (!overflow && found) {
            hf.resize(found); hx.resize(found); hy.resize(found); hn.resize(found); order.resize(found);
            ck(cudaMemcpyAsync(hf.data(), dfr.p, found * 8, cudaMemcpyDeviceToHost, st[s]));
            ck(cudaMemcpyAsync(hx.data(), dx.p, found * 8, cudaMemcpyDeviceToHost, st[s]));
            ck(cudaMemcpyAsync(hy.data(), dy.p, found * 8, cudaMemcpyDeviceToHost, st[s]));
            ck(cudaMemcpyAsync(hn.data(), dng.p, found * 4, cudaMemcpyDeviceToHost, st[s]));
            if (!ck(cudaStreamSynchronize(st[s]))) break;
            std::iota(order.begin(), order.end(), 0);
            std::sort(order.begin(), order.end(), [&](size_t p, size_t q) {
                if (hf[p] != hf[q]) return hf[p] < hf[q];
                if (hy[p] != hy[q]) return hy[p] < hy[q];
                return hx[p] < hx[q];
            });
            for (size_t k = 0; k < found; k++) {
                frame[total + k] = hf[order[k]]; x[total + k] = hx[order[k]];
                y[total + k] = hy[order[k]]; ng[total + k] = hn[order[k]];
            }
            ck(cudaMemcpyAsync(dfr.p, frame + total, found * 8, cudaMemcpyHostToDevice, st[s]));
            ck(cudaMemcpyAsync(dx.p, x + total, found * 8, cudaMemcpyHostToDevice, st[s]));
            if (!ck(cudaMemcpyAsync(dy.p, y + total, found * 8, cudaMemcpyHostToDevice, st[s]))) break;
            rc = pb_get_spots_dev(mv[s].p, dtype, nf, Y, X, frame_offset + (long long)f0, found,
                                  static_cast<long long*>(dfr.p), static_cast<long long*>(dx.p),
                                  static_cast<long long*>(dy.p), box, baseline, sensitivity, gain,
                                  static_cast<float*>(dsp.p), st[s]);
            if (rc != PB_OK) break;
            ck(cudaMemcpyAsync(spots + total * pix, dsp.p, found * pix * 4, cudaMemcpyDeviceToHost, st[s]));
            if (!ck(cudaStreamSynchronize(st[s]))) break;
        }
        total += found;
    }
    for (int s = 0; s < 2; s++)
        if (st[s]) { ck(cudaStreamSynchronize(st[s])); cudaStreamDestroy(st[s]); }
    ck(cudaGetLastError());
    if (rc != PB_OK) return rc;
    if (e != cudaSuccess) { pb_set_error("pb_identify_get_spots: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    *n_found = total;
    if (overflow) {
        pb_set_error("pb_identify_get_spots: found %zu spots, capacity %zu", total, capacity);
        return PB_ERR_CAPACITY;
    }
    return PB_OK;
}
