// picasso_b200/csrc/rcc.cu
//
// Redundant cross-correlation (RCC) drift estimation on B200 -- replaces the FFT
// part of picasso.imageprocess.xcorr / get_image_shift / rcc
// (reference picasso/imageprocess.py:27-217) used by postprocess.undrift
// (picasso/postprocess.py:2903-2961).
//
// The reference recomputes fft2(A), fft2(B) and one ifft2 for EVERY pair in
// float64 on the CPU (3 FFTs x n(n-1)/2 pairs).  Here:
//   * one batched cuFFT R2C per SEGMENT (half-spectra kept resident in HBM),
//   * per pair, when only a small centre window of the correlation is read (the reference
//     crops `roi` x `roi` pixels after the fact, imageprocess.py:88-101): a PRUNED inverse
//     transform -- stage 1 multiplies the two half-spectra and evaluates the inverse DFT
//     along y for the window rows only (each spectrum is read once, nothing is written
//     back but H x (X/2+1) coefficients), stage 2 evaluates the real inverse DFT along x
//     for the window columns with fftshift + the 1/(N*sqrt(N)) normalisation folded in,
//   * otherwise (full-size windows): a fused conj-multiply kernel -> batched cuFFT C2R ->
//     a crop kernel,
//   * 4 KB per pair go back to the host,
//   * the arg-max + 5x5 peak fit stay on the host (picasso_b200/imageprocess.py).
// float32 transforms: the peak is fitted from a 5x5 window of O(1)-relative
// values; cuFFT's 1e-6 relative error moves the fitted shift by << 1e-3 px.
// HBM-bound (cuFFT passes over 2*Y*X*4 B per pair).
#include <algorithm>
#include <atomic>
#include <mutex>
#include <cufft.h>
#include <stdlib.h>
#include <vector>

#include "pb_common.cuh"
#include "../../include/picasso_b200.h"

extern std::atomic<long long> g_pb_launches;

namespace {

#define PB_CUFFT_CHECK(expr)                                                       \
    do {                                                                           \
        cufftResult _r = (expr);                                                   \
        if (_r != CUFFT_SUCCESS) {                                                 \
            pb_set_error("%s failed: cufft error %d (%s:%d)", #expr, (int)_r, __FILE__, __LINE__); \
            return PB_ERR_CUFFT;                                                   \
        }                                                                          \
    } while (0)

// P[b] = F[i_b] * conj(F[j_b]) over the half spectrum (imageprocess.py:45-47)
__global__ void rcc_conj_mul_kernel(const float2* __restrict__ spectra, size_t spec_elems,
                                    const int* __restrict__ pi, const int* __restrict__ pj,
                                    int npairs, float2* __restrict__ out) {
    const size_t total = (size_t)npairs * spec_elems;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / spec_elems);
        const size_t k = idx - (size_t)b * spec_elems;
        const float2 a = __ldg(spectra + (size_t)pi[b] * spec_elems + k);
        const float2 c = __ldg(spectra + (size_t)pj[b] * spec_elems + k);
        // a * conj(c)
        out[idx] = make_float2(fmaf(a.x, c.x, a.y * c.y), fmaf(a.y, c.x, -a.x * c.y));
    }
}

// crop[b][r][c] = corr[b][(r + Y0 - Y/2) mod Y][(c + X0 - X/2) mod X] * scale
// (np.fft.fftshift rolls by N//2; scale = 1/(Y*X) from ifft2 and 1/sqrt(Y*X))
__global__ void rcc_crop_kernel(const float* __restrict__ corr, int Y, int X, int npairs, int Y0,
                                int X0, int H, int W, double scale, float* __restrict__ out) {
    const size_t total = (size_t)npairs * H * W;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % W);
        const int r = (int)((idx / W) % H);
        const int b = (int)(idx / ((size_t)W * H));
        int sr = (r + Y0 - Y / 2) % Y; if (sr < 0) sr += Y;
        int sc = (c + X0 - X / 2) % X; if (sc < 0) sc += X;
        out[idx] = (float)((double)corr[((size_t)b * Y + sr) * X + sc] * scale);
    }
}

// per-image sum in f64 (the reference tests np.sum(image) == 0, imageprocess.py:83)
__global__ void rcc_sum_kernel(const float* __restrict__ img, size_t elems, double* out) {
    const float* p = img + (size_t)blockIdx.y * elems;
    double s = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < elems;
         i += (size_t)gridDim.x * blockDim.x)
        s += (double)p[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) atomicAdd(out + blockIdx.y, s);
    }
}

// ---- pruned inverse transform (window rows / columns only) -----------------------------
// corr[y][x] = sum_ky sum_kx w_kx Re(P[ky][kx] e^{2 pi i (ky y / Y + kx x / X)}) over the half
// spectrum (w = 1 for kx = 0 and the Nyquist column, else 2), P = F_i conj(F_j).
//
// Stage 1: T[pair][r][kx] = sum_ky P[ky][kx] e^{2 pi i ky y_r / Y} for the window rows y_r.
// Block = 64 kx columns x 4 ky groups (each group owns a quarter of the rows; partial sums are
// combined through shared memory at the end -- also keeps the float32 sums short).  A warp is
// 32 adjacent columns of one group: 256-byte coalesced spectrum loads, twiddles broadcast from
// shared memory, 32 complex accumulators per thread in registers (128 FFMA per spectrum pair).
constexpr int kPR = 32;        // window rows per pass
constexpr int kKC = 16;        // ky per twiddle chunk
constexpr int kS1Cols = 64;
constexpr int kS1Groups = 4;

__global__ void __launch_bounds__(kS1Cols * kS1Groups, 2)
rcc_pruned_rows_kernel(const float2* __restrict__ spectra, size_t spec_elems,
                       const int* __restrict__ pi, const int* __restrict__ pj, int Y, int XH,
                       int ywin0, int row0, int nrows, int H, float2* __restrict__ T) {
    extern __shared__ __align__(16) unsigned char rcc_smem[];
    float2* tw = reinterpret_cast<float2*>(rcc_smem);            // [groups][kKC][kPR]
    float2* red = reinterpret_cast<float2*>(rcc_smem);           // [2][kPR][kS1Cols] (aliases tw)
    const int tx = threadIdx.x & (kS1Cols - 1);
    const int g = threadIdx.x / kS1Cols;
    const int kx = blockIdx.x * kS1Cols + tx;
    const bool col_ok = kx < XH;
    const int pair = blockIdx.y;
    const float2* A = spectra + (size_t)pi[pair] * spec_elems;
    const float2* B = spectra + (size_t)pj[pair] * spec_elems;
    const int per = (Y + kS1Groups - 1) / kS1Groups;
    const int k0 = g * per;
    const int k1 = k0 + per < Y ? k0 + per : Y;
    float accr[kPR], acci[kPR];
#pragma unroll
    for (int r = 0; r < kPR; r++) { accr[r] = 0.f; acci[r] = 0.f; }
    const float invY = 1.0f / (float)Y;
    for (int kb = 0; kb < per; kb += kKC) {
        __syncthreads();
        // twiddles of this chunk: e^{2 pi i ky y_r / Y}, argument reduced exactly in integers
        for (int e = tx; e < kKC * kPR; e += kS1Cols) {
            const int kc = e / kPR, r = e % kPR;
            const long long ky = k0 + kb + kc;
            const long long yy = (ywin0 + row0 + r) % Y;
            const int m = (int)((ky * yy) % Y);
            float sn, cs;
            sincospif(2.0f * (float)m * invY, &sn, &cs);
            tw[(g * kKC + kc) * kPR + r] = make_float2(cs, sn);
        }
        __syncthreads();
        if (col_ok) {
            const float2* twg = tw + g * kKC * kPR;
#pragma unroll 4
            for (int kc = 0; kc < kKC; kc++) {
                const int ky = k0 + kb + kc;
                if (ky < k1) {
                    const float2 a = __ldg(A + (size_t)ky * XH + kx);
                    const float2 b = __ldg(B + (size_t)ky * XH + kx);
                    const float px = fmaf(a.x, b.x, a.y * b.y);      // a * conj(b)
                    const float py = fmaf(a.y, b.x, -a.x * b.y);
                    const float4* t4 = reinterpret_cast<const float4*>(twg + kc * kPR);
#pragma unroll
                    for (int r = 0; r < kPR; r += 2) {
                        const float4 t = t4[r >> 1];               // (cos, sin) of rows r, r+1
                        accr[r] = fmaf(px, t.x, fmaf(-py, t.y, accr[r]));
                        acci[r] = fmaf(px, t.y, fmaf(py, t.x, acci[r]));
                        accr[r + 1] = fmaf(px, t.z, fmaf(-py, t.w, accr[r + 1]));
                        acci[r + 1] = fmaf(px, t.w, fmaf(py, t.z, acci[r + 1]));
                    }
                }
            }
        }
    }
    // combine the four ky groups: (2,3) -> scratch, (0,1) add; 1 -> scratch, 0 adds and stores
    __syncthreads();
    if (g >= 2) {
#pragma unroll
        for (int r = 0; r < kPR; r++)
            red[((g - 2) * kPR + r) * kS1Cols + tx] = make_float2(accr[r], acci[r]);
    }
    __syncthreads();
    if (g < 2) {
#pragma unroll
        for (int r = 0; r < kPR; r++) {
            const float2 v = red[(g * kPR + r) * kS1Cols + tx];
            accr[r] += v.x;
            acci[r] += v.y;
        }
    }
    __syncthreads();
    if (g == 1) {
#pragma unroll
        for (int r = 0; r < kPR; r++) red[r * kS1Cols + tx] = make_float2(accr[r], acci[r]);
    }
    __syncthreads();
    if (g == 0 && col_ok) {
        float2* out = T + ((size_t)pair * H + row0) * XH + kx;
#pragma unroll
        for (int r = 0; r < kPR; r++) {
            if (r < nrows) {
                const float2 v = red[r * kS1Cols + tx];
                out[(size_t)r * XH] = make_float2(accr[r] + v.x, acci[r] + v.y);
            }
        }
    }
}


// ---- FFT-structured stage 1 (Y % 32 == 0) --------------------------------------------------
// The 32 window rows are contiguous mod Y, so they cover every residue mod 32 exactly once.
// With ky = q + M p (M = Y / 32, q < M, p < 32):
//     T[y] = sum_q e^{2 pi i q y / Y} * S_q[y mod 32],   S_q[c] = sum_p P[q + M p] e^{2 pi i p c / 32}
// i.e. one 32-point inverse FFT per q (in registers, radix 2, decimation in frequency -- all
// indices and twiddles are compile-time constants after unrolling) followed by one complex
// multiply-add per window row: ~800 instructions per (column, q) instead of ~4800 for the
// direct sum.  A warp is 32 adjacent kx columns (256-byte coalesced loads of the rows q + M p);
// the 4 warps of a CTA split the q range and are combined through shared memory.
// `tw` holds e^{2 pi i q y(pos) / Y} for the row y(pos) whose residue ends up at FFT output
// position pos (bit-reversed), [M][32] float2, prepared by rcc_fft_twiddle_kernel.
constexpr int kFW = 4;   // warps per CTA

__host__ __device__ constexpr int rcc_bitrev5(int v) {
    return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
}
__device__ __forceinline__ constexpr float rcc_c32(int t) {   // cos(2 pi t / 32)
    switch (t) {
        case 0: return 1.0f;
        case 1: return 0.98078528040323044913f;
        case 2: return 0.92387953251128675613f;
        case 3: return 0.83146961230254523708f;
        case 4: return 0.70710678118654752440f;
        case 5: return 0.55557023301960222474f;
        case 6: return 0.38268343236508977173f;
        case 7: return 0.19509032201612826785f;
        case 8: return 0.0f;
        case 9: return -0.19509032201612826785f;
        case 10: return -0.38268343236508977173f;
        case 11: return -0.55557023301960222474f;
        case 12: return -0.70710678118654752440f;
        case 13: return -0.83146961230254523708f;
        case 14: return -0.92387953251128675613f;
        default: return -0.98078528040323044913f;
    }
}
__device__ __forceinline__ constexpr float rcc_s32(int t) {   // sin(2 pi t / 32), t < 16
    return t <= 8 ? rcc_c32(8 - t) : rcc_c32(t - 8);
}

// in-place 32-point inverse DFT (no scaling); position n holds X[bitrev5(n)] afterwards
__device__ __forceinline__ void rcc_fft32_inv(float (&xr)[32], float (&xi)[32]) {
#pragma unroll
    for (int sp = 16; sp >= 1; sp >>= 1) {
#pragma unroll
        for (int g = 0; g < 32; g += 2 * sp) {
#pragma unroll
            for (int k = 0; k < sp; k++) {
                const int i = g + k, j = i + sp;
                const int t = k * (16 / sp);           // twiddle e^{+2 pi i t / 32}
                const float ar = xr[i], ai = xi[i], br = xr[j], bi = xi[j];
                xr[i] = ar + br;
                xi[i] = ai + bi;
                const float tr = ar - br, ti = ai - bi;
                if (t == 0) { xr[j] = tr; xi[j] = ti; }
                else if (t == 8) { xr[j] = -ti; xi[j] = tr; }
                else {
                    const float c = rcc_c32(t), sn = rcc_s32(t);
                    xr[j] = fmaf(tr, c, -ti * sn);
                    xi[j] = fmaf(tr, sn, ti * c);
                }
            }
        }
    }
}

__global__ void rcc_fft_twiddle_kernel(int Y, int M, int y_first, float2* __restrict__ tw) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M * 32) return;
    const int q = e >> 5, pos = e & 31;
    const int c0 = y_first & 31;
    const int r = (rcc_bitrev5(pos) - c0) & 31;           // window row with this residue
    const long long y = (y_first + r) % Y;
    const long long m = ((long long)q * y) % Y;
    double sn, cs;
    sincospi(2.0 * (double)m / (double)Y, &sn, &cs);
    tw[e] = make_float2((float)cs, (float)sn);
}

template <int MINB>
__global__ void __launch_bounds__(32 * kFW, MINB)
rcc_fft_rows_kernel(const float2* __restrict__ spectra, size_t spec_elems, const int* __restrict__ pi,
                    const int* __restrict__ pj, int XH, int M, const float2* __restrict__ tw_g,
                    int c0, int row0, int nrows, int H, float2* __restrict__ T) {
    extern __shared__ __align__(16) unsigned char rcc_smem[];
    float2* tw = reinterpret_cast<float2*>(rcc_smem);            // [M][32]
    float2* red = reinterpret_cast<float2*>(rcc_smem);           // [kFW][32][32] (aliases tw)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int pair = blockIdx.x;
    const int kx_raw = blockIdx.y * 32 + lane;
    const int kx = kx_raw < XH ? kx_raw : XH - 1;
    {   // stage the twiddle table (same for every CTA; L2 resident)
        const float4* src = reinterpret_cast<const float4*>(tw_g);
        float4* dst = reinterpret_cast<float4*>(tw);
        for (int e = threadIdx.x; e < M * 16; e += 32 * kFW) dst[e] = __ldg(src + e);
    }
    __syncthreads();
    const float2* A = spectra + (size_t)pi[pair] * spec_elems + kx;
    const float2* B = spectra + (size_t)pj[pair] * spec_elems + kx;
    const size_t pstride = (size_t)M * XH;
    float accr[32], acci[32];
#pragma unroll
    for (int n = 0; n < 32; n++) { accr[n] = 0.f; acci[n] = 0.f; }
    for (int q = w; q < M; q += kFW) {
        float xr[32], xi[32];
        const float2* a = A + (size_t)q * XH;
        const float2* b = B + (size_t)q * XH;
#pragma unroll
        for (int p = 0; p < 32; p++) {
            const float2 u = __ldg(a + p * pstride);
            const float2 v = __ldg(b + p * pstride);
            xr[p] = fmaf(u.x, v.x, u.y * v.y);       // u * conj(v)
            xi[p] = fmaf(u.y, v.x, -u.x * v.y);
        }
        rcc_fft32_inv(xr, xi);
        const float4* t4 = reinterpret_cast<const float4*>(tw + q * 32);
#pragma unroll
        for (int n = 0; n < 32; n += 2) {
            const float4 t = t4[n >> 1];
            accr[n] = fmaf(xr[n], t.x, fmaf(-xi[n], t.y, accr[n]));
            acci[n] = fmaf(xr[n], t.y, fmaf(xi[n], t.x, acci[n]));
            accr[n + 1] = fmaf(xr[n + 1], t.z, fmaf(-xi[n + 1], t.w, accr[n + 1]));
            acci[n + 1] = fmaf(xr[n + 1], t.w, fmaf(xi[n + 1], t.z, acci[n + 1]));
        }
    }
    __syncthreads();     // twiddles no longer needed: reuse the buffer for the reduction
#pragma unroll
    for (int n = 0; n < 32; n++) red[(w * 32 + n) * 32 + lane] = make_float2(accr[n], acci[n]);
    __syncthreads();
    if (kx_raw < XH) {
#pragma unroll
        for (int k = 0; k < 32 / kFW; k++) {
            const int pos = w * (32 / kFW) + k;
            float sr = 0.f, si = 0.f;
#pragma unroll
            for (int u = 0; u < kFW; u++) {
                const float2 v = red[(u * 32 + pos) * 32 + lane];
                sr += v.x; si += v.y;
            }
            const int r = (rcc_bitrev5(pos) - c0) & 31;
            if (r < nrows) T[((size_t)pair * H + row0 + r) * XH + kx] = make_float2(sr, si);
        }
    }
}

// Stage 2: out[pair][r][c] = scale * sum_kx w_kx Re(T[pair][r][kx] e^{2 pi i kx x_c / X}).
// One block per (window row, pair): 32 columns x 8 kx slices, float64 accumulation.
__global__ void __launch_bounds__(256)
rcc_pruned_cols_kernel(const float2* __restrict__ T, int H, int W, int XH, int X, int xwin0,
                       double scale, float* __restrict__ out, const int* __restrict__ out_idx) {
    __shared__ double part[8][33];
    const int row = blockIdx.x, pair = blockIdx.y;
    const int c = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const float2* Trow = T + ((size_t)pair * H + row) * XH;
    const size_t opair = out_idx ? (size_t)out_idx[pair] : (size_t)pair;
    const float invX = 1.0f / (float)X;
    for (int c0 = 0; c0 < W; c0 += 32) {
        const int cc = c0 + c;
        const long long xx = (xwin0 + (cc < W ? cc : 0)) % X;
        double sum = 0.0;
        for (int kx = sl; kx < XH; kx += 8) {
            const float2 t = __ldg(Trow + kx);
            const int m = (int)(((long long)kx * xx) % X);
            float sn, cs;
            sincospif(2.0f * (float)m * invX, &sn, &cs);
            const float wgt = (kx == 0 || ((X & 1) == 0 && kx == X / 2)) ? 1.0f : 2.0f;
            sum += (double)(wgt * (t.x * cs - t.y * sn));
        }
        part[sl][c] = sum;
        __syncthreads();
        if (sl == 0 && cc < W) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < 8; q++) s += part[q][c];
            out[(opair * H + row) * W + cc] = (float)(s * scale);
        }
        __syncthreads();
    }
}



// Variant with asynchronous staging (LDGSTS): the 64 spectrum rows of the NEXT q stream into
// the warp's 16 KB shared-memory stage with 8-byte cp.async copies (half-spectrum rows are only
// 8-byte aligned: X/2+1 is odd) while the warp runs the FFT of the current q out of registers.
// No raw-load registers (128 in the register variant) -> 168 registers, 12 warps per SM, and
// the load latency is hidden behind ~800 instructions of arithmetic per warp.  Twiddles come
// through L1 (warp-uniform 16-byte loads).
__device__ __forceinline__ void rcc_cp_async8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(pb_smem_u32(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void rcc_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__global__ void __launch_bounds__(32 * kFW, 3)
rcc_fft_rows_async_kernel(const float2* __restrict__ spectra, size_t spec_elems, const int* __restrict__ pi,
                          const int* __restrict__ pj, int XH, int M, const float2* __restrict__ tw_g,
                          int c0, int row0, int nrows, int H, float2* __restrict__ T) {
    extern __shared__ __align__(16) unsigned char rcc_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float2* stage = reinterpret_cast<float2*>(rcc_smem) + (size_t)w * 64 * 32;   // [64 rows][32 lanes]
    float2* red = reinterpret_cast<float2*>(rcc_smem);                            // [kFW][32][32] after the loop
    const int pair = blockIdx.x;
    const int kx_raw = blockIdx.y * 32 + lane;
    const int kx = kx_raw < XH ? kx_raw : XH - 1;
    const float2* A = spectra + (size_t)pi[pair] * spec_elems + kx;
    const float2* B = spectra + (size_t)pj[pair] * spec_elems + kx;
    const size_t pstride = (size_t)M * XH;
    auto prefetch = [&](int q) {
        const float2* a = A + (size_t)q * XH;
        const float2* b = B + (size_t)q * XH;
#pragma unroll
        for (int p = 0; p < 32; p++) {
            rcc_cp_async8(stage + p * 32 + lane, a + p * pstride);
            rcc_cp_async8(stage + (32 + p) * 32 + lane, b + p * pstride);
        }
    };
    float accr[32], acci[32];
#pragma unroll
    for (int n = 0; n < 32; n++) { accr[n] = 0.f; acci[n] = 0.f; }
    if (w < M) prefetch(w);
    for (int q = w; q < M; q += kFW) {
        float xr[32], xi[32];
        rcc_cp_async_wait_all();
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 32; p++) {
            const float2 u = stage[p * 32 + lane];
            const float2 v = stage[(32 + p) * 32 + lane];
            xr[p] = fmaf(u.x, v.x, u.y * v.y);       // u * conj(v)
            xi[p] = fmaf(u.y, v.x, -u.x * v.y);
        }
        __syncwarp();
        if (q + kFW < M) prefetch(q + kFW);
        rcc_fft32_inv(xr, xi);
        const float4* t4 = reinterpret_cast<const float4*>(tw_g + q * 32);
#pragma unroll
        for (int n = 0; n < 32; n += 2) {
            const float4 t = __ldg(t4 + (n >> 1));
            accr[n] = fmaf(xr[n], t.x, fmaf(-xi[n], t.y, accr[n]));
            acci[n] = fmaf(xr[n], t.y, fmaf(xi[n], t.x, acci[n]));
            accr[n + 1] = fmaf(xr[n + 1], t.z, fmaf(-xi[n + 1], t.w, accr[n + 1]));
            acci[n + 1] = fmaf(xr[n + 1], t.w, fmaf(xi[n + 1], t.z, acci[n + 1]));
        }
    }
    __syncthreads();     // every warp is done with its stage: reuse the buffer for the reduction
#pragma unroll
    for (int n = 0; n < 32; n++) red[(w * 32 + n) * 32 + lane] = make_float2(accr[n], acci[n]);
    __syncthreads();
    if (kx_raw < XH) {
#pragma unroll
        for (int k = 0; k < 32 / kFW; k++) {
            const int pos = w * (32 / kFW) + k;
            float sr = 0.f, si = 0.f;
#pragma unroll
            for (int u = 0; u < kFW; u++) {
                const float2 v = red[(u * 32 + pos) * 32 + lane];
                sr += v.x; si += v.y;
            }
            const int r = (rcc_bitrev5(pos) - c0) & 31;
            if (r < nrows) T[((size_t)pair * H + row0 + r) * XH + kx] = make_float2(sr, si);
        }
    }
}

// Stage 2, table driven: out[pair][r][c] = scale * sum_kx Re(T[pair][r][kx] * E[kx][c]) with
// E[kx][c] = w_kx e^{2 pi i kx x_c / X} tabulated once per call (float64 sincospi, stored as
// (w cos, -w sin) so that the real part is two FMAs).  One CTA = 32 window rows x 32 window
// columns of one pair: warp = 4 rows (T broadcast from shared memory), lane = column.
constexpr int kC2K = 64;   // kx per shared-memory chunk

__global__ void rcc_cols_table_kernel(int X, int XH, int W, int xwin0, float2* __restrict__ E) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= XH * W) return;
    const int kx = e / W, c = e - kx * W;
    const long long xx = (xwin0 + c) % X;
    const long long m = ((long long)kx * xx) % X;
    double sn, cs;
    sincospi(2.0 * (double)m / (double)X, &sn, &cs);
    const double wgt = (kx == 0 || ((X & 1) == 0 && kx == X / 2)) ? 1.0 : 2.0;
    E[e] = make_float2((float)(wgt * cs), (float)(-wgt * sn));
}

__global__ void __launch_bounds__(256)
rcc_cols_gemm_kernel(const float2* __restrict__ T, int H, int W, int XH, const float2* __restrict__ E,
                     double scale, float* __restrict__ out, const int* __restrict__ out_idx) {
    __shared__ float2 Ts[32][kC2K + 1];
    __shared__ float2 Es[kC2K][32];
    const int pair = blockIdx.y;
    const int row0 = blockIdx.x * 32, col0 = blockIdx.z * 32;
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const float2* Tp = T + (size_t)pair * H * XH;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k0 = 0; k0 < XH; k0 += kC2K) {
        __syncthreads();
        for (int e = threadIdx.x; e < 32 * kC2K; e += 256) {
            const int r = e / kC2K, k = e - r * kC2K;
            const bool ok = (row0 + r < H) && (k0 + k < XH);
            Ts[r][k] = ok ? __ldg(Tp + (size_t)(row0 + r) * XH + k0 + k) : make_float2(0.f, 0.f);
        }
        for (int e = threadIdx.x; e < kC2K * 32; e += 256) {
            const int k = e >> 5, c = e & 31;
            const bool ok = (k0 + k < XH) && (col0 + c < W);
            Es[k][c] = ok ? __ldg(E + (size_t)(k0 + k) * W + col0 + c) : make_float2(0.f, 0.f);
        }
        __syncthreads();
        float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
        for (int k = 0; k < kC2K; k++) {
            const float2 ev = Es[k][lane];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float2 t = Ts[wp * 4 + u][k];
                part[u] = fmaf(t.x, ev.x, fmaf(t.y, ev.y, part[u]));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc[u] += (double)part[u];
    }
    const size_t opair = out_idx ? (size_t)out_idx[pair] : (size_t)pair;
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int r = row0 + wp * 4 + u, c = col0 + lane;
        if (r < H && c < W) out[(opair * H + r) * W + c] = (float)(acc[u] * scale);
    }
}


// ---- device-side peak fit of the correlation windows ---------------------------------------
// One thread per pair: arg-max of the H x W window (first maximum in row-major order, like
// np.argmax), the 5 x 5 cut-out around it (imageprocess.py:103-116) and a Levenberg-Marquardt fit
// of a * exp(-((x-xc)^2 + (y-yc)^2) / (2 s^2)) + b with the reference's start values
// [max, 0, 0, 1, min] and bounds a, s, b >= 0 (imageprocess.py:119-135) -- the same iteration as
// the host's vectorised solver (picasso_b200/imageprocess.py::_gauss_peak_fit_batch).  Record per
// pair (32 doubles): [0] status, [1] y_max, [2] x_max, [3] xc, [4] yc, [5..29] the 5 x 5 window.
//   status 0: fitted (interior optimum)      1: not settled -> the host re-fits the 5 x 5 window
//          2: cut-out empty / not square -> shift (0, 0)    3: cut-out square but not 5 x 5 -> host
constexpr int kRecDoubles = 32;

__device__ __forceinline__ bool rcc_solve5(double A[5][5], double b[5]) {
    for (int c = 0; c < 5; c++) {
        int piv = c;
        double best = fabs(A[c][c]);
        for (int r = c + 1; r < 5; r++) if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
        if (!(best > 0.0)) return false;
        if (piv != c) {
            for (int k = 0; k < 5; k++) { const double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
            const double t = b[c]; b[c] = b[piv]; b[piv] = t;
        }
        const double inv = 1.0 / A[c][c];
        for (int r = c + 1; r < 5; r++) {
            const double f = A[r][c] * inv;
            for (int k = c; k < 5; k++) A[r][k] -= f * A[c][k];
            b[r] -= f * b[c];
        }
    }
    for (int c = 4; c >= 0; c--) {
        double v = b[c];
        for (int k = c + 1; k < 5; k++) v -= A[c][k] * b[k];
        b[c] = v / A[c][c];
    }
    return true;
}

__device__ void rcc_model5(const double p[5], const double* d, double res[25], double J[25][5], double* cost) {
    const double a = p[0], xc = p[1], yc = p[2], sg = p[3], bb = p[4];
    const double s2 = sg * sg, s3 = s2 * sg;
    double c = 0.0;
    for (int i = 0; i < 25; i++) {
        const double dx = (double)(i % 5 - 2) - xc, dy = (double)(i / 5 - 2) - yc;
        const double r2 = dx * dx + dy * dy;
        const double E = exp(-0.5 * r2 / s2);
        const double m = a * E + bb;
        J[i][0] = E; J[i][1] = a * E * dx / s2; J[i][2] = a * E * dy / s2; J[i][3] = a * E * r2 / s3; J[i][4] = 1.0;
        res[i] = d[i] - m;
        c += res[i] * res[i];
    }
    *cost = c;
}

__global__ void __launch_bounds__(64)
rcc_peakfit_kernel(const float* __restrict__ windows, int n_pairs, int H, int W, double* __restrict__ rec) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= n_pairs) return;
    const float* w = windows + (size_t)pair * H * W;
    double* out = rec + (size_t)pair * kRecDoubles;
    int am = 0;
    float best = w[0];
    for (int i = 1; i < H * W; i++) {
        const float v = w[i];
        if (v > best || (best != best && v == v)) { best = v; am = i; }   // first max; NaN handling as np.argmax
    }
    if (w[0] != w[0]) am = 0;                                              // np.argmax returns the first NaN
    const int ym = am / W, xm = am % W;
    out[1] = ym; out[2] = xm; out[3] = 0.0; out[4] = 0.0;
    // python slice [ym-2 : ym+3]: a negative start wraps around (-> empty for ym < 2)
    auto span = [](int c, int n) { const int lo = c - 2, hi = min(c + 3, n); return lo < 0 ? 0 : max(hi - lo, 0); };
    const int ny = span(ym, H), nx = span(xm, W);
    if (ny == 0 || nx == 0 || ny != nx) { out[0] = 2.0; return; }
    if (ny != 5) { out[0] = 3.0; return; }
    double d[25];
    double dmax = -INFINITY, dmin = INFINITY;
    for (int i = 0; i < 25; i++) {
        d[i] = (double)w[(size_t)(ym - 2 + i / 5) * W + (xm - 2 + i % 5)];
        out[5 + i] = d[i];
        dmax = fmax(dmax, d[i]); dmin = fmin(dmin, d[i]);
    }
    double p[5] = {dmax, 0.0, 0.0, 1.0, dmin};
    double lam = 1e-3, cost, res[25], J[25][5];
    rcc_model5(p, d, res, J, &cost);
    bool converged = false, active = true;
    for (int it = 0; it < 60 && active; it++) {
        double A[5][5], g[5];
        for (int j = 0; j < 5; j++) {
            g[j] = 0.0;
            for (int k = 0; k < 5; k++) A[j][k] = 0.0;
        }
        for (int i = 0; i < 25; i++)
            for (int j = 0; j < 5; j++) {
                g[j] += J[i][j] * res[i];
                for (int k = 0; k < 5; k++) A[j][k] += J[i][j] * J[i][k];
            }
        for (int j = 0; j < 5; j++) A[j][j] += lam * fmax(A[j][j], 1e-300);
        if (!rcc_solve5(A, g)) { active = false; break; }
        double t[5], smax = 0.0, pmax = 0.0;
        for (int j = 0; j < 5; j++) { t[j] = p[j] + g[j]; smax = fmax(smax, fabs(g[j])); pmax = fmax(pmax, fabs(p[j])); }
        t[0] = fmax(t[0], 0.0); t[3] = fmax(t[3], 1e-12); t[4] = fmax(t[4], 0.0);
        double ct, rt[25], Jt[25][5];
        rcc_model5(t, d, rt, Jt, &ct);
        if (isfinite(ct) && ct <= cost) {
            const bool small = smax < 1e-10 * (1.0 + pmax);
            const bool flat = (cost - ct) <= 1e-14 * (cost + 1e-300);
            for (int j = 0; j < 5; j++) p[j] = t[j];
            for (int i = 0; i < 25; i++) { res[i] = rt[i]; for (int j = 0; j < 5; j++) J[i][j] = Jt[i][j]; }
            cost = ct;
            lam = fmax(lam * 0.3, 1e-12);
            if (small || flat) { converged = true; active = false; }
        } else {
            lam *= 10.0;
            if (lam > 1e12) active = false;
        }
    }
    bool fin = true;
    for (int j = 0; j < 5; j++) fin = fin && isfinite(p[j]);
    const bool ok = converged && fin && p[0] > 0.0 && p[4] > 0.0 && p[3] > 1e-6;
    out[0] = ok ? 0.0 : 1.0;
    out[3] = p[1]; out[4] = p[2];
}

struct CufftPlan {
    cufftHandle h = 0;
    bool ok = false;
    ~CufftPlan() { if (ok) cufftDestroy(h); }
};

// Small LRU cache of cuFFT plans (creating a batched 4096^2 plan costs tens of milliseconds -- more than
// the transforms of a rank's share of the segments on 8 GPUs).  Forward transforms run in groups of at
// most kR2CBatch segments, so one plan (and its bounded work area) serves any number of segments.
constexpr int kR2CBatch = 16;
struct PlanEntry { int dev, Y, X, batch, type; cufftHandle h; unsigned long long used; };
std::mutex g_plan_mutex;
std::vector<PlanEntry> g_plans;
unsigned long long g_plan_clock = 0;
int cached_plan(int Y, int X, int batch, cufftType type, cufftHandle* out) {
    int dev = 0;
    PB_CUDA_CHECK(cudaGetDevice(&dev));
    for (auto& e : g_plans)
        if (e.dev == dev && e.Y == Y && e.X == X && e.batch == batch && e.type == (int)type) {
            e.used = ++g_plan_clock;
            *out = e.h;
            return PB_OK;
        }
    if (g_plans.size() >= 8) {
        size_t old = 0;
        for (size_t k = 1; k < g_plans.size(); k++) if (g_plans[k].used < g_plans[old].used) old = k;
        cufftDestroy(g_plans[old].h);
        g_plans.erase(g_plans.begin() + old);
    }
    cufftHandle h;
    int dims[2] = {Y, X};
    PB_CUFFT_CHECK(cufftPlanMany(&h, 2, dims, nullptr, 1, 0, nullptr, 1, 0, type, batch));
    g_plans.push_back({dev, Y, X, batch, (int)type, h, ++g_plan_clock});
    *out = h;
    return PB_OK;
}

// -1 auto (pruned when the window covers at most a quarter of the image), 0 cuFFT, 1 pruned
std::atomic<int> g_rcc_mode{-2};
int rcc_mode() {
    int v = g_rcc_mode.load();
    if (v == -2) {
        v = -1;
        if (const char* e = getenv("PB_RCC_PRUNED")) v = atoi(e) ? 1 : 0;
        g_rcc_mode.store(v);
    }
    return v;
}

// PB_RCC_FFT=0 disables the FFT-structured stage 1 (A/B measurement); PB_RCC_TILE_MB sets the
// L2 budget of one pair tile (default 40 MB).
bool rcc_fft_enabled() {
    const char* e = getenv("PB_RCC_FFT");
    return !(e && atoi(e) == 0);
}
int rcc_tile_mb() {
    if (const char* e = getenv("PB_RCC_TILE_MB")) { const int v = atoi(e); if (v >= 1) return v; }
    return 40;
}

}  // namespace

// Segments per pair tile: 2 * TS spectrum slabs of 32 columns x Y rows stay L2 resident while the
// TS^2 pairs of a tile read them.  Multi-GPU callers hand each rank whole tiles (contiguous ranges
// of the pair list sorted by (i / TS, j / TS)) so the L2 reuse survives the partition.
extern "C" int pb_rcc_tile_segments(int Y) {
    const size_t slab = (size_t)32 * (size_t)std::max(Y, 1) * sizeof(float2);
    return (int)std::max<size_t>(2, std::min<size_t>(64, ((size_t)rcc_tile_mb() << 20) / (2 * slab)));
}

// arg-max + 5 x 5 cut-out + Gaussian peak fit of n_pairs correlation windows (H x W float32 each)
// on the device: 32 doubles per pair (status, arg-max y / x, xc, yc, the 5 x 5 window).
extern "C" int pb_rcc_peakfit_dev(int n_pairs, const float* d_windows, int H, int W, double* d_records,
                                  void* stream) {
    if (n_pairs <= 0) return PB_OK;
    if (!d_windows || !d_records || H < 1 || W < 1) { pb_set_error("pb_rcc_peakfit_dev: bad argument"); return PB_ERR_INVALID; }
    rcc_peakfit_kernel<<<(n_pairs + 63) / 64, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_windows, n_pairs, H, W, d_records);
    g_pb_launches++;
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_rcc_set_mode(int mode) {
    if (mode < -1 || mode > 1) { pb_set_error("pb_rcc_set_mode: mode must be -1, 0 or 1"); return PB_ERR_INVALID; }
    g_rcc_mode.store(mode);
    return PB_OK;
}

// Forward transforms of all segments: d_spectra[s] = rfft2(d_segments[s]) (unnormalised),
// plus per-segment sums.  d_spectra holds n_seg * Y * (X/2+1) float2.
extern "C" int pb_rcc_spectra_dev(int n_seg, int Y, int X, const float* d_segments,
                                  void* d_spectra, double* d_sums, void* stream) {
    if (n_seg <= 0) return PB_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    {
        std::lock_guard<std::mutex> lk(g_plan_mutex);
        const size_t img = (size_t)Y * X, spec = (size_t)Y * (X / 2 + 1);
        for (int s0 = 0; s0 < n_seg;) {
            const int nb = std::min(kR2CBatch, n_seg - s0);
            cufftHandle h;
            int rc = cached_plan(Y, X, nb, CUFFT_R2C, &h);
            if (rc != PB_OK) return rc;
            PB_CUFFT_CHECK(cufftSetStream(h, s));
            PB_CUFFT_CHECK(cufftExecR2C(h, const_cast<float*>(d_segments) + (size_t)s0 * img,
                                        static_cast<cufftComplex*>(d_spectra) + (size_t)s0 * spec));
            s0 += nb;
        }
    }
    g_pb_launches++;
    if (d_sums) {
        PB_CUDA_CHECK(cudaMemsetAsync(d_sums, 0, sizeof(double) * n_seg, s));
        dim3 grid(64, n_seg);
        rcc_sum_kernel<<<grid, 256, 0, s>>>(d_segments, (size_t)Y * X, d_sums);
        g_pb_launches++;
    }
    PB_CUDA_CHECK(cudaGetLastError());
    PB_CUDA_CHECK(cudaStreamSynchronize(s));
    return PB_OK;
}

// Correlation windows for a list of pairs: d_windows[p] = crop of
// fftshift(irfft2(F_i * conj(F_j))) / sqrt(Y*X), H x W window starting at (Y0, X0).
// Pairs are processed in batches of `batch` (workspace: batch * (spec + Y*X) floats).
extern "C" int pb_rcc_windows_dev(int n_pairs, const int* d_pair_i, const int* d_pair_j, int Y,
                                  int X, const void* d_spectra, int Y0, int X0, int H, int W,
                                  float* d_windows, int batch, void* d_workspace,
                                  size_t workspace_bytes, void* stream) {
    if (n_pairs <= 0) return PB_OK;
    if (batch < 1) batch = 1;
    const size_t spec = (size_t)Y * (X / 2 + 1);
    const size_t need = (size_t)batch * (spec * 8 + (size_t)Y * X * 4);
    if (!d_workspace || workspace_bytes < need) {
        pb_set_error("pb_rcc_windows_dev: workspace too small (%zu < %zu)", workspace_bytes, need);
        return PB_ERR_INVALID;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int mode = rcc_mode();
    if (mode == 1 || (mode == -1 && (size_t)H * W * 4 <= (size_t)Y * X)) {
        // pruned inverse transform: only the H x W window is evaluated
        const int XH = X / 2 + 1;
        const size_t t_per_pair = (size_t)H * XH * sizeof(float2);
        int pb = (int)std::min<size_t>(workspace_bytes / t_per_pair, 32768);
        pb = std::max(1, std::min(pb, n_pairs));
        float2* T = static_cast<float2*>(d_workspace);
        int ywin0 = (Y0 - Y / 2) % Y; if (ywin0 < 0) ywin0 += Y;
        int xwin0 = (X0 - X / 2) % X; if (xwin0 < 0) xwin0 += X;
        const double scale = 1.0 / ((double)Y * X) / sqrt((double)Y * X);
        const int M = Y / 32;
        const size_t tw_bytes = (size_t)M * 32 * sizeof(float2);
        if (Y % 32 == 0 && tw_bytes <= (size_t)160 * 1024 && rcc_fft_enabled()) {
            // FFT-structured stage 1, pairs walked in L2-sized (i-tile, j-tile) blocks: for one
            // block of 32 kx columns the 2 * TS spectrum slabs of a tile stay L2 resident while
            // all TS^2 pairs of the tile read them.
            std::vector<int> hi(n_pairs), hj(n_pairs);
            PB_CUDA_CHECK(cudaMemcpyAsync(hi.data(), d_pair_i, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, s));
            PB_CUDA_CHECK(cudaMemcpyAsync(hj.data(), d_pair_j, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, s));
            PB_CUDA_CHECK(cudaStreamSynchronize(s));
            const int TS = pb_rcc_tile_segments(Y);
            std::vector<int> order(n_pairs);
            for (int k = 0; k < n_pairs; k++) order[k] = k;
            auto key = [&](int k) { return ((long long)(hi[k] / TS) << 32) | (unsigned)(hj[k] / TS); };
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) < key(b); });
            std::vector<int> perm(3 * (size_t)n_pairs);
            for (int k = 0; k < n_pairs; k++) {
                perm[k] = hi[order[k]]; perm[n_pairs + k] = hj[order[k]]; perm[2 * (size_t)n_pairs + k] = order[k];
            }
            int* d_perm = nullptr;
            float2* d_tw = nullptr;
            PB_CUDA_CHECK(cudaMalloc(&d_perm, perm.size() * 4));
            if (cudaMalloc(&d_tw, tw_bytes) != cudaSuccess) { cudaFree(d_perm); pb_set_error("pb_rcc_windows_dev: out of memory"); return PB_ERR_CUDA; }
            cudaMemcpyAsync(d_perm, perm.data(), perm.size() * 4, cudaMemcpyHostToDevice, s);
            float2* d_E = nullptr;
            if (cudaMalloc(&d_E, (size_t)XH * W * sizeof(float2)) != cudaSuccess) {
                cudaFree(d_perm); cudaFree(d_tw);
                pb_set_error("pb_rcc_windows_dev: out of memory");
                return PB_ERR_CUDA;
            }
            rcc_cols_table_kernel<<<(unsigned)(((size_t)XH * W + 255) / 256), 256, 0, s>>>(X, XH, W, xwin0, d_E);
            g_pb_launches++;
            int occ = 2;   // 255 registers, no spills: measured 0.36 s vs 0.53 s (168 registers) on 19 900 pairs
            if (const char* e = getenv("PB_RCC_OCC")) occ = atoi(e) == 3 ? 3 : 2;
            auto rows_kernel = occ == 2 ? rcc_fft_rows_kernel<2> : rcc_fft_rows_kernel<3>;
            cudaFuncSetAttribute(rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(tw_bytes, 32768));
            int smem_fft = (int)std::max<size_t>(tw_bytes, (size_t)kFW * 32 * 32 * sizeof(float2));
            // PB_RCC_ROWS=regs selects the register-staged variant (A/B measurement); default: async staging
            bool use_async = true;
            if (const char* e = getenv("PB_RCC_ROWS")) use_async = strcmp(e, "regs") != 0;
            if (use_async) {
                smem_fft = kFW * 64 * 32 * (int)sizeof(float2);      // 64 KB: one 16 KB stage per warp
                cudaFuncSetAttribute(rcc_fft_rows_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fft);
            }
            const int nkb = (XH + 31) / 32;
            for (int p0 = 0; p0 < n_pairs;) {
                // batch = as many pairs as T fits into the workspace, cut at a tile boundary
                int p1 = std::min(n_pairs, p0 + pb);
                if (p1 < n_pairs) {
                    int b = p1;
                    while (b > p0 && key(order[b]) == key(order[b - 1])) b--;
                    if (b > p0) p1 = b;
                }
                for (int row0 = 0; row0 < H; row0 += 32) {
                    const int nrows = std::min(32, H - row0);
                    const int y_first = (ywin0 + row0) % Y;
                    rcc_fft_twiddle_kernel<<<(M * 32 + 255) / 256, 256, 0, s>>>(Y, M, y_first, d_tw);
                    g_pb_launches++;
                    // one launch per tile: pairs fastest, kx block slower
                    for (int t0 = p0; t0 < p1;) {
                        int t1 = t0;
                        const long long kk = key(order[t0]);
                        while (t1 < p1 && key(order[t1]) == kk) t1++;
                        dim3 g1(t1 - t0, nkb);
                        if (use_async)
                            rcc_fft_rows_async_kernel<<<g1, 32 * kFW, smem_fft, s>>>(
                                static_cast<const float2*>(d_spectra), spec, d_perm + t0, d_perm + n_pairs + t0,
                                XH, M, d_tw, y_first & 31, row0, nrows, H, T + (size_t)(t0 - p0) * H * XH);
                        else
                            rows_kernel<<<g1, 32 * kFW, smem_fft, s>>>(
                                static_cast<const float2*>(d_spectra), spec, d_perm + t0, d_perm + n_pairs + t0,
                                XH, M, d_tw, y_first & 31, row0, nrows, H, T + (size_t)(t0 - p0) * H * XH);
                        g_pb_launches++;
                        t0 = t1;
                    }
                }
                dim3 g2((H + 31) / 32, p1 - p0, (W + 31) / 32);
                rcc_cols_gemm_kernel<<<g2, 256, 0, s>>>(T, H, W, XH, d_E, scale, d_windows,
                                                        d_perm + 2 * (size_t)n_pairs + p0);
                g_pb_launches++;
                p0 = p1;
            }
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            cudaFree(d_perm); cudaFree(d_tw); cudaFree(d_E);
            if (e != cudaSuccess) { pb_set_error("pb_rcc_windows_dev: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
            return PB_OK;
        }
        const int smem1 = 2 * kPR * kS1Cols * (int)sizeof(float2);    // 32 KB (>= twiddle chunk)
        static_assert(kS1Groups * kKC * kPR <= 2 * kPR * kS1Cols, "scratch must hold the twiddles");
        for (int p0 = 0; p0 < n_pairs; p0 += pb) {
            const int nb = std::min(pb, n_pairs - p0);
            for (int row0 = 0; row0 < H; row0 += kPR) {
                const int nrows = std::min(kPR, H - row0);
                dim3 g1((XH + kS1Cols - 1) / kS1Cols, nb);
                rcc_pruned_rows_kernel<<<g1, kS1Cols * kS1Groups, smem1, s>>>(
                    static_cast<const float2*>(d_spectra), spec, d_pair_i + p0, d_pair_j + p0, Y, XH,
                    ywin0, row0, nrows, H, T);
                g_pb_launches++;
            }
            dim3 g2(H, nb);
            rcc_pruned_cols_kernel<<<g2, 256, 0, s>>>(T, H, W, XH, X, xwin0, scale,
                                                      d_windows + (size_t)p0 * H * W, nullptr);
            g_pb_launches++;
        }
        PB_CUDA_CHECK(cudaGetLastError());
        PB_CUDA_CHECK(cudaStreamSynchronize(s));
        return PB_OK;
    }
    float2* prod = static_cast<float2*>(d_workspace);
    float* corr = reinterpret_cast<float*>(static_cast<char*>(d_workspace) + (size_t)batch * spec * 8);
    CufftPlan plan, tail;
    int dims[2] = {Y, X};
    PB_CUFFT_CHECK(cufftPlanMany(&plan.h, 2, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, batch));
    plan.ok = true;
    PB_CUFFT_CHECK(cufftSetStream(plan.h, s));
    const double scale = 1.0 / ((double)Y * X) / sqrt((double)Y * X);
    for (int p0 = 0; p0 < n_pairs; p0 += batch) {
        const int nb = std::min(batch, n_pairs - p0);
        cufftHandle h = plan.h;
        if (nb != batch) {
            PB_CUFFT_CHECK(cufftPlanMany(&tail.h, 2, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, nb));
            tail.ok = true;
            PB_CUFFT_CHECK(cufftSetStream(tail.h, s));
            h = tail.h;
        }
        const size_t total = (size_t)nb * spec;
        int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 32);
        rcc_conj_mul_kernel<<<grid, 256, 0, s>>>(static_cast<const float2*>(d_spectra), spec,
                                                 d_pair_i + p0, d_pair_j + p0, nb, prod);
        PB_CUFFT_CHECK(cufftExecC2R(h, reinterpret_cast<cufftComplex*>(prod), corr));
        const size_t wt = (size_t)nb * H * W;
        int g2 = (int)std::min<size_t>((wt + 255) / 256, 148 * 32);
        rcc_crop_kernel<<<g2, 256, 0, s>>>(corr, Y, X, nb, Y0, X0, H, W, scale,
                                           d_windows + (size_t)p0 * H * W);
        g_pb_launches += 3;
    }
    PB_CUDA_CHECK(cudaGetLastError());
    PB_CUDA_CHECK(cudaStreamSynchronize(s));
    return PB_OK;
}

// Host-buffer variant for all i<j pairs of `segments` (n_seg, Y, X) float32:
// windows (n_pairs, H, W) float32 in the reference's pair order (i outer, j inner),
// sums (n_seg) float64.
extern "C" int pb_rcc_windows(int n_seg, int Y, int X, const float* segments, int Y0, int X0,
                              int H, int W, float* windows, double* sums) {
    if (n_seg < 1) return PB_OK;
    if (!segments || !sums || (n_seg > 1 && !windows)) { pb_set_error("pb_rcc_windows: null pointer"); return PB_ERR_INVALID; }
    if (Y < 1 || X < 1 || H < 1 || W < 1 || Y0 < 0 || X0 < 0 || Y0 + H > Y || X0 + W > X) {
        pb_set_error("pb_rcc_windows: bad window");
        return PB_ERR_INVALID;
    }
    const size_t img = (size_t)Y * X, spec = (size_t)Y * (X / 2 + 1);
    const int n_pairs = n_seg * (n_seg - 1) / 2;
    // batch sized for ~1 GB of workspace
    int batch = (int)std::max<size_t>(1, std::min<size_t>(64, ((size_t)1 << 30) / (spec * 8 + img * 4)));
    batch = std::max(1, std::min(batch, std::max(n_pairs, 1)));
    float *dseg = nullptr, *dwin = nullptr;
    void *dspec = nullptr, *dws = nullptr;
    double* dsum = nullptr;
    int *dpi = nullptr, *dpj = nullptr;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t err) { if (err != cudaSuccess && e == cudaSuccess) e = err; };
    const size_t wsb = (size_t)batch * (spec * 8 + img * 4);
    ok(cudaMalloc(&dseg, n_seg * img * 4));
    ok(cudaMalloc(&dspec, n_seg * spec * 8));
    ok(cudaMalloc(&dsum, n_seg * 8));
    ok(cudaMalloc(&dws, wsb));
    ok(cudaMalloc(&dwin, std::max<size_t>(1, (size_t)n_pairs * H * W * 4)));
    ok(cudaMalloc(&dpi, std::max(1, n_pairs) * 4));
    ok(cudaMalloc(&dpj, std::max(1, n_pairs) * 4));
    int rc = PB_OK;
    if (e == cudaSuccess) {
        std::vector<int> pi, pj;
        for (int i = 0; i < n_seg - 1; i++)
            for (int j = i + 1; j < n_seg; j++) { pi.push_back(i); pj.push_back(j); }
        if (pb_h2d(dseg, segments, n_seg * img * 4, nullptr) != PB_OK) ok(cudaErrorUnknown);   // threaded pinned staging
        if (n_pairs) {
            ok(cudaMemcpy(dpi, pi.data(), n_pairs * 4, cudaMemcpyHostToDevice));
            ok(cudaMemcpy(dpj, pj.data(), n_pairs * 4, cudaMemcpyHostToDevice));
        }
        if (e == cudaSuccess) rc = pb_rcc_spectra_dev(n_seg, Y, X, dseg, dspec, dsum, nullptr);
        if (e == cudaSuccess && rc == PB_OK && n_pairs)
            rc = pb_rcc_windows_dev(n_pairs, dpi, dpj, Y, X, dspec, Y0, X0, H, W, dwin, batch, dws,
                                    wsb, nullptr);
        if (e == cudaSuccess && rc == PB_OK) {
            ok(cudaMemcpy(sums, dsum, n_seg * 8, cudaMemcpyDeviceToHost));
            if (n_pairs && pb_d2h(windows, dwin, (size_t)n_pairs * H * W * 4, nullptr) != PB_OK) ok(cudaErrorUnknown);
        }
    }
    cudaFree(dseg); cudaFree(dspec); cudaFree(dsum); cudaFree(dws); cudaFree(dwin); cudaFree(dpi); cudaFree(dpj);
    if (e != cudaSuccess) { pb_set_error("pb_rcc_windows: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    return rc;
}

// Fused undrift front end (postprocess.undrift :2903-2961 = segment :2846-2900 + rcc
// :160-217): the segment images are rendered ON THE DEVICE straight into the segment
// stack (no host round trip of n_seg x Y x X images), then spectra and pair windows as in
// pb_rcc_windows.  Localisations arrive grouped by segment: segment i owns
// [seg_start[i], seg_start[i+1]) of x/y/lpx/lpy.  Rendering parameters are those of
// postprocess.segment: oversampling 1, full field of view, blur "gaussian" (mode 1).
extern "C" int pb_undrift_windows(int n_seg, const long long* seg_start, const float* x,
                                  const float* y, const float* lpx, const float* lpy, int Y, int X,
                                  double min_blur_width, int Y0, int X0, int H, int W,
                                  float* windows, double* sums, float* segments_out /*nullable*/) {
    return pb_undrift_windows_pairs(n_seg, seg_start, x, y, lpx, lpy, Y, X, min_blur_width, Y0, X0, H, W,
                                    -1, nullptr, nullptr, windows, sums, segments_out);
}

// Same for a SUBSET of the pairs (multi-GPU: every rank renders all segments and transforms them
// -- 33 ms for 200 x 4096^2 -- and correlates only its share of the pairs; no spectra exchange).
// n_pairs < 0: all i < j pairs in the reference's order.
static int undrift_impl(int n_seg, const long long* seg_start, const float* x, const float* y,
                        const float* lpx, const float* lpy, int Y, int X, double min_blur_width, int Y0,
                        int X0, int H, int W, int n_pairs_in, const int* pair_i, const int* pair_j,
                        float* windows, double* sums, float* segments_out, double* peak_records);

extern "C" int pb_undrift_windows_pairs(int n_seg, const long long* seg_start, const float* x,
                                        const float* y, const float* lpx, const float* lpy, int Y, int X,
                                        double min_blur_width, int Y0, int X0, int H, int W,
                                        int n_pairs_in, const int* pair_i, const int* pair_j,
                                        float* windows, double* sums, float* segments_out /*nullable*/) {
    return undrift_impl(n_seg, seg_start, x, y, lpx, lpy, Y, X, min_blur_width, Y0, X0, H, W, n_pairs_in,
                        pair_i, pair_j, windows, sums, segments_out, nullptr);
}

// Windows AND their peak fits on the device: per pair a 32-double record (see rcc_peakfit_kernel);
// `windows` may be NULL -- then only 256 bytes per pair come back instead of H x W floats.
extern "C" int pb_undrift_peaks_pairs(int n_seg, const long long* seg_start, const float* x,
                                      const float* y, const float* lpx, const float* lpy, int Y, int X,
                                      double min_blur_width, int Y0, int X0, int H, int W,
                                      int n_pairs_in, const int* pair_i, const int* pair_j,
                                      double* peak_records, double* sums, float* windows /*nullable*/) {
    if (!peak_records && n_pairs_in != 0 && n_seg > 1) { pb_set_error("pb_undrift_peaks_pairs: null records"); return PB_ERR_INVALID; }
    return undrift_impl(n_seg, seg_start, x, y, lpx, lpy, Y, X, min_blur_width, Y0, X0, H, W, n_pairs_in,
                        pair_i, pair_j, windows, sums, nullptr, peak_records);
}

static int undrift_impl(int n_seg, const long long* seg_start, const float* x, const float* y,
                        const float* lpx, const float* lpy, int Y, int X, double min_blur_width, int Y0,
                        int X0, int H, int W, int n_pairs_in, const int* pair_i, const int* pair_j,
                        float* windows, double* sums, float* segments_out, double* peak_records) {
    if (n_seg < 1) return PB_OK;
    if (n_pairs_in > 0 && (!pair_i || !pair_j)) { pb_set_error("pb_undrift_windows_pairs: null pair list"); return PB_ERR_INVALID; }
    if (!seg_start || !sums || (n_seg > 1 && n_pairs_in != 0 && !windows && !peak_records)) { pb_set_error("pb_undrift_windows: null pointer"); return PB_ERR_INVALID; }
    if (Y < 1 || X < 1 || H < 1 || W < 1 || Y0 < 0 || X0 < 0 || Y0 + H > Y || X0 + W > X) {
        pb_set_error("pb_undrift_windows: bad window");
        return PB_ERR_INVALID;
    }
    const size_t n_locs = (size_t)seg_start[n_seg];
    const size_t img = (size_t)Y * X, spec = (size_t)Y * (X / 2 + 1);
    const int n_pairs = n_pairs_in >= 0 ? n_pairs_in : n_seg * (n_seg - 1) / 2;
    int batch = (int)std::max<size_t>(1, std::min<size_t>(64, ((size_t)1 << 30) / (spec * 8 + img * 4)));
    batch = std::max(1, std::min(batch, std::max(n_pairs, 1)));
    size_t max_seg = 0;
    for (int i = 0; i < n_seg; i++) max_seg = std::max<size_t>(max_seg, (size_t)(seg_start[i + 1] - seg_start[i]));
    const size_t rws = pb_render_workspace_bytes(max_seg, Y, X);
    const size_t wsb = std::max((size_t)batch * (spec * 8 + img * 4), rws);
    float *dseg = nullptr, *dwin = nullptr, *dx = nullptr, *dy = nullptr, *dlx = nullptr, *dly = nullptr;
    void *dspec = nullptr, *dws = nullptr;
    double* dsum = nullptr;
    unsigned long long* dcnt = nullptr;
    int *dpi = nullptr, *dpj = nullptr;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t err) { if (err != cudaSuccess && e == cudaSuccess) e = err; };
    const size_t nb = std::max<size_t>(n_locs, 1) * 4;
    ok(cudaMalloc(&dseg, n_seg * img * 4));
    ok(cudaMalloc(&dspec, n_seg * spec * 8));
    ok(cudaMalloc(&dsum, n_seg * 8));
    ok(cudaMalloc(&dcnt, 8));
    ok(cudaMalloc(&dws, wsb));
    ok(cudaMalloc(&dwin, std::max<size_t>(1, (size_t)n_pairs * H * W * 4)));
    ok(cudaMalloc(&dpi, std::max(1, n_pairs) * 4));
    ok(cudaMalloc(&dpj, std::max(1, n_pairs) * 4));
    ok(cudaMalloc(&dx, nb)); ok(cudaMalloc(&dy, nb)); ok(cudaMalloc(&dlx, nb)); ok(cudaMalloc(&dly, nb));
    int rc = PB_OK;
    if (e == cudaSuccess) {
        if (n_locs) {
            if (pb_h2d(dx, x, n_locs * 4, nullptr) != PB_OK) ok(cudaErrorUnknown);
            if (pb_h2d(dy, y, n_locs * 4, nullptr) != PB_OK) ok(cudaErrorUnknown);
            if (pb_h2d(dlx, lpx, n_locs * 4, nullptr) != PB_OK) ok(cudaErrorUnknown);
            if (pb_h2d(dly, lpy, n_locs * 4, nullptr) != PB_OK) ok(cudaErrorUnknown);
        }
        for (int i = 0; i < n_seg && rc == PB_OK && e == cudaSuccess; i++) {
            const size_t a0 = (size_t)seg_start[i], m = (size_t)(seg_start[i + 1] - seg_start[i]);
            rc = pb_render_dev(m, dx + a0, dy + a0, dlx + a0, dly + a0, 1.0, 0.0, 0.0, (double)Y,
                               (double)X, min_blur_width, 1, dseg + (size_t)i * img, Y, X, dcnt, dws,
                               wsb, nullptr);
        }
        std::vector<int> pi, pj;
        if (n_pairs_in >= 0) {
            pi.assign(pair_i, pair_i + n_pairs);
            pj.assign(pair_j, pair_j + n_pairs);
            for (int k = 0; k < n_pairs; k++)
                if (pi[k] < 0 || pj[k] < 0 || pi[k] >= n_seg || pj[k] >= n_seg) {
                    pb_set_error("pb_undrift_windows_pairs: pair index out of range");
                    rc = PB_ERR_INVALID;
                }
        } else {
            for (int i = 0; i < n_seg - 1; i++)
                for (int j = i + 1; j < n_seg; j++) { pi.push_back(i); pj.push_back(j); }
        }
        if (n_pairs) {
            ok(cudaMemcpy(dpi, pi.data(), n_pairs * 4, cudaMemcpyHostToDevice));
            ok(cudaMemcpy(dpj, pj.data(), n_pairs * 4, cudaMemcpyHostToDevice));
        }
        if (e == cudaSuccess && rc == PB_OK) rc = pb_rcc_spectra_dev(n_seg, Y, X, dseg, dspec, dsum, nullptr);
        if (e == cudaSuccess && rc == PB_OK && n_pairs)
            rc = pb_rcc_windows_dev(n_pairs, dpi, dpj, Y, X, dspec, Y0, X0, H, W, dwin, batch, dws, wsb, nullptr);
        if (e == cudaSuccess && rc == PB_OK && n_pairs && peak_records) {
            // the spectra are no longer needed: their buffer holds the peak records
            double* drec = static_cast<double*>(dspec);
            if ((size_t)n_pairs * kRecDoubles * 8 > (size_t)n_seg * spec * 8) {
                pb_set_error("pb_undrift_peaks_pairs: record buffer too small");
                rc = PB_ERR_INVALID;
            } else {
                rcc_peakfit_kernel<<<(n_pairs + 63) / 64, 64>>>(dwin, n_pairs, H, W, drec);
                g_pb_launches++;
                ok(cudaGetLastError());
                ok(cudaMemcpy(peak_records, drec, (size_t)n_pairs * kRecDoubles * 8, cudaMemcpyDeviceToHost));
            }
        }
        if (e == cudaSuccess && rc == PB_OK) {
            ok(cudaMemcpy(sums, dsum, n_seg * 8, cudaMemcpyDeviceToHost));
            if (n_pairs && windows && pb_d2h(windows, dwin, (size_t)n_pairs * H * W * 4, nullptr) != PB_OK) ok(cudaErrorUnknown);
            if (segments_out && pb_d2h(segments_out, dseg, n_seg * img * 4, nullptr) != PB_OK) ok(cudaErrorUnknown);
        }
    }
    cudaFree(dseg); cudaFree(dspec); cudaFree(dsum); cudaFree(dcnt); cudaFree(dws); cudaFree(dwin);
    cudaFree(dpi); cudaFree(dpj); cudaFree(dx); cudaFree(dy); cudaFree(dlx); cudaFree(dly);
    if (e != cudaSuccess) { pb_set_error("pb_undrift_windows: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    return rc;
}
