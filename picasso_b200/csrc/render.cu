// picasso_b200/csrc/render.cu
//
// Super-resolution rendering on B200 -- replaces the unrotated paths of
// picasso.render.render: _render_setup + _fill (_render_hist) and
// _draw_gaussian_loc/_fill_gaussian (_render_gaussian, _render_gaussian_iso)
// (reference picasso/render.py:177-232, 451-575, 798-853, 1020-1216).
//
// Per-localisation values follow the reference exactly: strict viewport test,
// coordinates promoted to f64, float32 sigma = f32(oversampling) * max(lp,
// f32(min_blur_width)), window = np.int32 truncations of (c -/+ 3 sigma), 1-D
// kernels evaluated in f64 and rounded to f32, pixel value = gy[i] * gx[j] in
// f32.  Only the accumulation ORDER differs (atomics vs the reference's serial
// loop), so images agree to float32 summation rounding; the histogram (integer
// counts) is exact.
//
// Kernels:
//   render_hist_kernel     one thread per localisation, RED.ADD of 1.0f
//   render_splat_kernel    8 lanes per localisation: lane j owns window columns
//                          j, j+8, ...; a row's 8 consecutive pixels share one or
//                          two 32 B sectors, so the L2 atomic unit sees coalesced
//                          reductions.  Used for small batches.
//   render_strip_kernel    (pb_render_set_impl(1); measured slower, kept for the A/B) binned by 16-row x 64-column
//                          STRIP (counting sort on the device; a window that straddles two strips is
//                          filed in both).  One WARP owns one strip of its CTA's 64x64 shared-memory tile
//                          and draws its localisations one after the other, the 32 lanes covering an
//                          8-column x 8-row block of the window (two pixels per lane) with plain
//                          load / add / store: no two lanes ever touch the same pixel and no other warp
//                          touches the strip, so there are NO shared-memory atomics.  The tile is flushed
//                          once with coalesced reductions; window columns beyond the tile edge go
//                          straight to global atomics.
//   render_tiled_kernel    (default for dense images) localisations are binned by 64x64-pixel tile
//                          (counting sort on the device, the float4 records are scattered into tile
//                          order); one CTA accumulates its tile in shared memory, one THREAD per
//                          localisation, with shared-memory atomicAdd, and flushes it once with
//                          coalesced reductions; window parts that cross the tile edge go straight
//                          to global atomics.
#include <algorithm>
#include <atomic>
#include <stdlib.h>

#include <cub/device/device_scan.cuh>

#include "pb_common.cuh"
#include "../../include/picasso_b200.h"

extern std::atomic<long long> g_pb_launches;

namespace {

struct RenderArgs {
    const float* x;
    const float* y;
    const float* lpx;
    const float* lpy;
    long long n;
    double os, y_min, x_min, y_max, x_max;
    float osf, mbw;
    int mode;              // 0 hist, 1 gaussian, 2 gaussian_iso
    float* image;                // rows [row0, row0 + nrows) of the npy x npx image
    int npy, npx;
    unsigned long long* count;   // localisations in view (whose centre row lies in the band)
    int row0, nrows;             // row band held by `image` (full render: 0, npy)
};

// A localisation is COUNTED by the band that holds its centre pixel row, so that the counts of
// disjoint bands add up to the reference's n; it is DRAWN by every band its window reaches.
__device__ __forceinline__ bool centre_in_band(const RenderArgs& a, double y_) {
    int py = (int)y_;
    py = min(max(py, 0), a.npy - 1);
    return py >= a.row0 && py < a.row0 + a.nrows;
}

struct Splat {            // everything _draw_gaussian_loc derives for one localisation
    double x_, y_, inv_2sx2, inv_2sy2, norm;
    int i_min, i_max, j_min, j_max;
    bool in_view;
};

__device__ __forceinline__ Splat make_splat(const RenderArgs& a, long long k) {
    Splat s;
    const double xv = (double)a.x[k], yv = (double)a.y[k];
    s.in_view = (xv > a.x_min) && (yv > a.y_min) && (xv < a.x_max) && (yv < a.y_max);
    s.x_ = a.os * (xv - a.x_min);
    s.y_ = a.os * (yv - a.y_min);
    s.i_min = s.i_max = s.j_min = s.j_max = 0;
    s.inv_2sx2 = s.inv_2sy2 = s.norm = 0.0;
    if (!s.in_view || a.mode == 0) return s;
    const float bw = __fmul_rn(a.osf, fmaxf(a.lpx[k], a.mbw));
    const float bh = __fmul_rn(a.osf, fmaxf(a.lpy[k], a.mbw));
    float sx, sy;
    if (a.mode == 2) { sy = __fdiv_rn(__fadd_rn(bh, bw), 2.0f); sx = sy; }
    else { sx = bw; sy = bh; }
    const double moy = 3.0 * (double)sy, mox = 3.0 * (double)sx;
    s.i_min = max((int)(s.y_ - moy), 0);
    s.i_max = min((int)(s.y_ + moy + 1.0), a.npy);
    s.j_min = max((int)(s.x_ - mox), 0);
    s.j_max = min((int)(s.x_ + mox) + 1, a.npx);
    s.inv_2sx2 = 1.0 / (2.0 * (double)sx * (double)sx);
    s.inv_2sy2 = 1.0 / (2.0 * (double)sy * (double)sy);
    s.norm = 1.0 / (2.0 * 3.141592653589793 * (double)sx * (double)sy);
    return s;
}

__device__ __forceinline__ float splat_gx(const Splat& s, int j) {
    const double dx = (double)j + 0.5 - s.x_;
    return (float)exp(-dx * dx * s.inv_2sx2);
}
__device__ __forceinline__ float splat_gy(const Splat& s, int i) {
    const double dy = (double)i + 0.5 - s.y_;
    return (float)(s.norm * exp(-dy * dy * s.inv_2sy2));
}

__global__ void render_hist_kernel(const RenderArgs a) {
    unsigned long long local = 0;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < a.n;
         k += (long long)gridDim.x * blockDim.x) {
        const Splat s = make_splat(a, k);
        if (s.in_view) {
            if (centre_in_band(a, s.y_)) local++;
            const int i = (int)s.x_, j = (int)s.y_;      // x.astype(int32): truncation
            if (j >= a.row0 && j < a.row0 + a.nrows && j < a.npy && i >= 0 && i < a.npx)
                atomicAdd(a.image + (size_t)(j - a.row0) * a.npx + i, 1.0f);
        }
    }
    // block-aggregated count
    __shared__ unsigned long long blk;
    if (threadIdx.x == 0) blk = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&blk, local);
    __syncthreads();
    if (threadIdx.x == 0 && blk) atomicAdd(a.count, blk);
}

constexpr int kLanes = 8;     // lanes per localisation in the direct splat kernel

__global__ void render_splat_kernel(const RenderArgs a) {
    const int lane = threadIdx.x & 31;
    const int g = lane & (kLanes - 1);
    const unsigned gmask = 0xffu << (lane & ~(kLanes - 1));   // the 8 lanes sharing a localisation
    const long long groups = ((long long)gridDim.x * blockDim.x) / kLanes;
    const long long gid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / kLanes;
    unsigned long long local = 0;
    const long long nround = (a.n + groups - 1) / groups;
    for (long long r = 0; r < nround; r++) {
        const long long k = r * groups + gid;
        Splat s;
        s.in_view = false;
        if (k < a.n) s = make_splat(a, k);
        if (s.in_view && g == 0 && centre_in_band(a, s.y_)) local++;
        if (!s.in_view) continue;
        s.i_min = max(s.i_min, a.row0);                 // clip the window rows to the band
        s.i_max = min(s.i_max, a.row0 + a.nrows);
        const int nx = s.j_max - s.j_min, ny = s.i_max - s.i_min;
        if (nx <= 0 || ny <= 0) continue;
        // lane g evaluates ONE column kernel value and ONE row kernel value per 8x8 chunk;
        // row values are passed around the 8-lane group with shuffles
        for (int i0 = 0; i0 < ny; i0 += kLanes) {
            const float gy_mine = (i0 + g < ny) ? splat_gy(s, s.i_min + i0 + g) : 0.0f;
            const int nr = min(kLanes, ny - i0);
            for (int j0 = 0; j0 < nx; j0 += kLanes) {
                const int j = s.j_min + j0 + g;
                const bool jok = (j0 + g) < nx;
                const float gx = jok ? splat_gx(s, j) : 0.0f;
                for (int rr = 0; rr < nr; rr++) {
                    const float gy = __shfl_sync(gmask, gy_mine, rr, kLanes);
                    if (jok)
                        atomicAdd(a.image + (size_t)(s.i_min + i0 + rr - a.row0) * a.npx + j, __fmul_rn(gy, gx));
                }
            }
        }
    }
    __shared__ unsigned long long blk;
    if (threadIdx.x == 0) blk = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (lane == 0 && local) atomicAdd(&blk, local);
    __syncthreads();
    if (threadIdx.x == 0 && blk) atomicAdd(a.count, blk);
}

// ---- tiled path ---------------------------------------------------------------
constexpr int kTile = 64;                 // tile edge in image pixels (16 KB of smem)
constexpr int kWinClass = 16;             // window heights 1..16 get their own class (taller: the last one)
constexpr int kSizeClasses = kWinClass * 2;           // bins per tile: window rows x (narrow | wide columns)

// pass 1: bin per localisation (-1 = not in view) + histogram of bin sizes.  bin = (tile, window class): the
// accumulation pass runs one thread per localisation, and a warp takes as long as its largest window --
// with four classes by blur width only 16.5 of 32 lanes were busy (windows are 3..11 pixels per axis at
// config 4, rows and columns independently).  Classes = exact window height x (narrow | wide): 20 lanes busy.
// (256 exact (rows, columns) classes reach 23 -- bins of 8 records, warps span several -- but scatter the
// records over 6.5 M bins, more write sectors than L2 holds: scatter pass 1.3 -> 2.3 ms.)
__global__ void render_bin_kernel(const RenderArgs a, int tiles_x, int* __restrict__ tile_of,
                                  unsigned int* __restrict__ tile_count) {
    unsigned long long local = 0;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < a.n;
         k += (long long)gridDim.x * blockDim.x) {
        const double xv = (double)a.x[k], yv = (double)a.y[k];
        const bool in_view = (xv > a.x_min) && (yv > a.y_min) && (xv < a.x_max) && (yv < a.y_max);
        int t = -1;
        if (in_view) {
            const double x_ = a.os * (xv - a.x_min), y_ = a.os * (yv - a.y_min);
            int px = (int)x_, py = (int)y_;
            px = min(max(px, 0), a.npx - 1);
            py = min(max(py, 0), a.npy - 1);
            if (py >= a.row0 && py < a.row0 + a.nrows) local++;
            // window exactly as the accumulation pass derives it (render.py:505-525)
            const float bw = __fmul_rn(a.osf, fmaxf(a.lpx[k], a.mbw));
            const float bh = __fmul_rn(a.osf, fmaxf(a.lpy[k], a.mbw));
            float sx, sy;
            if (a.mode == 2) { sy = __fdiv_rn(__fadd_rn(bh, bw), 2.0f); sx = sy; }
            else { sx = bw; sy = bh; }
            const double moy = 3.0 * (double)sy, mox = 3.0 * (double)sx;
            const int ny = min((int)(y_ + moy + 1.0), a.npy) - max((int)(y_ - moy), 0);
            const int nx = min((int)(x_ + mox) + 1, a.npx) - max((int)(x_ - mox), 0);
            const int cls = (min(max(ny, 1), kWinClass) - 1) * 2 + (nx > 6 ? 1 : 0);
            // row band: localisations centred outside the band whose 3-sigma window reaches into it
            // (conservative test; the accumulation pass clips exactly) go to the nearest band tile
            const double reach = 3.0 * (double)fmaxf(sx, sy) + 2.0;
            if (y_ + reach >= (double)a.row0 && y_ - reach <= (double)(a.row0 + a.nrows)) {
                const int pyb = min(max(py, a.row0), a.row0 + a.nrows - 1) - a.row0;
                const int tile = (pyb / kTile) * tiles_x + (px / kTile);
                t = tile * kSizeClasses + cls;
                atomicAdd(tile_count + t, 1u);
            }
        }
        tile_of[k] = t;
    }
    __shared__ unsigned long long blk;
    if (threadIdx.x == 0) blk = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&blk, local);
    __syncthreads();
    if (threadIdx.x == 0 && blk) atomicAdd(a.count, blk);
}

// pass 2: exclusive scan of the bin histogram -> start[0 .. nbins] (start[nbins] = total; count[nbins] is kept
// zero for that) and a second copy, cursor, for the scatter pass.  cub::DeviceScan (decoupled look-back, all
// SMs, ~10 us for 100 k bins); the first version was a single-CTA loop with strided reads (180 us).
constexpr size_t kScanTempBytes = 1 << 20;
int render_scan(unsigned int* count, unsigned int* start, unsigned int* cursor, long long nbins, void* temp,
                cudaStream_t s) {
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, count, start, (int)(nbins + 1), s);
    if (need > kScanTempBytes) { pb_set_error("pb_render: scan workspace too small (%zu bytes)", need); return PB_ERR_INVALID; }
    size_t tb = kScanTempBytes;
    if (cub::DeviceScan::ExclusiveSum(temp, tb, count, start, (int)(nbins + 1), s) != cudaSuccess) {
        pb_set_error("pb_render: scan failed: %s", cudaGetErrorString(cudaGetLastError()));
        return PB_ERR_CUDA;
    }
    PB_CUDA_CHECK(cudaMemcpyAsync(cursor, start, (size_t)nbins * 4, cudaMemcpyDeviceToDevice, s));
    return PB_OK;
}

// pass 3: scatter the localisations themselves (x, y, lpx, lpy as one float4) into bin order, so the
// accumulation pass streams them with coalesced 16 B loads.  Slots are handed out by a cursor per bin at
// scatter time: records that are written close in time land next to each other and L2 merges them into
// full sectors (taking the slot from a rank recorded by the bin pass instead -- no atomics here -- was
// measured slower: 2.2 ms against 1.3 ms).
__global__ void render_scatter_kernel(const RenderArgs a, const int* __restrict__ tile_of,
                                      unsigned int* __restrict__ cursor,
                                      float4* __restrict__ sorted) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < a.n;
         k += (long long)gridDim.x * blockDim.x) {
        const int t = tile_of[k];
        if (t >= 0)
            sorted[atomicAdd(cursor + t, 1u)] = make_float4(a.x[k], a.y[k], a.lpx[k], a.lpy[k]);
    }
}

// pass 4: one CTA per tile; ONE THREAD per localisation.  The thread evaluates its 1-D
// column kernel once into shared memory (gxs[jj][tid], conflict-free), then walks the
// window rows adding gy * gx into the CTA's shared tile with shared-memory atomics; window
// parts outside the tile go to global atomics.  The window geometry and sigma follow the
// reference in float64 / float32 exactly; the two exponentials per pixel row/column are
// evaluated with expf on a float64-computed offset (per-pixel relative deviation ~1e-6,
// far inside the 1e-4 render tolerance).
constexpr int kMaxWin = 16;               // windows wider than this take the generic path

__global__ void __launch_bounds__(256) render_tiled_kernel(const RenderArgs a, int tiles_x,
                                                           const unsigned int* __restrict__ start,
                                                           const float4* __restrict__ sorted) {
    __shared__ float acc[kTile * kTile];
    __shared__ float gxs[kMaxWin][256];
    const int tile = blockIdx.x;
    // the tile's localisations: its kSizeClasses consecutive bins, short windows first
    const unsigned int first = start[tile * kSizeClasses], last = start[(tile + 1) * kSizeClasses];
    if (first == last) return;
    const int ty0 = a.row0 + (tile / tiles_x) * kTile, tx0 = (tile % tiles_x) * kTile;
    const int band_end = a.row0 + a.nrows;
    for (int q = threadIdx.x; q < kTile * kTile; q += blockDim.x) acc[q] = 0.0f;      // (0.0f == 0u)
    __syncthreads();
    const int tid = threadIdx.x;
    for (unsigned int q = first + tid; q < last; q += blockDim.x) {
        const float4 L = sorted[q];
        // ---- window and sigmas exactly as _draw_gaussian_loc (render.py:505-525) ----
        const double x_ = a.os * ((double)L.x - a.x_min);
        const double y_ = a.os * ((double)L.y - a.y_min);
        const float bw = __fmul_rn(a.osf, fmaxf(L.z, a.mbw));
        const float bh = __fmul_rn(a.osf, fmaxf(L.w, a.mbw));
        float sx, sy;
        if (a.mode == 2) { sy = __fdiv_rn(__fadd_rn(bh, bw), 2.0f); sx = sy; }
        else { sx = bw; sy = bh; }
        const double moy = 3.0 * (double)sy, mox = 3.0 * (double)sx;
        const int i_min = max((int)(y_ - moy), 0);
        const int i_max = min((int)(y_ + moy + 1.0), a.npy);
        const int j_min = max((int)(x_ - mox), 0);
        const int j_max = min((int)(x_ + mox) + 1, a.npx);
        const int nx = j_max - j_min, ny = i_max - i_min;
        if (nx <= 0 || ny <= 0) continue;
        const int ii0 = max(a.row0 - i_min, 0);                   // window rows inside the band
        const int ii1 = min(band_end - i_min, ny);
        if (ii0 >= ii1) continue;
        const float inv_2sx2 = 1.0f / (2.0f * sx * sx);
        const float inv_2sy2 = 1.0f / (2.0f * sy * sy);
        const float norm = 1.0f / (6.2831853071795862f * sx * sy);
        const float dx0 = (float)((double)j_min + 0.5 - x_);   // offsets relative to the centre,
        const float dy0 = (float)((double)i_min + 0.5 - y_);   // formed in f64 (x_ can be ~1e4)
        if (nx <= kMaxWin) {
#pragma unroll 1
            for (int jj = 0; jj < nx; jj++) {
                const float dx = dx0 + (float)jj;
                gxs[jj][tid] = expf(-dx * dx * inv_2sx2);
            }
#pragma unroll 1
            for (int ii = ii0; ii < ii1; ii++) {
                const float dy = dy0 + (float)ii;
                const float gy = norm * expf(-dy * dy * inv_2sy2);
                const int i = i_min + ii;
                const bool iin = (i >= ty0) && (i < ty0 + kTile);
                float* grow = a.image + (size_t)(i - a.row0) * a.npx;
                float* srow = acc + (i - ty0) * kTile - tx0;
#pragma unroll 1
                for (int jj = 0; jj < nx; jj++) {
                    const int j = j_min + jj;
                    const float g = gxs[jj][tid];
                    if (iin && j >= tx0 && j < tx0 + kTile) {
                        atomicAdd(srow + j, gy * g);
                    } else {
                        atomicAdd(grow + j, gy * g);
                    }
                }
            }
        } else {   // very wide kernels (sigma > 2.5 display px): no column cache
            for (int ii = ii0; ii < ii1; ii++) {
                const float dy = dy0 + (float)ii;
                const float gy = norm * expf(-dy * dy * inv_2sy2);
                for (int jj = 0; jj < nx; jj++) {
                    const float dx = dx0 + (float)jj;
                    atomicAdd(a.image + (size_t)(i_min + ii - a.row0) * a.npx + j_min + jj,
                              gy * expf(-dx * dx * inv_2sx2));
                }
            }
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < kTile * kTile; q += blockDim.x) {
        const int i = ty0 + q / kTile, j = tx0 + q % kTile;
        const float v = acc[q];
        if (v != 0.0f && i < band_end && j < a.npx) atomicAdd(a.image + (size_t)(i - a.row0) * a.npx + j, v);
    }
}

// ---- strip path: warp-owned strips, no shared-memory atomics -------------------------------------
constexpr int kStripRows = 16;            // rows per strip (4 strips = 4 warps per 64x64 tile)
constexpr int kStripsPerTile = kTile / kStripRows;
constexpr int kAccStride = kTile + 8;     // padded tile row: the 4 rows x 8 columns a warp touches per step
                                          // fall into 32 different banks
constexpr int kTwoStrips = 1 << 30;       // tile_of flag: the window continues in the strip below

__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// pass 1: strip bin(s) per localisation + histogram of bin sizes.  tile_of: -1 nothing to draw,
// -2 window taller than two strips (drawn directly by the scatter pass), else first bin
// (| kTwoStrips).  Window geometry as make_splat, rows clipped to the band.
__global__ void render_strip_bin_kernel(const RenderArgs a, int tiles_x, int* __restrict__ tile_of,
                                        unsigned int* __restrict__ bin_count) {
    unsigned long long local = 0;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < a.n;
         k += (long long)gridDim.x * blockDim.x) {
        const Splat s = make_splat(a, k);
        int t = -1;
        if (s.in_view) {
            if (centre_in_band(a, s.y_)) local++;
            const int lo = max(s.i_min, a.row0), hi = min(s.i_max, a.row0 + a.nrows);
            if (lo < hi && s.j_min < s.j_max) {
                const int s0 = (lo - a.row0) / kStripRows, s1 = (hi - 1 - a.row0) / kStripRows;
                if (s1 - s0 > 1) {
                    t = -2;
                } else {
                    int px = (int)s.x_;
                    px = min(max(px, 0), a.npx - 1);
                    t = s0 * tiles_x + px / kTile;
                    atomicAdd(bin_count + t, 1u);
                    if (s1 > s0) {
                        atomicAdd(bin_count + t + tiles_x, 1u);
                        t |= kTwoStrips;
                    }
                }
            }
        }
        tile_of[k] = t;
    }
    __shared__ unsigned long long blk;
    if (threadIdx.x == 0) blk = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&blk, local);
    __syncthreads();
    if (threadIdx.x == 0 && blk) atomicAdd(a.count, blk);
}

// pass 3: scatter the records into bin order (a straddling window into both strips); the rare very
// tall windows are drawn here, one thread each, with global atomics and float64 kernels
__global__ void render_strip_scatter_kernel(const RenderArgs a, int tiles_x, const int* __restrict__ tile_of,
                                            unsigned int* __restrict__ cursor, float4* __restrict__ sorted) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < a.n;
         k += (long long)gridDim.x * blockDim.x) {
        const int t = tile_of[k];
        if (t >= 0) {
            const float4 rec = make_float4(a.x[k], a.y[k], a.lpx[k], a.lpy[k]);
            const int bin = t & (kTwoStrips - 1);
            sorted[atomicAdd(cursor + bin, 1u)] = rec;
            if (t & kTwoStrips) sorted[atomicAdd(cursor + bin + tiles_x, 1u)] = rec;
        } else if (t == -2) {
            const Splat s = make_splat(a, k);
            const int lo = max(s.i_min, a.row0), hi = min(s.i_max, a.row0 + a.nrows);
            for (int i = lo; i < hi; i++) {
                const float gy = splat_gy(s, i);
                for (int j = s.j_min; j < s.j_max; j++)
                    atomicAdd(a.image + (size_t)(i - a.row0) * a.npx + j, __fmul_rn(gy, splat_gx(s, j)));
            }
        }
    }
}

// pass 4: one CTA per 64x64 tile, one warp per strip.  Per batch of 32 records every lane derives the
// window of one record (float64 coordinates, float32 sigma, np.int32 truncation: render.py:505-525)
// and parks it in shared memory; then the warp draws the records one after the other.  For each
// 8-column x 8-row block of a window lanes 0-7 evaluate the column kernel, lanes 8-15 the row kernel
// (exp2 of an offset formed in float64; relative deviation from the reference's float64 exp ~1e-6,
// render tolerance 1e-4), shuffles hand them out and every lane updates its two pixels with plain
// shared-memory load / add / store.
__global__ void __launch_bounds__(128) render_strip_kernel(const RenderArgs a, int tiles_x, int strips_y,
                                                           const unsigned int* __restrict__ start,
                                                           const float4* __restrict__ sorted) {
    __shared__ float acc[kTile * kAccStride];
    __shared__ int4 geo_a[kStripsPerTile][32];     // j_min, first row, columns, rows (inside the strip)
    __shared__ float4 geo_b[kStripsPerTile][32];   // dx0, dy0, -log2e / (2 sx^2), -log2e / (2 sy^2)
    __shared__ float geo_c[kStripsPerTile][32];    // 1 / (2 pi sx sy)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_row = blockIdx.x / tiles_x, txi = blockIdx.x % tiles_x;
    const int ty0 = a.row0 + tile_row * kTile, tx0 = txi * kTile;
    const int band_end = a.row0 + a.nrows;
    const int strip = tile_row * kStripsPerTile + warp;
    unsigned int first = 0, last = 0;
    if (strip < strips_y) {
        const long long bin = (long long)strip * tiles_x + txi;
        first = start[bin];
        last = start[bin + 1];
    }
    if (!__syncthreads_or(first != last)) return;
    for (int q = threadIdx.x; q < kTile * kAccStride; q += blockDim.x) acc[q] = 0.0f;
    __syncthreads();
    const int sy0 = ty0 + warp * kStripRows, sy1 = min(sy0 + kStripRows, band_end);
    const int c = lane & 7, r2 = lane >> 3;
    const bool is_row_lane = (lane & 8) != 0;
    for (unsigned int q0 = first; q0 < last; q0 += 32) {
        const int cnt = (int)min(32u, last - q0);
        if (lane < cnt) {
            const float4 L = sorted[q0 + lane];
            const double x_ = a.os * ((double)L.x - a.x_min);
            const double y_ = a.os * ((double)L.y - a.y_min);
            const float bw = __fmul_rn(a.osf, fmaxf(L.z, a.mbw));
            const float bh = __fmul_rn(a.osf, fmaxf(L.w, a.mbw));
            float sx, sy;
            if (a.mode == 2) { sy = __fdiv_rn(__fadd_rn(bh, bw), 2.0f); sx = sy; }
            else { sx = bw; sy = bh; }
            const double moy = 3.0 * (double)sy, mox = 3.0 * (double)sx;
            const int i_min = max((int)(y_ - moy), 0);
            const int i_max = min((int)(y_ + moy + 1.0), a.npy);
            const int j_min = max((int)(x_ - mox), 0);
            const int j_max = min((int)(x_ + mox) + 1, a.npx);
            const int lo = max(i_min, sy0), hi = min(i_max, sy1);
            const int nx = j_max - j_min, nr = hi - lo;
            geo_a[warp][lane] = make_int4(j_min, lo, (nx > 0 && nr > 0) ? nx : 0, nr > 0 ? nr : 0);
            geo_b[warp][lane] = make_float4((float)((double)j_min + 0.5 - x_), (float)((double)lo + 0.5 - y_),
                                            -1.4426950408889634f / (2.0f * sx * sx),
                                            -1.4426950408889634f / (2.0f * sy * sy));
            geo_c[warp][lane] = 1.0f / (6.2831853071795862f * sx * sy);
        }
        __syncwarp();
        for (int b = 0; b < cnt; b++) {
            const int4 A = geo_a[warp][b];
            const float4 B = geo_b[warp][b];
            const float nrm = geo_c[warp][b];
            for (int rr = 0; rr < A.w; rr += 8) {
                for (int cc = 0; cc < A.z; cc += 8) {
                    // lanes 0-7 (and 16-23): column kernel of column cc + c; lanes 8-15 (24-31): row kernel
                    const float off = (is_row_lane ? B.y : B.x) + (float)((is_row_lane ? rr : cc) + c);
                    float g = ex2_approx(off * off * (is_row_lane ? B.w : B.z));
                    if (is_row_lane) g *= nrm;
                    const float gx = __shfl_sync(0xffffffffu, g, c);
                    const float gy0 = __shfl_sync(0xffffffffu, g, 8 + r2);
                    const float gy1 = __shfl_sync(0xffffffffu, g, 12 + r2);
                    const int col = cc + c;
                    if (col < A.z) {
                        const int j = A.x + col;
                        const bool intile = (unsigned)(j - tx0) < (unsigned)kTile;
                        const int ra = rr + r2, rb = rr + r2 + 4;
                        if (ra < A.w) {
                            const int i = A.y + ra;
                            const float v = gy0 * gx;
                            if (intile) acc[(i - ty0) * kAccStride + (j - tx0)] += v;
                            else atomicAdd(a.image + (size_t)(i - a.row0) * a.npx + j, v);
                        }
                        if (rb < A.w) {
                            const int i = A.y + rb;
                            const float v = gy1 * gx;
                            if (intile) acc[(i - ty0) * kAccStride + (j - tx0)] += v;
                            else atomicAdd(a.image + (size_t)(i - a.row0) * a.npx + j, v);
                        }
                    }
                    __syncwarp();      // the next block / record may touch the same pixels from other lanes
                }
            }
        }
        __syncwarp();                  // geo_* are rewritten by the next batch
    }
    __syncthreads();
    for (int q = threadIdx.x; q < kTile * kTile; q += blockDim.x) {
        const int r = q / kTile, cq = q % kTile;
        const int i = ty0 + r, j = tx0 + cq;
        const float v = acc[r * kAccStride + cq];
        if (v != 0.0f && i < band_end && j < a.npx) atomicAdd(a.image + (size_t)(i - a.row0) * a.npx + j, v);
    }
}

// ---- multi-GPU: bucket localisations by destination row band ---------------------------------
constexpr int kMaxBands = 64;
struct BandArgs {
    int n_bands;
    int rows[kMaxBands + 1];          // band b = image rows [rows[b], rows[b + 1])
};

// Every in-view localisation goes to each band its window can reach (the same conservative
// reach as render_bin_kernel: 3 sigma + 2 rows); mode 0 (histogram) has no blur.
__device__ __forceinline__ void band_range(const RenderArgs& a, const BandArgs& b, long long k, int& b0, int& b1) {
    const double xv = (double)a.x[k], yv = (double)a.y[k];
    b0 = 0; b1 = -1;
    if (!((xv > a.x_min) && (yv > a.y_min) && (xv < a.x_max) && (yv < a.y_max))) return;
    const double y_ = a.os * (yv - a.y_min);
    double reach = 0.0;
    if (a.mode) {
        const float sg = __fmul_rn(a.osf, fmaxf(fmaxf(a.lpx[k], a.lpy[k]), a.mbw));
        reach = 3.0 * (double)sg + 2.0;
    }
    int py = (int)y_;
    py = min(max(py, 0), a.npy - 1);
    b0 = b.n_bands; b1 = -1;
    for (int q = 0; q < b.n_bands; q++) {
        const bool owns = py >= b.rows[q] && py < b.rows[q + 1];
        const bool reaches = b.rows[q + 1] > b.rows[q] && y_ + reach >= (double)b.rows[q] &&
                             y_ - reach <= (double)b.rows[q + 1];
        if (owns || reaches) { b0 = min(b0, q); b1 = max(b1, q); }
    }
}

// counts per band: warp-aggregated (one ballot per band) into shared counters, one global atomic
// per band and CTA
__global__ void __launch_bounds__(256) render_band_count_kernel(const RenderArgs a, const BandArgs b,
                                                                unsigned long long* __restrict__ counts) {
    __shared__ unsigned int sc[kMaxBands];
    for (int q = threadIdx.x; q < kMaxBands; q += blockDim.x) sc[q] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long rounds = (a.n + stride - 1) / stride;
    for (long long r = 0; r < rounds; r++) {
        const long long k = r * stride + blockIdx.x * (long long)blockDim.x + threadIdx.x;
        int b0 = 0, b1 = -1;
        if (k < a.n) band_range(a, b, k, b0, b1);
        for (int q = 0; q < b.n_bands; q++) {
            const unsigned m = __ballot_sync(0xffffffffu, q >= b0 && q <= b1);
            if (lane == 0 && m) atomicAdd(&sc[q], (unsigned)__popc(m));
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < b.n_bands; q += blockDim.x)
        if (sc[q]) atomicAdd(counts + q, (unsigned long long)sc[q]);
}

// Append (x, y, lpx, lpy) records to the destination band's slice of the send buffer.  Per tile of
// 256 localisations the CTA ranks its records per band in shared memory (warp ballots), reserves one
// range per band with a single global atomic and writes the float4 records -- 2 * n_bands global
// atomics per 256 localisations instead of one per record on two hot addresses.
__global__ void __launch_bounds__(256) render_band_scatter_kernel(const RenderArgs a, const BandArgs b,
                                                                  const unsigned long long* __restrict__ offsets,
                                                                  unsigned long long* __restrict__ cursor,
                                                                  float4* __restrict__ out) {
    __shared__ unsigned int warp_cnt[kMaxBands][8];      // per band, per warp of the CTA
    __shared__ unsigned long long base[kMaxBands];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long rounds = (a.n + stride - 1) / stride;
    for (long long r = 0; r < rounds; r++) {
        const long long k = r * stride + blockIdx.x * (long long)blockDim.x + threadIdx.x;
        int b0 = 0, b1 = -1;
        if (k < a.n) band_range(a, b, k, b0, b1);
        for (int q = 0; q < b.n_bands; q++) {
            const unsigned m = __ballot_sync(0xffffffffu, q >= b0 && q <= b1);
            if (lane == 0) warp_cnt[q][warp] = __popc(m);
        }
        __syncthreads();
        if (threadIdx.x < b.n_bands) {
            const int q = threadIdx.x;
            unsigned tot = 0;
            for (int w = 0; w < 8; w++) { const unsigned c = warp_cnt[q][w]; warp_cnt[q][w] = tot; tot += c; }
            base[q] = tot ? offsets[q] + atomicAdd(cursor + q, (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < a.n && b1 >= b0) rec = make_float4(a.x[k], a.y[k], a.mode ? a.lpx[k] : 0.f, a.mode ? a.lpy[k] : 0.f);
        for (int q = 0; q < b.n_bands; q++) {           // same ballots again: rank inside the warp
            const bool in = q >= b0 && q <= b1;
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (in) out[base[q] + warp_cnt[q][warp] + __popc(m & ((1u << lane) - 1u))] = rec;
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int pb_render_band_count_dev(size_t n, const float* d_x, const float* d_y, const float* d_lpx,
                                        const float* d_lpy, double oversampling, double y_min, double x_min,
                                        double y_max, double x_max, double min_blur_width, int mode,
                                        int n_pixel_y, int n_pixel_x, int n_bands, const int* band_rows,
                                        unsigned long long* d_counts, void* stream) {
    if (mode < 0 || mode > 2) { pb_set_error("blur_method not understood."); return PB_ERR_INVALID; }
    if (n_bands < 1 || n_bands > kMaxBands || !band_rows || !d_counts) { pb_set_error("pb_render_band_count_dev: bad band list"); return PB_ERR_INVALID; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    PB_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, 8 * (size_t)n_bands, s));
    if (n == 0) return PB_OK;
    RenderArgs a{d_x, d_y, d_lpx, d_lpy, (long long)n, oversampling, y_min, x_min, y_max, x_max,
                 (float)oversampling, (float)min_blur_width, mode, nullptr, n_pixel_y, n_pixel_x, nullptr, 0, n_pixel_y};
    BandArgs b;
    b.n_bands = n_bands;
    for (int q = 0; q <= n_bands; q++) b.rows[q] = band_rows[q];
    const int grid = (int)std::min<long long>(((long long)n + 255) / 256, 148 * 16);
    render_band_count_kernel<<<grid, 256, 0, s>>>(a, b, d_counts);
    g_pb_launches++;
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_render_band_scatter_dev(size_t n, const float* d_x, const float* d_y, const float* d_lpx,
                                          const float* d_lpy, double oversampling, double y_min, double x_min,
                                          double y_max, double x_max, double min_blur_width, int mode,
                                          int n_pixel_y, int n_pixel_x, int n_bands, const int* band_rows,
                                          const unsigned long long* d_offsets, unsigned long long* d_cursor,
                                          float* d_records, void* stream) {
    if (mode < 0 || mode > 2) { pb_set_error("blur_method not understood."); return PB_ERR_INVALID; }
    if (n_bands < 1 || n_bands > kMaxBands || !band_rows || !d_offsets || !d_cursor) { pb_set_error("pb_render_band_scatter_dev: bad band list"); return PB_ERR_INVALID; }
    if (reinterpret_cast<uintptr_t>(d_records) & 15) { pb_set_error("pb_render_band_scatter_dev: records must be 16-byte aligned"); return PB_ERR_INVALID; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    PB_CUDA_CHECK(cudaMemsetAsync(d_cursor, 0, 8 * (size_t)n_bands, s));
    if (n == 0) return PB_OK;
    RenderArgs a{d_x, d_y, d_lpx, d_lpy, (long long)n, oversampling, y_min, x_min, y_max, x_max,
                 (float)oversampling, (float)min_blur_width, mode, nullptr, n_pixel_y, n_pixel_x, nullptr, 0, n_pixel_y};
    BandArgs b;
    b.n_bands = n_bands;
    for (int q = 0; q <= n_bands; q++) b.rows[q] = band_rows[q];
    const int grid = (int)std::min<long long>(((long long)n + 255) / 256, 148 * 8);
    render_band_scatter_kernel<<<grid, 256, 0, s>>>(a, b, d_offsets, d_cursor, reinterpret_cast<float4*>(d_records));
    g_pb_launches++;
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

// (x, y, lpx, lpy) float4 records -> the four float32 columns the renderer reads
__global__ void __launch_bounds__(256) render_unpack_kernel(const float4* __restrict__ rec, long long n,
                                                            float* __restrict__ x, float* __restrict__ y,
                                                            float* __restrict__ lpx, float* __restrict__ lpy) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const float4 r = rec[k];
        x[k] = r.x; y[k] = r.y; lpx[k] = r.z; lpy[k] = r.w;
    }
}
extern "C" int pb_render_unpack_records_dev(size_t n, const float* d_records, float* d_x, float* d_y, float* d_lpx,
                                            float* d_lpy, void* stream) {
    if (n == 0) return PB_OK;
    if (!d_records || !d_x || !d_y || !d_lpx || !d_lpy || (reinterpret_cast<uintptr_t>(d_records) & 15)) {
        pb_set_error("pb_render_unpack_records_dev: bad pointer");
        return PB_ERR_INVALID;
    }
    const int grid = (int)std::min<long long>(((long long)n + 255) / 256, 148 * 16);
    render_unpack_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(d_records), (long long)n, d_x, d_y, d_lpx, d_lpy);
    g_pb_launches++;
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

// Workspace (bytes) pb_render_dev needs for the binned paths: up to two float4 records per localisation
// (strip path: a window straddling two strips), one int32 bin word per localisation, and count / start /
// cursor per bin (32 window-size classes per 64x64 tile, or 16-row x 64-column strips).
extern "C" size_t pb_render_workspace_bytes(size_t n, int n_pixel_y, int n_pixel_x) {
    const size_t tiles_x = (size_t)((n_pixel_x + kTile - 1) / kTile);
    const size_t strips = (size_t)((n_pixel_y + kStripRows - 1) / kStripRows) * tiles_x;
    const size_t tiles4 = (size_t)((n_pixel_y + kTile - 1) / kTile) * tiles_x * kSizeClasses;
    return n * 36 + (std::max(strips, tiles4) + 2) * 12 + 256 + 256 + kScanTempBytes;
}

// Accumulation pass of the binned path: 0 = 64x64 tiles, one thread per localisation, shared-memory
// atomicAdd (default); 1 = warp-owned 16-row strips, plain load / add / store (no shared-memory atomics).
// Measured on B200, 50 M localisations -> 10240^2 (profiles/r02_summary.md): 7.15 ms vs 12.4 ms -- the
// strip pass needs 205 warp instructions per localisation (8x8 blocks of 3..11-pixel windows keep 38 % of
// the lanes busy, ~60 instructions of geometry, kernel evaluation, shuffles and index arithmetic per
// block) against 141 for the atomic pass, so removing the atomics does not pay.  Kept selectable
// (pb_render_set_impl / PB_RENDER_IMPL) so the A/B stays reproducible.
static std::atomic<int> g_render_impl{-1};
static int render_impl() {
    int v = g_render_impl.load();
    if (v < 0) {
        v = 0;
        if (const char* e = getenv("PB_RENDER_IMPL")) v = atoi(e) == 1 ? 1 : 0;
        g_render_impl.store(v);
    }
    return v;
}
extern "C" int pb_render_set_impl(int impl) {
    if (impl != 0 && impl != 1) { pb_set_error("pb_render_set_impl: impl must be 0 or 1"); return PB_ERR_INVALID; }
    g_render_impl.store(impl);
    return PB_OK;
}
extern "C" int pb_render_get_impl(void) { return render_impl(); }

extern "C" int pb_render_dev(size_t n, const float* d_x, const float* d_y, const float* d_lpx,
                             const float* d_lpy, double oversampling, double y_min, double x_min,
                             double y_max, double x_max, double min_blur_width, int mode,
                             float* d_image, int n_pixel_y, int n_pixel_x,
                             unsigned long long* d_count, void* d_workspace,
                             size_t workspace_bytes, void* stream) {
    return pb_render_band_dev(n, d_x, d_y, d_lpx, d_lpy, oversampling, y_min, x_min, y_max, x_max,
                              min_blur_width, mode, d_image, n_pixel_y, n_pixel_x, 0, n_pixel_y, d_count,
                              d_workspace, workspace_bytes, stream);
}

// Row band [row0, row0 + n_rows) of the same image: `d_image` holds n_rows x n_pixel_x floats.
// Multi-GPU rendering shards the image by such bands (SURVEY.md 8e option B): every rank draws
// the localisations whose 3-sigma window reaches its band, clipped to the band, and counts those
// whose centre row it owns -- the bands concatenate to the full image and the counts add up to n.
extern "C" int pb_render_band_dev(size_t n, const float* d_x, const float* d_y, const float* d_lpx,
                                  const float* d_lpy, double oversampling, double y_min, double x_min,
                                  double y_max, double x_max, double min_blur_width, int mode,
                                  float* d_image, int n_pixel_y, int n_pixel_x, int row0, int n_rows,
                                  unsigned long long* d_count, void* d_workspace,
                                  size_t workspace_bytes, void* stream) {
    if (mode < 0 || mode > 2) { pb_set_error("blur_method not understood."); return PB_ERR_INVALID; }
    if (n_pixel_y <= 0 || n_pixel_x <= 0) return PB_OK;
    if (row0 < 0 || n_rows < 0 || row0 + n_rows > n_pixel_y) { pb_set_error("pb_render_band_dev: bad row band"); return PB_ERR_INVALID; }
    if (!d_count || (n_rows && !d_image)) { pb_set_error("pb_render_dev: null pointer"); return PB_ERR_INVALID; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    PB_CUDA_CHECK(cudaMemsetAsync(d_count, 0, 8, s));
    if (n_rows == 0) return PB_OK;
    PB_CUDA_CHECK(cudaMemsetAsync(d_image, 0, sizeof(float) * (size_t)n_rows * n_pixel_x, s));
    if (n == 0) return PB_OK;
    RenderArgs a{d_x, d_y, d_lpx, d_lpy, (long long)n, oversampling, y_min, x_min, y_max, x_max,
                 (float)oversampling, (float)min_blur_width, mode, d_image, n_pixel_y, n_pixel_x,
                 d_count, row0, n_rows};
    const int threads = 256;
    if (mode == 0) {
        int grid = (int)std::min<long long>(((long long)n + threads - 1) / threads, 148 * 16);
        render_hist_kernel<<<grid, threads, 0, s>>>(a);
        g_pb_launches++;
        PB_CUDA_CHECK(cudaGetLastError());
        return PB_OK;
    }
    const int tiles_x = (n_pixel_x + kTile - 1) / kTile, tiles_y = (n_rows + kTile - 1) / kTile;
    const long long ntiles = (long long)tiles_x * tiles_y;
    const size_t need = pb_render_workspace_bytes(n, n_rows, n_pixel_x);
    // The tiled pass costs ~10 ns per 64 x 64 tile (zeroing and flushing its shared copy, mostly idle
    // CTAs) plus 0.13 ns per localisation; the direct splat 0.32 ns per localisation.  Sparse images --
    // the 100 k-localisation segments of an undrift run on 4096^2 pixels: 24 per tile -- are faster
    // direct (measured: 0.24 -> 0.04 ms per segment, 48 -> 7.6 ms for the 200 segments of config 5); dense
    // ones (config 4: 1953 per tile) tiled.
    const bool tiled = d_workspace && workspace_bytes >= need && n >= 65536 && n < 0x7fffffffull &&
                       ntiles >= 64 && (long long)n >= 64 * ntiles &&
                       ((reinterpret_cast<uintptr_t>(d_workspace) & 15) == 0);
    if (!tiled) {
        long long want = ((long long)n * kLanes + threads - 1) / threads;
        int grid = (int)std::min<long long>(want, 148 * 16);
        render_splat_kernel<<<grid, threads, 0, s>>>(a);
        g_pb_launches++;
        PB_CUDA_CHECK(cudaGetLastError());
        return PB_OK;
    }
    // workspace layout: sorted[2n] float4 | tile_of[n] i32 | count[B+1] | start[B+1] | cursor[B] | scan temp
    char* w = static_cast<char*>(d_workspace);
    float4* sorted = reinterpret_cast<float4*>(w);
    int* tile_of = reinterpret_cast<int*>(w + n * 32);
    unsigned int* tcount = reinterpret_cast<unsigned int*>(w + n * 36);
    int grid = (int)std::min<long long>(((long long)n + threads - 1) / threads, 148 * 16);
    // bins: count[B + 1] (last entry stays 0) | start[B + 1] | cursor[B] | scan temp (256-byte aligned)
    const int strips_y = (n_rows + kStripRows - 1) / kStripRows;
    const long long nbins = render_impl() == 0 ? ntiles * kSizeClasses : (long long)strips_y * tiles_x;
    unsigned int* tstart = tcount + nbins + 1;
    unsigned int* tcursor = tstart + nbins + 1;
    void* scan_temp = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(tcursor + nbins) + 255) & ~(uintptr_t)255);
    PB_CUDA_CHECK(cudaMemsetAsync(tcount, 0, (nbins + 1) * 4, s));
    int rc;
    if (render_impl() == 0) {
        render_bin_kernel<<<grid, threads, 0, s>>>(a, tiles_x, tile_of, tcount);
        if ((rc = render_scan(tcount, tstart, tcursor, nbins, scan_temp, s)) != PB_OK) return rc;
        render_scatter_kernel<<<grid, threads, 0, s>>>(a, tile_of, tcursor, sorted);
        render_tiled_kernel<<<(unsigned)ntiles, 256, 0, s>>>(a, tiles_x, tstart, sorted);
    } else {
        render_strip_bin_kernel<<<grid, threads, 0, s>>>(a, tiles_x, tile_of, tcount);
        if ((rc = render_scan(tcount, tstart, tcursor, nbins, scan_temp, s)) != PB_OK) return rc;
        render_strip_scatter_kernel<<<grid, threads, 0, s>>>(a, tiles_x, tile_of, tcursor, sorted);
        render_strip_kernel<<<(unsigned)ntiles, 128, 0, s>>>(a, tiles_x, strips_y, tstart, sorted);
    }
    g_pb_launches += 5;      // bin, scan (2 cub kernels), scatter, accumulate
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

// Host-buffer variant: H2D of (x, y, lpx, lpy), render, D2H of the image.
extern "C" int pb_render(size_t n, const float* x, const float* y, const float* lpx,
                         const float* lpy, double oversampling, double y_min, double x_min,
                         double y_max, double x_max, double min_blur_width, int mode, float* image,
                         int n_pixel_y, int n_pixel_x, long long* n_in_view) {
    if (mode < 0 || mode > 2) { pb_set_error("blur_method not understood."); return PB_ERR_INVALID; }
    if (n_in_view) *n_in_view = 0;
    const size_t npix = (size_t)std::max(n_pixel_y, 0) * std::max(n_pixel_x, 0);
    if (npix == 0) return PB_OK;
    if (!image || (n && (!x || !y)) || (n && mode && (!lpx || !lpy))) {
        pb_set_error("pb_render: null pointer");
        return PB_ERR_INVALID;
    }
    float *dx = nullptr, *dy = nullptr, *dlx = nullptr, *dly = nullptr, *dimg = nullptr;
    unsigned long long* dcnt = nullptr;
    void* dws = nullptr;
    const size_t nb = std::max<size_t>(n, 1) * 4;
    const size_t wsb = pb_render_workspace_bytes(n, n_pixel_y, n_pixel_x);
    int rc = PB_OK;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t err) { if (err != cudaSuccess && e == cudaSuccess) e = err; return err == cudaSuccess; };
    ok(cudaMalloc(&dx, nb)); ok(cudaMalloc(&dy, nb));
    if (mode) { ok(cudaMalloc(&dlx, nb)); ok(cudaMalloc(&dly, nb)); ok(cudaMalloc(&dws, wsb)); }
    ok(cudaMalloc(&dimg, npix * 4)); ok(cudaMalloc(&dcnt, 8));
    if (e == cudaSuccess && n) {
        // pageable numpy columns: threaded pinned staging (csrc/transfer.cu)
        if (rc == PB_OK) rc = pb_h2d(dx, x, n * 4, nullptr);
        if (rc == PB_OK) rc = pb_h2d(dy, y, n * 4, nullptr);
        if (mode) {
            if (rc == PB_OK) rc = pb_h2d(dlx, lpx, n * 4, nullptr);
            if (rc == PB_OK) rc = pb_h2d(dly, lpy, n * 4, nullptr);
        }
    }
    if (e == cudaSuccess && rc == PB_OK)
        rc = pb_render_dev(n, dx, dy, dlx, dly, oversampling, y_min, x_min, y_max, x_max,
                           min_blur_width, mode, dimg, n_pixel_y, n_pixel_x, dcnt, dws, wsb, nullptr);
    unsigned long long cnt = 0;
    if (e == cudaSuccess && rc == PB_OK) {
        rc = pb_d2h(image, dimg, npix * 4, nullptr);
        ok(cudaMemcpy(&cnt, dcnt, 8, cudaMemcpyDeviceToHost));
    }
    cudaFree(dx); cudaFree(dy); cudaFree(dlx); cudaFree(dly); cudaFree(dimg); cudaFree(dcnt); cudaFree(dws);
    if (e != cudaSuccess) { pb_set_error("pb_render: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    if (n_in_view) *n_in_view = (long long)cnt;
    return rc;
}
