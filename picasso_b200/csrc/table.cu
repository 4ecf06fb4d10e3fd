// picasso_b200/csrc/table.cu -- localisation-table post-processing on the device (sm_100a).
//
// SURVEY.md 8(f) rank 2, the output side of the hot path:
//   lib.ensure_sanity           (picasso/lib.py:1786-1832)  inf / NaN rows out, x < Width, y < Height,
//                                                           x, y, lpx, lpy, lpz, photons, ellipticity,
//                                                           sx, sy >= 0
//   zfit._fit_z tail            (picasso/zfit.py:356-383)   z / d_zcalib / lpz columns appended, then
//   zfit.filter_z_fits          (picasso/zfit.py:675-704)   d_zcalib <= range * sqrt(nanmean(d_zcalib^2))
//   io.save_locs record packing (picasso/io.py:2089-2110)   locs.to_records(index=False)
//
// The reference copies the whole table for each of its eleven boolean filters; the Python layer of
// round 1 still built one mask and gathered every column on the host (0.8-1.4 s for 10 M rows
// around a 1.3 ms z-fit kernel).  Here the table crosses PCIe once in each direction: all 4-byte
// columns are uploaded, the (optional) z fit runs on the resident columns, one kernel evaluates the
// sanity mask, the kept row indices come from a CUB select, the RMSD threshold reproduces numpy's
// float32 pairwise summation bit for bit (leaf sums on the device, the combine tree on the host), and
// one gather writes the compacted columns -- or the packed records -- that are downloaded.
#include <algorithm>
#include <atomic>
#include <cub/device/device_select.cuh>
#include <math.h>
#include <vector>

#include "pb_common.cuh"
#include "../../include/picasso_b200.h"

extern std::atomic<long long> g_pb_launches;

namespace {

constexpr int kMaxCols = 64;

struct TableArgs {
    const unsigned* cols;      // (ncols, ld) 4-byte elements on the device
    long long n, ld;
    int ncols;
    unsigned long long float_mask;    // bit k: column k is float32 (checked for inf / NaN)
    unsigned long long nonneg_mask;   // bit k: keep only rows with column k >= 0
    int ix, iy;                       // x / y columns (-1: absent)
    float width, height;
};

// pandas compares the float32 column with the Python number from the metadata: float32 < float64
// in numpy 2 promotes per NEP 50 to float32 for Python scalars -- the comparison runs in float32.
__global__ void __launch_bounds__(256) sanity_mask_kernel(const TableArgs a, unsigned char* __restrict__ keep) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    bool ok = true;
    for (int k = 0; k < a.ncols; k++) {
        if (!((a.float_mask >> k) & 1ull)) continue;
        const float v = __uint_as_float(a.cols[(long long)k * a.ld + i]);
        ok = ok && isfinite(v);
        if ((a.nonneg_mask >> k) & 1ull) ok = ok && (v >= 0.0f);
        if (k == a.ix) ok = ok && (v < a.width);
        if (k == a.iy) ok = ok && (v < a.height);
    }
    keep[i] = ok ? 1 : 0;
}

__global__ void iota_kernel(long long n, long long* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum_@TYPE@) splits
// [0, n) recursively at n/2 rounded down to a multiple of 8 until a block has <= 128 elements; a block
// of >= 8 elements is summed with eight strided accumulators combined as
// ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)) plus the tail in order.  One thread per leaf; the
// leaf boundaries come from the host (they depend on n only).
__global__ void __launch_bounds__(128) pairwise_leaf_kernel(const float* __restrict__ col,
                                                            const long long* __restrict__ idx,
                                                            const long long* __restrict__ leaf_start,
                                                            long long n_leaves, float* __restrict__ out) {
    const long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (l >= n_leaves) return;
    const long long b = leaf_start[l], e = leaf_start[l + 1];
    const long long n = e - b;
    auto sq = [&](long long k) { const float v = col[idx[b + k]]; return __fmul_rn(v, v); };
    float res;
    if (n < 8) {
        res = 0.0f;
        for (long long k = 0; k < n; k++) res = __fadd_rn(res, sq(k));
    } else {
        float r[8];
        for (int q = 0; q < 8; q++) r[q] = sq(q);
        long long k = 8;
        for (; k < n - (n % 8); k += 8)
            for (int q = 0; q < 8; q++) r[q] = __fadd_rn(r[q], sq(k + q));
        res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; k < n; k++) res = __fadd_rn(res, sq(k));
    }
    out[l] = res;
}

void pairwise_leaves(long long b, long long n, std::vector<long long>& starts) {
    if (n <= 128) { starts.push_back(b); return; }
    long long n2 = n / 2;
    n2 -= n2 % 8;
    pairwise_leaves(b, n2, starts);
    pairwise_leaves(b + n2, n - n2, starts);
}
float pairwise_combine(long long n, const float*& leaf) {
    if (n <= 128) return *leaf++;
    long long n2 = n / 2;
    n2 -= n2 % 8;
    const float a = pairwise_combine(n2, leaf);
    const float b = pairwise_combine(n - n2, leaf);
    return a + b;      // (host code is compiled without -ffast-math: one IEEE float32 add)
}

__global__ void __launch_bounds__(256) thresh_mask_kernel(const float* __restrict__ col,
                                                          const long long* __restrict__ idx, long long n,
                                                          float thresh, unsigned char* __restrict__ keep) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) keep[i] = (col[idx[i]] <= thresh) ? 1 : 0;
}

// out[k][i] = cols[k][idx[i]] (column block) or out[i][k] (packed records)
__global__ void __launch_bounds__(256) gather_kernel(const unsigned* __restrict__ cols, long long ld, int ncols,
                                                     const long long* __restrict__ idx, long long m,
                                                     unsigned* __restrict__ out, long long out_ld, int records) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m) return;
    const long long src = idx[i];
    if (records) {
        for (int k = 0; k < ncols; k++) out[i * ncols + k] = cols[(long long)k * ld + src];
    } else {
        for (int k = 0; k < ncols; k++) out[(long long)k * out_ld + i] = cols[(long long)k * ld + src];
    }
}

struct Dev {
    void* p = nullptr;
    ~Dev() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { PB_CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 1)); return PB_OK; }
};

int select_flagged(const long long* d_in, const unsigned char* d_flags, long long n, long long* d_out,
                   long long* d_count, cudaStream_t s) {
    size_t tb = 0;
    cub::DeviceSelect::Flagged(nullptr, tb, d_in, d_flags, d_out, d_count, n, s);
    Dev tmp;
    int rc = tmp.alloc(tb);
    if (rc) return rc;
    if (cub::DeviceSelect::Flagged(tmp.p, tb, d_in, d_flags, d_out, d_count, n, s) != cudaSuccess) {
        pb_set_error("pb_locs_filter: select failed: %s", cudaGetErrorString(cudaGetLastError()));
        return PB_ERR_CUDA;
    }
    g_pb_launches++;
    PB_CUDA_CHECK(cudaStreamSynchronize(s));      // tmp is freed on return
    return PB_OK;
}

}  // namespace

extern "C" int pb_locs_filter(size_t n, int ncols_in, const void* const* cols, const int* is_float, int ix,
                              int iy, const int* nonneg_cols, int n_nonneg, double width, double height,
                              const PbZfitSpec* zfit, int out_records, void* out, size_t capacity,
                              long long* kept_index, size_t* n_kept) {
    if (!n_kept) { pb_set_error("pb_locs_filter: n_kept is null"); return PB_ERR_INVALID; }
    *n_kept = 0;
    const int extra = zfit ? 3 : 0;
    const int ncols = ncols_in + extra;
    if (ncols_in < 1 || ncols > kMaxCols) { pb_set_error("pb_locs_filter: 1..%d columns", kMaxCols - 3); return PB_ERR_INVALID; }
    if (n == 0) return PB_OK;
    if (!cols || !is_float || (capacity && !out)) { pb_set_error("pb_locs_filter: null pointer"); return PB_ERR_INVALID; }
    if (ix >= ncols || iy >= ncols) { pb_set_error("pb_locs_filter: x / y column out of range"); return PB_ERR_INVALID; }
    const long long N = (long long)n;
    Dev block, keep, idx0, idx1, idx2, cnt;
    int rc;
    if ((rc = block.alloc((size_t)ncols * n * 4)) || (rc = keep.alloc(n)) || (rc = idx0.alloc(n * 8)) ||
        (rc = idx1.alloc(n * 8)) || (rc = cnt.alloc(8)))
        return rc;
    unsigned* d_cols = static_cast<unsigned*>(block.p);
    cudaStream_t s = nullptr;
    for (int k = 0; k < ncols_in; k++) {
        if (!cols[k]) { pb_set_error("pb_locs_filter: column %d is null", k); return PB_ERR_INVALID; }
        if ((rc = pb_h2d(d_cols + (size_t)k * n, cols[k], n * 4, s)) != PB_OK) return rc;
    }
    TableArgs a{};
    a.cols = d_cols; a.n = N; a.ld = N; a.ncols = ncols;
    a.ix = ix; a.iy = iy; a.width = (float)width; a.height = (float)height;
    for (int k = 0; k < ncols_in; k++) if (is_float[k]) a.float_mask |= 1ull << k;
    for (int q = 0; q < n_nonneg; q++) {
        if (nonneg_cols[q] < 0 || nonneg_cols[q] >= ncols) { pb_set_error("pb_locs_filter: bad non-negative column"); return PB_ERR_INVALID; }
        a.nonneg_mask |= 1ull << nonneg_cols[q];
    }
    int i_dz = -1;
    if (zfit) {
        // z, d_zcalib, lpz become the last three columns (zfit.py:356-358); lpz joins the >= 0 list
        const PbZfitSpec& z = *zfit;
        const int need[4] = {z.i_sx, z.i_sy, z.i_photons, z.i_bg};
        for (int q = 0; q < 4; q++)
            if (need[q] < 0 || need[q] >= ncols_in) { pb_set_error("pb_locs_filter: z-fit input column out of range"); return PB_ERR_INVALID; }
        auto colp = [&](int k) { return k >= 0 ? reinterpret_cast<const float*>(d_cols + (size_t)k * n) : nullptr; };
        float* dz = reinterpret_cast<float*>(d_cols + (size_t)ncols_in * n);
        rc = pb_zfit_dev(n, colp(z.i_sx), colp(z.i_sy), colp(z.i_photons), colp(z.i_bg), colp(z.i_sx_unc),
                         colp(z.i_sy_unc), z.cx, z.cy, z.magnification, z.pixelsize, z.method, dz, dz + n,
                         dz + 2 * n, nullptr, s);
        if (rc != PB_OK) return rc;
        for (int k = ncols_in; k < ncols; k++) a.float_mask |= 1ull << k;
        a.nonneg_mask |= 1ull << (ncols_in + 2);
        i_dz = ncols_in + 1;
    }
    const unsigned g = (unsigned)((N + 255) / 256);
    sanity_mask_kernel<<<g, 256, 0, s>>>(a, static_cast<unsigned char*>(keep.p));
    iota_kernel<<<g, 256, 0, s>>>(N, static_cast<long long*>(idx0.p));
    g_pb_launches += 2;
    PB_CUDA_CHECK(cudaGetLastError());
    if ((rc = select_flagged(static_cast<long long*>(idx0.p), static_cast<unsigned char*>(keep.p), N,
                             static_cast<long long*>(idx1.p), static_cast<long long*>(cnt.p), s)))
        return rc;
    long long m = 0;
    PB_CUDA_CHECK(cudaMemcpy(&m, cnt.p, 8, cudaMemcpyDeviceToHost));
    const long long* d_idx = static_cast<long long*>(idx1.p);
    if (zfit && zfit->filter_range > 0 && m > 0) {
        // rmsd = np.sqrt(np.nanmean(locs["d_zcalib"] ** 2)) on the sane rows, float32 like numpy
        std::vector<long long> starts;
        pairwise_leaves(0, m, starts);
        starts.push_back(m);
        const long long nl = (long long)starts.size() - 1;
        Dev dstart, dleaf;
        if ((rc = dstart.alloc(starts.size() * 8)) || (rc = dleaf.alloc(nl * 4))) return rc;
        PB_CUDA_CHECK(cudaMemcpy(dstart.p, starts.data(), starts.size() * 8, cudaMemcpyHostToDevice));
        const float* dzc = reinterpret_cast<const float*>(d_cols + (size_t)i_dz * n);
        pairwise_leaf_kernel<<<(unsigned)((nl + 127) / 128), 128, 0, s>>>(dzc, d_idx, static_cast<long long*>(dstart.p),
                                                                          nl, static_cast<float*>(dleaf.p));
        g_pb_launches++;
        std::vector<float> leaves(nl);
        PB_CUDA_CHECK(cudaMemcpy(leaves.data(), dleaf.p, nl * 4, cudaMemcpyDeviceToHost));
        const float* lp = leaves.data();
        const float total = pairwise_combine(m, lp);
        // np.nanmean: np.float32 sum / np.intp count is evaluated in float64 and cast back to float32
        const float mean = (float)((double)total / (double)m);
        const float rmsd = sqrtf(mean);
        const float thresh = (float)zfit->filter_range * rmsd;   // python int * np.float32 -> float32
        if ((rc = idx2.alloc((size_t)m * 8))) return rc;
        thresh_mask_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(dzc, d_idx, m, thresh,
                                                                       static_cast<unsigned char*>(keep.p));
        g_pb_launches++;
        if ((rc = select_flagged(d_idx, static_cast<unsigned char*>(keep.p), m, static_cast<long long*>(idx2.p),
                                 static_cast<long long*>(cnt.p), s)))
            return rc;
        PB_CUDA_CHECK(cudaMemcpy(&m, cnt.p, 8, cudaMemcpyDeviceToHost));
        d_idx = static_cast<long long*>(idx2.p);
    }
    *n_kept = (size_t)m;
    if ((size_t)m > capacity) {
        pb_set_error("pb_locs_filter: %lld rows kept, capacity %zu", m, capacity);
        return PB_ERR_CAPACITY;
    }
    if (m == 0) return PB_OK;
    Dev outd;
    if ((rc = outd.alloc((size_t)ncols * m * 4))) return rc;
    gather_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(d_cols, N, ncols, d_idx, m, static_cast<unsigned*>(outd.p),
                                                              m, out_records ? 1 : 0);
    g_pb_launches++;
    PB_CUDA_CHECK(cudaGetLastError());
    if (out_records || capacity == (size_t)m) {
        if ((rc = pb_d2h(out, outd.p, (size_t)ncols * m * 4, s)) != PB_OK) return rc;
    } else {
        for (int k = 0; k < ncols; k++)
            if ((rc = pb_d2h(static_cast<char*>(out) + (size_t)k * capacity * 4,
                             static_cast<char*>(outd.p) + (size_t)k * m * 4, (size_t)m * 4, s)) != PB_OK)
                return rc;
    }
    if (kept_index && (rc = pb_d2h(kept_index, d_idx, (size_t)m * 8, s)) != PB_OK) return rc;
    PB_CUDA_CHECK(cudaStreamSynchronize(s));
    return PB_OK;
}
