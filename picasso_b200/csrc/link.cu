// picasso_b200/csrc/link.cu -- linking localizations into binding events (sm_100a).
//
// Replaces the numba loops behind picasso.postprocess.link (reference picasso/postprocess.py):
// `_get_link_groups` :2440-2507 / `_get_next_loc_index_in_link_group` :2510-2552 and the
// per-group reductions `_link_group_count/sum/min_max/last` :2555-2661 used by
// `_link_loc_groups` :2680-2821.
//
// The reference is a sequential greedy: localizations are visited in frame order; an unlinked
// one starts a chain, and the chain repeatedly takes the FIRST not yet linked localization of the
// same `group` within d_max in the next max_dark_time + 1 frames.  Every such step follows an
// edge of the graph "j is a candidate successor of i", so chains never leave a connected
// component of that graph, and the greedy restricted to a component (its members in index order)
// reproduces the global result exactly.  Hence:
//   1. edges kernel: one thread per localization scans its candidate index window (frames are
//      sorted, so the window is a contiguous index range) and unions itself with every candidate
//      (lock-free union-find, smaller index wins);
//   2. labels are flattened, localizations radix-sorted by component (stable: index order kept);
//   3. greedy kernel: one thread per component runs the reference's loop over its members;
//   4. chain starts are numbered by an exclusive scan in index order = the reference's
//      `current_link_group` counter.
// Distances use the coordinate dtype (float32 or float64) with the reference's three-step test
// dx^2 <= d^2, dy^2 <= d^2, dx^2 + dy^2 <= d^2 against the float64 d_max^2.  The reference's
// end-of-data quirk (when no later frame exists, `min_index` stays at N-1, so the very last
// localization is a candidate even within the same frame) is reproduced.
// Integer results are bit-exact; the per-group sums accumulate sequentially in the column's
// dtype in localization order like the numba loops.
#include <algorithm>
#include <atomic>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <vector>

#include "pb_common.cuh"
#include "../../include/picasso_b200.h"

extern std::atomic<long long> g_pb_launches;

namespace {

struct LinkArgs {
    long long n;
    const long long* frame;        // sorted ascending
    const void* x;
    const void* y;
    int f64;
    const int* group;
    double d2;                     // d_max^2
    long long max_dark;
    long long fmin, fmax;          // first / last frame value
    const long long* fstart;       // fstart[f - fmin] = first index with frame >= f, f in [fmin, fmax + 1]
};

template <typename T>
__device__ __forceinline__ bool link_near(const T* x, const T* y, long long i, long long j, double d2) {
    const T dx = x[i] - x[j];
    const T dx2 = dx * dx;
    if (!((double)dx2 <= d2)) return false;
    const T dy = y[i] - y[j];
    const T dy2 = dy * dy;
    if (!((double)dy2 <= d2)) return false;
    return (double)(dx2 + dy2) <= d2;
}

__device__ __forceinline__ bool link_is_near(const LinkArgs& a, long long i, long long j) {
    return a.f64 ? link_near(static_cast<const double*>(a.x), static_cast<const double*>(a.y), i, j, a.d2)
                 : link_near(static_cast<const float*>(a.x), static_cast<const float*>(a.y), i, j, a.d2);
}

// candidate index window [lo, hi) of localization i (postprocess.py:2528-2540)
__device__ __forceinline__ void link_window(const LinkArgs& a, long long i, long long* lo, long long* hi) {
    const long long f = a.frame[i];
    if (i + 1 >= a.n) { *lo = *hi = 0; return; }
    if (f >= a.fmax) { *lo = a.n - 1; *hi = a.n; return; }          // no later frame: min_index stays N - 1
    *lo = a.fstart[f + 1 - a.fmin];
    const long long fm = f + a.max_dark + 1;                          // last admissible frame
    *hi = fm >= a.fmax ? a.n : a.fstart[fm + 1 - a.fmin];
}

__device__ __forceinline__ int uf_find(int* parent, int v) {
    int p = parent[v];
    while (p != v) {
        const int gp = parent[p];
        if (gp != p) atomicCAS(parent + v, p, gp);     // path halving (benign race)
        v = p;
        p = parent[v];
    }
    return v;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }        // hook the larger root under the smaller
        if (atomicCAS(parent + a, a, b) == a) return;
    }
}

__global__ void link_init_kernel(int* parent, int* chain, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) { parent[i] = (int)i; chain[i] = -1; }
}

__global__ void link_edges_kernel(const LinkArgs a, int* parent) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    long long lo, hi;
    link_window(a, i, &lo, &hi);
    const int g = a.group[i];
    for (long long j = lo; j < hi; j++)
        if (a.group[j] == g && link_is_near(a, i, j)) uf_union(parent, (int)i, (int)j);
}

__global__ void link_label_kernel(int* parent, int* label, int* idx, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) { label[i] = uf_find(parent, (int)i); idx[i] = (int)i; }
}

__global__ void link_heads_kernel(const int* sorted_label, int* head, long long n) {
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k < n) head[k] = (k == 0 || sorted_label[k] != sorted_label[k - 1]) ? 1 : 0;
}

// comp_start[c] = position of the c-th head (positions where head == 1); scan = exclusive scan of head
__global__ void link_comp_start_kernel(const int* head, const int* scan, int* comp_start, long long n) {
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k < n && head[k]) comp_start[scan[k]] = (int)k;
}

// one thread per component: the reference's greedy over the component's members (ascending index)
__global__ void link_greedy_kernel(const LinkArgs a, const int* __restrict__ members,
                                   const int* __restrict__ comp_start, int n_comp, int* chain) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_comp) return;
    const int p0 = comp_start[c];
    const int p1 = (c + 1 < n_comp) ? comp_start[c + 1] : (int)a.n;
    for (int p = p0; p < p1; p++) {
        const int s = members[p];
        if (chain[s] != -1) continue;
        chain[s] = s;
        int cur = s, cp = p;
        for (;;) {
            long long lo, hi;
            link_window(a, cur, &lo, &hi);
            const int g = a.group[cur];
            int nxt = -1, np_ = -1;
            for (int q = cp + 1; q < p1; q++) {
                const int j = members[q];
                if (j >= hi) break;
                if (j < lo || a.group[j] != g || chain[j] != -1) continue;
                if (link_is_near(a, cur, j)) { nxt = j; np_ = q; break; }
            }
            if (nxt < 0) break;
            chain[nxt] = s;
            cur = nxt;
            cp = np_;
        }
    }
}

__global__ void link_start_flags_kernel(const int* chain, int* flag, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) flag[i] = chain[i] == (int)i ? 1 : 0;
}
__global__ void link_assign_kernel(const int* chain, const int* scan, int* link_group, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) link_group[i] = scan[chain[i]];
}

// ---- per-group reductions -----------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void reduce_one(const T* col, const int* members, int p0, int p1, int op, T* out, int g) {
    T acc;
    if (op == 0) {
        acc = (T)0;
        for (int p = p0; p < p1; p++) acc += col[members[p]];          // sequential, localization order
    } else if (op == 3) {
        acc = col[members[p1 - 1]];
    } else {
        acc = col[members[p0]];
        for (int p = p0 + 1; p < p1; p++) {
            const T v = col[members[p]];
            if (op == 1 ? v < acc : v > acc) acc = v;
        }
    }
    out[g] = acc;
}
__global__ void link_reduce_kernel(const void* col, int dtype, int op, const int* __restrict__ members,
                                   const int* __restrict__ start, int n_groups, long long n, void* out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const int p0 = start[g], p1 = (g + 1 < n_groups) ? start[g + 1] : (int)n;
    switch (dtype) {
        case 0: reduce_one(static_cast<const float*>(col), members, p0, p1, op, static_cast<float*>(out), g); break;
        case 1: reduce_one(static_cast<const double*>(col), members, p0, p1, op, static_cast<double*>(out), g); break;
        case 2: reduce_one(static_cast<const unsigned*>(col), members, p0, p1, op, static_cast<unsigned*>(out), g); break;
        default: reduce_one(static_cast<const int*>(col), members, p0, p1, op, static_cast<int*>(out), g); break;
    }
}

struct DBuf {
    void* p = nullptr;
    ~DBuf() { if (p) cudaFree(p); }
    int alloc(size_t b) { PB_CUDA_CHECK(cudaMalloc(&p, b ? b : 1)); return PB_OK; }
    template <typename T> T* as() { return static_cast<T*>(p); }
};

inline unsigned blocks(long long n) { return (unsigned)((n + 255) / 256); }

// sort (key, index) pairs, mark run heads, return start offsets of the runs
int group_by_key(int* d_key, int* d_idx, long long n, DBuf& key_out, DBuf& idx_out, DBuf& start, int* n_runs,
                 int end_bit) {
    DBuf tmp, head, scan;
    int rc;
    if ((rc = key_out.alloc(n * 4)) || (rc = idx_out.alloc(n * 4)) || (rc = head.alloc(n * 4)) ||
        (rc = scan.alloc(n * 4)) || (rc = start.alloc(n * 4)))
        return rc;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, d_key, key_out.as<int>(), d_idx, idx_out.as<int>(), (int)n, 0, end_bit);
    size_t sb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, sb, head.as<int>(), scan.as<int>(), (int)n);
    if ((rc = tmp.alloc(std::max(tb, sb)))) return rc;
    cub::DeviceRadixSort::SortPairs(tmp.p, tb, d_key, key_out.as<int>(), d_idx, idx_out.as<int>(), (int)n, 0, end_bit);
    link_heads_kernel<<<blocks(n), 256>>>(key_out.as<int>(), head.as<int>(), n);
    cub::DeviceScan::ExclusiveSum(tmp.p, sb, head.as<int>(), scan.as<int>(), (int)n);
    link_comp_start_kernel<<<blocks(n), 256>>>(head.as<int>(), scan.as<int>(), start.as<int>(), n);
    g_pb_launches += 4;
    int last_scan = 0, last_head = 0;
    PB_CUDA_CHECK(cudaMemcpy(&last_scan, scan.as<int>() + (n - 1), 4, cudaMemcpyDeviceToHost));
    PB_CUDA_CHECK(cudaMemcpy(&last_head, head.as<int>() + (n - 1), 4, cudaMemcpyDeviceToHost));
    *n_runs = last_scan + last_head;
    return PB_OK;
}

}  // namespace

extern "C" int pb_link_groups(size_t n, const long long* frame, const void* x, const void* y, int xy_f64,
                              const int* group, double d_max, long long max_dark_time, int* link_group,
                              int* n_groups) {
    if (n_groups) *n_groups = 0;
    if (n == 0) return PB_OK;
    if (!frame || !x || !y || !group || !link_group) { pb_set_error("pb_link_groups: null pointer"); return PB_ERR_INVALID; }
    if (n >= ((size_t)1 << 31)) { pb_set_error("pb_link_groups: too many localizations"); return PB_ERR_INVALID; }
    for (size_t i = 1; i < n; i++)
        if (frame[i] < frame[i - 1]) { pb_set_error("pb_link_groups: frames must be sorted ascending"); return PB_ERR_INVALID; }
    const long long N = (long long)n;
    const long long fmin = frame[0], fmax = frame[n - 1];
    if (fmax - fmin > ((long long)1 << 28)) { pb_set_error("pb_link_groups: frame range too large"); return PB_ERR_INVALID; }
    std::vector<long long> fstart((size_t)(fmax - fmin + 2));
    {
        size_t i = 0;
        for (long long f = fmin; f <= fmax + 1; f++) {
            while (i < n && frame[i] < f) i++;
            fstart[(size_t)(f - fmin)] = (long long)i;
        }
    }
    const size_t cb = n * (xy_f64 ? 8 : 4);
    DBuf dfr, dx, dy, dg, dfs, parent, chain, label, idx, flag, scan, dlg, tmp;
    int rc;
    if ((rc = dfr.alloc(n * 8)) || (rc = dx.alloc(cb)) || (rc = dy.alloc(cb)) || (rc = dg.alloc(n * 4)) ||
        (rc = dfs.alloc(fstart.size() * 8)) || (rc = parent.alloc(n * 4)) || (rc = chain.alloc(n * 4)) ||
        (rc = label.alloc(n * 4)) || (rc = idx.alloc(n * 4)) || (rc = flag.alloc(n * 4)) ||
        (rc = scan.alloc(n * 4)) || (rc = dlg.alloc(n * 4)))
        return rc;
    if ((rc = pb_h2d(dfr.p, frame, n * 8, nullptr)) || (rc = pb_h2d(dx.p, x, cb, nullptr)) ||
        (rc = pb_h2d(dy.p, y, cb, nullptr)) || (rc = pb_h2d(dg.p, group, n * 4, nullptr)))
        return rc;
    PB_CUDA_CHECK(cudaMemcpy(dfs.p, fstart.data(), fstart.size() * 8, cudaMemcpyHostToDevice));
    LinkArgs a{N, dfr.as<long long>(), dx.p, dy.p, xy_f64 != 0, dg.as<int>(), d_max * d_max, max_dark_time,
               fmin, fmax, dfs.as<long long>()};
    link_init_kernel<<<blocks(N), 256>>>(parent.as<int>(), chain.as<int>(), N);
    link_edges_kernel<<<blocks(N), 256>>>(a, parent.as<int>());
    link_label_kernel<<<blocks(N), 256>>>(parent.as<int>(), label.as<int>(), idx.as<int>(), N);
    g_pb_launches += 3;
    DBuf skey, sidx, cstart;
    int n_comp = 0;
    if ((rc = group_by_key(label.as<int>(), idx.as<int>(), N, skey, sidx, cstart, &n_comp, 32))) return rc;
    link_greedy_kernel<<<(n_comp + 63) / 64, 64>>>(a, sidx.as<int>(), cstart.as<int>(), n_comp, chain.as<int>());
    link_start_flags_kernel<<<blocks(N), 256>>>(chain.as<int>(), flag.as<int>(), N);
    size_t sb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, sb, flag.as<int>(), scan.as<int>(), (int)N);
    if ((rc = tmp.alloc(sb))) return rc;
    cub::DeviceScan::ExclusiveSum(tmp.p, sb, flag.as<int>(), scan.as<int>(), (int)N);
    link_assign_kernel<<<blocks(N), 256>>>(chain.as<int>(), scan.as<int>(), dlg.as<int>(), N);
    g_pb_launches += 4;
    PB_CUDA_CHECK(cudaGetLastError());
    if ((rc = pb_d2h(link_group, dlg.p, n * 4, nullptr))) return rc;
    int ls = 0, lf = 0;
    PB_CUDA_CHECK(cudaMemcpy(&ls, scan.as<int>() + (n - 1), 4, cudaMemcpyDeviceToHost));
    PB_CUDA_CHECK(cudaMemcpy(&lf, flag.as<int>() + (n - 1), 4, cudaMemcpyDeviceToHost));
    if (n_groups) *n_groups = ls + lf;
    return PB_OK;
}

extern "C" int pb_link_reduce(size_t n, const int* link_group, int n_groups, int n_cols,
                              const void* const* cols, const int* dtypes, const int* ops, void* const* outs) {
    if (n == 0 || n_cols == 0 || n_groups == 0) return PB_OK;
    if (!link_group || !cols || !dtypes || !ops || !outs) { pb_set_error("pb_link_reduce: null pointer"); return PB_ERR_INVALID; }
    for (int k = 0; k < n_cols; k++)
        if (!cols[k] || !outs[k] || dtypes[k] < 0 || dtypes[k] > 3 || ops[k] < 0 || ops[k] > 3) {
            pb_set_error("pb_link_reduce: bad column %d (dtype 0 f32, 1 f64, 2 u32, 3 i32; op 0 sum, 1 min, 2 max, 3 last)", k);
            return PB_ERR_INVALID;
        }
    const long long N = (long long)n;
    DBuf dlg, idx, skey, sidx, start, dcol, dout;
    int rc;
    if ((rc = dlg.alloc(n * 4)) || (rc = idx.alloc(n * 4)) || (rc = dcol.alloc(n * 8)) || (rc = dout.alloc((size_t)n_groups * 8)))
        return rc;
    if ((rc = pb_h2d(dlg.p, link_group, n * 4, nullptr))) return rc;
    {
        // idx = 0..n-1 (reuse the label kernel's iota through a tiny lambda kernel is overkill: init + copy)
        std::vector<int> iota(n);
        for (size_t i = 0; i < n; i++) iota[i] = (int)i;
        if ((rc = pb_h2d(idx.p, iota.data(), n * 4, nullptr))) return rc;
        PB_CUDA_CHECK(cudaDeviceSynchronize());
    }
    int runs = 0;
    if ((rc = group_by_key(dlg.as<int>(), idx.as<int>(), N, skey, sidx, start, &runs, 32))) return rc;
    if (runs != n_groups) {
        pb_set_error("pb_link_reduce: link_group has %d distinct values, expected n_groups = %d (0 .. n_groups-1, all used)", runs, n_groups);
        return PB_ERR_INVALID;
    }
    for (int k = 0; k < n_cols; k++) {
        const size_t eb = dtypes[k] == 1 ? 8 : 4;
        if ((rc = pb_h2d(dcol.p, cols[k], n * eb, nullptr))) return rc;
        link_reduce_kernel<<<(n_groups + 127) / 128, 128>>>(dcol.p, dtypes[k], ops[k], sidx.as<int>(), start.as<int>(),
                                                            n_groups, N, dout.p);
        g_pb_launches++;
        PB_CUDA_CHECK(cudaGetLastError());
        if ((rc = pb_d2h(outs[k], dout.p, (size_t)n_groups * eb, nullptr))) return rc;
    }
    return PB_OK;
}
