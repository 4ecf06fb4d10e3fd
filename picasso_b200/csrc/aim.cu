// picasso_b200/csrc/aim.cu -- AIM drift correction: intersection counting on the GPU (sm_100a).
//
// Replaces the counting core of picasso.aim (reference picasso/aim.py): `_point_intersect_2d`
// :297-344 / `_point_intersect_3d` :377-431 (quantise the segment's coordinates to the 1-D
// index x + y*width_units [+ z*width_units*height_units], np.unique with counts) and
// `_run_intersections(_multithread)` :148-266 / `_count_intersections` :89-126 (for every shift
// of the local search region: sum over coordinates common to the reference and the shifted
// target of min(count0, count1)).  The reference sorts the concatenated coordinate arrays once
// per shift and segment (np.argsort of ~N elements x 49 shifts x n_segments); here
//   * the reference coordinates go once per round into an open-addressing hash table
//     (key = int32 1-D index, value = multiplicity) resident in HBM,
//   * per segment: a small hash table of the target multiplicities (insert kernel), then one
//     probe kernel that walks its occupied slots and looks every shifted key up in the
//     reference table, accumulating min(count0, count1) per shift in shared memory.
// Integer results are bit-exact.  The index arithmetic reproduces the reference's dtypes per
// coordinate array (float32 pandas columns stay float32 through `+= drift`, `/ intersect_d`,
// np.round and the products with width_units; float64 after the first round), including the
// float32 rounding of indices above 2^24 and the int32 truncation.
// The sub-pixel peak (phase of the first Fourier coefficients of the 7x7 count array), the
// spline and the final subtraction stay on the host (picasso_b200/aim.py).
#include <atomic>
#include <limits.h>
#include <math.h>
#include <vector>

#include "pb_common.cuh"
#include "../../include/picasso_b200.h"

extern std::atomic<long long> g_pb_launches;

namespace {

constexpr int kEmpty = INT_MIN;      // never a valid slot key: INT_MIN itself is counted separately

struct Table {
    int* keys = nullptr;
    int* counts = nullptr;
    unsigned cap = 0;                // power of two
    int* special = nullptr;          // multiplicity of the key INT_MIN (np.int32 overflow value)
};

struct Coord {
    const void* p;
    int f64;                          // 0 float32, 1 float64
};

struct QuantArgs {
    Coord x, y, z;                    // z.p == nullptr: 2-D
    long long first, count;
    double rel_x, rel_y, rel_z;       // added before quantisation (`x1 += rel_drift_x`)
    double d, wu, hu;                 // intersect_d, width_units, height_units
    Table t;
};

__device__ __forceinline__ unsigned aim_hash(int key, unsigned cap) {
    unsigned h = (unsigned)key * 2654435761u;
    h ^= h >> 15;
    return h & (cap - 1);
}

// np.round(v / d) in the array's own dtype; `rel` is added first in that dtype (weak scalar)
__device__ __forceinline__ void units(const Coord& c, long long i, double rel, bool add_rel, double d,
                                      float* uf, double* ud) {
    if (c.f64) {
        double v = static_cast<const double*>(c.p)[i];
        if (add_rel) v = __dadd_rn(v, rel);
        *ud = rint(__ddiv_rn(v, d));
    } else {
        float v = static_cast<const float*>(c.p)[i];
        if (add_rel) v = __fadd_rn(v, (float)rel);
        *uf = rintf(__fdiv_rn(v, (float)d));
    }
}

// np.int32(...) of a float: C truncation; out-of-range / NaN -> INT_MIN (x86 cvttss2si/cvttsd2si)
__device__ __forceinline__ int to_int32(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return (int)v;
}

__device__ __forceinline__ int quantise(const QuantArgs& a, long long i, bool add_rel_xy, bool add_rel_z) {
    float xf = 0.f, yf = 0.f, zf = 0.f;
    double xd = 0.0, yd = 0.0, zd = 0.0;
    units(a.x, i, a.rel_x, add_rel_xy, a.d, &xf, &xd);
    units(a.y, i, a.rel_y, add_rel_xy, a.d, &yf, &yd);
    // y_units * width_units in y's dtype, x_units + (...) in the promoted dtype
    double s;
    bool s64;
    if (a.y.f64) {
        const double ty = __dmul_rn(yd, a.wu);
        s = a.x.f64 ? __dadd_rn(xd, ty) : __dadd_rn((double)xf, ty);
        s64 = true;
    } else {
        const float ty = __fmul_rn(yf, (float)a.wu);
        if (a.x.f64) { s = __dadd_rn(xd, (double)ty); s64 = true; }
        else { s = (double)__fadd_rn(xf, ty); s64 = false; }
    }
    if (a.z.p) {
        units(a.z, i, a.rel_z, add_rel_z, a.d, &zf, &zd);
        // z_units * width_units * height_units, left to right in z's dtype
        if (a.z.f64) {
            const double tz = __dmul_rn(__dmul_rn(zd, a.wu), a.hu);
            s = __dadd_rn(s, tz);
        } else {
            const float tz = __fmul_rn(__fmul_rn(zf, (float)a.wu), (float)a.hu);
            s = s64 ? __dadd_rn(s, (double)tz) : (double)__fadd_rn((float)s, tz);
        }
    }
    return to_int32(s);
}

__device__ __forceinline__ void table_add(const Table& t, int key) {
    if (key == kEmpty) { atomicAdd(t.special, 1); return; }
    unsigned h = aim_hash(key, t.cap);
    for (;;) {
        const int cur = t.keys[h];
        if (cur == key) break;
        if (cur == kEmpty) {
            const int old = atomicCAS(t.keys + h, kEmpty, key);
            if (old == kEmpty || old == key) break;
        }
        h = (h + 1) & (t.cap - 1);
    }
    atomicAdd(t.counts + h, 1);
}

__device__ __forceinline__ int table_get(const Table& t, int key) {
    if (key == kEmpty) return *t.special;
    unsigned h = aim_hash(key, t.cap);
    for (;;) {
        const int cur = t.keys[h];
        if (cur == key) return t.counts[h];
        if (cur == kEmpty) return 0;
        h = (h + 1) & (t.cap - 1);
    }
}

__global__ void aim_clear_kernel(Table t) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < t.cap; i += gridDim.x * blockDim.x) {
        t.keys[i] = kEmpty;
        t.counts[i] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *t.special = 0;
}

__global__ void aim_insert_kernel(const QuantArgs a, int add_rel_xy, int add_rel_z) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.count;
         i += (long long)gridDim.x * blockDim.x)
        table_add(a.t, quantise(a, a.first + i, add_rel_xy != 0, add_rel_z != 0));
}

// roi[j] = sum over target keys c of min(ref[c + shift_j], target[c]).  int_shifts: the shifted
// key is an int32 sum with wrap-around (2-D, int32 + int32 arrays); otherwise int32 + float64
// (3-D): only integer-valued sums can equal an int32 reference key.
__global__ void __launch_bounds__(256)
aim_probe_kernel(Table seg, Table ref, const double* __restrict__ shifts, int n_shifts, int int_shifts,
                 int* __restrict__ roi) {
    extern __shared__ int acc[];
    for (int j = threadIdx.x; j < n_shifts; j += blockDim.x) acc[j] = 0;
    __syncthreads();
    const unsigned total = seg.cap + 1;           // last index = the INT_MIN counter
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int key, k1;
        if (i < seg.cap) { key = seg.keys[i]; k1 = seg.counts[i]; if (key == kEmpty) continue; }
        else { key = kEmpty; k1 = *seg.special; if (k1 == 0) continue; }
        for (int j = 0; j < n_shifts; j++) {
            int probe;
            if (int_shifts) {
                probe = (int)((unsigned)key + (unsigned)(int)shifts[j]);
            } else {
                const double v = (double)key + shifts[j];
                if (v != floor(v) || !(v >= -2147483648.0 && v <= 2147483647.0)) continue;
                probe = (int)v;
            }
            const int k0 = table_get(ref, probe);
            if (k0) atomicAdd(acc + j, k0 < k1 ? k0 : k1);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n_shifts; j += blockDim.x)
        if (acc[j]) atomicAdd(roi + j, acc[j]);
}

struct DevArr {
    void* p = nullptr;
    size_t bytes = 0;
    int grow(size_t want) {
        if (want <= bytes) return PB_OK;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        PB_CUDA_CHECK(cudaMalloc(&p, want ? want : 1));
        bytes = want;
        return PB_OK;
    }
    ~DevArr() { if (p) cudaFree(p); }
};

struct AimState {
    DevArr tx, ty, tz;                 // targets (raw dtype)
    int fx = 0, fy = 0, fz = 0;
    bool has_z = false;
    size_t n = 0;
    DevArr rk, rc, sk, sc, misc;       // reference / segment tables, misc = special counters + roi + shifts
    Table ref, seg;
    bool ref_3d = false;
    double d = 0, wu = 0, hu = 0;
    cudaStream_t stream = nullptr;
    int* h_roi = nullptr;              // pinned
    size_t h_roi_cap = 0;
    ~AimState() {
        if (stream) cudaStreamDestroy(stream);
        if (h_roi) cudaFreeHost(h_roi);
    }
};

constexpr int kMaxShifts = 4096;

unsigned pow2_at_least(size_t v) {
    unsigned c = 1024;
    while (c < v && c < (1u << 30)) c <<= 1;
    return c;
}

int upload(DevArr& dst, const void* src, size_t n, int f64, cudaStream_t s) {
    const size_t b = n * (f64 ? 8 : 4);
    int rc = dst.grow(b);
    if (rc) return rc;
    if (b) PB_CUDA_CHECK(cudaMemcpyAsync(dst.p, src, b, cudaMemcpyHostToDevice, s));
    return PB_OK;
}

int ensure_misc(AimState* st) {
    // layout: [0] ref special, [1] seg special, [16 .. 16+kMaxShifts) roi (int), then shifts (double)
    return st->misc.grow(64 + (size_t)kMaxShifts * 4 + (size_t)kMaxShifts * 8);
}
int* misc_roi(AimState* st) { return reinterpret_cast<int*>(static_cast<char*>(st->misc.p) + 64); }
double* misc_shifts(AimState* st) {
    return reinterpret_cast<double*>(static_cast<char*>(st->misc.p) + 64 + (size_t)kMaxShifts * 4);
}

}  // namespace

extern "C" int pb_aim_create(void** handle) {
    if (!handle) { pb_set_error("pb_aim_create: null handle pointer"); return PB_ERR_INVALID; }
    AimState* st = new AimState();
    cudaError_t e = cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        pb_set_error("pb_aim_create: %s", cudaGetErrorString(e));
        st->stream = nullptr;
        delete st;
        return PB_ERR_CUDA;
    }
    *handle = st;
    return PB_OK;
}

extern "C" int pb_aim_destroy(void* handle) {
    if (!handle) return PB_OK;
    AimState* st = static_cast<AimState*>(handle);
    cudaStreamSynchronize(st->stream);
    delete st;
    return PB_OK;
}

extern "C" int pb_aim_set_targets(void* handle, size_t n, const void* x, int x_f64, const void* y, int y_f64,
                                  const void* z, int z_f64) {
    AimState* st = static_cast<AimState*>(handle);
    if (!st || (n && (!x || !y))) { pb_set_error("pb_aim_set_targets: null pointer"); return PB_ERR_INVALID; }
    int rc;
    PB_CUDA_CHECK(cudaStreamSynchronize(st->stream));
    if ((rc = upload(st->tx, x, n, x_f64, st->stream)) || (rc = upload(st->ty, y, n, y_f64, st->stream))) return rc;
    st->has_z = z != nullptr;
    if (z && (rc = upload(st->tz, z, n, z_f64, st->stream))) return rc;
    st->fx = x_f64 != 0; st->fy = y_f64 != 0; st->fz = z_f64 != 0; st->n = n;
    PB_CUDA_CHECK(cudaStreamSynchronize(st->stream));   // host arrays may be released after return
    return PB_OK;
}

extern "C" int pb_aim_set_reference(void* handle, size_t n_ref, const void* rx, int rx_f64, const void* ry,
                                    int ry_f64, const void* rz, int rz_f64, double intersect_d,
                                    double width_units, double height_units) {
    AimState* st = static_cast<AimState*>(handle);
    if (!st || (n_ref && (!rx || !ry))) { pb_set_error("pb_aim_set_reference: null pointer"); return PB_ERR_INVALID; }
    if (!(intersect_d > 0)) { pb_set_error("pb_aim_set_reference: intersect_d must be positive"); return PB_ERR_INVALID; }
    int rc;
    if ((rc = ensure_misc(st))) return rc;
    DevArr ux, uy, uz;
    if ((rc = upload(ux, rx, n_ref, rx_f64, st->stream)) || (rc = upload(uy, ry, n_ref, ry_f64, st->stream))) return rc;
    if (rz && (rc = upload(uz, rz, n_ref, rz_f64, st->stream))) return rc;
    const unsigned cap = pow2_at_least(2 * n_ref + 16);
    if ((rc = st->rk.grow((size_t)cap * 4)) || (rc = st->rc.grow((size_t)cap * 4))) return rc;
    st->ref.keys = static_cast<int*>(st->rk.p);
    st->ref.counts = static_cast<int*>(st->rc.p);
    st->ref.cap = cap;
    st->ref.special = static_cast<int*>(st->misc.p);
    st->d = intersect_d; st->wu = width_units; st->hu = height_units; st->ref_3d = rz != nullptr;
    aim_clear_kernel<<<148 * 4, 256, 0, st->stream>>>(st->ref);
    if (n_ref) {
        QuantArgs a{};
        a.x = Coord{ux.p, rx_f64 != 0}; a.y = Coord{uy.p, ry_f64 != 0};
        a.z = Coord{rz ? uz.p : nullptr, rz_f64 != 0};
        a.first = 0; a.count = (long long)n_ref;
        a.d = intersect_d; a.wu = width_units; a.hu = height_units; a.t = st->ref;
        const int grid = (int)std::min<size_t>((n_ref + 255) / 256, 148 * 16);
        aim_insert_kernel<<<grid, 256, 0, st->stream>>>(a, 0, 0);
        g_pb_launches++;
    }
    g_pb_launches++;
    PB_CUDA_CHECK(cudaGetLastError());
    PB_CUDA_CHECK(cudaStreamSynchronize(st->stream));   // ux/uy/uz are freed on return
    return PB_OK;
}

// Intersection counts of targets [first, first + count) against the reference for every shift.
// 2-D reference: rel_x / rel_y are added to x / y before quantisation and the shifts are int32
// (aim.py:598-616); 3-D reference: rel_z is added to z only and the shifts are float64
// (aim.py:726-744).  roi_cc receives n_shifts int32 counts.
extern "C" int pb_aim_count(void* handle, size_t first, size_t count, double rel_x, double rel_y, double rel_z,
                            int n_shifts, const double* shifts, int* roi_cc) {
    AimState* st = static_cast<AimState*>(handle);
    if (!st || !shifts || !roi_cc) { pb_set_error("pb_aim_count: null pointer"); return PB_ERR_INVALID; }
    if (!st->ref.keys) { pb_set_error("pb_aim_count: no reference set"); return PB_ERR_INVALID; }
    if (n_shifts < 1 || n_shifts > kMaxShifts) { pb_set_error("pb_aim_count: n_shifts must be 1..%d", kMaxShifts); return PB_ERR_INVALID; }
    if (first + count > st->n) { pb_set_error("pb_aim_count: target range out of bounds"); return PB_ERR_INVALID; }
    if (st->ref_3d && !st->has_z) { pb_set_error("pb_aim_count: 3-D reference but targets without z"); return PB_ERR_INVALID; }
    for (int j = 0; j < n_shifts; j++) roi_cc[j] = 0;
    if (count == 0) return PB_OK;
    int rc;
    const unsigned cap = pow2_at_least(2 * count + 16);
    if ((rc = st->sk.grow((size_t)cap * 4)) || (rc = st->sc.grow((size_t)cap * 4))) return rc;
    if (st->h_roi_cap < (size_t)n_shifts) {
        if (st->h_roi) cudaFreeHost(st->h_roi);
        st->h_roi = nullptr;
        PB_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&st->h_roi), (size_t)kMaxShifts * 4, cudaHostAllocDefault));
        st->h_roi_cap = kMaxShifts;
    }
    st->seg.keys = static_cast<int*>(st->sk.p);
    st->seg.counts = static_cast<int*>(st->sc.p);
    st->seg.cap = cap;
    st->seg.special = static_cast<int*>(st->misc.p) + 1;
    cudaStream_t s = st->stream;
    int* d_roi = misc_roi(st);
    double* d_shifts = misc_shifts(st);
    PB_CUDA_CHECK(cudaMemcpyAsync(d_shifts, shifts, (size_t)n_shifts * 8, cudaMemcpyHostToDevice, s));
    PB_CUDA_CHECK(cudaMemsetAsync(d_roi, 0, (size_t)n_shifts * 4, s));
    const int cgrid = (int)std::min<unsigned>((cap + 255) / 256, 148 * 4);
    aim_clear_kernel<<<cgrid, 256, 0, s>>>(st->seg);
    QuantArgs a{};
    a.x = Coord{st->tx.p, st->fx}; a.y = Coord{st->ty.p, st->fy};
    a.z = Coord{st->ref_3d ? st->tz.p : nullptr, st->fz};
    a.first = (long long)first; a.count = (long long)count;
    a.rel_x = rel_x; a.rel_y = rel_y; a.rel_z = rel_z;
    a.d = st->d; a.wu = st->wu; a.hu = st->hu; a.t = st->seg;
    const int igrid = (int)std::min<size_t>((count + 255) / 256, 148 * 16);
    aim_insert_kernel<<<igrid, 256, 0, s>>>(a, st->ref_3d ? 0 : 1, st->ref_3d ? 1 : 0);
    aim_probe_kernel<<<cgrid, 256, (size_t)n_shifts * 4, s>>>(st->seg, st->ref, d_shifts, n_shifts,
                                                            st->ref_3d ? 0 : 1, d_roi);
    g_pb_launches += 3;
    PB_CUDA_CHECK(cudaGetLastError());
    PB_CUDA_CHECK(cudaMemcpyAsync(st->h_roi, d_roi, (size_t)n_shifts * 4, cudaMemcpyDeviceToHost, s));
    PB_CUDA_CHECK(cudaStreamSynchronize(s));
    for (int j = 0; j < n_shifts; j++) roi_cc[j] = st->h_roi[j];
    return PB_OK;
}
