// picasso_b200/csrc/localize.cu -- fused movie -> localisation-table pipeline (sm_100a).
//
// Replaces, in one pass over the movie, the reference's
//   localize.identify (localize.py:639-712)  -> identifications
//   localize.get_spots (:1115-1145)          -> ROIs
//   gaussmle.gaussmle (gaussmle.py:409-475) | gausslq.fit_spots (gausslq.py:247-289)
//   gaussmle.locs_from_fits (gaussmle.py:957-1037) | gausslq.locs_from_fits (gausslq.py:404-484)
// i.e. the body of localize.localize (localize.py:1682-1815).  Each frame chunk crosses PCIe
// once; identifications, ROIs, theta and CRLB never leave the GPU -- only the finished
// localisation columns (4 bytes x 17 or 11 per spot) are copied back.
//
//   movie chunk (pageable) --threads--> pinned staging --DMA--> HBM        (copy stream)
//   identify -> key/sort (frame, y, x) -> gather ids -> cut ROIs -> fit -> columns -> D2H
//                                                                           (compute stream)
// The column arithmetic reproduces numpy's evaluation order and dtypes (float32 IEEE ops,
// float64 only where pandas promotes: theta + int64 pixel index).
#include <algorithm>
#include <atomic>
#include <cub/device/device_radix_sort.cuh>
#include <mutex>
#include <thread>
#include <vector>

#include "pb_common.cuh"
#include "../../include/picasso_b200.h"

extern std::atomic<long long> g_pb_launches;

namespace {

constexpr int kColsMle = 17;
constexpr int kColsLq = 11;

struct ColArgs {
    long long n;
    long long ld;           // elements between two columns of `out`
    int kind;               // 0/1 MLE, 2 LQ, 3 theta in the Gpufit column layout, 4 LQ theta -> Gpufit-path table
    int half;               // box // 2
    int em;
    const long long* frame;
    const long long* x;
    const long long* y;
    const float* ng;
    const float* theta;     // (n, 6)
    const float* crlb;      // (n, 6)   MLE only
    const float* loglik;    // (n)      MLE only
    const int* iterations;  // (n)      MLE only
    void* out;              // (ncols, ld) 4-byte elements
};

// np.maximum / np.minimum propagate NaN (fmaxf would drop it)
__device__ __forceinline__ float np_max(float a, float b) { return (a != a || b != b) ? (a + b) : fmaxf(a, b); }
__device__ __forceinline__ float np_min(float a, float b) { return (a != a || b != b) ? (a + b) : fminf(a, b); }

// gausslq.localization_precision (gausslq.py:547-589), all float32, numpy evaluation order:
//   sa2 = s**2 + 1/12; sa = sa2**0.5; sa_orth = (s_orth**2 + 1/12)**0.5
//   v = sa2 * (16/9 + (8*pi*sa*sa_orth*bg)/photons) / photons;  em: v *= 2;  sqrt(v)
__device__ __forceinline__ float lq_precision(float photons, float s, float so, float bg, int em) {
    const float c12 = (float)(1.0 / 12.0);
    const float c169 = (float)(16.0 / 9.0);
    const float c8pi = (float)(8.0 * 3.141592653589793);
    const float sa2 = __fadd_rn(__fmul_rn(s, s), c12);
    const float sa = __fsqrt_rn(sa2);
    const float sao = __fsqrt_rn(__fadd_rn(__fmul_rn(so, so), c12));
    const float t = __fmul_rn(__fmul_rn(__fmul_rn(c8pi, sa), sao), bg);
    float v = __fdiv_rn(__fmul_rn(sa2, __fadd_rn(c169, __fdiv_rn(t, photons))), photons);
    if (em) v = __fmul_rn(v, 2.0f);
    return __fsqrt_rn(v);
}

__global__ void __launch_bounds__(256) locs_columns_kernel(const ColArgs a) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    float* of = static_cast<float*>(a.out) + i;
    unsigned* ou = static_cast<unsigned*>(a.out) + i;
    const long long ld = a.ld;
    const float* th = a.theta + i * 6;
    const float t0 = th[0], t1 = th[1], t2 = th[2], t3 = th[3], t4 = th[4], t5 = th[5];
    const double xi = (double)a.x[i], yi = (double)a.y[i];
    ou[0] = (unsigned)a.frame[i];
    float x, y, photons, sx, sy, bg;
    if (a.kind <= 1) {          // gaussmle.locs_from_fits: theta + ids - box//2 in float64
        x = (float)(((double)t0 + xi) - (double)a.half);
        y = (float)(((double)t1 + yi) - (double)a.half);
        photons = t2; bg = t3; sx = t4; sy = t5;
    } else if (a.kind == 2) {   // gausslq.locs_from_fits: no box offset
        x = (float)((double)t0 + xi);
        y = (float)((double)t1 + yi);
        photons = t2; bg = t3; sx = t4; sy = t5;
    } else if (a.kind == 3) {   // Gpufit column layout [photons, x, y, sx, sy, bg] (gausslq.py:487-544)
        x = (float)(((double)t1 + xi) - (double)a.half);
        y = (float)(((double)t2 + yi) - (double)a.half);
        photons = t0; sx = t3; sy = t4; bg = t5;
    } else {                    // kind 4 (pb_localize, fit 3): lmdif theta; fit_spots_gpufit adds box//2 in
                                // float32 (gausslq.py:346-395), locs_from_fits_gpufit removes it in float64
        const float xg = __fadd_rn(t0, (float)a.half), yg = __fadd_rn(t1, (float)a.half);
        x = (float)(((double)xg + xi) - (double)a.half);
        y = (float)(((double)yg + yi) - (double)a.half);
        photons = t2; bg = t3; sx = t4; sy = t5;
    }
    of[1 * ld] = x;
    of[2 * ld] = y;
    of[3 * ld] = photons;
    of[4 * ld] = sx;
    of[5 * ld] = sy;
    of[6 * ld] = bg;
    const float big = np_max(sx, sy), small = np_min(sx, sy);
    of[9 * ld] = __fdiv_rn(__fsub_rn(big, small), big);
    of[10 * ld] = a.ng[i];
    if (a.kind <= 1) {
        const float* c = a.crlb + i * 6;
        of[7 * ld] = __fsqrt_rn(c[0]);
        of[8 * ld] = __fsqrt_rn(c[1]);
        of[11 * ld] = a.loglik[i];
        ou[12 * ld] = (unsigned)a.iterations[i];
        of[13 * ld] = __fsqrt_rn(c[2]);
        of[14 * ld] = __fsqrt_rn(c[3]);
        of[15 * ld] = __fsqrt_rn(c[4]);
        of[16 * ld] = __fsqrt_rn(c[5]);
    } else {
        of[7 * ld] = lq_precision(photons, sx, sy, bg, a.em);
        of[8 * ld] = lq_precision(photons, sy, sx, bg, a.em);
    }
}

// ---- (frame, y, x) ordering of the identifications of one chunk ------------------
// key = ((frame - frame_offset) * Y + y) * X + x: the linear pixel index inside the chunk, so the radix
// sort only has to look at ceil(log2(chunk frames * Y * X)) bits (29 for 2000 x 512 x 512: 4 passes, not 8)
__global__ void sort_keys_kernel(const long long* __restrict__ frame, const long long* __restrict__ x,
                                 const long long* __restrict__ y, long long frame_offset, unsigned n,
                                 unsigned long long Y, unsigned long long X,
                                 unsigned long long* __restrict__ keys, unsigned* __restrict__ idx) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = ((unsigned long long)(frame[i] - frame_offset) * Y + (unsigned long long)y[i]) * X +
              (unsigned long long)x[i];
    idx[i] = i;
}
__global__ void gather_ids_kernel(const unsigned* __restrict__ idx, unsigned n,
                                  const long long* __restrict__ frame, const long long* __restrict__ x,
                                  const long long* __restrict__ y, const float* __restrict__ ng,
                                  long long* __restrict__ sframe, long long* __restrict__ sx,
                                  long long* __restrict__ sy, float* __restrict__ sng) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned j = idx[i];
    sframe[i] = frame[j]; sx[i] = x[j]; sy[i] = y[j]; sng[i] = ng[j];
}

// ---- grow-only cached buffers (per device) -----------------------------------
struct Buf {
    void* p = nullptr;
    size_t bytes = 0;
    bool pinned = false;
    int grow(size_t want) {
        if (want <= bytes) return PB_OK;
        if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); p = nullptr; bytes = 0; }
        if (pinned) PB_CUDA_CHECK(cudaHostAlloc(&p, want, cudaHostAllocDefault));
        else PB_CUDA_CHECK(cudaMalloc(&p, want));
        bytes = want;
        return PB_OK;
    }
};

struct Pipeline {
    Buf mv[2], stage[2];
    Buf ids;        // unsorted frame,x,y (i64) + ng | sorted frame,x,y + ng
    Buf keys;       // keys in/out (u64), idx in/out (u32)
    Buf cubtmp, spots, fit, cols, counter;
    Buf hcols[2];   // pinned landing buffers of the finished columns (copied out one chunk later)
    Buf hcount;     // pinned 8 bytes
    cudaStream_t copy = nullptr, comp = nullptr;
    cudaEvent_t up[2] = {nullptr, nullptr}, staged[2] = {nullptr, nullptr},
                cut[2] = {nullptr, nullptr}, d2h[2] = {nullptr, nullptr}, cnt = nullptr;
    int init() {
        if (copy) return PB_OK;
        PB_CUDA_CHECK(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
        PB_CUDA_CHECK(cudaStreamCreateWithFlags(&comp, cudaStreamNonBlocking));
        for (int s = 0; s < 2; s++) {
            PB_CUDA_CHECK(cudaEventCreateWithFlags(&up[s], cudaEventDisableTiming));
            PB_CUDA_CHECK(cudaEventCreateWithFlags(&staged[s], cudaEventDisableTiming));
            PB_CUDA_CHECK(cudaEventCreateWithFlags(&cut[s], cudaEventDisableTiming));
            PB_CUDA_CHECK(cudaEventCreateWithFlags(&d2h[s], cudaEventDisableTiming));
            stage[s].pinned = true;
            hcols[s].pinned = true;
        }
        PB_CUDA_CHECK(cudaEventCreateWithFlags(&cnt, cudaEventDisableTiming));
        hcount.pinned = true;
        return PB_OK;
    }
};
std::mutex g_pipe_mutex;
std::vector<Pipeline*> g_pipes;   // indexed by device

int get_pipeline(Pipeline** out) {
    int dev = 0;
    PB_CUDA_CHECK(cudaGetDevice(&dev));
    if ((int)g_pipes.size() <= dev) g_pipes.resize(dev + 1, nullptr);
    if (!g_pipes[dev]) g_pipes[dev] = new Pipeline();
    *out = g_pipes[dev];
    return g_pipes[dev]->init();
}

inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

int fit_columns(int fit) { return fit <= 1 ? kColsMle : kColsLq; }

}  // namespace

extern "C" int pb_locs_columns(int fit) {
    if (fit < 0 || fit > 3) { pb_set_error("pb_locs_columns: fit must be 0..3"); return -1; }
    return fit_columns(fit);
}

extern "C" int pb_locs_from_fits_dev(size_t n, int fit, int box, int em, const long long* d_frame,
                                     const long long* d_x, const long long* d_y, const float* d_ng,
                                     const float* d_thetas, const float* d_crlbs,
                                     const float* d_logliks, const int* d_iterations,
                                     void* d_columns, size_t ld, void* stream) {
    if (fit < 0 || fit > 4) { pb_set_error("pb_locs_from_fits: fit must be 0..3"); return PB_ERR_INVALID; }
    if (n == 0) return PB_OK;
    if (!d_frame || !d_x || !d_y || !d_ng || !d_thetas || !d_columns ||
        (fit <= 1 && (!d_crlbs || !d_logliks || !d_iterations))) {
        pb_set_error("pb_locs_from_fits: null pointer");
        return PB_ERR_INVALID;
    }
    if (ld < n) { pb_set_error("pb_locs_from_fits: column stride smaller than n"); return PB_ERR_INVALID; }
    ColArgs a{(long long)n, (long long)ld, fit, box / 2, em, d_frame, d_x, d_y, d_ng, d_thetas,
              d_crlbs, d_logliks, d_iterations, d_columns};
    locs_columns_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    g_pb_launches++;
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

// host-pointer variant (the column arithmetic alone; used by the parity tests and by callers
// that hold fit results on the host)
extern "C" int pb_locs_from_fits(size_t n, int fit, int box, int em, const long long* frame,
                                 const long long* x, const long long* y, const float* ng,
                                 const float* thetas, const float* crlbs, const float* logliks,
                                 const int* iterations, void* columns) {
    if (fit < 0 || fit > 3) { pb_set_error("pb_locs_from_fits: fit must be 0..3"); return PB_ERR_INVALID; }
    if (n == 0) return PB_OK;
    const int nc = fit_columns(fit);
    const bool mle = fit <= 1;
    const size_t o_f = 0, o_x = o_f + up256(n * 8), o_y = o_x + up256(n * 8), o_ng = o_y + up256(n * 8),
                 o_th = o_ng + up256(n * 4), o_cr = o_th + up256(n * 24), o_ll = o_cr + up256(n * 24),
                 o_it = o_ll + up256(n * 4), o_out = o_it + up256(n * 4), total = o_out + (size_t)nc * n * 4;
    char* d = nullptr;
    PB_CUDA_CHECK(cudaMalloc(&d, total));
    auto H2D = [&](size_t o, const void* src, size_t b) { return cudaMemcpy(d + o, src, b, cudaMemcpyHostToDevice); };
    cudaError_t e = H2D(o_f, frame, n * 8);
    if (e == cudaSuccess) e = H2D(o_x, x, n * 8);
    if (e == cudaSuccess) e = H2D(o_y, y, n * 8);
    if (e == cudaSuccess) e = H2D(o_ng, ng, n * 4);
    if (e == cudaSuccess) e = H2D(o_th, thetas, n * 24);
    if (mle && e == cudaSuccess) e = H2D(o_cr, crlbs, n * 24);
    if (mle && e == cudaSuccess) e = H2D(o_ll, logliks, n * 4);
    if (mle && e == cudaSuccess) e = H2D(o_it, iterations, n * 4);
    int rc = PB_OK;
    if (e == cudaSuccess)
        rc = pb_locs_from_fits_dev(n, fit, box, em, (const long long*)(d + o_f), (const long long*)(d + o_x),
                                   (const long long*)(d + o_y), (const float*)(d + o_ng),
                                   (const float*)(d + o_th), (const float*)(d + o_cr),
                                   (const float*)(d + o_ll), (const int*)(d + o_it), d + o_out, n, nullptr);
    if (rc == PB_OK && e == cudaSuccess) e = cudaMemcpy(columns, d + o_out, (size_t)nc * n * 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (rc != PB_OK) return rc;
    if (e != cudaSuccess) { pb_set_error("pb_locs_from_fits: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    return PB_OK;
}

// on_device: `movie` and `columns` are device pointers (pb_localize_dev): no upload, the finished
// columns are copied device -> device; everything else is the same pipeline.
static int localize_impl(bool on_device, const void* movie, int dtype, size_t n_frames, int Y, int X,
                         long long frame_offset, int box, double min_ng, const int* roi,
                         float baseline, float sensitivity, float gain, int fit, double eps,
                         int max_it, int em, void* columns, size_t capacity, size_t* n_found) {
    if (!n_found) { pb_set_error("pb_localize: n_found is null"); return PB_ERR_INVALID; }
    *n_found = 0;
    if (fit < 0 || fit > 3) { pb_set_error("pb_localize: fit must be 0 (MLE sigma), 1 (MLE sigmaxy), 2 (LQ) or 3 (LQ, Gpufit layout)"); return PB_ERR_INVALID; }
    if (dtype != PB_DTYPE_U16 && dtype != PB_DTYPE_F32) { pb_set_error("unsupported movie dtype %d (0 = uint16, 1 = float32)", dtype); return PB_ERR_INVALID; }
    if (box < 5 || box > 15 || (box & 1) == 0) { pb_set_error("unsupported box size %d for pb_localize (odd 5..15)", box); return PB_ERR_INVALID; }
    if (Y <= 0 || X <= 0 || Y >= (1 << 20) || X >= (1 << 20)) { pb_set_error("pb_localize: frame size out of range"); return PB_ERR_INVALID; }
    if (n_frames == 0) return PB_OK;
    if (!movie || (!columns && capacity)) { pb_set_error("pb_localize: null pointer"); return PB_ERR_INVALID; }

    std::lock_guard<std::mutex> lk(g_pipe_mutex);
    Pipeline* P = nullptr;
    int rc = get_pipeline(&P);
    if (rc != PB_OK) return rc;
    const size_t fsz = (size_t)Y * X * (dtype == PB_DTYPE_U16 ? 2 : 4);
    const size_t pix = (size_t)box * box;
    const int ncols = fit_columns(fit);
    size_t chunk = std::max<size_t>(1, ((size_t)64 << 20) / fsz);
    if (const char* e = getenv("PB_LOCALIZE_CHUNK_FRAMES")) {   // test / tuning override
        const long v = atol(e);
        if (v >= 1) chunk = (size_t)v;
    }
    // resident movie: no staging limit; one chunk per 2 GB keeps the per-chunk count read-backs rare and gives
    // the fit kernel all of a block's spots at once (config 3: one identify + one fit launch instead of two
    // of each, 3.0 -> 2.0 ms; the id / ROI / fit workspace is ~400 B per expected spot = 0.85 GB)
    if (on_device) chunk = std::max<size_t>(chunk, ((size_t)2048 << 20) / fsz);
    chunk = std::min(chunk, std::min<size_t>(n_frames, (size_t)1 << 22));
    const bool pinned_src = on_device || pb_host_is_pinned(movie);
    for (int s = 0; s < 2 && !on_device; s++) {
        if ((rc = P->mv[s].grow(chunk * fsz))) return rc;
        if (!pinned_src && (rc = P->stage[s].grow(chunk * fsz))) return rc;
    }
    if ((rc = P->counter.grow(8)) || (rc = P->hcount.grow(8))) return rc;
    size_t dcap = 0;
    long long *uf = nullptr, *ux = nullptr, *uy = nullptr, *sf = nullptr, *sx = nullptr, *sy = nullptr;
    float *ung = nullptr, *sng = nullptr, *d_sp = nullptr, *d_th = nullptr, *d_cr = nullptr, *d_ll = nullptr;
    int *d_it = nullptr, *d_st = nullptr;
    unsigned long long *k_in = nullptr, *k_out = nullptr;
    unsigned *i_in = nullptr, *i_out = nullptr;
    size_t cub_bytes = 0;
    auto size_for = [&](size_t cap) -> int {
        int r;
        const size_t a8 = up256(cap * 8), a4 = up256(cap * 4);
        if ((r = P->ids.grow(6 * a8 + 2 * a4))) return r;
        char* b = static_cast<char*>(P->ids.p);
        uf = (long long*)b; ux = (long long*)(b + a8); uy = (long long*)(b + 2 * a8);
        sf = (long long*)(b + 3 * a8); sx = (long long*)(b + 4 * a8); sy = (long long*)(b + 5 * a8);
        ung = (float*)(b + 6 * a8); sng = (float*)(b + 6 * a8 + a4);
        if ((r = P->keys.grow(2 * a8 + 2 * a4))) return r;
        b = static_cast<char*>(P->keys.p);
        k_in = (unsigned long long*)b; k_out = (unsigned long long*)(b + a8);
        i_in = (unsigned*)(b + 2 * a8); i_out = (unsigned*)(b + 2 * a8 + a4);
        cub_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, k_in, k_out, i_in, i_out, (int)cap);
        if ((r = P->cubtmp.grow(cub_bytes + 256))) return r;
        if ((r = P->spots.grow(cap * pix * 4))) return r;
        d_sp = static_cast<float*>(P->spots.p);
        const size_t a24 = up256(cap * 24);
        if ((r = P->fit.grow(2 * a24 + 3 * a4))) return r;
        b = static_cast<char*>(P->fit.p);
        d_th = (float*)b; d_cr = (float*)(b + a24); d_ll = (float*)(b + 2 * a24);
        d_it = (int*)(b + 2 * a24 + a4); d_st = (int*)(b + 2 * a24 + 2 * a4);
        if ((r = P->cols.grow((size_t)ncols * cap * 4))) return r;
        dcap = cap;
        return PB_OK;
    };
    if ((rc = size_for(std::max<size_t>(8192, chunk * 512)))) return rc;

    const size_t nchunks = (n_frames + chunk - 1) / chunk;
    auto upload = [&](size_t c) -> int {
        if (c >= nchunks || on_device) return PB_OK;
        const int s = (int)(c & 1);
        const size_t f0 = c * chunk, nf = std::min(chunk, n_frames - f0);
        const char* src = static_cast<const char*>(movie) + f0 * fsz;
        if (!pinned_src) {
            if (c >= 2) PB_CUDA_CHECK(cudaEventSynchronize(P->staged[s]));   // staging buffer drained
            pb_parallel_memcpy(P->stage[s].p, src, nf * fsz);
            src = static_cast<const char*>(P->stage[s].p);
        }
        if (c >= 2) PB_CUDA_CHECK(cudaStreamWaitEvent(P->copy, P->cut[s], 0));   // chunk c-2 no longer read
        PB_CUDA_CHECK(cudaMemcpyAsync(P->mv[s].p, src, nf * fsz, cudaMemcpyHostToDevice, P->copy));
        PB_CUDA_CHECK(cudaEventRecord(P->up[s], P->copy));
        PB_CUDA_CHECK(cudaEventRecord(P->staged[s], P->copy));
        return PB_OK;
    };
    if ((rc = upload(0))) return rc;
    size_t total = 0;
    bool overflow = false;
    // finished columns land in pinned memory and are copied to the caller one chunk later, so the
    // host never blocks on a pageable D2H while the next chunk could be staged
    struct Pending { bool on = false; size_t at = 0, n = 0, pitch = 0; } pend[2];
    auto drain = [&](int s) -> int {
        if (!pend[s].on) return PB_OK;
        PB_CUDA_CHECK(cudaEventSynchronize(P->d2h[s]));
        const char* h = static_cast<const char*>(P->hcols[s].p);
        for (int k = 0; k < ncols; k++)
            memcpy(static_cast<char*>(columns) + ((size_t)k * capacity + pend[s].at) * 4,
                   h + (size_t)k * pend[s].pitch * 4, pend[s].n * 4);
        pend[s].on = false;
        return PB_OK;
    };
    volatile unsigned long long* hcount = static_cast<unsigned long long*>(P->hcount.p);
    const cudaStream_t cs = P->comp;
    for (size_t c = 0; c < nchunks; c++) {
        const int s = (int)(c & 1);
        const size_t f0 = c * chunk, nf = std::min(chunk, n_frames - f0);
        const long long foff = frame_offset + (long long)f0;
        const void* mvp = on_device ? static_cast<const void*>(static_cast<const char*>(movie) + f0 * fsz)
                                    : P->mv[s].p;
        if (!on_device) PB_CUDA_CHECK(cudaStreamWaitEvent(cs, P->up[s], 0));
        bool next_uploaded = false;
        unsigned long long found = 0;
        for (;;) {   // retried only when the per-chunk device capacity was too small
            PB_CUDA_CHECK(cudaMemsetAsync(P->counter.p, 0, 8, cs));
            rc = pb_identify_dev(mvp, dtype, nf, Y, X, foff, box, min_ng, roi, uf, ux, uy, ung,
                                 dcap, static_cast<unsigned long long*>(P->counter.p), cs);
            if (rc != PB_OK) return rc;
            PB_CUDA_CHECK(cudaMemcpyAsync((void*)hcount, P->counter.p, 8, cudaMemcpyDeviceToHost, cs));
            PB_CUDA_CHECK(cudaEventRecord(P->cnt, cs));
            if (!next_uploaded) {   // stage + enqueue the next chunk while this one is searched
                if ((rc = upload(c + 1))) return rc;
                next_uploaded = true;
            }
            PB_CUDA_CHECK(cudaEventSynchronize(P->cnt));
            found = *hcount;
            if (found <= dcap) break;
            PB_CUDA_CHECK(cudaStreamSynchronize(cs));
            if ((rc = size_for((size_t)found + (size_t)found / 8))) return rc;
        }
        if (total + found > capacity) overflow = true;
        if (!overflow && found) {
            const unsigned n = (unsigned)found;
            const unsigned g = (n + 255) / 256;
            sort_keys_kernel<<<g, 256, 0, cs>>>(uf, ux, uy, foff, n, (unsigned long long)Y, (unsigned long long)X,
                                                k_in, i_in);
            int key_bits = 1;
            while (key_bits < 64 && ((unsigned long long)nf * Y * X - 1) >> key_bits) key_bits++;
            size_t tb = cub_bytes;
            if (cub::DeviceRadixSort::SortPairs(P->cubtmp.p, tb, k_in, k_out, i_in, i_out, (int)n, 0, key_bits, cs) !=
                cudaSuccess) {
                pb_set_error("pb_localize: radix sort failed: %s", cudaGetErrorString(cudaGetLastError()));
                return PB_ERR_CUDA;
            }
            gather_ids_kernel<<<g, 256, 0, cs>>>(i_out, n, uf, ux, uy, ung, sf, sx, sy, sng);
            g_pb_launches += 3;
            if ((rc = pb_get_spots_dev(mvp, dtype, nf, Y, X, foff, n, sf, sx, sy, box, baseline,
                                       sensitivity, gain, d_sp, cs)))
                return rc;
            PB_CUDA_CHECK(cudaEventRecord(P->cut[s], cs));
            if (fit <= 1) rc = pb_mle_fit_dev(n, box, d_sp, eps, max_it, fit, d_th, d_cr, d_ll, d_it, d_st, cs);
            else if (fit == 2) rc = pb_lq_fit_dev(n, box, d_sp, d_th, d_it, d_st, cs);
            else rc = pb_gpufit_fit_dev(n, box, d_sp, 1e-2f, 20, d_th, d_st, d_ll, d_it, cs);   // "gausslq-gpu"
            if (rc != PB_OK) return rc;
            // fit 3: theta is in Gpufit's layout [photons, x, y, sx, sy, bg] -> locs_from_fits_gpufit (kind 3)
            if ((rc = pb_locs_from_fits_dev(n, fit, box, em, sf, sx, sy, sng, d_th, d_cr, d_ll, d_it,
                                            P->cols.p, dcap, cs)))
                return rc;
            if (on_device) {
                PB_CUDA_CHECK(cudaMemcpy2DAsync(static_cast<char*>(columns) + total * 4, capacity * 4, P->cols.p,
                                                dcap * 4, (size_t)n * 4, ncols, cudaMemcpyDeviceToDevice, cs));
            } else {
            if ((rc = drain(s))) return rc;                       // chunk c-2 used this landing buffer
            if ((rc = P->hcols[s].grow((size_t)ncols * n * 4))) return rc;
            PB_CUDA_CHECK(cudaMemcpy2DAsync(P->hcols[s].p, (size_t)n * 4, P->cols.p, dcap * 4, (size_t)n * 4,
                                            ncols, cudaMemcpyDeviceToHost, cs));
            PB_CUDA_CHECK(cudaEventRecord(P->d2h[s], cs));
            pend[s].on = true; pend[s].at = total; pend[s].n = n; pend[s].pitch = n;
            if ((rc = drain(s ^ 1))) return rc;                   // previous chunk: long finished
            }
            // the id / fit / column buffers are reused by the next chunk on the same stream: in order
        } else {
            PB_CUDA_CHECK(cudaEventRecord(P->cut[s], cs));
        }
        total += found;
    }
    PB_CUDA_CHECK(cudaStreamSynchronize(cs));
    PB_CUDA_CHECK(cudaStreamSynchronize(P->copy));
    PB_CUDA_CHECK(cudaGetLastError());
    if ((rc = drain(0)) || (rc = drain(1))) return rc;
    *n_found = total;
    if (overflow) {
        pb_set_error("pb_localize: found %zu spots, capacity %zu", total, capacity);
        return PB_ERR_CAPACITY;
    }
    return PB_OK;
}

extern "C" int pb_localize(const void* movie, int dtype, size_t n_frames, int Y, int X,
                           long long frame_offset, int box, double min_ng, const int* roi,
                           float baseline, float sensitivity, float gain, int fit, double eps,
                           int max_it, int em, void* columns, size_t capacity, size_t* n_found) {
    return localize_impl(false, movie, dtype, n_frames, Y, X, frame_offset, box, min_ng, roi, baseline,
                         sensitivity, gain, fit, eps, max_it, em, columns, capacity, n_found);
}

// Device-resident variant: `d_movie` (n_frames, Y, X) and `d_columns` (ncols, capacity) live in
// HBM.  Synchronous like pb_localize (the identification count of every chunk is read back).
extern "C" int pb_localize_dev(const void* d_movie, int dtype, size_t n_frames, int Y, int X,
                               long long frame_offset, int box, double min_ng, const int* roi,
                               float baseline, float sensitivity, float gain, int fit, double eps,
                               int max_it, int em, void* d_columns, size_t capacity, size_t* n_found) {
    return localize_impl(true, d_movie, dtype, n_frames, Y, X, frame_offset, box, min_ng, roi, baseline,
                         sensitivity, gain, fit, eps, max_it, em, d_columns, capacity, n_found);
}
