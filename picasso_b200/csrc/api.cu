// picasso_b200/csrc/api.cu -- C-ABI plumbing: error strings, device management,
// cached device workspaces and the host-buffer (H2D -> kernel -> D2H) pipelines.
#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <stdlib.h>
#include <vector>

#include "pb_common.cuh"
#include "../../include/picasso_b200.h"

static thread_local char g_err[512] = "";
std::atomic<long long> g_pb_launches{0};

void pb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" const char* pb_last_error(void) { return g_err; }
extern "C" const char* pb_version(void) { return "picasso_b200 0.1.0 (sm_100a)"; }
extern "C" long long pb_launch_count(void) { return g_pb_launches.load(); }

extern "C" int pb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; i++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess &&
            major == 10)
            ok++;
    }
    return ok;
}

extern "C" int pb_set_device(int device) {
    PB_CUDA_CHECK(cudaSetDevice(device));
    return PB_OK;
}

extern "C" int pb_synchronize(void) {
    PB_CUDA_CHECK(cudaDeviceSynchronize());
    return PB_OK;
}

extern "C" int pb_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) { pb_set_error("pb_host_alloc: null out pointer"); return PB_ERR_INVALID; }
    PB_CUDA_CHECK(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return PB_OK;
}
extern "C" int pb_host_free(void* ptr) {
    if (ptr) PB_CUDA_CHECK(cudaFreeHost(ptr));
    return PB_OK;
}

// ---- cached device workspaces (per device, per slot) ----------------------
// The host-buffer entry points stream chunks through kSlots slots (own stream each);
// the device buffers are kept between calls (cudaMalloc costs ~1 ms per 100 MB).
namespace {
struct Slot {
    void* buf = nullptr;
    size_t bytes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
};
constexpr int kSlots = 3;
struct DevWs {
    Slot slots[kSlots];
};
std::mutex g_ws_mutex;
std::vector<DevWs> g_ws;   // indexed by device

int ws_get(int slot, size_t bytes, Slot** out) {
    int dev = 0;
    PB_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_ws_mutex);
    if ((int)g_ws.size() <= dev) g_ws.resize(dev + 1);
    Slot& s = g_ws[dev].slots[slot];
    if (!s.stream) {
        PB_CUDA_CHECK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        PB_CUDA_CHECK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    if (s.bytes < bytes) {
        if (s.buf) PB_CUDA_CHECK(cudaFree(s.buf));
        s.buf = nullptr;
        s.bytes = 0;
        PB_CUDA_CHECK(cudaMalloc(&s.buf, bytes));
        s.bytes = bytes;
    }
    *out = &s;
    return PB_OK;
}
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
}  // namespace

// Serialises host-buffer calls that share the cached workspaces.
static std::mutex g_host_call_mutex;

extern "C" int pb_mle_fit(size_t n, int box, const float* spots, double eps, int max_it, int method,
                          float* thetas, float* crlbs, float* logliks, int* iterations, int* status,
                          volatile long long* progress) {
    if (method != 0 && method != 1) {
        pb_set_error("Method not available.");
        return PB_ERR_INVALID;
    }
    if (box < 5 || box > 21 || (box & 1) == 0) {
        pb_set_error("unsupported box size %d (supported: odd 5..21)", box);
        return PB_ERR_INVALID;
    }
    if (progress) *progress = 0;
    if (n == 0) return PB_OK;
    if (!spots || !thetas || !crlbs || !logliks || !iterations) {
        pb_set_error("pb_mle_fit: null pointer");
        return PB_ERR_INVALID;
    }
    std::lock_guard<std::mutex> call_lk(g_host_call_mutex);
    const size_t pix = (size_t)box * box;
    // chunk: ~64 MB of ROIs (PB_MLE_CHUNK_MB overrides), a multiple of 4096 spots
    size_t chunk_mb = 64;
    if (const char* e = getenv("PB_MLE_CHUNK_MB")) {
        const long v = atol(e);
        if (v >= 1 && v <= 4096) chunk_mb = (size_t)v;
    }
    size_t chunk = (chunk_mb << 20) / (pix * 4);
    chunk = chunk / 4096 * 4096;
    if (chunk < 4096) chunk = 4096;
    if (chunk > n) chunk = align_up(n, 4);
    // slot layout: spots | thetas | crlbs | logliks | iterations | status
    const size_t o_sp = 0;
    const size_t o_th = align_up(o_sp + chunk * pix * 4, 256);
    const size_t o_cr = align_up(o_th + chunk * 24, 256);
    const size_t o_ll = align_up(o_cr + chunk * 24, 256);
    const size_t o_it = align_up(o_ll + chunk * 4, 256);
    const size_t o_st = align_up(o_it + chunk * 4, 256);
    const size_t total = align_up(o_st + chunk * 4, 256);
    Slot* sl[kSlots];
    int rc;
    for (int s = 0; s < kSlots; s++)
        if ((rc = ws_get(s, total, &sl[s])) != PB_OK) return rc;

    // kSlots chunks in flight: the H2D of chunk c+1/c+2 and the D2H of chunk c-1 overlap
    // the kernel of chunk c (separate streams -> separate copy engines)
    size_t done_spots[kSlots] = {0};
    bool busy[kSlots] = {false};
    size_t c = 0;
    for (size_t first = 0; first < n; first += chunk, c++) {
        const int s = (int)(c % kSlots);
        Slot* S = sl[s];
        if (busy[s]) {
            PB_CUDA_CHECK(cudaEventSynchronize(S->done));
            if (progress) *progress = (long long)done_spots[s];
        }
        const size_t m = (n - first < chunk) ? n - first : chunk;
        char* base = static_cast<char*>(S->buf);
        PB_CUDA_CHECK(cudaMemcpyAsync(base + o_sp, spots + first * pix, m * pix * 4,
                                      cudaMemcpyHostToDevice, S->stream));
        rc = pb_mle_fit_dev(m, box, reinterpret_cast<float*>(base + o_sp), eps, max_it, method,
                            reinterpret_cast<float*>(base + o_th),
                            reinterpret_cast<float*>(base + o_cr),
                            reinterpret_cast<float*>(base + o_ll),
                            reinterpret_cast<int*>(base + o_it),
                            reinterpret_cast<int*>(base + o_st), S->stream);
        if (rc != PB_OK) return rc;
        PB_CUDA_CHECK(cudaMemcpyAsync(thetas + first * 6, base + o_th, m * 24,
                                      cudaMemcpyDeviceToHost, S->stream));
        PB_CUDA_CHECK(cudaMemcpyAsync(crlbs + first * 6, base + o_cr, m * 24,
                                      cudaMemcpyDeviceToHost, S->stream));
        PB_CUDA_CHECK(cudaMemcpyAsync(logliks + first, base + o_ll, m * 4, cudaMemcpyDeviceToHost,
                                      S->stream));
        PB_CUDA_CHECK(cudaMemcpyAsync(iterations + first, base + o_it, m * 4,
                                      cudaMemcpyDeviceToHost, S->stream));
        if (status)
            PB_CUDA_CHECK(cudaMemcpyAsync(status + first, base + o_st, m * 4,
                                          cudaMemcpyDeviceToHost, S->stream));
        PB_CUDA_CHECK(cudaEventRecord(S->done, S->stream));
        busy[s] = true;
        done_spots[s] = first + m;
    }
    for (int s = 0; s < kSlots; s++)
        if (busy[s]) PB_CUDA_CHECK(cudaEventSynchronize(sl[s]->done));
    if (progress) *progress = (long long)n;
    return PB_OK;
}
