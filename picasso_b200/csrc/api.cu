// picasso_b200/csrc/api.cu -- C-ABI plumbing: error strings, device management,
// cached device workspaces and the host-buffer (H2D -> kernel -> D2H) pipelines.
#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <thread>
#include <utility>
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
// The SHA-256 of the sources this library was built from (csrc/, include/, flags) is embedded by
// picasso_b200/build.py: build() rebuilds whenever it differs from the sources on disk.
#ifndef PB_SOURCE_HASH
#define PB_SOURCE_HASH "0000000000000000000000000000000000000000000000000000000000000000"
#endif
extern "C" const char* pb_version(void) {
    return "picasso_b200 0.2.0 (sm_100a) PB_SRC_HASH=" PB_SOURCE_HASH;
}
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

extern "C" int pb_get_device(int* device) {
    if (!device) { pb_set_error("pb_get_device: null pointer"); return PB_ERR_INVALID; }
    PB_CUDA_CHECK(cudaGetDevice(device));
    return PB_OK;
}

extern "C" int pb_synchronize(void) {
    PB_CUDA_CHECK(cudaDeviceSynchronize());
    return PB_OK;
}

// Page-locked host memory.  cudaHostAlloc pins 4 KB pages one by one (measured on the B200 hosts:
// 350 ms for the 560 MB of results of a 10 M-spot fit -- eight times the fit itself).  Large blocks are
// therefore taken from anonymous memory backed by transparent huge pages (2 MB, madvise), faulted in by
// several threads and then registered with cudaHostRegister: 512x fewer pages to pin.  Falls back to
// cudaHostAlloc when huge pages or the registration are unavailable (PB_HOST_HUGEPAGES=0 forces it).
namespace {
struct HugeBlock { void* base; size_t mapped; };
std::mutex g_huge_mutex;
std::vector<std::pair<void*, HugeBlock>> g_huge_blocks;
constexpr size_t kHuge = (size_t)2 << 20;

bool huge_enabled() {
    static const bool on = [] {
        if (const char* e = getenv("PB_HOST_HUGEPAGES")) return atoi(e) != 0;
        return true;
    }();
    return on;
}

void* huge_alloc(size_t bytes) {
    const size_t size = (bytes + kHuge - 1) / kHuge * kHuge;
    void* base = mmap(nullptr, size + kHuge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == MAP_FAILED) return nullptr;
    char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(base) + kHuge - 1) / kHuge * kHuge);
    madvise(p, size, MADV_HUGEPAGE);
    // first touch from several threads (one write per 4 KB page faults the whole huge page in)
    const unsigned nt = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    std::vector<std::thread> ts;
    const size_t per = (size / kHuge + nt - 1) / nt * kHuge;
    for (unsigned t = 0; t < nt; t++)
        ts.emplace_back([=] {
            const size_t lo = std::min(size, t * per), hi = std::min(size, lo + per);
            for (size_t o = lo; o < hi; o += 4096) p[o] = 0;
        });
    for (auto& t : ts) t.join();
    if (cudaHostRegister(p, size, cudaHostRegisterDefault) != cudaSuccess) {
        cudaGetLastError();
        munmap(base, size + kHuge);
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_huge_mutex);
    g_huge_blocks.push_back({p, HugeBlock{base, size + kHuge}});
    return p;
}
}  // namespace

extern "C" int pb_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) { pb_set_error("pb_host_alloc: null out pointer"); return PB_ERR_INVALID; }
    if (bytes >= ((size_t)8 << 20) && huge_enabled()) {
        if (void* p = huge_alloc(bytes)) { *ptr = p; return PB_OK; }
    }
    PB_CUDA_CHECK(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return PB_OK;
}
// Pageable-friendly transfers for callers that keep their own device buffers (csrc/transfer.cu).
extern "C" int pb_copy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream) {
    if (bytes && (!d_dst || !h_src)) { pb_set_error("pb_copy_h2d: null pointer"); return PB_ERR_INVALID; }
    return pb_h2d(d_dst, h_src, bytes, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int pb_copy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream) {
    if (bytes && (!h_dst || !d_src)) { pb_set_error("pb_copy_d2h: null pointer"); return PB_ERR_INVALID; }
    return pb_d2h(h_dst, d_src, bytes, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int pb_host_free(void* ptr) {
    if (!ptr) return PB_OK;
    {
        std::lock_guard<std::mutex> lk(g_huge_mutex);
        for (size_t k = 0; k < g_huge_blocks.size(); k++)
            if (g_huge_blocks[k].first == ptr) {
                const HugeBlock b = g_huge_blocks[k].second;
                g_huge_blocks.erase(g_huge_blocks.begin() + k);
                cudaHostUnregister(ptr);
                munmap(b.base, b.mapped);
                return PB_OK;
            }
    }
    PB_CUDA_CHECK(cudaFreeHost(ptr));
    return PB_OK;
}

// ---- cached device workspaces (per device, per slot) ----------------------
// The host-buffer fit entry points stream chunks through kSlots slots (own stream each); the
// device buffers and the pinned output staging are kept between calls (cudaMalloc costs ~1 ms
// per 100 MB).
namespace {
struct Slot {
    void* buf = nullptr;
    size_t bytes = 0;
    void* hout = nullptr;      // pinned staging for the outputs of one chunk (pageable callers)
    size_t hbytes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
};
constexpr int kSlots = 3;
constexpr int kMaxDevices = 64;
struct DevWs {
    Slot slots[kSlots];
    std::mutex call_mutex;     // serialises the host-buffer fit calls that share these slots
};
std::mutex g_ws_mutex;
DevWs g_ws[kMaxDevices];       // indexed by device (fixed storage: Slot pointers stay valid)

int ws_device(int* dev) {
    PB_CUDA_CHECK(cudaGetDevice(dev));
    if (*dev < 0 || *dev >= kMaxDevices) { pb_set_error("device index %d out of range", *dev); return PB_ERR_INVALID; }
    return PB_OK;
}

int ws_get(int slot, size_t bytes, size_t host_bytes, Slot** out) {
    int dev = 0, rc0 = ws_device(&dev);
    if (rc0) return rc0;
    std::lock_guard<std::mutex> lk(g_ws_mutex);
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
    if (s.hbytes < host_bytes) {
        if (s.hout) PB_CUDA_CHECK(cudaFreeHost(s.hout));
        s.hout = nullptr;
        s.hbytes = 0;
        PB_CUDA_CHECK(cudaHostAlloc(&s.hout, host_bytes, cudaHostAllocDefault));
        s.hbytes = host_bytes;
    }
    *out = &s;
    return PB_OK;
}
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// One per-item output array of a fit: `host` (nullable) receives `stride` bytes per item.
struct OutSpec {
    void* host;
    size_t stride;
    bool device_only;      // allocated in the slot, never copied back (host == nullptr)
    size_t off;            // byte offset inside the slot (filled in by the pipeline)
};

// Chunked H2D -> kernel -> D2H pipeline shared by pb_mle_fit and pb_lq_fit: kSlots chunks in
// flight on separate streams, so the upload of chunk c+1/c+2 and the download of chunk c-1
// overlap the kernel of chunk c.  Pageable inputs go through pb_h2d (threaded pinned staging);
// pageable outputs come back with ONE DMA per chunk into the slot's pinned staging and are
// copied out when the slot is recycled, so no call blocks on a pageable cudaMemcpy.
template <class Launch>
int fit_pipeline(size_t n, size_t in_stride, const void* in, OutSpec* outs, int n_outs, size_t chunk,
                 Launch launch, volatile long long* progress) {
    size_t off = align_up(chunk * in_stride, 256);
    const size_t out0 = off;
    for (int k = 0; k < n_outs; k++) { outs[k].off = off; off = align_up(off + chunk * outs[k].stride, 256); }
    const size_t total = off;
    bool staged = false;
    for (int k = 0; k < n_outs; k++)
        if (outs[k].host && !pb_host_is_pinned(outs[k].host)) staged = true;
    Slot* sl[kSlots];
    int rc;
    for (int s = 0; s < kSlots; s++)
        if ((rc = ws_get(s, total, staged ? total - out0 : 0, &sl[s])) != PB_OK) return rc;
    size_t first_of[kSlots] = {0}, m_of[kSlots] = {0};
    bool busy[kSlots] = {false};
    // On ANY error the other slots may still have kernels and D2H copies into the caller's arrays
    // in flight: wait for every slot stream before returning, so that the caller may free or reuse
    // its buffers as soon as the call is back ("the library never keeps a pointer after the call").
    struct Quiesce {
        Slot** sl; bool armed = true;
        ~Quiesce() {
            if (!armed) return;
            for (int s = 0; s < kSlots; s++)
                if (sl[s] && sl[s]->stream) cudaStreamSynchronize(sl[s]->stream);
            cudaGetLastError();
        }
    } quiesce{sl};
    auto retire = [&](int s) -> int {       // chunk in slot s finished: hand its outputs to the caller
        PB_CUDA_CHECK(cudaEventSynchronize(sl[s]->done));
        if (staged) {
            const char* h = static_cast<const char*>(sl[s]->hout);
            for (int k = 0; k < n_outs; k++)
                if (outs[k].host)
                    pb_parallel_memcpy(static_cast<char*>(outs[k].host) + first_of[s] * outs[k].stride,
                                       h + (outs[k].off - out0), m_of[s] * outs[k].stride);
        }
        busy[s] = false;
        if (progress) *progress = (long long)(first_of[s] + m_of[s]);
        return PB_OK;
    };
    size_t c = 0;
    for (size_t first = 0; first < n; first += chunk, c++) {
        const int s = (int)(c % kSlots);
        Slot* S = sl[s];
        if (busy[s] && (rc = retire(s)) != PB_OK) return rc;
        const size_t m = (n - first < chunk) ? n - first : chunk;
        char* base = static_cast<char*>(S->buf);
        if ((rc = pb_h2d(base, static_cast<const char*>(in) + first * in_stride, m * in_stride, S->stream)))
            return rc;
        if ((rc = launch(m, base, outs, S->stream)) != PB_OK) return rc;
        if (staged) {
            PB_CUDA_CHECK(cudaMemcpyAsync(S->hout, base + out0, total - out0, cudaMemcpyDeviceToHost, S->stream));
        } else {
            for (int k = 0; k < n_outs; k++)
                if (outs[k].host)
                    PB_CUDA_CHECK(cudaMemcpyAsync(static_cast<char*>(outs[k].host) + first * outs[k].stride,
                                                  base + outs[k].off, m * outs[k].stride,
                                                  cudaMemcpyDeviceToHost, S->stream));
        }
        PB_CUDA_CHECK(cudaEventRecord(S->done, S->stream));
        busy[s] = true;
        first_of[s] = first;
        m_of[s] = m;
    }
    for (size_t k = 0; k < (size_t)kSlots; k++) {      // remaining chunks, oldest first
        const int s = (int)((c + k) % kSlots);
        if (busy[s] && (rc = retire(s)) != PB_OK) return rc;
    }
    if (progress) *progress = (long long)n;
    quiesce.armed = false;     // every slot was retired: nothing in flight
    return PB_OK;
}

// Host-buffer fit calls on the same device share that device's cached slots and are serialised;
// calls on different devices, and the render / identify / undrift entry points, are not.
std::mutex* device_call_mutex() {
    int dev = 0;
    if (ws_device(&dev) != PB_OK) return nullptr;
    return &g_ws[dev].call_mutex;
}
}  // namespace

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
    std::mutex* cm = device_call_mutex();
    if (!cm) return PB_ERR_CUDA;
    std::lock_guard<std::mutex> call_lk(*cm);
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
    OutSpec outs[5] = {{thetas, 24, false, 0}, {crlbs, 24, false, 0}, {logliks, 4, false, 0},
                       {iterations, 4, false, 0}, {status, 4, status == nullptr, 0}};
    auto launch = [&](size_t m, char* base, const OutSpec* o, cudaStream_t st) {
        return pb_mle_fit_dev(m, box, reinterpret_cast<float*>(base), eps, max_it, method,
                              reinterpret_cast<float*>(base + o[0].off), reinterpret_cast<float*>(base + o[1].off),
                              reinterpret_cast<float*>(base + o[2].off), reinterpret_cast<int*>(base + o[3].off),
                              reinterpret_cast<int*>(base + o[4].off), st);
    };
    return fit_pipeline(n, pix * 4, spots, outs, 5, chunk, launch, progress);
}

// Host-buffer LQ fit through the same pipeline (declared in include/picasso_b200.h).
extern "C" int pb_lq_fit(size_t n, int box, const float* spots, float* thetas, int* infos, int* nfevs) {
    if (n == 0) return PB_OK;
    if (!spots || !thetas) { pb_set_error("pb_lq_fit: null pointer"); return PB_ERR_INVALID; }
    if (box < 5 || box > 15 || !(box & 1)) {
        pb_set_error("unsupported box size %d for LQ fit (odd 5..15)", box);
        return PB_ERR_INVALID;
    }
    std::mutex* cm = device_call_mutex();
    if (!cm) return PB_ERR_CUDA;
    std::lock_guard<std::mutex> call_lk(*cm);
    const size_t pix = (size_t)box * box;
    size_t chunk = ((size_t)64 << 20) / (pix * 4);
    chunk = chunk / 4096 * 4096;
    if (chunk < 4096) chunk = 4096;
    if (chunk > n) chunk = align_up(n, 4);
    OutSpec outs[3] = {{thetas, 24, false, 0}, {infos, 4, infos == nullptr, 0}, {nfevs, 4, nfevs == nullptr, 0}};
    auto launch = [&](size_t m, char* base, const OutSpec* o, cudaStream_t st) {
        return pb_lq_fit_dev(m, box, reinterpret_cast<float*>(base), reinterpret_cast<float*>(base + o[0].off),
                             reinterpret_cast<int*>(base + o[1].off), reinterpret_cast<int*>(base + o[2].off), st);
    };
    return fit_pipeline(n, pix * 4, spots, outs, 3, chunk, launch, nullptr);
}

// Host-buffer Gpufit-path fit (csrc/gpufit_lm.cu) through the same pipeline.
extern "C" int pb_gpufit_fit(size_t n, int box, const float* spots, float tolerance, int max_iterations,
                             float* params, int* states, float* chi2, int* n_iterations) {
    if (n == 0) return PB_OK;
    if (!spots || !params) { pb_set_error("pb_gpufit_fit: null pointer"); return PB_ERR_INVALID; }
    if (box < 5 || box > 15 || !(box & 1)) {
        pb_set_error("unsupported box size %d for the Gpufit-path fit (odd 5..15)", box);
        return PB_ERR_INVALID;
    }
    std::mutex* cm = device_call_mutex();
    if (!cm) return PB_ERR_CUDA;
    std::lock_guard<std::mutex> call_lk(*cm);
    const size_t pix = (size_t)box * box;
    size_t chunk = ((size_t)64 << 20) / (pix * 4);
    chunk = chunk / 4096 * 4096;
    if (chunk < 4096) chunk = 4096;
    if (chunk > n) chunk = align_up(n, 4);
    OutSpec outs[4] = {{params, 24, false, 0}, {states, 4, states == nullptr, 0}, {chi2, 4, chi2 == nullptr, 0},
                       {n_iterations, 4, n_iterations == nullptr, 0}};
    auto launch = [&](size_t m, char* base, const OutSpec* o, cudaStream_t st) {
        return pb_gpufit_fit_dev(m, box, reinterpret_cast<float*>(base), tolerance, max_iterations,
                                 reinterpret_cast<float*>(base + o[0].off), reinterpret_cast<int*>(base + o[1].off),
                                 reinterpret_cast<float*>(base + o[2].off), reinterpret_cast<int*>(base + o[3].off), st);
    };
    return fit_pipeline(n, pix * 4, spots, outs, 4, chunk, launch, nullptr);
}
