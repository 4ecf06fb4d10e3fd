// picasso_b200/csrc/transfer.cu -- pageable host memory <-> device at close to PCIe speed.
//
// The reference's callers hand numpy arrays (pageable memory) to every entry point.  A plain
// cudaMemcpy from / to pageable memory goes through the driver's single-threaded staging copy
// (measured 5-10 GB/s on the B200 hosts, against 52 GB/s for pinned memory).  Here transfers are
// cut into 32 MB chunks that alternate between two pinned staging buffers: several host
// threads copy chunk c+1 into (out of) its buffer while the DMA engine moves chunk c.
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <stdlib.h>
#include <thread>
#include <vector>

#include "pb_common.cuh"

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define PB_HAVE_STREAM_COPY 1
#endif

namespace {

constexpr size_t kStageBytes = (size_t)32 << 20;

// Slice copy of the staging workers.  Each worker moves a few MB -- below glibc's non-temporal
// threshold, so plain memcpy writes through the cache and pays a read-for-ownership of every
// destination line (3 bytes of memory traffic per byte copied).  The staging buffers are only ever
// read by the DMA engine (H2D) or the data is consumed much later (D2H), so the destination lines
// are written with streaming stores: 2 bytes of traffic per byte.  PB_COPY_STREAM=0 disables it.
#ifdef PB_HAVE_STREAM_COPY
__attribute__((target("avx2"))) void stream_copy_avx2(char* dst, const char* src, size_t n) {
    size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
    if (head > n) head = n;
    if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
    }
    _mm_sfence();
    if (i < n) memcpy(dst + i, src + i, n - i);
}
bool stream_copy_enabled() {
    static const bool on = [] {
        if (const char* e = getenv("PB_COPY_STREAM")) if (atoi(e) == 0) return false;
        return __builtin_cpu_supports("avx2") != 0;
    }();
    return on;
}
#endif

inline void slice_copy(char* dst, const char* src, size_t n) {
#ifdef PB_HAVE_STREAM_COPY
    if (n >= ((size_t)256 << 10) && stream_copy_enabled()) { stream_copy_avx2(dst, src, n); return; }
#endif
    memcpy(dst, src, n);
}

struct Stage {
    void* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int init() {
        if (buf[0]) return PB_OK;
        for (int b = 0; b < 2; b++) {
            PB_CUDA_CHECK(cudaHostAlloc(&buf[b], kStageBytes, cudaHostAllocDefault));
            PB_CUDA_CHECK(cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming));
        }
        return PB_OK;
    }
};
// One staging set per device and direction: the pinned buffers are not tied to a device, but the
// events recorded on the caller's stream are (cudaEventRecord needs event and stream on the same
// device), and a process may drive several GPUs (pb_set_device).  Uploads and downloads have
// separate locks so that one thread's H2D overlaps another thread's D2H.
constexpr int kMaxDevices = 64;
struct DevStage {
    std::mutex up_mutex, down_mutex;
    Stage up, down;
};
DevStage g_stage[kMaxDevices];

int current_stage(DevStage** out) {
    int dev = 0;
    PB_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) { pb_set_error("device index %d out of range", dev); return PB_ERR_INVALID; }
    *out = &g_stage[dev];
    return PB_OK;
}

unsigned copy_threads() {
    static const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    unsigned nt = std::min(12u, hw);
    if (const char* e = getenv("PB_COPY_THREADS")) nt = (unsigned)std::max(1, atoi(e));
    return std::min(nt, 64u);
}

}  // namespace

bool pb_host_is_pinned(const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// Persistent copy workers: spawning ~11 threads per 32 MB chunk cost ~1 ms per fit chunk
// (44 thread creations); the pool is created on first use and lives for the process (heap
// allocated and never destroyed, so no static-destruction race with blocked workers).
namespace {
struct CopyPool {
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    std::vector<std::thread> workers;
    // current job
    char* dst = nullptr;
    const char* src = nullptr;
    size_t bytes = 0, per = 0;
    unsigned parts = 0, next = 0, done = 0;
    unsigned long long generation = 0;
    std::mutex submit;      // one job at a time

    void worker() {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(m);
            cv_job.wait(lk, [&] { return generation != seen && next < parts; });
            const unsigned long long gen = generation;
            while (next < parts) {
                const unsigned k = next++;
                lk.unlock();
                const size_t o = (size_t)k * per;
                slice_copy(dst + o, src + o, std::min(per, bytes - o));
                lk.lock();
                if (++done == parts) cv_done.notify_all();
            }
            seen = gen;
        }
    }
    void run(void* d, const void* s_, size_t n, unsigned nt) {
        std::lock_guard<std::mutex> one(submit);
        if (workers.size() + 1 < nt) {
            std::lock_guard<std::mutex> lk(m);
            while (workers.size() + 1 < nt) {
                workers.emplace_back([this] { worker(); });
                workers.back().detach();
            }
        }
        const size_t p = ((n + nt - 1) / nt + 4095) / 4096 * 4096;
        const unsigned np_ = (unsigned)((n + p - 1) / p);
        {
            std::lock_guard<std::mutex> lk(m);
            dst = static_cast<char*>(d); src = static_cast<const char*>(s_);
            bytes = n; per = p; parts = np_; next = 1; done = 0;     // part 0 is copied by the caller
            generation++;
        }
        cv_job.notify_all();
        slice_copy(static_cast<char*>(d), static_cast<const char*>(s_), std::min(p, n));
        std::unique_lock<std::mutex> lk(m);
        if (++done != parts) cv_done.wait(lk, [&] { return done == parts; });
        parts = 0;
    }
};
CopyPool* copy_pool() {
    static CopyPool* p = new CopyPool();
    return p;
}
}  // namespace

void pb_parallel_memcpy(void* dst, const void* src, size_t bytes) {
    const unsigned nt = copy_threads();
    if (bytes < ((size_t)2 << 20) || nt == 1) { memcpy(dst, src, bytes); return; }
    copy_pool()->run(dst, src, bytes, nt);
}

int pb_h2d(void* d_dst, const void* h_src, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return PB_OK;
    if (bytes < ((size_t)1 << 20) || pb_host_is_pinned(h_src)) {
        PB_CUDA_CHECK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, s));
        return PB_OK;
    }
    DevStage* ds = nullptr;
    int rc = current_stage(&ds);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ds->up_mutex);
    Stage& g_up = ds->up;
    if ((rc = g_up.init())) return rc;
    size_t c = 0;
    for (size_t off = 0; off < bytes; off += kStageBytes, c++) {
        const int b = (int)(c & 1);
        const size_t len = std::min(kStageBytes, bytes - off);
        PB_CUDA_CHECK(cudaEventSynchronize(g_up.ev[b]));      // earlier DMA out of this buffer finished
        pb_parallel_memcpy(g_up.buf[b], static_cast<const char*>(h_src) + off, len);
        PB_CUDA_CHECK(cudaMemcpyAsync(static_cast<char*>(d_dst) + off, g_up.buf[b], len,
                                      cudaMemcpyHostToDevice, s));
        PB_CUDA_CHECK(cudaEventRecord(g_up.ev[b], s));
    }
    return PB_OK;
}

int pb_d2h(void* h_dst, const void* d_src, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return PB_OK;
    if (bytes < ((size_t)1 << 20) || pb_host_is_pinned(h_dst)) {
        PB_CUDA_CHECK(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, s));
        PB_CUDA_CHECK(cudaStreamSynchronize(s));
        return PB_OK;
    }
    DevStage* ds = nullptr;
    int rc = current_stage(&ds);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ds->down_mutex);
    Stage& g_down = ds->down;
    if ((rc = g_down.init())) return rc;
    const size_t nchunks = (bytes + kStageBytes - 1) / kStageBytes;
    for (size_t c = 0; c <= nchunks; c++) {
        if (c < nchunks) {
            const int b = (int)(c & 1);
            const size_t off = c * kStageBytes, len = std::min(kStageBytes, bytes - off);
            PB_CUDA_CHECK(cudaMemcpyAsync(g_down.buf[b], static_cast<const char*>(d_src) + off, len,
                                          cudaMemcpyDeviceToHost, s));
            PB_CUDA_CHECK(cudaEventRecord(g_down.ev[b], s));
        }
        if (c >= 1) {       // drain the previous chunk while this one is in flight
            const int b = (int)((c - 1) & 1);
            const size_t off = (c - 1) * kStageBytes, len = std::min(kStageBytes, bytes - off);
            PB_CUDA_CHECK(cudaEventSynchronize(g_down.ev[b]));
            pb_parallel_memcpy(static_cast<char*>(h_dst) + off, g_down.buf[b], len);
        }
    }
    return PB_OK;
}
