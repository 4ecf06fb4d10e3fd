// picasso_b200/csrc/multicast.cu -- NVSwitch multicast (NVLS) buffers for the fused fit + all-gather.
//
// the one exchange of the sharded fit is the all-gather of every rank's packed output
// block (56 B / spot).  With unicast copies each rank sends its 560 MB block seven times (3.9 GB of
// NVLink egress per step at 8 GPUs, and seven re-reads of the block from HBM).  An NVSwitch MULTICAST
// object turns that into ONE store stream: memory of all N GPUs is bound to the object at the same
// offsets, and a `multimem.st` to the multicast address is replicated by the switch into every GPU's
// copy -- 0.56 GB of egress per rank and step, issued by the kernel that produces the values (the CRLB
// / log-likelihood kernel, csrc/mle_tps.cu): compute and collective in one kernel, no copy engine, no
// extra kernel, nothing re-read.
//
// The driver entry points are fetched with cudaGetDriverEntryPoint (the library does not link libcuda:
// it must load on the CPU-only build box).  One process per GPU: rank 0 creates the multicast object and
// exports a POSIX file descriptor, which the Python layer passes to the other ranks over a Unix socket
// (picasso_b200/distributed.py, MulticastBuffer).
#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>

#include <mutex>

#include "pb_common.cuh"
#include "../../include/picasso_b200.h"

namespace {

struct Drv {
    bool ok = false;
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                                 unsigned long long) = nullptr;
    CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
    CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType,
                                           unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
};

template <class F>
bool entry(const char* name, F* fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return false;
    }
    *fn = reinterpret_cast<F>(p);
    return true;
}

Drv& drv() {
    static Drv d;
    static std::once_flag once;
    std::call_once(once, [] {
        cudaFree(nullptr);      // make sure the runtime (and with it the driver) is initialised
        d.ok = entry("cuDeviceGet", &d.DeviceGet) && entry("cuDeviceGetAttribute", &d.DeviceGetAttribute) &&
               entry("cuMulticastCreate", &d.MulticastCreate) && entry("cuMulticastAddDevice", &d.MulticastAddDevice) &&
               entry("cuMulticastBindMem", &d.MulticastBindMem) && entry("cuMulticastUnbind", &d.MulticastUnbind) &&
               entry("cuMulticastGetGranularity", &d.MulticastGetGranularity) && entry("cuMemCreate", &d.MemCreate) &&
               entry("cuMemRelease", &d.MemRelease) &&
               entry("cuMemGetAllocationGranularity", &d.MemGetAllocationGranularity) &&
               entry("cuMemAddressReserve", &d.MemAddressReserve) && entry("cuMemAddressFree", &d.MemAddressFree) &&
               entry("cuMemMap", &d.MemMap) && entry("cuMemUnmap", &d.MemUnmap) && entry("cuMemSetAccess", &d.MemSetAccess) &&
               entry("cuMemExportToShareableHandle", &d.MemExportToShareableHandle) &&
               entry("cuMemImportFromShareableHandle", &d.MemImportFromShareableHandle) &&
               entry("cuGetErrorString", &d.GetErrorString);
    });
    return d;
}

struct Mc {
    CUmemGenericAllocationHandle mc = 0, mem = 0;
    CUdeviceptr uc_ptr = 0, mc_ptr = 0;
    size_t size = 0;
    int n_devices = 0;
    CUdevice dev = 0;
    bool bound = false;
};

#define PB_DRV_CHECK(expr)                                                                     \
    do {                                                                                       \
        CUresult _r = (expr);                                                                  \
        if (_r != CUDA_SUCCESS) {                                                              \
            const char* _m = nullptr;                                                          \
            drv().GetErrorString(_r, &_m);                                                     \
            pb_set_error("%s failed: %s (%s:%d)", #expr, _m ? _m : "?", __FILE__, __LINE__);   \
            return PB_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

int current_cudev(CUdevice* out) {
    int dev = 0;
    PB_CUDA_CHECK(cudaGetDevice(&dev));
    PB_DRV_CHECK(drv().DeviceGet(out, dev));
    return PB_OK;
}

CUmulticastObjectProp mc_prop(int n_devices, size_t size) {
    CUmulticastObjectProp p{};
    p.numDevices = (unsigned)n_devices;
    p.size = size;
    p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    p.flags = 0;
    return p;
}

}  // namespace

// 1 when the current device and driver support multicast objects (NVSwitch / NVLS), else 0
extern "C" int pb_mc_supported(void) {
    if (!drv().ok) return 0;
    CUdevice dev;
    if (current_cudev(&dev) != PB_OK) return 0;
    int v = 0;
    if (drv().DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS) return 0;
    return v ? 1 : 0;
}

// Size every rank must use for `bytes` of payload (multiple of the recommended multicast granularity).
extern "C" int pb_mc_padded_size(size_t bytes, int n_devices, size_t* padded) {
    if (!drv().ok) { pb_set_error("multicast: driver entry points unavailable"); return PB_ERR_CUDA; }
    if (!padded || n_devices < 2) { pb_set_error("pb_mc_padded_size: bad argument"); return PB_ERR_INVALID; }
    CUmulticastObjectProp p = mc_prop(n_devices, bytes);
    size_t g = 0;
    PB_DRV_CHECK(drv().MulticastGetGranularity(&g, &p, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    *padded = (bytes + g - 1) / g * g;
    return PB_OK;
}

// Rank 0: create the multicast object and export it as a POSIX file descriptor.
extern "C" int pb_mc_create(size_t padded_bytes, int n_devices, void** handle, int* export_fd) {
    if (!drv().ok) { pb_set_error("multicast: driver entry points unavailable"); return PB_ERR_CUDA; }
    if (!handle || !export_fd) { pb_set_error("pb_mc_create: null pointer"); return PB_ERR_INVALID; }
    Mc* m = new Mc();
    m->size = padded_bytes; m->n_devices = n_devices;
    CUmulticastObjectProp p = mc_prop(n_devices, padded_bytes);
    CUresult r = drv().MulticastCreate(&m->mc, &p);
    if (r != CUDA_SUCCESS) { delete m; PB_DRV_CHECK(r); }
    int fd = -1;
    r = drv().MemExportToShareableHandle(&fd, m->mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) { drv().MemRelease(m->mc); delete m; PB_DRV_CHECK(r); }
    *export_fd = fd;
    *handle = m;
    return PB_OK;
}

// Other ranks: import the object from the descriptor received from rank 0 (the fd is closed here).
extern "C" int pb_mc_import(int fd, size_t padded_bytes, int n_devices, void** handle) {
    if (!drv().ok) { pb_set_error("multicast: driver entry points unavailable"); return PB_ERR_CUDA; }
    if (!handle) { pb_set_error("pb_mc_import: null pointer"); return PB_ERR_INVALID; }
    Mc* m = new Mc();
    m->size = padded_bytes; m->n_devices = n_devices;
    CUresult r = drv().MemImportFromShareableHandle(&m->mc, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    close(fd);
    if (r != CUDA_SUCCESS) { delete m; PB_DRV_CHECK(r); }
    *handle = m;
    return PB_OK;
}

// Every rank: join the team with the current device.  ALL ranks must have returned from this call
// (barrier in the caller) before any rank calls pb_mc_bind_map.
extern "C" int pb_mc_add_device(void* handle) {
    Mc* m = static_cast<Mc*>(handle);
    if (!m) { pb_set_error("pb_mc_add_device: null handle"); return PB_ERR_INVALID; }
    int rc = current_cudev(&m->dev);
    if (rc != PB_OK) return rc;
    PB_DRV_CHECK(drv().MulticastAddDevice(m->mc, m->dev));
    return PB_OK;
}

// Every rank: allocate `size` bytes of physical memory on the current device, bind it to the object
// at offset 0 and map (a) the local memory (unicast pointer, ordinary loads / stores / memcpy) and
// (b) the multicast object (stores through it reach the bound memory of ALL devices).
extern "C" int pb_mc_bind_map(void* handle, void** uc_ptr, void** mc_ptr) {
    Mc* m = static_cast<Mc*>(handle);
    if (!m || !uc_ptr || !mc_ptr) { pb_set_error("pb_mc_bind_map: null pointer"); return PB_ERR_INVALID; }
    int dev = 0;
    PB_CUDA_CHECK(cudaGetDevice(&dev));
    CUmemAllocationProp ap{};
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = dev;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t g = 0;
    PB_DRV_CHECK(drv().MemGetAllocationGranularity(&g, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    if (m->size % g) { pb_set_error("pb_mc_bind_map: size %zu is not a multiple of the allocation granularity %zu", m->size, g); return PB_ERR_INVALID; }
    PB_DRV_CHECK(drv().MemCreate(&m->mem, m->size, &ap, 0));
    PB_DRV_CHECK(drv().MulticastBindMem(m->mc, 0, m->mem, 0, m->size, 0));
    m->bound = true;
    CUmemAccessDesc ad{};
    ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ad.location.id = dev;
    ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    PB_DRV_CHECK(drv().MemAddressReserve(&m->uc_ptr, m->size, g, 0, 0));
    PB_DRV_CHECK(drv().MemMap(m->uc_ptr, m->size, 0, m->mem, 0));
    PB_DRV_CHECK(drv().MemSetAccess(m->uc_ptr, m->size, &ad, 1));
    PB_DRV_CHECK(drv().MemAddressReserve(&m->mc_ptr, m->size, g, 0, 0));
    PB_DRV_CHECK(drv().MemMap(m->mc_ptr, m->size, 0, m->mc, 0));
    PB_DRV_CHECK(drv().MemSetAccess(m->mc_ptr, m->size, &ad, 1));
    *uc_ptr = reinterpret_cast<void*>(m->uc_ptr);
    *mc_ptr = reinterpret_cast<void*>(m->mc_ptr);
    return PB_OK;
}

extern "C" int pb_mc_destroy(void* handle) {
    Mc* m = static_cast<Mc*>(handle);
    if (!m) return PB_OK;
    cudaDeviceSynchronize();
    if (m->mc_ptr) { drv().MemUnmap(m->mc_ptr, m->size); drv().MemAddressFree(m->mc_ptr, m->size); }
    if (m->uc_ptr) { drv().MemUnmap(m->uc_ptr, m->size); drv().MemAddressFree(m->uc_ptr, m->size); }
    if (m->bound) drv().MulticastUnbind(m->mc, m->dev, 0, m->size);
    if (m->mem) drv().MemRelease(m->mem);
    if (m->mc) drv().MemRelease(m->mc);
    delete m;
    return PB_OK;
}

// ---- a plain multicast copy (tests, and gathers of buffers produced by other kernels) ----------
namespace {
__global__ void __launch_bounds__(256) mc_copy_kernel(const float4* __restrict__ src, float4* mc_dst, size_t n16) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = src[i];
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + i), "f"(v.x),
                     "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
    }
}
}  // namespace

// Store `bytes` (multiple of 16, 16-byte aligned) from local device memory through a multicast address:
// the data lands in every team member's bound memory.  `n_ctas` bounds the SMs taken from concurrent work.
extern "C" int pb_mc_copy_async(void* mc_dst, const void* d_src, size_t bytes, int n_ctas, void* stream) {
    if (bytes == 0) return PB_OK;
    if (!mc_dst || !d_src || (bytes & 15) || (reinterpret_cast<uintptr_t>(mc_dst) & 15) ||
        (reinterpret_cast<uintptr_t>(d_src) & 15)) {
        pb_set_error("pb_mc_copy_async: pointers and size must be 16-byte aligned");
        return PB_ERR_INVALID;
    }
    if (n_ctas < 1) n_ctas = 16;
    mc_copy_kernel<<<n_ctas, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        static_cast<const float4*>(d_src), static_cast<float4*>(mc_dst), bytes / 16);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}
