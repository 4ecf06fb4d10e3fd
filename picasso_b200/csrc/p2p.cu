// picasso_b200/csrc/p2p.cu -- peer-to-peer plumbing for the multi-GPU gather of fit results.
//
// the only exchange of the sharded fit is the gather of every rank's packed output
// block.  An NCCL all-gather runs as a kernel and competes with the persistent fit kernel of the
// next step for SMs (measured: 19.3 ms/step at 4 GPUs against 18.0 ms alone).  With one process
// per GPU on an NVLink / NVSwitch box the same exchange is N-1 peer copies per rank executed by
// the COPY ENGINES: the destination buffers are allocated here (cudaMalloc, so that the whole
// allocation can be exported), shared through CUDA IPC handles, and written with
// cudaMemcpyAsync on side streams -- no SM is used, the fit keeps the whole chip.
#include <cuda_runtime.h>

#include "pb_common.cuh"
#include "../../include/picasso_b200.h"

extern "C" int pb_dev_alloc(void** ptr, size_t bytes) {
    if (!ptr) { pb_set_error("pb_dev_alloc: null out pointer"); return PB_ERR_INVALID; }
    PB_CUDA_CHECK(cudaMalloc(ptr, bytes ? bytes : 1));
    return PB_OK;
}
extern "C" int pb_dev_free(void* ptr) {
    if (ptr) PB_CUDA_CHECK(cudaFree(ptr));
    return PB_OK;
}
extern "C" int pb_ipc_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }
extern "C" int pb_ipc_export(const void* d_ptr, void* handle) {
    if (!d_ptr || !handle) { pb_set_error("pb_ipc_export: null pointer"); return PB_ERR_INVALID; }
    cudaIpcMemHandle_t h;
    PB_CUDA_CHECK(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
    memcpy(handle, &h, sizeof h);
    return PB_OK;
}
extern "C" int pb_ipc_open(const void* handle, void** d_ptr) {
    if (!d_ptr || !handle) { pb_set_error("pb_ipc_open: null pointer"); return PB_ERR_INVALID; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    PB_CUDA_CHECK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PB_OK;
}
extern "C" int pb_ipc_close(void* d_ptr) {
    if (d_ptr) PB_CUDA_CHECK(cudaIpcCloseMemHandle(d_ptr));
    return PB_OK;
}
// device -> device copy (own or peer memory) on `stream`: executed by a copy engine
extern "C" int pb_copy_d2d_async(void* dst, const void* src, size_t bytes, void* stream) {
    if (bytes == 0) return PB_OK;
    if (!dst || !src) { pb_set_error("pb_copy_d2d_async: null pointer"); return PB_ERR_INVALID; }
    PB_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(stream)));
    return PB_OK;
}
