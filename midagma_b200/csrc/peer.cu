// Peer-visible device memory for the kernels that exchange data between the GPUs of one box by plain loads / stores
// over NVLink (one process per GPU): a cudaMalloc allocation of its own (an IPC handle names a whole allocation),
// exported as a 64-byte handle that the other processes import into their address space.  The exchange protocols
// live in the kernels that use the memory (csrc/lin_iter.cu: partial score products of the row-sharded logistic loss).
#include "common.cuh"
#include "../../include/dagma_b200.h"

using namespace dagma;

static_assert(sizeof(cudaIpcMemHandle_t) == DAGMA_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");

extern "C" int dagma_peer_alloc(size_t bytes, void** out_dev) {
    DAGMA_REQUIRE(out_dev && bytes > 0, "bad arguments");
    void* p = nullptr;
    DAGMA_CUDA_OK(cudaMalloc(&p, bytes));
    DAGMA_CUDA_OK(cudaMemset(p, 0, bytes));
    DAGMA_CUDA_OK(cudaDeviceSynchronize());
    *out_dev = p;
    return 0;
}

extern "C" int dagma_peer_free(void* dev) {
    if (dev) DAGMA_CUDA_OK(cudaFree(dev));
    return 0;
}

extern "C" int dagma_peer_export(void* dev, unsigned char* handle_out) {
    DAGMA_REQUIRE(dev && handle_out, "null pointer");
    cudaIpcMemHandle_t h;
    DAGMA_CUDA_OK(cudaIpcGetMemHandle(&h, dev));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

extern "C" int dagma_peer_import(const unsigned char* handle, void** out_dev) {
    DAGMA_REQUIRE(handle && out_dev, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    DAGMA_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *out_dev = p;
    return 0;
}

extern "C" int dagma_peer_release(void* imported_dev) {
    if (imported_dev) DAGMA_CUDA_OK(cudaIpcCloseMemHandle(imported_dev));
    return 0;
}

// A stream of its own for a batch lane (linear._run_lanes): created when the lane starts, so that consecutive lanes get
// consecutive hardware work queues -- streams handed out by a framework's pool were created long before, interleaved
// with other streams, and several of them can share a queue, which serialises the lanes' long persistent kernels.
extern "C" int dagma_stream_create(void** out_stream) {
    DAGMA_REQUIRE(out_stream, "null pointer");
    cudaStream_t st = nullptr;
    DAGMA_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *out_stream = (void*)st;
    return 0;
}

extern "C" int dagma_stream_destroy(void* stream) {
    if (stream) DAGMA_CUDA_OK(cudaStreamDestroy((cudaStream_t)stream));
    return 0;
}
