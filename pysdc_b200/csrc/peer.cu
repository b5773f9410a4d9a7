// Peer-mapped device memory for the slab-decomposed solver: each rank allocates one segment, exports a cudaIpc handle,
// and maps the segments of the other ranks of the node (the handles travel through any host channel, e.g.
// torch.distributed).  Kernels then read / write the neighbours' memory directly over NVLink.
#include "common.cuh"

using namespace sdcb200;

extern "C" {

int sdcb200_peer_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64) {
    SDC_REQUIRE(dev_ptr != nullptr && handle64 != nullptr && bytes > 0, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == SDCB200_IPC_HANDLE_BYTES, "unexpected cudaIpcMemHandle_t size");
    void* p = nullptr;
    SDC_CUDA_OK(cudaMalloc(&p, bytes));
    SDC_CUDA_OK(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail_cuda("sdcb200_peer_alloc", e);
    }
    memcpy(handle64, &h, sizeof(h));
    *dev_ptr = p;
    return 0;
}

int sdcb200_peer_open(const unsigned char* handle64, void** peer_ptr) {
    SDC_REQUIRE(handle64 != nullptr && peer_ptr != nullptr, "bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    SDC_CUDA_OK(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int sdcb200_peer_close(void* peer_ptr) {
    SDC_CUDA_OK(cudaIpcCloseMemHandle(peer_ptr));
    return 0;
}

int sdcb200_peer_free(void* dev_ptr) {
    SDC_CUDA_OK(cudaFree(dev_ptr));
    return 0;
}

}  // extern "C"
