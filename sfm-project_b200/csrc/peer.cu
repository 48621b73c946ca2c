// peer.cu -- result regions in peer memory: the multi-GPU gather of the pair loop's output.
//
// The reference appends every pair's matches to one Python list (code/pipeline.py:43-47).  Sharded over the GPUs of
// a box, each rank pushes the packed rows of its pair block into a region of the gathering rank's HBM with plain
// device-to-device copies over NVLink (copy engines; the SMs keep matching the next batch).  The region is
// allocated here with cudaMalloc (an IPC handle needs the base of an allocation, which a caching allocator's
// sub-block is not) and shared through cudaIpc*: one process per GPU, same box.
#include "common.cuh"

using namespace sfm;

static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI carries IPC handles as 64 opaque bytes");

extern "C" {

int sfm_peer_alloc(int device, size_t bytes, void** out_ptr, uint8_t out_handle[64])
{
    SFM_REQUIRE(out_ptr && out_handle && bytes > 0, "sfm_peer_alloc: bad argument");
    SFM_ON_DEVICE(device);
    void* p = nullptr;
    SFM_CUDA_CHECK(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return SFM_ERR_CUDA;
    }
    memcpy(out_handle, &h, 64);
    *out_ptr = p;
    return SFM_OK;
}

int sfm_peer_open(int device, const uint8_t handle[64], void** out_ptr)
{
    SFM_REQUIRE(out_ptr && handle, "sfm_peer_open: bad argument");
    SFM_ON_DEVICE(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    SFM_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    // The lazy flag only covers the mapping itself.  cudaMemcpyAsync into the region takes the direct NVLink path only when
    // peer access between the two DEVICES is enabled in this process; without it the copy is staged through host memory
    // (measured in round 2: 32 GB/s instead of NVLink speed).
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e == cudaSuccess && attr.device != device) {
        int can = 0;
        e = cudaDeviceCanAccessPeer(&can, device, attr.device);
        if (e == cudaSuccess && !can) {
            // No direct path between the two GPUs (different NVLink domains, P2P disabled): refuse, so that the caller
            // switches every rank to the send/recv transport instead of pushing through host memory.
            cudaIpcCloseMemHandle(p);
            set_error("sfm_peer_open: device %d has no direct peer access to device %d", device, attr.device);
            return SFM_ERR_DEVICE;
        }
        if (e == cudaSuccess) {
            e = cudaDeviceEnablePeerAccess(attr.device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
                e = cudaSuccess;
            }
        }
    }
    if (e != cudaSuccess) {
        cudaIpcCloseMemHandle(p);
        set_error("sfm_peer_open: enabling peer access failed: %s", cudaGetErrorString(e));
        return SFM_ERR_CUDA;
    }
    *out_ptr = p;
    return SFM_OK;
}

int sfm_peer_close(int device, void* ptr)
{
    SFM_REQUIRE(ptr != nullptr, "sfm_peer_close: NULL pointer");
    SFM_ON_DEVICE(device);
    SFM_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
    return SFM_OK;
}

int sfm_peer_free(int device, void* ptr)
{
    SFM_REQUIRE(ptr != nullptr, "sfm_peer_free: NULL pointer");
    SFM_ON_DEVICE(device);
    SFM_CUDA_CHECK(cudaFree(ptr));
    return SFM_OK;
}

int sfm_copy_async(void* dst, const void* src, size_t bytes, void* stream)
{
    SFM_REQUIRE(dst && src, "sfm_copy_async: NULL pointer");
    if (bytes == 0) return SFM_OK;
    SFM_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return SFM_OK;
}

}  // extern "C"
