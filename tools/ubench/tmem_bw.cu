// Microbenchmark: TMEM -> register read bandwidth per SM for several tcgen05.ld shapes and warp counts.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int SHAPE>
__device__ __forceinline__ uint32_t ld_once(uint32_t taddr);

// 32x32b.x32 : 32 regs, 4 KB per warp
template <> __device__ __forceinline__ uint32_t ld_once<0>(uint32_t taddr) {
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
          "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) x ^= v[i];
    return x;
}
// 32x32b.x8 : 8 regs, 1 KB per warp
template <> __device__ __forceinline__ uint32_t ld_once<1>(uint32_t taddr) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return v[0]^v[1]^v[2]^v[3]^v[4]^v[5]^v[6]^v[7];
}
// 16x256b.x4 : 16 regs? (16 lanes x 256 bit x4 = 2 KB per warp) -> 16 regs per thread
template <> __device__ __forceinline__ uint32_t ld_once<2>(uint32_t taddr) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) x ^= v[i];
    return x;
}
// 32x32b.x32 issued twice before one wait (two loads in flight)
template <> __device__ __forceinline__ uint32_t ld_once<3>(uint32_t taddr) {
    uint32_t v[32], w[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
          "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(w[0]),"=r"(w[1]),"=r"(w[2]),"=r"(w[3]),"=r"(w[4]),"=r"(w[5]),"=r"(w[6]),"=r"(w[7]),"=r"(w[8]),"=r"(w[9]),"=r"(w[10]),"=r"(w[11]),"=r"(w[12]),"=r"(w[13]),"=r"(w[14]),"=r"(w[15]),
          "=r"(w[16]),"=r"(w[17]),"=r"(w[18]),"=r"(w[19]),"=r"(w[20]),"=r"(w[21]),"=r"(w[22]),"=r"(w[23]),"=r"(w[24]),"=r"(w[25]),"=r"(w[26]),"=r"(w[27]),"=r"(w[28]),"=r"(w[29]),"=r"(w[30]),"=r"(w[31])
        : "r"(taddr + 32) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) x ^= v[i] ^ w[i];
    return x;
}

template <int SHAPE>
__global__ void bench(int iters, long long* cycles, uint32_t* sink)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t x = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x ^= ld_once<SHAPE>(base + (uint32_t)((i * 64) & 255));
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    if (x == 0x12345678u) sink[threadIdx.x] = x;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(slot), "r"(512) : "memory");
}

int main()
{
    long long* dcyc; uint32_t* sink;
    cudaMalloc(&dcyc, 8); cudaMalloc(&sink, 4096);
    const int iters = 2000;
    const double bytes_per_warp_iter[4] = {4096, 1024, 2048, 8192};
    const char* names[4] = {"32x32b.x32", "32x32b.x8", "16x256b.x4", "2 x 32x32b.x32 in flight"};
    for (int shape = 0; shape < 4; ++shape)
        for (int warps : {4, 8, 16}) {
            long long cyc = 0;
            for (int rep = 0; rep < 2; ++rep) {
                switch (shape) {
                    case 0: bench<0><<<148, warps * 32>>>(iters, dcyc, sink); break;
                    case 1: bench<1><<<148, warps * 32>>>(iters, dcyc, sink); break;
                    case 2: bench<2><<<148, warps * 32>>>(iters, dcyc, sink); break;
                    case 3: bench<3><<<148, warps * 32>>>(iters, dcyc, sink); break;
                }
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
            const double bytes = bytes_per_warp_iter[shape] * warps * iters;
            printf("%-28s warps=%2d  cycles=%9lld  %.1f B/clk/SM  (%.1f cycles per warp-load)\n", names[shape], warps, cyc,
                   bytes / cyc, (double)cyc / iters);
        }
    return 0;
}
