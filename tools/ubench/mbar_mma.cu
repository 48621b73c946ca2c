// Microbenchmarks: (1) latency of mbarrier.try_wait on an already-complete phase, alone and while other warps
// poll an incomplete barrier; (2) cycles the issuing thread spends on n back-to-back tcgen05.mma (queue depth).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// mode 0: try_wait complete, nobody else; mode 1: + pollers using try_wait on an incomplete barrier;
// mode 2: + pollers using test_wait + nanosleep(32); mode 3: measured thread uses test_wait, pollers try_wait
__global__ void mbar_bench(int mode, int pollers, long long* out)
{
    __shared__ alignas(8) uint64_t done_bar, never_bar;
    __shared__ volatile int stop;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&done_bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&never_bar)));
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&done_bar)) : "memory");   // phase 0 complete
        stop = 0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        if (threadIdx.x == 0) {
            long long t0 = clock64();
            int ok = 0;
            for (int i = 0; i < 2000; ++i) ok += (mode == 3) ? test_wait(smem_u32(&done_bar), 0) : try_wait(smem_u32(&done_bar), 0);
            long long t1 = clock64();
            out[0] = t1 - t0;
            out[1] = ok;
            stop = 1;
        }
    } else if (warp <= pollers) {
        while (!stop) {
            if (mode == 2) { test_wait(smem_u32(&never_bar), 0); __nanosleep(32); }
            else try_wait(smem_u32(&never_bar), 0);
        }
    }
}

__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void issue_bench(int n_mma, long long* out)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t slot;
    __shared__ alignas(8) uint64_t bar;
    for (int e = threadIdx.x; e < 32768 / 4; e += blockDim.x) ((uint32_t*)smem)[e] = 0x01010101u;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (16u << 17) | (8u << 24);
        const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t a_lo = ((smem_u32(smem) & 0x3FFFF) >> 4) | (1u << 16);
        const uint32_t b_lo = (((smem_u32(smem) + 16384) & 0x3FFFF) >> 4) | (1u << 16);
        long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i)
            mma_i8(tm + (uint32_t)((i & 3) * 128), ((uint64_t)hi << 32) | (a_lo + 2 * (i & 3)), ((uint64_t)hi << 32) | (b_lo + 2 * (i & 3)), idesc, i & 1);
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        long long t2 = clock64();
        while (!try_wait(smem_u32(&bar), 0)) {}
        long long t3 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512) : "memory");
}

int main()
{
    long long* d; cudaMalloc(&d, 64);
    long long h[4];
    for (int mode = 0; mode < 4; ++mode)
        for (int pollers : {0, 4, 10}) {
            if (mode == 0 && pollers) continue;
            if (mode != 0 && !pollers) continue;
            mbar_bench<<<1, 32 * 12>>>(mode, pollers, d);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("mbar mode %d pollers %2d: %.1f cycles per satisfied wait (ok=%lld)\n", mode, pollers, h[0] / 2000.0, h[1]);
        }
    cudaFuncSetAttribute(issue_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    for (int n : {1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24, 32, 64}) {
        for (int rep = 0; rep < 2; ++rep) {
            issue_bench<<<1, 128, 40000>>>(n, d);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        }
        cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
        printf("issue n=%2d: issue loop %5lld cycles, commit %4lld, until complete %5lld (ideal exec %d)\n", n, h[0], h[1], h[2], n * 64);
    }
    return 0;
}
