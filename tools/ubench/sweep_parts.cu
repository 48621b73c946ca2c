// Microbenchmarks behind the round-2 sweep redesign: the two sides of the matcher pipeline measured in isolation.
//   part A  tensor side : int8 tcgen05.mma issue rate for the tile shapes under consideration, operands in shared memory,
//                         no epilogue: cta_group::1 M128 N128 / N256, cta_group::2 M256 N128 / N256 (2-CTA clusters).
//   part B  epilogue side: TMEM -> registers -> sub-group maxima -> tile maximum with the production arithmetic, no MMA and no
//                         handshakes: 8 warps software-pipelined (the round-1 epilogue), 16 warps with one 32-column buffer,
//                         16 warps with two 16-column buffers.
// All numbers are cycles per "B tile" = 256 query rows x 128 train rows per SM (the accounting unit of DESIGN.md: 10 MMAs of
// 64 cycles = 640 ideal with the K-extension, 512 without).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sweep_parts sweep_parts.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 20000;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (++spins > (1u << 16)) __trap();
    }
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) { return ((uint64_t)kDescHiSw128 << 32) | (((saddr & 0x3FFFFu) >> 4) | (1u << 16)); }
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

template <int CG>
__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// ---------------------------------------------------------------------------------------------- part A
// One issuing thread; per "step" kMmas MMAs of K = 32 into one of the accumulator stages (4 data + 1 extension in the matcher).
// CG = 2: launched as 2-CTA clusters, the leader CTA issues for both SMs (M = 256: 128 rows per CTA; each CTA provides half of B).
template <int CG, int N>
__global__ void __launch_bounds__(128, 1) mma_rate(int steps, int mmas_per_step, long long* cycles, int rotate)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    __shared__ uint32_t tmem_slot;
    __shared__ alignas(8) uint64_t done_bar;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0;
    constexpr int kBRows = (CG == 2) ? N / 2 : N;                     // rows of B this CTA holds
    constexpr int kBytes = 128 * 128 + kBRows * 128;                  // A tile (128 rows x 128 B) + B tile
    for (int e = threadIdx.x; e < kBytes / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem)[e] = 0x01010101u;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&done_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0 && rank == 0) {
        constexpr uint32_t id = idesc_i8(CG == 2 ? 256 : 128, N);
        const uint32_t a_addr = sbase, b_addr = sbase + 128 * 128;
        constexpr int kStages = 512 / N;
        t0 = clock64();
        for (int s = 0; s < steps; ++s) {
            const uint32_t d = tmem_base + (uint32_t)((s % kStages) * N);
            // rotate = 1: consecutive MMAs go to DIFFERENT accumulators (no back-to-back accumulation into one TMEM tile)
            for (int k = 0; k < mmas_per_step; ++k)
                mma_i8<CG>(rotate ? tmem_base + (uint32_t)(((s + k) % kStages) * N) : d, desc_sw128(a_addr + 32 * (k & 3)), desc_sw128(b_addr + 32 * (k & 3)), id, k > 0);
        }
        if (CG == 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
        else
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
        mbar_wait(smem_u32(&done_bar), 0);
        t1 = clock64();
        if (blockIdx.x == 0) *cycles = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- part B
#define LD32(v, addr)                                                                                                      \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                 \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                 \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),  \
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), \
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])              \
                 : "r"(addr) : "memory")
#define LD16(v, addr)                                                                                                      \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])               \
                 : "r"(addr) : "memory")
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int NV>
__device__ __forceinline__ void submax(const uint32_t (&u)[NV], int* c)
{
#pragma unroll
    for (int g = 0; g < NV / 8; ++g) {
        int m = __vimax3_s32((int)u[8 * g], (int)u[8 * g + 1], (int)u[8 * g + 2]);
        m = __vimax3_s32(m, (int)u[8 * g + 3], (int)u[8 * g + 4]);
        m = __vimax3_s32(m, (int)u[8 * g + 5], (int)u[8 * g + 6]);
        c[g] = max(m, (int)u[8 * g + 7]);
    }
}

struct Top3 {
    int M1, M2, M3, k1, k2, k3;
    bool tie4;
    __device__ void init() { M1 = -2147483645; M2 = -2147483646; M3 = -2147483647; k1 = 0xFFFF; k2 = 0xFFFF | (1 << 16); k3 = 0xFFFF | (2 << 16); tie4 = false; }
    __device__ __forceinline__ void tile(const int (&c)[16], int t, int4* park, int stride)
    {
        int m = __vimax3_s32(c[0], c[1], c[2]);
        m = __vimax3_s32(m, c[3], c[4]);
        m = __vimax3_s32(m, c[5], c[6]);
        m = __vimax3_s32(m, c[7], c[8]);
        m = __vimax3_s32(m, c[9], c[10]);
        m = __vimax3_s32(m, c[11], c[12]);
        m = __vimax3_s32(m, c[13], c[14]);
        m = max(m, c[15]);
        if (m >= M3) {
            if (m == M3) { tie4 = true; }
            else {
                const int slot = k3 >> 16;
                int4* dst = park + (slot * 4) * stride;
                dst[0] = make_int4(c[0], c[1], c[2], c[3]);
                dst[stride] = make_int4(c[4], c[5], c[6], c[7]);
                dst[2 * stride] = make_int4(c[8], c[9], c[10], c[11]);
                dst[3 * stride] = make_int4(c[12], c[13], c[14], c[15]);
                const int key = t | (slot << 16);
                if (m > M2) { tie4 = (M2 == M3); M3 = M2; k3 = k2; if (m > M1) { M2 = M1; k2 = k1; M1 = m; k1 = key; } else { M2 = m; k2 = key; } }
                else { tie4 = false; M3 = m; k3 = key; }
            }
        }
    }
};

// MODE 0: 8 warps, each 32 lanes x 128 columns per tile, three 32-register buffers, loads one chunk ahead (round-1 epilogue)
// MODE 1: 16 warps, each 32 lanes x 128 columns of every OTHER tile, one 32-register buffer, no intra-warp pipelining
// MODE 2: 16 warps, same split, two 16-register buffers (load 16 columns ahead)
// MODE 3: 16 warps, each 32 lanes x 64 columns of every tile (two warps share a row's tile), one 32-register buffer
// Every warp hands a "tile" back through a named barrier of its 4-warp group (stands in for the mbarrier handshake).
template <int MODE>
__global__ void __launch_bounds__(MODE == 0 ? 256 : 512, 1) epi_rate(int tiles, long long* cycles, int* sink, int seed)
{
    extern __shared__ int4 park_all[];
    __shared__ uint32_t tmem_slot;
    __shared__ alignas(8) uint64_t rel_bar[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&rel_bar[i]), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane_base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
    // fill this warp's lanes with pseudo-random values so that insertions into the top-3 happen at a realistic (falling) rate
    {
        uint32_t x = (uint32_t)(seed + threadIdx.x * 7919 + blockIdx.x * 104729);
        for (int col = 0; col < 512; col += 8) {
            if ((warp >> 2) == ((col >> 3) & ((blockDim.x >> 7) - 1))) {
                uint32_t v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { x = x * 1664525u + 1013904223u; v[j] = (x >> 12); }
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_base + col), "r"(v[0]), "r"(v[1]),
                             "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    Top3 T;
    T.init();
    int4* park = park_all + threadIdx.x;
    const int stride = blockDim.x;
    const long long t0 = clock64();
    if (MODE == 0) {
        // accumulator (tile & 1, row block = warp >> 2): columns ((tile & 1) * 2 + rb) * 128
        const int rb = warp >> 2;
        uint32_t va[32], vb[32], vc[32];
        int c[16];
        LD32(va, lane_base + (uint32_t)(rb * 128));
        for (int t = 0; t < tiles; ++t) {
            const uint32_t taddr = lane_base + (uint32_t)((((t & 1) * 2) + rb) * 128);
            wait_ld();
            LD32(vb, taddr + 32);
            LD32(vc, taddr + 64);
            submax<32>(va, c);
            wait_ld();
            LD32(va, taddr + 96);
            submax<32>(vb, c + 4);
            submax<32>(vc, c + 8);
            wait_ld();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&rel_bar[rb])) : "memory");   // release (4 arrivals complete a phase)
            submax<32>(va, c + 12);
            LD32(va, lane_base + (uint32_t)(((((t + 1) & 1) * 2) + rb) * 128));
            T.tile(c, t, park, stride);
        }
        wait_ld();
    } else if (MODE == 1 || MODE == 2) {
        const int rb = (warp >> 2) & 1, par = warp >> 3;               // this warp takes tiles with (t & 1) == par
        int c[16];
        for (int t = par; t < tiles; t += 2) {
            const uint32_t taddr = lane_base + (uint32_t)((((t & 1) * 2) + rb) * 128);
            if (MODE == 1) {
                uint32_t v[32];
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    LD32(v, taddr + 32 * ch);
                    wait_ld();
                    submax<32>(v, c + 4 * ch);
                }
            } else {
                uint32_t v0[16], v1[16];
                LD16(v0, taddr);
#pragma unroll
                for (int ch = 0; ch < 8; ch += 2) {
                    wait_ld();
                    LD16(v1, taddr + 16 * (ch + 1));
                    submax<16>(v0, c + 2 * ch);
                    wait_ld();
                    if (ch + 2 < 8) LD16(v0, taddr + 16 * (ch + 2));
                    submax<16>(v1, c + 2 * ch + 2);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&rel_bar[rb * 2 + par])) : "memory");
            T.tile(c, t, park, stride);
        }
    } else {
        // MODE 3: warp = (quadrant, rb, column half): 64 columns of every tile; the tile maximum is per half-tile ("tile" ids 2t + half)
        const int rb = (warp >> 2) & 1, half = warp >> 3;
        int c[16];
#pragma unroll
        for (int j = 8; j < 16; ++j) c[j] = -2147483647;
        for (int t = 0; t < tiles; ++t) {
            const uint32_t taddr = lane_base + (uint32_t)((((t & 1) * 2) + rb) * 128 + half * 64);
            uint32_t v[32];
            LD32(v, taddr);
            wait_ld();
            submax<32>(v, c);
            LD32(v, taddr + 32);
            wait_ld();
            submax<32>(v, c + 4);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&rel_bar[rb * 2 + half])) : "memory");
            T.tile(c, 2 * t + half, park, stride);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    if (T.M1 == 0x12345678 || T.tie4) sink[threadIdx.x] = T.M2 + T.k3;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512) : "memory");
}

template <typename K, typename... Args>
static void launch_cluster2(K k, int grid, int block, int smem, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, args...);
    if (e != cudaSuccess) printf("cluster launch failed: %s\n", cudaGetErrorString(e));
}

template <typename K>
static void set_smem(K k, int bytes) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); }

int main()
{
    long long* dcyc;
    int* sink;
    cudaMalloc(&dcyc, 8);
    cudaMalloc(&sink, 4096);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    auto report = [&](const char* name, double btiles_per_cta) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-82s ERROR %s\n", name, cudaGetErrorString(e)); exit(1); }
        long long cyc = 0;
        cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
        printf("%-82s %9lld cycles  -> %7.1f cycles per B tile (256 q x 128 t per SM)\n", name, cyc, (double)cyc / btiles_per_cta);
    };
    // ---- part A.  "steps" accumulators of N columns; a B tile = 2 row blocks x 128 columns x (4 or 5) MMAs
    const int steps = 4096;
    for (int rot : {0, 1})
    for (int mm : {5, 4}) {
        if (rot && mm == 4) continue;
        char nm[128];
        const int smem = 128 * 128 + 256 * 128 + 2048;
        set_smem(mma_rate<1, 128>, smem); set_smem(mma_rate<1, 256>, smem); set_smem(mma_rate<2, 128>, smem); set_smem(mma_rate<2, 256>, smem);
        for (int rep = 0; rep < 2; ++rep) mma_rate<1, 128><<<sms, 128, smem>>>(steps, mm, dcyc, rot);
        snprintf(nm, sizeof nm, "A: cta_group::1 M128 N128, %d MMAs / accumulator%s", mm, rot ? ", rotating accumulators" : "");
        report(nm, steps / 2.0);                                        // 2 accumulators (row blocks) per B tile
        for (int rep = 0; rep < 2; ++rep) mma_rate<1, 256><<<sms, 128, smem>>>(steps, mm, dcyc, rot);
        snprintf(nm, sizeof nm, "A: cta_group::1 M128 N256, %d MMAs / accumulator%s", mm, rot ? ", rotating accumulators" : "");
        report(nm, steps / 1.0);                                        // 128 rows x 256 columns = one B tile's worth of work
        for (int rep = 0; rep < 2; ++rep) launch_cluster2(mma_rate<2, 128>, sms & ~1, 128, smem, steps, mm, dcyc, rot);
        snprintf(nm, sizeof nm, "A: cta_group::2 M256 N128, %d MMAs / accumulator%s", mm, rot ? ", rotating accumulators" : "");
        report(nm, steps / 2.0);                                        // per SM: 128 rows x 128 columns per step
        for (int rep = 0; rep < 2; ++rep) launch_cluster2(mma_rate<2, 256>, sms & ~1, 128, smem, steps, mm, dcyc, rot);
        snprintf(nm, sizeof nm, "A: cta_group::2 M256 N256, %d MMAs / accumulator%s", mm, rot ? ", rotating accumulators" : "");
        report(nm, steps / 1.0);
    }
    // ---- part B.  tiles = 128-column accumulators per row block; a B tile = one tile of both row blocks
    const int tiles = 4096;
    set_smem(epi_rate<0>, 3 * 4 * 256 * 16); set_smem(epi_rate<1>, 3 * 4 * 512 * 16); set_smem(epi_rate<2>, 3 * 4 * 512 * 16); set_smem(epi_rate<3>, 3 * 4 * 512 * 16);
    for (int rep = 0; rep < 2; ++rep) epi_rate<0><<<sms, 256, 3 * 4 * 256 * 16>>>(tiles, dcyc, sink, 1);
    report("B: 8 warps, 3 x 32-col buffers, pipelined (round 1)", tiles);
    for (int rep = 0; rep < 2; ++rep) epi_rate<1><<<sms, 512, 3 * 4 * 512 * 16>>>(tiles, dcyc, sink, 1);
    report("B: 16 warps, alternate tiles, one 32-col buffer", tiles);
    for (int rep = 0; rep < 2; ++rep) epi_rate<2><<<sms, 512, 3 * 4 * 512 * 16>>>(tiles, dcyc, sink, 1);
    report("B: 16 warps, alternate tiles, two 16-col buffers", tiles);
    for (int rep = 0; rep < 2; ++rep) epi_rate<3><<<sms, 512, 3 * 4 * 512 * 16>>>(tiles, dcyc, sink, 1);
    report("B: 16 warps, 64 columns of every tile, one 32-col buffer", tiles);
    return 0;
}
