// match_tc.cu -- tcgen05 int8 distance-GEMM matcher with fused chunk-maximum epilogue (K2).
//
// Replaces cv2.BFMatcher(NORM_L2).knnMatch(k=2) for a list of image pairs (north-star workload;
// displaces bf.match at code/feature_matching.py:50 inside the pair loop code/pipeline.py:38-41).
//
// Work unit  = (pair, block of 256 query rows) swept over every 128-row tile of the train image.
// Per tile   : acc[q,t] = a_s(q).b_s(t) + (H0 - floor(|b_s(t)|^2/2))     (int32, exact)
//              = 4 x tcgen05.mma kind::i8 s8*s8 (K=128, SWIZZLE_128B operands from TMA)
//              + 1 x tcgen05.mma kind::i8 u8*u8 on the 32-byte K-extension (no-swizzle tiles),
//              so the squared distance is D = |a_s|^2 + 2*H0 - 2*acc + (|b_s|^2 & 1): larger acc <=> smaller D.
// Epilogue   : one thread per query row keeps the two largest 32-column chunk maxima (value, chunk) and a
//              tie flag -- 0.5 VIMNMX3 per distance, no per-element index work.
// Refinement : the exact top-2 (distance, index) lies inside those two chunks unless the flag is set
//              (proof in DESIGN.md); 4 refine warps recompute the 64 candidates exactly with dp4a, rows
//              with a tie are brute-forced over the whole train image.  Results are bit-exact.
//
// Warp roles (512 threads, 1 CTA / SM, persistent over units):
//   warp 0      TMA producer         warp 1      MMA issuer + TMEM owner
//   warps 4-7   epilogue row block 0 warps 8-11  epilogue row block 1
//   warps 12-15 exact refinement     (warps 2,3 idle)
#include "match_common.cuh"

namespace sfm {

constexpr int kUnitRows = 256;
constexpr int kStages = 5;
constexpr int kTileBytes = kTileRows * kDescDim;             // 16384
constexpr int kBStageBytes = kTileBytes + kExtTileBytes;     // 20480
constexpr int kABufBytes = 2 * kTileBytes;                   // 32768
constexpr int kTcThreads = 512;
constexpr int kChunk = 32;
constexpr int kTmemCols = 512;
constexpr int kMaskedAcc = -2147483646;                      // INT_MIN + 2: below every real accumulator

struct TcSmem {
    static constexpr int kA = 0;
    static constexpr int kB = kA + 2 * kABufBytes;
    static constexpr int kAext = kB + kStages * kBStageBytes;
    static constexpr int kRef = kAext + kExtTileBytes;
    static constexpr int kBar = kRef + 2 * kUnitRows * 8;
    static constexpr int kNumBar = 2 + 2 + 2 * kStages + 4 + 4 + 2 + 2;
    static constexpr int kTmemSlot = kBar + kNumBar * 8;
    static constexpr int kTotal = kTmemSlot + 16;
};
constexpr int kTcSmemBytes = TcSmem::kTotal + 1024;          // + alignment slack

// ------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (sticky error reported to the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {
            printf("sfm_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B operand: rows of 128 B, 8-row groups 1024 B apart (SBO), descriptor version 1.
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// K-major no-swizzle operand [k-chunk][row][16 B]: LBO = 2048 B between the two K chunks, SBO = 128 B between 8-row groups.
__device__ __forceinline__ uint64_t desc_ext(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((kTileRows * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
           ((uint64_t)1 << 46);
}
// kind::i8 instruction descriptor: D = s32, M = 128, N = 128, both operands K-major.
__host__ __device__ constexpr uint32_t idesc_i8(uint32_t a_signed, uint32_t b_signed)
{
    return (2u << 4) | (a_signed << 7) | (b_signed << 10) | ((uint32_t)(kTileRows >> 3) << 17) | ((uint32_t)(kTileRows >> 4) << 24);
}

__device__ __forceinline__ int max32(const uint32_t (&u)[32])
{
    int c[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        int m = __vimax3_s32((int)u[8 * g], (int)u[8 * g + 1], (int)u[8 * g + 2]);
        m = __vimax3_s32(m, (int)u[8 * g + 3], (int)u[8 * g + 4]);
        m = __vimax3_s32(m, (int)u[8 * g + 5], (int)u[8 * g + 6]);
        c[g] = max(m, (int)u[8 * g + 7]);
    }
    return max(__vimax3_s32(c[0], c[1], c[2]), c[3]);
}

struct UnitInfo {
    int pair, qblk, img_q, img_t, nq, nt, tiles;
    bool live;
};

__device__ __forceinline__ UnitInfo decode_unit(int u, int units_per_pair, const int32_t* __restrict__ pairs,
                                                const int32_t* __restrict__ count)
{
    UnitInfo I;
    I.pair = u / units_per_pair;
    I.qblk = u - I.pair * units_per_pair;
    I.img_q = __ldg(pairs + 2 * I.pair);
    I.img_t = __ldg(pairs + 2 * I.pair + 1);
    I.nq = __ldg(count + I.img_q);
    I.nt = __ldg(count + I.img_t);
    I.tiles = (I.nt + kTileRows - 1) / kTileRows;
    I.live = (I.qblk * kUnitRows < I.nq) && I.nt > 0;
    return I;
}

__global__ void __launch_bounds__(kTcThreads, 1) match_tc_kernel(
    const __grid_constant__ CUtensorMap tmap_desc, const int8_t* __restrict__ desc, const int8_t* __restrict__ ext,
    const int32_t* __restrict__ norm, const int32_t* __restrict__ count, const int32_t* __restrict__ pairs, int n_pairs,
    int feat_stride, int32_t* __restrict__ knn_out, int32_t* __restrict__ dbg_acc, int dbg_mode)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + TcSmem::kBar;
    // barrier indices
    auto bar_a_full = [&](int i) { return bar0 + 8 * (0 + i); };
    auto bar_a_empty = [&](int i) { return bar0 + 8 * (2 + i); };
    auto bar_b_full = [&](int i) { return bar0 + 8 * (4 + i); };
    auto bar_b_empty = [&](int i) { return bar0 + 8 * (4 + kStages + i); };
    auto bar_t_full = [&](int st, int rb) { return bar0 + 8 * (4 + 2 * kStages + st * 2 + rb); };
    auto bar_t_empty = [&](int st, int rb) { return bar0 + 8 * (8 + 2 * kStages + st * 2 + rb); };
    auto bar_r_full = [&](int i) { return bar0 + 8 * (12 + 2 * kStages + i); };
    auto bar_r_empty = [&](int i) { return bar0 + 8 * (14 + 2 * kStages + i); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + TcSmem::kTmemSlot);
    int2* ref_buf = reinterpret_cast<int2*>(smem + TcSmem::kRef);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units_per_pair = feat_stride / kUnitRows;
    const int total_units = n_pairs * units_per_pair;

    // ---- one-time setup
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(bar_a_full(i), 1); mbar_init(bar_a_empty(i), 1); }
        for (int i = 0; i < kStages; ++i) { mbar_init(bar_b_full(i), 1); mbar_init(bar_b_empty(i), 1); }
        for (int st = 0; st < 2; ++st)
            for (int rb = 0; rb < 2; ++rb) { mbar_init(bar_t_full(st, rb), 1); mbar_init(bar_t_empty(st, rb), 4); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_r_full(i), 8); mbar_init(bar_r_empty(i), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // constant A-side K-extension tile: weights 255 x 24, 1, 0 x 7 for every query row
    for (int e = threadIdx.x; e < kExtTileBytes / 4; e += kTcThreads) {
        const int chunk = e / (kTileRows * 4), w = e & 3;       // word w of the 16-byte row slice
        uint32_t val = 0xFFFFFFFFu;
        if (chunk == 1) val = (w < 2) ? 0xFFFFFFFFu : (w == 2 ? 0x00000001u : 0u);
        reinterpret_cast<uint32_t*>(smem + TcSmem::kAext)[e] = val;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= TMA producer
        if (lane == 0) {
            int ucount = 0, bit = 0;
            for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
                const UnitInfo I = decode_unit(u, units_per_pair, pairs, count);
                if (!I.live) continue;
                const int abuf = ucount & 1, aph = (ucount >> 1) & 1;
                const int qrow0 = I.img_q * feat_stride + I.qblk * kUnitRows;
                mbar_wait(bar_a_empty(abuf), aph ^ 1);
                mbar_expect_tx(bar_a_full(abuf), kABufBytes);
                tma_load_2d(sbase + TcSmem::kA + abuf * kABufBytes, &tmap_desc, 0, qrow0, bar_a_full(abuf));
                tma_load_2d(sbase + TcSmem::kA + abuf * kABufBytes + kTileBytes, &tmap_desc, 0, qrow0 + kTileRows, bar_a_full(abuf));
                const int trow0 = I.img_t * feat_stride;
                for (int t = 0; t < I.tiles; ++t, ++bit) {
                    const int s = bit % kStages, ph = (bit / kStages) & 1;
                    mbar_wait(bar_b_empty(s), ph ^ 1);
                    mbar_expect_tx(bar_b_full(s), kBStageBytes);
                    const int row = trow0 + t * kTileRows;
                    const uint32_t dst = sbase + TcSmem::kB + s * kBStageBytes;
                    tma_load_2d(dst, &tmap_desc, 0, row, bar_b_full(s));
                    bulk_load(dst + kTileBytes, ext + (long long)(row / kTileRows) * kExtTileBytes, kExtTileBytes, bar_b_full(s));
                }
                ++ucount;
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        if (lane == 0) {
            constexpr uint32_t id_main = idesc_i8(1, 1);
            constexpr uint32_t id_ext = idesc_i8(0, 0);
            const uint64_t aext_desc = desc_ext(sbase + TcSmem::kAext);
            int ucount = 0, bit = 0, tcount = 0;
            for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
                const UnitInfo I = decode_unit(u, units_per_pair, pairs, count);
                if (!I.live) continue;
                const int abuf = ucount & 1, aph = (ucount >> 1) & 1;
                mbar_wait(bar_a_full(abuf), aph);
                tc_fence_after();
                for (int t = 0; t < I.tiles; ++t, ++bit, ++tcount) {
                    const int s = bit % kStages, ph = (bit / kStages) & 1;
                    const int st = tcount & 1, tph = (tcount >> 1) & 1;
                    mbar_wait(bar_b_full(s), ph);
                    tc_fence_after();
                    const uint32_t b_addr = sbase + TcSmem::kB + s * kBStageBytes;
                    const uint64_t bext_desc = desc_ext(b_addr + kTileBytes);
#pragma unroll
                    for (int rb = 0; rb < 2; ++rb) {
                        mbar_wait(bar_t_empty(st, rb), tph ^ 1);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + (uint32_t)((st * 2 + rb) * kTileRows);
                        const uint32_t a_addr = sbase + TcSmem::kA + abuf * kABufBytes + rb * kTileBytes;
                        if (dbg_mode != 2) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                tc_mma_i8(d_tmem, desc_sw128(a_addr + 32 * k), desc_sw128(b_addr + 32 * k), id_main, k > 0);
                        }
                        if (dbg_mode != 1) tc_mma_i8(d_tmem, aext_desc, bext_desc, id_ext, dbg_mode != 2);
                        tc_commit(bar_t_full(st, rb));
                    }
                    tc_commit(bar_b_empty(s));
                }
                tc_commit(bar_a_empty(abuf));
                ++ucount;
            }
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 12) {
        // ================================================================= epilogue: chunk maxima
        const int rb = (warp - 4) >> 2, wq = warp & 3;
        const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
        int ucount = 0, tcount = 0;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
            const UnitInfo I = decode_unit(u, units_per_pair, pairs, count);
            if (!I.live) continue;
            int M1 = kMaskedAcc + 1, M2 = kMaskedAcc, c1 = -1, c2 = -1;
            bool tie = false;
            for (int t = 0; t < I.tiles; ++t, ++tcount) {
                const int st = tcount & 1, tph = (tcount >> 1) & 1;
                mbar_wait(bar_t_full(st, rb), tph);
                tc_fence_after();
                const int valid = min(kTileRows, I.nt - t * kTileRows);
                const int nch = (valid + kChunk - 1) / kChunk;
                const uint32_t taddr = lane_base + (uint32_t)((st * 2 + rb) * kTileRows);
                for (int ch = 0; ch < nch; ++ch) {
                    uint32_t v[32];
                    tc_ld32(taddr + ch * kChunk, v);
                    tc_wait_ld();
                    if (ch == nch - 1) {              // all TMEM reads of this stage are done: hand it back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_t_empty(st, rb));
                    }
                    if (dbg_acc != nullptr && ucount == 0 && t == 0 && blockIdx.x == 0) {
                        int32_t* o = dbg_acc + (long long)(rb * kTileRows + wq * 32 + lane) * kTileRows + ch * kChunk;
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[j] = (int)v[j];
                    }
                    if ((ch + 1) * kChunk > valid) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (ch * kChunk + j >= valid) v[j] = (uint32_t)kMaskedAcc;
                    }
                    const int m = max32(v);
                    const int c = t * (kTileRows / kChunk) + ch;
                    if (m > M1) { tie = (M2 == M1); M2 = M1; c2 = c1; M1 = m; c1 = c; }
                    else if (m > M2) { M2 = m; c2 = c; tie = false; }
                    else if (m == M2) tie = true;
                }
            }
            const int ubuf = ucount & 1, uph = (ucount >> 1) & 1;
            mbar_wait(bar_r_empty(ubuf), uph ^ 1);
            ref_buf[ubuf * kUnitRows + rb * kTileRows + wq * 32 + lane] = make_int2(c1, tie ? -2 : c2);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_r_full(ubuf));
            ++ucount;
        }
    } else if (warp >= 12) {
        // ================================================================= exact refinement
        const int wr = warp - 12;
        int ucount = 0;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
            const UnitInfo I = decode_unit(u, units_per_pair, pairs, count);
            if (!I.live) continue;
            const int ubuf = ucount & 1, uph = (ucount >> 1) & 1;
            mbar_wait(bar_r_full(ubuf), uph);
            const long long qrow0 = (long long)I.img_q * feat_stride + I.qblk * kUnitRows;
            const long long trow0 = (long long)I.img_t * feat_stride;
            const int rows = min(kUnitRows, I.nq - I.qblk * kUnitRows);
            for (int r = wr; r < rows; r += 4) {
                const int2 cand = ref_buf[ubuf * kUnitRows + r];
                int a[32];
                load_query_row(a, desc, qrow0 + r);
                const int na = __ldg(norm + qrow0 + r);
                Top2 best;
                if (cand.y == -2) {
                    best = warp_bruteforce_row(a, na, desc, norm, trow0, I.nt, lane);
                } else {
                    best.clear();
                    const int t1 = cand.x * kChunk + lane;
                    if (cand.x >= 0 && t1 < I.nt) best.push(exact_sqdist(a, na, desc, norm, trow0 + t1), t1);
                    const int t2 = cand.y * kChunk + lane;
                    if (cand.y >= 0 && t2 < I.nt) best.push(exact_sqdist(a, na, desc, norm, trow0 + t2), t2);
                    best = warp_merge(best);
                }
                if (lane == 0) store_knn(knn_out + ((long long)I.pair * feat_stride + I.qblk * kUnitRows + r) * 4, best);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_r_empty(ubuf));
            ++ucount;
        }
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

int launch_match_tc(const sfm_bank* b, const int32_t* pairs, int n_pairs, int grid_req, int32_t* knn_out, int32_t* dbg_acc,
                    int dbg_mode, cudaStream_t st)
{
    if (!b->tmap_ready) {
        set_error("bank has no descriptor tensor map (metric must be L2)");
        return SFM_ERR_STATE;
    }
    static bool attr_set = false;
    if (!attr_set) {
        SFM_CUDA_CHECK(cudaFuncSetAttribute(match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
        attr_set = true;
    }
    const long long units = (long long)n_pairs * (b->L.feat_stride / kUnitRows);
    int grid = grid_req > 0 ? grid_req : b->sm_count;
    if (grid > units) grid = (int)units;
    if (grid < 1) grid = 1;
    match_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, st>>>(b->tmap_desc, b->desc, b->ext, b->norm, b->count, pairs, n_pairs,
                                                           (int)b->L.feat_stride, knn_out, dbg_acc, dbg_mode);
    SFM_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return SFM_OK;
}

// ------------------------------------------------------------------------------------ tensor-pipe probe
// Same MMA shape and operand layouts as the matcher, no epilogue: the attainable int8 tensor rate.
__global__ void __launch_bounds__(128, 1) probe_int8_kernel(int n_tiles)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    __shared__ uint32_t tmem_slot;
    __shared__ alignas(8) uint64_t done_bar;
    const int warp = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < (2 * kTileBytes + 2 * kExtTileBytes) / 4; e += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[e] = 0x01010101u;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&done_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t id_main = idesc_i8(1, 1);
        constexpr uint32_t id_ext = idesc_i8(0, 0);
        const uint32_t a_addr = sbase, b_addr = sbase + kTileBytes;
        const uint64_t ae = desc_ext(sbase + 2 * kTileBytes), be = desc_ext(sbase + 2 * kTileBytes + kExtTileBytes);
        for (int t = 0; t < n_tiles; ++t) {
            const uint32_t d = tmem_base + (uint32_t)((t & 3) * kTileRows);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_i8(d, desc_sw128(a_addr + 32 * k), desc_sw128(b_addr + 32 * k), id_main, k > 0);
            tc_mma_i8(d, ae, be, id_ext, 1);
        }
        tc_commit(smem_u32(&done_bar));
        mbar_wait(smem_u32(&done_bar), 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

}  // namespace sfm

using namespace sfm;

extern "C" int sfm_probe_int8_mma(int device, int n_tiles, float* out_ms, double* out_ops)
{
    SFM_REQUIRE(out_ms && out_ops && n_tiles > 0, "sfm_probe_int8_mma: bad argument");
    cudaDeviceProp p;
    SFM_CUDA_CHECK(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) {
        set_error("device %d is sm_%d%d; sm_100a required", device, p.major, p.minor);
        return SFM_ERR_DEVICE;
    }
    SFM_CUDA_CHECK(cudaSetDevice(device));
    const int smem = 2 * kTileBytes + 2 * kExtTileBytes + 1024;
    SFM_CUDA_CHECK(cudaFuncSetAttribute(probe_int8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    SFM_CUDA_CHECK(cudaEventCreate(&e0));
    SFM_CUDA_CHECK(cudaEventCreate(&e1));
    probe_int8_kernel<<<p.multiProcessorCount, 128, smem>>>(64);          // warm-up
    SFM_CUDA_CHECK(cudaEventRecord(e0));
    probe_int8_kernel<<<p.multiProcessorCount, 128, smem>>>(n_tiles);
    SFM_CUDA_CHECK(cudaEventRecord(e1));
    SFM_CUDA_CHECK(cudaEventSynchronize(e1));
    SFM_CUDA_CHECK(cudaGetLastError());
    count_launch(2);
    SFM_CUDA_CHECK(cudaEventElapsedTime(out_ms, e0, e1));
    // algorithmic ops credited per tile: 2 * 128 * 128 * 128 (the K-extension is overhead, not credited)
    *out_ops = (double)p.multiProcessorCount * n_tiles * 2.0 * kTileRows * kTileRows * kDescDim;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return SFM_OK;
}
