// match_tc2.cu -- the tcgen05 matcher sweep on CTA PAIRS: tcgen05.mma.cta_group::2 (K2, variant "pair").
//
// Same arithmetic, same candidate records and same refinement as match_tc.cu (it replaces the same reference call,
// bf.match at code/feature_matching.py:50 inside the pair loop code/pipeline.py:38-41); what changes is the mapping.
// Round-2 measurements (tools/ubench/sweep_parts.cu -> profiles/r02_ubench_sweep_parts.log, clock64 timelines ->
// profiles/r02_pair_trace_*.log) showed that neither the tensor pipe (640 cycles per 256 x 128 block of distances) nor the
// epilogue arithmetic (440) bounds the one-CTA kernel (880-914): the latency of one accumulator's round trip does
// (issue -> MMA complete -> epilogue wake-up -> four dependent TMEM loads at ~100 cycles each -> release -> issuer wake-up)
// with only 1,280 cycles of MMA work in flight in the 512 TMEM columns and two accumulators per row block.  Hence:
//   * a unit (pair, 256 query rows) is shared by the two CTAs of a cluster; ONE thread of the leader CTA issues
//     M = 256 MMAs for both SMs (cta_group::2: CTA r owns query rows [128 r, 128 r + 128) and provides train rows
//     [64 r, 64 r + 64) of each 128-row tile), i.e. half the instructions, commits and barrier waits per unit of work;
//   * with one row block per CTA the 512 TMEM columns hold FOUR accumulator stages of 128 columns, and each stage has its
//     OWN set of four epilogue warps (16 epilogue warps per CTA): a set has four tile periods for its dependent chain of
//     TMEM loads, sub-group maxima and bookkeeping, so the chain no longer paces the tensor pipe;
//   * every train tile is read from L2 once per pair of SMs (each CTA loads its half, 8 KB descriptors + 2 KB K-extension,
//     two cp.async.bulk.tensor...cta_group::2 per tile completing on the LEADER's "full" barrier; two producer warps take
//     alternate tiles because one thread issues a TMA only every ~80 cycles); "stage empty" / "accumulator full" travel to
//     both CTAs with tcgen05.commit...multicast::cluster; the peer's epilogue warps release accumulators with a remote
//     mbarrier arrive (CTA-scope release: a .release.cluster arrive costs > 1,000 cycles) on the leader's barrier;
//   * the four threads of a query row (one per set) each see every fourth tile; they merge their top-3 tile maxima through
//     shared memory at the end of the unit and the set-0 thread writes the 16-byte candidate record of match_tc.cu.
#include "tc_ptx.cuh"

namespace sfm {

int launch_refine(const sfm_bank* b, const int32_t* pairs, int n_pairs, int32_t* knn_out, cudaStream_t st);

#ifndef SFM_TC2_STAGES
#define SFM_TC2_STAGES 6
#endif
#ifndef SFM_TC2_TRACE
#define SFM_TC2_TRACE 0          // diagnostics build: clock64 timeline of the first tiles of cluster 0 (sfm_debug_pair_trace)
#endif
constexpr int k2Stages = SFM_TC2_STAGES;                      // B ring depth (10 KB per stage and CTA)
constexpr int k2AccStages = 4;
constexpr int k2EpiWarps = 4 * k2AccStages;                   // one set of 4 warps (TMEM lane quadrants) per accumulator stage
constexpr int k2EpiThreads = 32 * k2EpiWarps;                 // 512
constexpr int k2WarpProd0 = k2EpiWarps, k2WarpProd1 = k2EpiWarps + 1, k2WarpIssue0 = k2EpiWarps + 2;
constexpr int k2Threads = 32 * (k2EpiWarps + 2 + k2AccStages);  // 704: epilogue, 2 TMA producers, 4 MMA issuers (leader CTA only)
constexpr int k2ABytes = kTileBytes;                          // one row block: 128 query rows
constexpr int k2HalfRows = kTileRows / 2;                     // 64 train rows per CTA and tile
constexpr int k2BDesc = k2HalfRows * kDescDim;                // 8192
constexpr int k2BExt = 2 * k2HalfRows * 16;                   // 2048: [2 K chunks][64 rows][16 B]
constexpr int k2BStage = k2BDesc + k2BExt;                    // 10240

struct Tc2Smem {
    static constexpr int kA = 0;
    static constexpr int kB = kA + 2 * k2ABytes;
    static constexpr int kAext = kB + ((k2Stages * k2BStage + 1023) / 1024) * 1024;
    static constexpr int kSub = kAext + kExtTileBytes;                     // [3 slots][4 quads][512 epilogue threads] int4
    static constexpr int kXchg = kSub + 3 * 4 * k2EpiThreads * 16;         // [512 epilogue threads][2] int4
    static constexpr int kBar = kXchg + k2EpiThreads * 32;
    static constexpr int kNumBar = 2 + 2 + 2 * k2Stages + 2 * k2AccStages;
    static constexpr int kTmemSlot = kBar + kNumBar * 8;
    static constexpr int kTotal = kTmemSlot + 16;
};
constexpr int k2SmemBytes = Tc2Smem::kTotal + 1024;
static_assert(k2SmemBytes <= 227 * 1024, "pair kernel exceeds the shared memory of an SM");
static_assert(k2Stages % 2 == 0, "each B stage must always be refilled by the same producer warp (tile parity == stage parity)");

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// TMA loads of a CTA pair: the bytes land in THIS CTA's shared memory, complete_tx goes to `bar`, which may live in the peer
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar_cluster)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(tmap), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint32_t bar_cluster)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(tmap), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster)
{
    // CTA-scope release (the default): the hazard this handshake orders is on TMEM (tcgen05.wait::ld + fence::before_thread_sync
    // before it), not on generic memory.  A .release.cluster arrive costs 1,000-1,400 cycles per call (round-2 trace, profiles/).
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void tc_commit_pair_mc(uint32_t bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_i8_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// kind::i8 instruction descriptor of the pair MMA: D = s32, M = 256 (128 rows per CTA), N = 128, both operands K-major
__host__ __device__ constexpr uint32_t idesc_i8_pair(uint32_t a_signed, uint32_t b_signed)
{
    return (2u << 4) | (a_signed << 7) | (b_signed << 10) | ((uint32_t)(kTileRows >> 3) << 17) | ((uint32_t)((2 * kTileRows) >> 4) << 24);
}
// K-major no-swizzle operand [k-chunk][64 rows][16 B]: LBO = 1024 B between the two K chunks, SBO = 128 B between 8-row groups
__device__ __forceinline__ uint32_t desc_lo_ext64(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | ((uint32_t)((k2HalfRows * 16) >> 4) << 16); }

#if SFM_TC2_TRACE
constexpr int k2TraceTiles = 96, k2TraceRoles = 8, k2TraceEvents = 6;
__device__ long long g_tc2_trace[k2TraceRoles * k2TraceTiles * k2TraceEvents];
#define TR2(role, tile, ev)                                                                                    \
    do {                                                                                                       \
        if (cluster_id == 0 && (tile) < k2TraceTiles) g_tc2_trace[((role) * k2TraceTiles + (tile)) * k2TraceEvents + (ev)] = clock64(); \
    } while (0)
#else
#define TR2(role, tile, ev) do { } while (0)
#endif

// two sub-group maxima (8 columns each) of a 16-column chunk
__device__ __forceinline__ void submax2(const uint32_t (&u)[16], int& c0, int& c1)
{
    int m = __vimax3_s32((int)u[0], (int)u[1], (int)u[2]);
    m = __vimax3_s32(m, (int)u[3], (int)u[4]);
    m = __vimax3_s32(m, (int)u[5], (int)u[6]);
    c0 = max(m, (int)u[7]);
    m = __vimax3_s32((int)u[8], (int)u[9], (int)u[10]);
    m = __vimax3_s32(m, (int)u[11], (int)u[12]);
    m = __vimax3_s32(m, (int)u[13], (int)u[14]);
    c1 = max(m, (int)u[15]);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1) match_tc2_kernel(
    const __grid_constant__ CUtensorMap tmap_desc, const __grid_constant__ CUtensorMap tmap_desc64,
    const __grid_constant__ CUtensorMap tmap_ext64, const int32_t* __restrict__ count, const int32_t* __restrict__ pairs, int n_pairs,
    int feat_stride, int32_t* __restrict__ knn_out, const int32_t* __restrict__ norm, const Prefilter pf, const int dbg_mode)
{
    extern __shared__ uint8_t smem_raw[];
    // both CTAs of the cluster must use the same CTA-relative offsets: the dynamic shared window starts at the same
    // offset in every CTA of a kernel, so the same round-up gives the same layout
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + Tc2Smem::kBar;
    auto bar_a_full = [&](int i) { return bar0 + 8 * (0 + i); };                            // leader only
    auto bar_a_empty = [&](int i) { return bar0 + 8 * (2 + i); };                           // both CTAs (multicast commit)
    auto bar_b_full = [&](int i) { return bar0 + 8 * (4 + i); };                            // leader only
    auto bar_b_empty = [&](int i) { return bar0 + 8 * (4 + k2Stages + i); };                // both CTAs
    auto bar_t_full = [&](int st) { return bar0 + 8 * (4 + 2 * k2Stages + st); };           // both CTAs
    auto bar_t_empty = [&](int st) { return bar0 + 8 * (4 + 2 * k2Stages + k2AccStages + st); };   // leader only
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Tc2Smem::kTmemSlot);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int units_per_pair = feat_stride / kUnitRows;
    const int total_units = n_pairs * units_per_pair;

    // ---- one-time setup
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(bar_a_full(i), 1); mbar_init(bar_a_empty(i), k2AccStages); }
        for (int i = 0; i < k2Stages; ++i) { mbar_init(bar_b_full(i), 1); mbar_init(bar_b_empty(i), 1); }
        for (int st = 0; st < k2AccStages; ++st) { mbar_init(bar_t_full(st), 1); mbar_init(bar_t_empty(st), 8); }   // 4 warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // constant A-side K-extension tile: weights 255 x 24, 1, 0 x 7 for every query row
    for (int e = threadIdx.x; e < kExtTileBytes / 4; e += k2Threads) {
        const int chunk = e / (kTileRows * 4), w = e & 3;
        uint32_t val = 0xFFFFFFFFu;
        if (chunk == 1) val = (w < 2) ? 0xFFFFFFFFu : (w == 2 ? 0x00000001u : 0u);
        reinterpret_cast<uint32_t*>(smem + Tc2Smem::kAext)[e] = val;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == k2WarpIssue0) {
        // both CTAs of the pair issue the allocation (same warp id, same destination offset)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // the peer's barriers are initialised before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == k2WarpProd0 || warp == k2WarpProd1) {
        // ================================================================= TMA producers (two per CTA: even / odd tiles; the even one
        // also loads the CTA's query rows).  Every load completes on the LEADER's "full" barrier, which expects both CTAs' bytes.
        if (lane == 0) {
            const int my_par = warp - k2WarpProd0;
            int ucount = 0, bit = 0;
            for (int u = cluster_id; u < total_units; u += n_clusters) {
                const UnitInfo I = decode_unit(u, units_per_pair, pairs, count);
                if (!I.live) continue;
                if (my_par == 0) {
                    const int abuf = ucount & 1, aph = (ucount >> 1) & 1;
                    const int qrow0 = I.img_q * feat_stride + I.qblk * kUnitRows + rank * kTileRows;
                    mbar_wait(bar_a_empty(abuf), aph ^ 1);
                    if (leader) mbar_expect_tx(bar_a_full(abuf), 2 * k2ABytes);
                    tma_load_2d_pair(sbase + Tc2Smem::kA + abuf * k2ABytes, &tmap_desc, 0, qrow0, mapa_u32(bar_a_full(abuf), 0));
                }
                const int trow0 = I.img_t * feat_stride;
                for (int t = 0; t < I.tiles; ++t, ++bit) {
                    if ((bit & 1) != my_par) continue;
                    const int s = bit % k2Stages, ph = (bit / k2Stages) & 1;
                    TR2(3 + 4 * rank, bit, 0);
                    mbar_wait(bar_b_empty(s), ph ^ 1);                 // the pair's MMAs have finished with stage s
                    TR2(3 + 4 * rank, bit, 1);
                    if (leader) mbar_expect_tx(bar_b_full(s), 2 * k2BStage);
                    const int row = trow0 + t * kTileRows;
                    const uint32_t dst = sbase + Tc2Smem::kB + s * k2BStage;
                    const uint32_t full = mapa_u32(bar_b_full(s), 0);
                    tma_load_2d_pair(dst, &tmap_desc64, 0, row + rank * k2HalfRows, full);
                    tma_load_3d_pair(dst + k2BDesc, &tmap_ext64, 0, rank * k2HalfRows, 2 * (row / kTileRows), full);
                    TR2(3 + 4 * rank, bit, 2);
                }
                ++ucount;
            }
            // tail: the leader's last commits arrive on THIS CTA's "empty" barriers asynchronously; do not leave before they landed
            for (int i = 0; i < k2Stages && i < bit; ++i) {
                const int idx = bit - 1 - i;
                if ((idx & 1) == my_par) mbar_wait(bar_b_empty(idx % k2Stages), (idx / k2Stages) & 1);
            }
            if (my_par == 0)
                for (int i = 0; i < 2 && i < ucount; ++i) {
                    const int idx = ucount - 1 - i;
                    mbar_wait(bar_a_empty(idx & 1), (idx >> 1) & 1);
                }
        }
    } else if (warp >= k2WarpIssue0) {
        // ================================================================= MMA issuers (leader CTA): one thread per accumulator stage
        if (lane == 0 && leader) {
            const int my_st = warp - k2WarpIssue0;
            constexpr uint32_t id_main = idesc_i8_pair(1, 1);
            constexpr uint32_t id_ext = idesc_i8_pair(0, 0);
            const uint64_t aext_desc = desc_ext(sbase + Tc2Smem::kAext);
            const uint32_t a_lo0 = desc_lo_sw128(sbase + Tc2Smem::kA);
            const uint32_t b_lo0 = desc_lo_sw128(sbase + Tc2Smem::kB);
            const uint32_t be_lo0 = desc_lo_ext64(sbase + Tc2Smem::kB + k2BDesc);
            int ucount = 0, bit = 0, tcount = 0;
            for (int u = cluster_id; u < total_units; u += n_clusters) {
                const UnitInfo I = decode_unit(u, units_per_pair, pairs, count);
                if (!I.live) continue;
                const int abuf = ucount & 1, aph = (ucount >> 1) & 1;
                mbar_wait(bar_a_full(abuf), aph);
                const uint32_t a_lo = a_lo0 + (uint32_t)(abuf * (k2ABytes >> 4));
                for (int t = 0; t < I.tiles; ++t, ++bit, ++tcount) {
                    const int st = tcount & (k2AccStages - 1), tph = (tcount >> 2) & 1;
                    if (st != my_st) continue;
                    const int s = bit % k2Stages, ph = (bit / k2Stages) & 1;
                    TR2(0, tcount, 0);
                    mbar_wait(bar_b_full(s), ph);
                    TR2(0, tcount, 1);
                    mbar_wait(bar_t_empty(st), tph ^ 1);
                    tc_fence_after();
                    TR2(0, tcount, 2);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(st * kTileRows);
                    const uint32_t b_lo = b_lo0 + (uint32_t)(s * (k2BStage >> 4));
                    if (dbg_mode != 5) {                               // diagnostics 5: K-extension MMA only (epilogue-bound rate)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma_i8_pair(d_tmem, mk_desc(kDescHiSw128, a_lo + 2 * k), mk_desc(kDescHiSw128, b_lo + 2 * k), id_main, k > 0);
                    }
                    tc_mma_i8_pair(d_tmem, aext_desc, mk_desc(kDescHiExt, be_lo0 + (uint32_t)(s * (k2BStage >> 4))), id_ext, dbg_mode != 5);
                    tc_commit_pair_mc(bar_t_full(st), (uint16_t)3);    // accumulator full: both CTAs' epilogues
                    tc_commit_pair_mc(bar_b_empty(s), (uint16_t)3);    // stage free: both CTAs' producers
                    TR2(0, tcount, 3);
                }
                tc_commit_pair_mc(bar_a_empty(abuf), (uint16_t)3);
                ++ucount;
            }
        }
        __syncwarp();
    } else {
        // ================================================================= epilogue: tile maxima -> candidate records.
        // warp = 4 * set + quadrant; set s owns accumulator stage s, i.e. the tiles with (tile counter & 3) == s
        const int wq = warp & 3, set = warp >> 2;
        const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(set * kTileRows);
        const int row_in_unit = rank * kTileRows + wq * 32 + lane;
        const int eth = threadIdx.x;                                    // 0..511
        int4* sub = reinterpret_cast<int4*>(smem + Tc2Smem::kSub);      // [(slot * 4 + quad) * 512 + eth]
        int4* xchg = reinterpret_cast<int4*>(smem + Tc2Smem::kXchg);    // [eth * 2 + {0, 1}]
        const int row_bar = 1 + wq;                                     // named barrier shared by the four warps of a quadrant
        const uint32_t my_full = bar_t_full(set);
        const uint32_t my_empty = mapa_u32(bar_t_empty(set), 0);        // the leader's "accumulator empty" barrier of this stage
        int tcount = 0, uses = 0;                                       // uses = tiles this set has drained (phase of its barriers)
        for (int u = cluster_id; u < total_units; u += n_clusters) {
            const UnitInfo I = decode_unit(u, units_per_pair, pairs, count);
            if (!I.live) continue;
            int M1 = kMaskedAcc + 3, M2 = kMaskedAcc + 2, M3 = kMaskedAcc + 1;
            int k1 = kInvalidTile | (0 << 16), k2 = kInvalidTile | (1 << 16), k3 = kInvalidTile | (2 << 16);
            bool tie4 = false;
            uint32_t v0[16], v1[16];
            auto mask_tail = [&](uint32_t (&v)[16], int col0, int valid) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (col0 + j >= valid) v[j] = (uint32_t)kMaskedAcc;
            };
            const int first = (set - tcount) & (k2AccStages - 1);       // first tile of this unit that lands in this set's stage
            for (int t = first; t < I.tiles; t += k2AccStages, ++uses) {
                const int tc = tcount + t;
                const int valid = I.nt - t * kTileRows;                 // >= 128 for full tiles
                const bool partial = valid < kTileRows;
                const bool tr = SFM_TC2_TRACE && lane == 0 && wq == 0;
                if (tr) TR2(1 + (set & 1) + 4 * rank, tc, 0);
                mbar_wait(my_full, uses & 1);
                tc_fence_after();
                if (tr) TR2(1 + (set & 1) + 4 * rank, tc, 1);
                int c[16];
                if (dbg_mode == 6) {                                   // diagnostics 6: hand the accumulator straight back (MMA-bound rate)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(my_empty);
                    continue;
                }
                // eight 16-column chunks, double buffered: chunk k + 1 is in flight while chunk k is reduced
                tc_ld16(lane_base, v0);
#pragma unroll
                for (int ch = 0; ch < 8; ch += 2) {
                    tc_wait_ld();
                    tc_ld16(lane_base + 16 * (ch + 1), v1);
                    if (partial) mask_tail(v0, 16 * ch, valid);
                    submax2(v0, c[2 * ch], c[2 * ch + 1]);
                    tc_wait_ld();
                    if (ch + 2 < 8) {
                        tc_ld16(lane_base + 16 * (ch + 2), v0);
                    } else {
                        // every TMEM read of this accumulator has landed: hand it back before the last chunk is reduced
                        if (tr) TR2(1 + (set & 1) + 4 * rank, tc, 2);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(my_empty);
                        if (tr) TR2(1 + (set & 1) + 4 * rank, tc, 3);
                    }
                    if (partial) mask_tail(v1, 16 * (ch + 1), valid);
                    submax2(v1, c[2 * ch + 2], c[2 * ch + 3]);
                }
                int m = __vimax3_s32(c[0], c[1], c[2]);
                m = __vimax3_s32(m, c[3], c[4]);
                m = __vimax3_s32(m, c[5], c[6]);
                m = __vimax3_s32(m, c[7], c[8]);
                m = __vimax3_s32(m, c[9], c[10]);
                m = __vimax3_s32(m, c[11], c[12]);
                m = __vimax3_s32(m, c[13], c[14]);
                m = max(m, c[15]);
                // top-3 update of (tile maximum, tile | slot << 16), select form (the branchy form of match_tc.cu runs all three paths
                // whenever any lane of the warp inserts, which is most tiles of a unit)
                const bool ins = m > M3, gt2 = m > M2, gt1 = m > M1;
                const int slot = k3 >> 16;                              // the evicted entry's parking slot is reused
                if (ins) {
                    int4* dst = sub + (slot * 4) * k2EpiThreads + eth;
                    dst[0] = make_int4(c[0], c[1], c[2], c[3]);
                    dst[k2EpiThreads] = make_int4(c[4], c[5], c[6], c[7]);
                    dst[2 * k2EpiThreads] = make_int4(c[8], c[9], c[10], c[11]);
                    dst[3 * k2EpiThreads] = make_int4(c[12], c[13], c[14], c[15]);
                }
                const int key = t | (slot << 16);
                const bool ntie = gt2 ? (M2 == M3) : (ins ? false : (tie4 || m == M3));
                const int nM3 = gt2 ? M2 : (ins ? m : M3), nk3 = gt2 ? k2 : (ins ? key : k3);
                const int nM2 = gt1 ? M1 : (gt2 ? m : M2), nk2 = gt1 ? k1 : (gt2 ? key : k2);
                M1 = gt1 ? m : M1; k1 = gt1 ? key : k1;
                M2 = nM2; k2 = nk2; M3 = nM3; k3 = nk3; tie4 = ntie;
                if (tr) TR2(1 + (set & 1) + 4 * rank, tc, 4);
            }
            tcount += I.tiles;
            // ---- the four threads of a row merge their top-3 tile maxima; the set-0 thread writes the record
            xchg[eth * 2 + 0] = make_int4(M1, M2, M3, tie4 ? 1 : 0);
            xchg[eth * 2 + 1] = make_int4(k1, k2, k3, 0);
            asm volatile("bar.sync %0, 128;" ::"r"(row_bar) : "memory");
            const int q = I.qblk * kUnitRows + row_in_unit;
            if (set == 0 && q < I.nq) {
                // top four of the twelve (value, key, owner) by insertion into a descending list; which of several equal values
                // comes first does not matter (every tile whose maximum reaches the second value is recorded or flagged)
                int Ms[4] = {kMaskedAcc, kMaskedAcc, kMaskedAcc, kMaskedAcc}, Ks[4] = {kInvalidTile, kInvalidTile, kInvalidTile, kInvalidTile};
                int Os[4] = {eth, eth, eth, eth};
                int tieM[k2AccStages], tieF[k2AccStages];
#pragma unroll
                for (int j = 0; j < k2AccStages; ++j) {
                    const int o = (eth & 127) + 128 * j;
                    const int4 pm = xchg[o * 2 + 0], pk = xchg[o * 2 + 1];
                    tieM[j] = pm.z;
                    tieF[j] = pm.w;
                    const int vv[3] = {pm.x, pm.y, pm.z}, kk[3] = {pk.x, pk.y, pk.z};
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        int v = vv[e], k = kk[e], ow = o;
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            if (v > Ms[p]) {
                                const int tv = Ms[p], tk = Ks[p], to = Os[p];
                                Ms[p] = v; Ks[p] = k; Os[p] = ow;
                                v = tv; k = tk; ow = to;
                            }
                        }
                    }
                }
                const int M2s = Ms[1];
                const bool valid3 = (Ks[2] & 0xFFFF) != kInvalidTile;
                const bool use3 = (Ms[2] == M2s) && valid3;
                // a tile outside the merged top three with maximum == the third value: the fourth of the twelve, or an untracked tile of
                // one of the four threads (its tie flag refers to its own third value)
                bool tie_more = (Ms[3] == Ms[2]);
#pragma unroll
                for (int j = 0; j < k2AccStages; ++j) tie_more |= (tieF[j] != 0 && tieM[j] == Ms[2]);
                int rec[3];
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    const int tile = Ks[e] & 0xFFFF, slot = Ks[e] >> 16;
                    int mask = 0;
                    if (tile != kInvalidTile && (e < 2 || use3)) {
                        const int4* src = sub + (slot * 4) * k2EpiThreads + Os[e];
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) {
                            const int4 w = src[qd * k2EpiThreads];
                            mask |= ((int)(w.x >= M2s) << (4 * qd)) | ((int)(w.y >= M2s) << (4 * qd + 1)) |
                                    ((int)(w.z >= M2s) << (4 * qd + 2)) | ((int)(w.w >= M2s) << (4 * qd + 3));
                        }
                    }
                    rec[e] = tile | (mask << 16);
                }
                int flags = (use3 && tie_more) ? 2 : 0;
                if (pf.mode != SFM_RATIO_NONE && (Ks[1] & 0xFFFF) != kInvalidTile) {
                    // prefilter on distance bounds, see match_tc.cu / DESIGN.md
                    const int cq = __ldg(norm + (long long)I.img_q * feat_stride + q) + 2 * kExtOffset;
                    if (!ratio_keep(max(cq - 2 * Ms[0], 0), cq - 2 * M2s + 1, pf.mode, pf.ratio, pf.num2, pf.den2)) flags |= 4;
                }
                *reinterpret_cast<int4*>(knn_out + ((long long)I.pair * feat_stride + q) * 4) =
                    (flags & 4) ? make_int4(-1, -1, -1, -1) : make_int4(rec[0], rec[1], rec[2], flags);      // prefiltered rows are final
            }
            asm volatile("bar.sync %0, 128;" ::"r"(row_bar) : "memory");  // parking slots and the exchange area are reused by the next unit
        }
    }

    // ---- teardown: nobody leaves while the peer may still arrive on this CTA's barriers or the pair's MMAs may still run
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == k2WarpIssue0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

int launch_match_tc2(const sfm_bank* b, const int32_t* pairs, int n_pairs, int grid_req, int32_t* knn_out, int sweep_only,
                     const Prefilter& pf, cudaStream_t st)
{
    if (!b->tmap_ready) {
        set_error("bank has no descriptor tensor map (metric must be L2)");
        return SFM_ERR_STATE;
    }
    static SmemAttrTable attr;                               // per device (one process may drive several GPUs)
    SFM_CUDA_CHECK(ensure_dyn_smem(match_tc2_kernel, k2SmemBytes, b->device, attr));
    const long long units = (long long)n_pairs * (b->L.feat_stride / kUnitRows);
    int grid = grid_req > 0 ? grid_req : b->sm_count;
    if (grid > 2 * units) grid = (int)(2 * units);
    grid &= ~1;                                                  // whole clusters
    if (grid < 2) grid = 2;
    match_tc2_kernel<<<grid, k2Threads, k2SmemBytes, st>>>(b->tmap_desc, b->tmap_desc64, b->tmap_ext64, b->count, pairs, n_pairs,
                                                           (int)b->L.feat_stride, knn_out, b->norm, pf, sweep_only > 4 ? sweep_only : 0);
    SFM_CUDA_CHECK(cudaGetLastError());
    count_launch();
    if (!sweep_only) return launch_refine(b, pairs, n_pairs, knn_out, st);
    return SFM_OK;
}

}  // namespace sfm

// Diagnostics (builds with -DSFM_TC2_TRACE=1 only): clock64 stamps [role][tile][event] of cluster 0's first tiles.
//   role 0 issuers (leader): top, B full, accumulator empty, issued + committed
//   role 1 / 2 leader epilogue warps of the even / odd sets: top, accumulator full, all chunks landed, released, bookkeeping done
//   role 3 leader producers: top, stage empty, loads issued;   roles 5, 6, 7: the same for the peer CTA (its own SM clock)
extern "C" int sfm_debug_pair_trace(long long* out, int n)
{
#if SFM_TC2_TRACE
    const int total = sfm::k2TraceRoles * sfm::k2TraceTiles * sfm::k2TraceEvents;
    SFM_REQUIRE(out && n >= total, "sfm_debug_pair_trace: buffer too small");
    SFM_CUDA_CHECK(cudaMemcpyFromSymbol(out, sfm::g_tc2_trace, sizeof(long long) * total));
    return total;
#else
    (void)out;
    (void)n;
    return 0;
#endif
}
