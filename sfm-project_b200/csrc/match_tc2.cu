// match_tc2.cu -- the tcgen05 matcher sweep on 2-CTA clusters (K2, variant "cluster").
//
// Same arithmetic, same candidate records and same refinement as match_tc.cu (it replaces the same reference call,
// bf.match at code/feature_matching.py:50 inside the pair loop code/pipeline.py:38-41); what changes is the mapping:
//   * a unit (pair, 256 query rows) is shared by the two CTAs of a cluster: CTA r owns row block r (128 rows);
//   * every train tile (16 KB descriptors + 4 KB K-extension) is fetched from L2 ONCE per cluster: CTA r issues the TMA
//     for descriptor rows [64 r, 64 r + 64) and K-extension chunk r with .multicast::cluster, so both halves land in both
//     CTAs' shared memory and complete_tx on both CTAs' "full" barriers;
//   * with one row block per CTA the 512 TMEM columns hold FOUR accumulator stages instead of 2 x 2: the MMA issuer runs
//     up to three tiles ahead of the epilogue, which removes the overlap loss of the 2-deep ring;
//   * a B stage is refilled only after BOTH CTAs' MMAs have read it: tcgen05.commit ... .multicast::cluster arrives on the
//     "empty" barrier of both CTAs;
//   * epilogue: 8 warps on 128 rows = two warp sets per TMEM lane quadrant, set h takes the tiles with (tile counter & 1)
//     == h (accumulator stages {h, h + 2}); the two threads of a row merge their top-3 tile maxima through shared memory at
//     the end of the unit and the set-0 thread writes the 16-byte candidate record in the format of match_tc.cu.
#include "tc_ptx.cuh"

namespace sfm {

int launch_refine(const sfm_bank* b, const int32_t* pairs, int n_pairs, int32_t* knn_out, cudaStream_t st);

#ifndef SFM_TC2_NO_MC
#define SFM_TC2_NO_MC 0          // experiment: 1 = every CTA fetches whole tiles itself (no multicast, no cross-CTA barrier)
#endif
constexpr int k2Stages = 6;
constexpr int k2AccStages = 4;
constexpr int k2Threads = 448;                                // warps 0-7 epilogue, 8 and 13 TMA producers (even / odd tiles), 9-12 MMA issuers
constexpr int k2ABytes = kTileBytes;                          // one row block
constexpr int k2HalfTile = kTileBytes / 2;                    // 8192: the 64 descriptor rows one CTA multicasts
constexpr int k2HalfExt = kExtTileBytes / 2;                  // 2048: one K chunk of the extension tile

struct Tc2Smem {
    static constexpr int kA = 0;
    static constexpr int kB = kA + 2 * k2ABytes;
    static constexpr int kAext = kB + k2Stages * kBStageBytes;
    static constexpr int kSub = kAext + kExtTileBytes;             // [3 slots][4 quads][256 epilogue threads] int4
    static constexpr int kXchg = kSub + 3 * 4 * 256 * 16;          // [256 epilogue threads][2] int4
    static constexpr int kBar = kXchg + 256 * 32;
    static constexpr int kNumBar = 2 + 2 + 2 * k2Stages + 2 * k2AccStages;
    static constexpr int kTmemSlot = kBar + kNumBar * 8;
    static constexpr int kTotal = kTmemSlot + 16;
};
constexpr int k2SmemBytes = Tc2Smem::kTotal + 1024;

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
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar, uint16_t mask)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
        "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void bulk_load_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1) match_tc2_kernel(
    const __grid_constant__ CUtensorMap tmap_desc, const __grid_constant__ CUtensorMap tmap_desc64, const int8_t* __restrict__ ext,
    const int32_t* __restrict__ count, const int32_t* __restrict__ pairs, int n_pairs, int feat_stride, int32_t* __restrict__ knn_out,
    const int32_t* __restrict__ norm, const Prefilter pf, const int dbg_mode)
{
    extern __shared__ uint8_t smem_raw[];
    // both CTAs of the cluster must use the same CTA-relative offsets: the dynamic shared window starts at the same
    // offset in every CTA of a kernel, so the same round-up gives the same layout
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + Tc2Smem::kBar;
    auto bar_a_full = [&](int i) { return bar0 + 8 * (0 + i); };
    auto bar_a_empty = [&](int i) { return bar0 + 8 * (2 + i); };
    auto bar_b_full = [&](int i) { return bar0 + 8 * (4 + i); };
    auto bar_b_empty = [&](int i) { return bar0 + 8 * (4 + k2Stages + i); };
    auto bar_t_full = [&](int st) { return bar0 + 8 * (4 + 2 * k2Stages + st); };
    auto bar_t_empty = [&](int st) { return bar0 + 8 * (4 + 2 * k2Stages + k2AccStages + st); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Tc2Smem::kTmemSlot);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int units_per_pair = feat_stride / kUnitRows;
    const int total_units = n_pairs * units_per_pair;

    // ---- one-time setup
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(bar_a_full(i), 1); mbar_init(bar_a_empty(i), k2AccStages); }
        for (int i = 0; i < k2Stages; ++i) { mbar_init(bar_b_full(i), 1); mbar_init(bar_b_empty(i), SFM_TC2_NO_MC ? 1 : 2); }     // one MMA commit per CTA
        for (int st = 0; st < k2AccStages; ++st) { mbar_init(bar_t_full(st), 1); mbar_init(bar_t_empty(st), 4); }
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
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // the peer's barriers are initialised before anything is multicast into its shared memory
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8 || warp == 13) {
        // ================================================================= TMA producers: warp 8 takes the even tiles (and the A
        // tiles), warp 13 the odd ones
        if (lane == 0) {
            const int my_par = (warp == 13) ? 1 : 0;
            int ucount = 0, bit = 0;
            for (int u = cluster_id; u < total_units; u += n_clusters) {
                const UnitInfo I = decode_unit(u, units_per_pair, pairs, count);
                if (!I.live) continue;
                const int abuf = ucount & 1, aph = (ucount >> 1) & 1;
                const int qrow0 = I.img_q * feat_stride + I.qblk * kUnitRows + rank * kTileRows;
                if (my_par == 0) {
                    mbar_wait(bar_a_empty(abuf), aph ^ 1);
                    mbar_expect_tx(bar_a_full(abuf), k2ABytes);
                    tma_load_2d(sbase + Tc2Smem::kA + abuf * k2ABytes, &tmap_desc, 0, qrow0, bar_a_full(abuf));
                }
                const int trow0 = I.img_t * feat_stride;
                for (int t = 0; t < I.tiles; ++t, ++bit) {
                    if ((bit & 1) != my_par) continue;
                    const int s = bit % k2Stages, ph = (bit / k2Stages) & 1;
                    mbar_wait(bar_b_empty(s), ph ^ 1);                 // BOTH CTAs' MMAs have finished with stage s
                    mbar_expect_tx(bar_b_full(s), kBStageBytes);       // my half + the peer's half
                    const int row = trow0 + t * kTileRows;
                    const uint32_t dst = sbase + Tc2Smem::kB + s * kBStageBytes;
#if SFM_TC2_NO_MC
                    tma_load_2d(dst, &tmap_desc, 0, row, bar_b_full(s));
                    bulk_load(dst + kTileBytes, ext + (long long)(row / kTileRows) * kExtTileBytes, kExtTileBytes, bar_b_full(s));
#else
                    tma_load_2d_mc(dst + rank * k2HalfTile, &tmap_desc64, 0, row + rank * (kTileRows / 2), bar_b_full(s), (uint16_t)3);
                    bulk_load_mc(dst + kTileBytes + rank * k2HalfExt,
                                 ext + (long long)(row / kTileRows) * kExtTileBytes + rank * k2HalfExt, k2HalfExt, bar_b_full(s), (uint16_t)3);
#endif
                }
                ++ucount;
            }
            // tail: the peer's last commits arrive on MY barriers asynchronously; do not leave before they have landed
            for (int i = 0; i < k2Stages && i < bit; ++i) {
                const int idx = bit - 1 - i;
                if ((idx & 1) == my_par) mbar_wait(bar_b_empty(idx % k2Stages), (idx / k2Stages) & 1);
            }
        }
    } else if (warp >= 9 && warp <= 12) {
        // ================================================================= MMA issuers: one thread per accumulator stage
        if (lane == 0) {
            const int my_st = warp - 9;
            constexpr uint32_t id_main = idesc_i8(1, 1);
            constexpr uint32_t id_ext = idesc_i8(0, 0);
            const uint64_t aext_desc = desc_ext(sbase + Tc2Smem::kAext);
            const uint32_t a_lo0 = desc_lo_sw128(sbase + Tc2Smem::kA);
            const uint32_t b_lo0 = desc_lo_sw128(sbase + Tc2Smem::kB);
            const uint32_t be_lo0 = desc_lo_ext(sbase + Tc2Smem::kB + kTileBytes);
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
                    mbar_wait(bar_b_full(s), ph);
                    mbar_wait(bar_t_empty(st), tph ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(st * kTileRows);
                    const uint32_t b_lo = b_lo0 + (uint32_t)(s * (kBStageBytes >> 4));
                    if (dbg_mode != 5) {                               // diagnostics 5: K-extension MMA only (epilogue-bound rate)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma_i8(d_tmem, mk_desc(kDescHiSw128, a_lo + 2 * k), mk_desc(kDescHiSw128, b_lo + 2 * k), id_main, k > 0);
                    }
                    tc_mma_i8(d_tmem, aext_desc, mk_desc(kDescHiExt, be_lo0 + (uint32_t)(s * (kBStageBytes >> 4))), id_ext, dbg_mode != 5);
                    tc_commit(bar_t_full(st));
#if SFM_TC2_NO_MC
                    tc_commit(bar_b_empty(s));
#else
                    tc_commit_mc(bar_b_empty(s), (uint16_t)3);         // frees the stage in both CTAs' producers
#endif
                }
                tc_commit(bar_a_empty(abuf));
                ++ucount;
            }
        }
        __syncwarp();
    } else if (warp < 8) {
        // ================================================================= epilogue: tile maxima -> candidate records
        const int wq = warp & 3, set = warp >> 2;
        const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
        const int row_in_unit = rank * kTileRows + wq * 32 + lane;
        const int eth = threadIdx.x;                                    // 0..255
        int4* sub = reinterpret_cast<int4*>(smem + Tc2Smem::kSub);      // [(slot * 4 + quad) * 256 + eth]
        int4* xchg = reinterpret_cast<int4*>(smem + Tc2Smem::kXchg);    // [eth * 2 + {0, 1}]
        const int pair_bar = 1 + wq;                                    // named barrier shared by warps wq and wq + 4
        int tcount = 0;
        for (int u = cluster_id; u < total_units; u += n_clusters) {
            const UnitInfo I = decode_unit(u, units_per_pair, pairs, count);
            if (!I.live) continue;
            int M1 = kMaskedAcc + 3, M2 = kMaskedAcc + 2, M3 = kMaskedAcc + 1;
            int k1 = kInvalidTile | (0 << 16), k2 = kInvalidTile | (1 << 16), k3 = kInvalidTile | (2 << 16);
            bool tie4 = false;
            uint32_t va[32], vb[32], vc[32];
            auto acc_addr = [&](int tc) { return lane_base + (uint32_t)((tc & (k2AccStages - 1)) * kTileRows); };
            auto wait_full = [&](int tc) {
                mbar_wait(bar_t_full(tc & (k2AccStages - 1)), (tc >> 2) & 1);
                tc_fence_after();
            };
            auto mask_tail = [&](uint32_t (&v)[32], int col0, int valid) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col0 + j >= valid) v[j] = (uint32_t)kMaskedAcc;
            };
            const int first = (set ^ tcount) & 1;                       // this set takes the tiles with (tile counter & 1) == set
            if (first < I.tiles) {
                wait_full(tcount + first);
                tc_ld32(acc_addr(tcount + first), va);
            }
            for (int t = first; t < I.tiles; t += 2) {
                const int tc = tcount + t;
                const uint32_t taddr = acc_addr(tc);
                const int valid = I.nt - t * kTileRows;                 // >= 128 for full tiles
                const bool partial = valid < kTileRows;
                int c[16];
                if (dbg_mode == 6) {                                   // diagnostics 6: hand the accumulator straight back (MMA-bound rate)
                    tc_wait_ld();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_t_empty(tc & (k2AccStages - 1)));
                    if (t + 2 < I.tiles) { wait_full(tc + 2); tc_ld32(acc_addr(tc + 2), va); }
                    continue;
                }
                // three TMEM round trips per tile: [c0 prefetched] -> {c1,c2} -> c3 -> (release, prefetch the set's next c0)
                tc_wait_ld();
                tc_ld32(taddr + 32, vb);
                tc_ld32(taddr + 64, vc);
                if (partial) mask_tail(va, 0, valid);
                submax4(va, c, 0);
                tc_wait_ld();
                tc_ld32(taddr + 96, va);
                if (partial) { mask_tail(vb, 32, valid); mask_tail(vc, 64, valid); }
                submax4(vb, c, 4);
                submax4(vc, c, 8);
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_t_empty(tc & (k2AccStages - 1)));
                if (partial) mask_tail(va, 96, valid);
                submax4(va, c, 12);
                if (t + 2 < I.tiles) {
                    wait_full(tc + 2);
                    tc_ld32(acc_addr(tc + 2), va);
                }
                int m = __vimax3_s32(c[0], c[1], c[2]);
                m = __vimax3_s32(m, c[3], c[4]);
                m = __vimax3_s32(m, c[5], c[6]);
                m = __vimax3_s32(m, c[7], c[8]);
                m = __vimax3_s32(m, c[9], c[10]);
                m = __vimax3_s32(m, c[11], c[12]);
                m = __vimax3_s32(m, c[13], c[14]);
                m = max(m, c[15]);
                if (m >= M3) {
                    if (m == M3) {
                        tie4 = true;
                    } else {
                        const int slot = k3 >> 16;                       // the evicted entry's slot is reused
                        int4* dst = sub + (slot * 4) * 256 + eth;
                        dst[0] = make_int4(c[0], c[1], c[2], c[3]);
                        dst[256] = make_int4(c[4], c[5], c[6], c[7]);
                        dst[512] = make_int4(c[8], c[9], c[10], c[11]);
                        dst[768] = make_int4(c[12], c[13], c[14], c[15]);
                        const int key = t | (slot << 16);
                        if (m > M2) {
                            tie4 = (M2 == M3);
                            M3 = M2; k3 = k2;
                            if (m > M1) { M2 = M1; k2 = k1; M1 = m; k1 = key; }
                            else { M2 = m; k2 = key; }
                        } else {
                            tie4 = false;
                            M3 = m; k3 = key;
                        }
                    }
                }
            }
            tcount += I.tiles;
            // ---- the two threads of a row (warps wq and wq + 4) merge their top-3 tile maxima; set 0 writes the record
            xchg[eth * 2 + 0] = make_int4(M1, M2, M3, tie4 ? 1 : 0);
            xchg[eth * 2 + 1] = make_int4(k1, k2, k3, 0);
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            const int q = I.qblk * kUnitRows + row_in_unit;
            if (set == 0 && q < I.nq) {
                const int peth = eth + 128;
                const int4 pm = xchg[peth * 2 + 0], pk = xchg[peth * 2 + 1];
                const int Ma[3] = {M1, M2, M3}, Ka[3] = {k1, k2, k3};
                const int Mb[3] = {pm.x, pm.y, pm.z}, Kb[3] = {pk.x, pk.y, pk.z};
                // merge of two descending triples: top-3 of the six (value, key, owner) plus the value of the fourth
                int ia = 0, ib = 0;
                int Ms[4], Ks[3], Os[3];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int va_ = (ia < 3) ? Ma[ia] : kMaskedAcc, vb_ = (ib < 3) ? Mb[ib] : kMaskedAcc;
                    const bool take_a = va_ >= vb_;
                    Ms[e] = take_a ? va_ : vb_;
                    if (e < 3) {
                        Ks[e] = take_a ? Ka[ia < 3 ? ia : 2] : Kb[ib < 3 ? ib : 2];
                        Os[e] = take_a ? eth : peth;
                    }
                    if (take_a) ++ia; else ++ib;
                }
                const int M2s = Ms[1];
                const bool valid3 = (Ks[2] & 0xFFFF) != kInvalidTile;
                const bool use3 = (Ms[2] == M2s) && valid3;
                // a tile outside the merged top three with maximum == M3*: the fourth of the six, or an untracked tile of either thread
                const bool tie_more = (Ms[3] == Ms[2]) || (tie4 && M3 == Ms[2]) || (pm.w != 0 && pm.z == Ms[2]);
                int rec[3];
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    const int tile = Ks[e] & 0xFFFF, slot = Ks[e] >> 16;
                    int mask = 0;
                    if (tile != kInvalidTile && (e < 2 || use3)) {
                        const int4* src = sub + (slot * 4) * 256 + Os[e];
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) {
                            const int4 w = src[qd * 256];
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
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");   // parking slots and the exchange area are reused by the next unit
        }
    }

    // ---- teardown: nobody leaves while the peer may still multicast into this CTA or arrive on its barriers
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

int launch_match_tc2(const sfm_bank* b, const int32_t* pairs, int n_pairs, int grid_req, int32_t* knn_out, int sweep_only,
                     const Prefilter& pf, cudaStream_t st)
{
    if (!b->tmap_ready) {
        set_error("bank has no descriptor tensor map (metric must be L2)");
        return SFM_ERR_STATE;
    }
    static bool attr_set[64] = {};                           // per device (one process may drive several GPUs)
    if (!attr_set[b->device & 63]) {
        SFM_CUDA_CHECK(cudaFuncSetAttribute(match_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k2SmemBytes));
        attr_set[b->device & 63] = true;
    }
    const long long units = (long long)n_pairs * (b->L.feat_stride / kUnitRows);
    int grid = grid_req > 0 ? grid_req : b->sm_count;
    if (grid > 2 * units) grid = (int)(2 * units);
    grid &= ~1;                                                  // whole clusters
    if (grid < 2) grid = 2;
    match_tc2_kernel<<<grid, k2Threads, k2SmemBytes, st>>>(b->tmap_desc, b->tmap_desc64, b->ext, b->count, pairs, n_pairs,
                                                           (int)b->L.feat_stride, knn_out, b->norm, pf, sweep_only > 1 ? sweep_only : 0);
    SFM_CUDA_CHECK(cudaGetLastError());
    count_launch();
    if (!sweep_only) return launch_refine(b, pairs, n_pairs, knn_out, st);
    return SFM_OK;
}

}  // namespace sfm
