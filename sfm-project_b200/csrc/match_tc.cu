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
// Epilogue   : one thread per query row keeps the three largest 32-column chunk maxima (value, chunk, mask of
//              8-column sub-groups that can still matter) and a tie flag -- ~0.5 VIMNMX3 per distance, no
//              per-element index work; TMEM loads are software-pipelined one chunk ahead.
// Refinement : every element that can be one of the exact top-2 (distance, index) lies inside the recorded
//              sub-groups unless the flag is set (proof in DESIGN.md).  The sweep writes a 16-byte candidate
//              record per query row into knn_out; refine_kernel (one thread per row, full occupancy)
//              recomputes those ~16 candidates exactly with dp4a and overwrites the record with the
//              result; flagged rows are brute-forced by a whole warp.  Results are bit-exact.
//
// Warp roles of the sweep (384 threads, 1 CTA / SM, persistent over units):
//   warps 0-3   epilogue row block 0 warps 4-7   epilogue row block 1   (TMEM lane quadrant = warp % 4)
//   warp 8      TMA producer
//   warps 9-12  MMA issuers (warp 9 owns the TMEM allocation): one thread per accumulator stage issues row block 0's chain, then row
//               block 1's (SFM_TC_SEQ, warps 9-10); the original form has one thread per accumulator (stage, row block).
//               tcgen05.mma issue is execution-paced (~65 cycles each, queue depth 1-2) and every mbarrier
//               wait costs the issuing thread ~150-250 cycles even when already satisfied, so one thread's
//               wait -> wait -> 5 x issue -> commit chain (~970 cycles) only fits the 1280-cycle budget of
//               "its" accumulator when four threads take turns (tools/ubench/mbar_mma.cu, clock64 trace)
#include "tc_ptx.cuh"

namespace sfm {

#ifndef SFM_TC_STAGES
#define SFM_TC_STAGES 3      // B-ring depth: 3 measured ~0.8 % faster than 4 / 5 (tools/ab_sweep.sh), 2 is 3 % slower
#endif
#ifndef SFM_TC_SEQ
#define SFM_TC_SEQ 1         // 1: one issuer thread per stage, the two row blocks' MMA chains one after the other (round-2 A/B on one
                             //    box, three repetitions each: 1.420 / 1.420 / 1.419 ms per 224 pairs against 1.428 / 1.430 / 1.429)
#endif
#ifndef SFM_TC_THREADS
#define SFM_TC_THREADS (SFM_TC_SEQ ? 352 : 416)   // 8 epilogue warps + TMA producer + 2 (SEQ) or 4 issuer warps
#endif
constexpr int kStages = SFM_TC_STAGES;
constexpr int kABufBytes = 2 * kTileBytes;                   // 32768
constexpr int kTcThreads = SFM_TC_THREADS;

struct TcSmem {
    static constexpr int kA = 0;
    static constexpr int kB = kA + 2 * kABufBytes;
    static constexpr int kAext = kB + kStages * kBStageBytes;
    static constexpr int kSub = kAext + kExtTileBytes;             // [3 slots][4 quads][256 epilogue threads] int4
    static constexpr int kBar = kSub + 3 * 4 * 256 * 16;
    static constexpr int kNumBar = 2 + 2 + 2 * kStages + 4 + 4;
    static constexpr int kTmemSlot = kBar + kNumBar * 8;
    static constexpr int kTotal = kTmemSlot + 16;
};
constexpr int kTcSmemBytes = TcSmem::kTotal + 1024;          // + alignment slack

template <bool kDbg>
__global__ void __launch_bounds__(kTcThreads, 1) match_tc_kernel(
    const __grid_constant__ CUtensorMap tmap_desc, const int8_t* __restrict__ ext, const int32_t* __restrict__ count,
    const int32_t* __restrict__ pairs, int n_pairs, int feat_stride, int32_t* __restrict__ knn_out,
    int32_t* __restrict__ dbg_acc, int dbg_mode, const int32_t* __restrict__ norm, const Prefilter pf)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + TcSmem::kBar;
    auto bar_a_full = [&](int i) { return bar0 + 8 * (0 + i); };
    auto bar_a_empty = [&](int i) { return bar0 + 8 * (2 + i); };
    auto bar_b_full = [&](int i) { return bar0 + 8 * (4 + i); };
    auto bar_b_empty = [&](int i) { return bar0 + 8 * (4 + kStages + i); };
    auto bar_t_full = [&](int st, int rb) { return bar0 + 8 * (4 + 2 * kStages + st * 2 + rb); };
    auto bar_t_empty = [&](int st, int rb) { return bar0 + 8 * (8 + 2 * kStages + st * 2 + rb); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + TcSmem::kTmemSlot);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units_per_pair = feat_stride / kUnitRows;
    const int total_units = n_pairs * units_per_pair;

    // ---- one-time setup
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(bar_a_full(i), 1); mbar_init(bar_a_empty(i), 4); }
        for (int i = 0; i < kStages; ++i) { mbar_init(bar_b_full(i), 1); mbar_init(bar_b_empty(i), 2); }
        for (int st = 0; st < 2; ++st)
            for (int rb = 0; rb < 2; ++rb) { mbar_init(bar_t_full(st, rb), 1); mbar_init(bar_t_empty(st, rb), 4); }
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
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
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
                    SFM_TRACE(3, bit, 0);
                    mbar_wait(bar_b_empty(s), ph ^ 1);
                    SFM_TRACE(3, bit, 1);
                    mbar_expect_tx(bar_b_full(s), kBStageBytes);
                    const int row = trow0 + t * kTileRows;
                    const uint32_t dst = sbase + TcSmem::kB + s * kBStageBytes;
                    tma_load_2d(dst, &tmap_desc, 0, row, bar_b_full(s));
                    bulk_load(dst + kTileBytes, ext + (long long)(row / kTileRows) * kExtTileBytes, kExtTileBytes, bar_b_full(s));
                }
                ++ucount;
            }
        }
    } else if (warp >= 9) {
        // ================================================================= MMA issuers
#if SFM_TC_SEQ
        // One thread per accumulator STAGE issues row block 0's five MMAs, commits, then row block 1's: the sweep's period is half
        // the round trip of an accumulator, and row block 0's round trip then no longer contains row block 1's MMAs
        // (with one thread per accumulator the two chains interleave in the tensor pipe and both complete late).
        if (lane == 0 && warp < 11) {
            const int my_st = warp - 9;
            constexpr uint32_t id_main = idesc_i8(1, 1);
            constexpr uint32_t id_ext = idesc_i8(0, 0);
            const uint64_t aext_desc = desc_ext(sbase + TcSmem::kAext);
            const uint32_t a_lo0 = desc_lo_sw128(sbase + TcSmem::kA);
            const uint32_t b_lo0 = desc_lo_sw128(sbase + TcSmem::kB);
            const uint32_t be_lo0 = desc_lo_ext(sbase + TcSmem::kB + kTileBytes);
            int ucount = 0, bit = 0, tcount = 0;
            for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
                const UnitInfo I = decode_unit(u, units_per_pair, pairs, count);
                if (!I.live) continue;
                const int abuf = ucount & 1, aph = (ucount >> 1) & 1;
                mbar_wait(bar_a_full(abuf), aph);
                const uint32_t a_lo = a_lo0 + (uint32_t)(abuf * (kABufBytes >> 4));
                for (int t = 0; t < I.tiles; ++t, ++bit, ++tcount) {
                    const int st = tcount & 1, tph = (tcount >> 1) & 1;
                    if (st != my_st) continue;                       // the other stage's issuer takes this tile
                    const int s = bit % kStages, ph = (bit / kStages) & 1;
                    SFM_TRACE(0, tcount, 0);
                    mbar_wait(bar_b_full(s), ph);
                    SFM_TRACE(0, tcount, 1);
                    const uint32_t b_lo = b_lo0 + (uint32_t)(s * (kBStageBytes >> 4));
                    const uint64_t be = mk_desc(kDescHiExt, be_lo0 + (uint32_t)(s * (kBStageBytes >> 4)));
#pragma unroll
                    for (int rb = 0; rb < 2; ++rb) {
                        mbar_wait(bar_t_empty(st, rb), tph ^ 1);
                        tc_fence_after();
                        SFM_TRACE(0, tcount, 2 + 4 * rb);
                        const uint32_t d_tmem = tmem_base + (uint32_t)((st * 2 + rb) * kTileRows);
                        const uint32_t a_rb = a_lo + (uint32_t)(rb * (kTileBytes >> 4));
                        const bool with_main = dbg_mode != 5 && (!kDbg || dbg_mode != 2);     // diagnostics: 5 / 2 = K-extension only
                        if (with_main) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                tc_mma_i8(d_tmem, mk_desc(kDescHiSw128, a_rb + 2 * k), mk_desc(kDescHiSw128, b_lo + 2 * k), id_main, k > 0);
                        }
                        if (!kDbg || dbg_mode != 1) tc_mma_i8(d_tmem, aext_desc, be, id_ext, with_main);   // (1 = descriptor MMAs only)
                        tc_commit(bar_t_full(st, rb));
                        tc_commit(bar_b_empty(s));                   // (the barrier counts one arrival per row block)
                        SFM_TRACE(0, tcount, 3 + 4 * rb);
                    }
                }
                tc_commit(bar_a_empty(abuf));
                tc_commit(bar_a_empty(abuf));                        // (four arrivals per unit: two per issuer)
                ++ucount;
            }
        }
        __syncwarp();
#else
        // one thread per accumulator (stage, row block)
        if (lane == 0) {
            const int rb = (warp - 9) & 1, my_st = (warp - 9) >> 1;
            constexpr uint32_t id_main = idesc_i8(1, 1);
            constexpr uint32_t id_ext = idesc_i8(0, 0);
            const uint64_t aext_desc = desc_ext(sbase + TcSmem::kAext);
            const uint32_t a_lo0 = desc_lo_sw128(sbase + TcSmem::kA + rb * kTileBytes);
            const uint32_t b_lo0 = desc_lo_sw128(sbase + TcSmem::kB);
            const uint32_t be_lo0 = desc_lo_ext(sbase + TcSmem::kB + kTileBytes);
            int ucount = 0, bit = 0, tcount = 0;
            for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
                const UnitInfo I = decode_unit(u, units_per_pair, pairs, count);
                if (!I.live) continue;
                const int abuf = ucount & 1, aph = (ucount >> 1) & 1;
                mbar_wait(bar_a_full(abuf), aph);
                const uint32_t a_lo = a_lo0 + (uint32_t)(abuf * (kABufBytes >> 4));
                for (int t = 0; t < I.tiles; ++t, ++bit, ++tcount) {
                    const int st = tcount & 1, tph = (tcount >> 1) & 1;
                    if (st != my_st) continue;                       // the other stage's issuers take this tile
                    const int s = bit % kStages, ph = (bit / kStages) & 1;
                    SFM_TRACE(0, tcount, 0 + 4 * rb);
                    if (SFM_SPIN_MASK & 4) mbar_wait_spin(bar_b_full(s), ph); else mbar_wait(bar_b_full(s), ph);
                    SFM_TRACE(0, tcount, 1 + 4 * rb);
                    if (SFM_SPIN_MASK & 1) mbar_wait_spin(bar_t_empty(st, rb), tph ^ 1); else mbar_wait(bar_t_empty(st, rb), tph ^ 1);
                    tc_fence_after();
                    SFM_TRACE(0, tcount, 2 + 4 * rb);
                    const uint32_t d_tmem = tmem_base + (uint32_t)((st * 2 + rb) * kTileRows);
                    const uint32_t b_lo = b_lo0 + (uint32_t)(s * (kBStageBytes >> 4));
                    if (dbg_mode != 5 && (!kDbg || dbg_mode != 2)) {   // (mode 4 = trace: full MMA; 5 = K-extension only, production epilogue)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma_i8(d_tmem, mk_desc(kDescHiSw128, a_lo + 2 * k), mk_desc(kDescHiSw128, b_lo + 2 * k), id_main, k > 0);
                    }
                    if (!kDbg || dbg_mode != 1)
                        tc_mma_i8(d_tmem, aext_desc, mk_desc(kDescHiExt, be_lo0 + (uint32_t)(s * (kBStageBytes >> 4))), id_ext,
                                  dbg_mode != 5 && (!kDbg || dbg_mode != 2));
                    tc_commit(bar_t_full(st, rb));
                    tc_commit(bar_b_empty(s));
                    SFM_TRACE(0, tcount, 3 + 4 * rb);
                }
                tc_commit(bar_a_empty(abuf));
                ++ucount;
            }
        }
        __syncwarp();
#endif
    } else if (warp < 8) {
        // ================================================================= epilogue: tile maxima -> candidate records
        const int rb = warp >> 2, wq = warp & 3;
        const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
        const int row_in_unit = rb * kTileRows + wq * 32 + lane;
        const int eth = threadIdx.x;                                    // 0..255
        int4* sub = reinterpret_cast<int4*>(smem + TcSmem::kSub);       // [(slot * 4 + quad) * 256 + eth]
        int tcount = 0;
        bool first_unit = true;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
            const UnitInfo I = decode_unit(u, units_per_pair, pairs, count);
            if (!I.live) continue;
            int M1 = kMaskedAcc + 3, M2 = kMaskedAcc + 2, M3 = kMaskedAcc + 1;
            int k1 = kInvalidTile | (0 << 16), k2 = kInvalidTile | (1 << 16), k3 = kInvalidTile | (2 << 16);
            bool tie4 = false;
            uint32_t va[32], vb[32], vc[32];
            auto acc_addr = [&](int tc) { return lane_base + (uint32_t)((((tc & 1) * 2) + rb) * kTileRows); };
            auto wait_full = [&](int tc) {
                if (SFM_SPIN_MASK & 2) mbar_wait_spin(bar_t_full(tc & 1, rb), (tc >> 1) & 1); else mbar_wait(bar_t_full(tc & 1, rb), (tc >> 1) & 1);
                tc_fence_after();
            };
            auto mask_tail = [&](uint32_t (&v)[32], int col0, int valid) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col0 + j >= valid) v[j] = (uint32_t)kMaskedAcc;
            };
            auto dump = [&](const uint32_t (&v)[32], int col0) {
                int32_t* o = dbg_acc + (long long)row_in_unit * kTileRows + col0;
#pragma unroll
                for (int j = 0; j < 32; ++j) o[j] = (int)v[j];
            };
            wait_full(tcount);
            tc_ld32(acc_addr(tcount), va);
            for (int t = 0; t < I.tiles; ++t) {
                const int tc = tcount + t;
                const uint32_t taddr = acc_addr(tc);
                const int valid = I.nt - t * kTileRows;                 // >= 128 for full tiles
                const bool partial = valid < kTileRows;
                const bool dbg_now = kDbg && dbg_acc != nullptr && dbg_mode < 4 && first_unit && t == 0 && blockIdx.x == 0;
                int c[16];
                const bool tr = kDbg && lane == 0 && wq == 2;
                // three TMEM round trips per tile: [c0 prefetched] -> {c1,c2} -> c3 -> (release, prefetch next c0)
                if (kDbg && dbg_mode >= 6) {
                    // diagnostics on the full grid: 6 = hand the accumulator straight back (MMA + handshake bound),
                    //                               7 = read all of it from TMEM but do no arithmetic
                    tc_wait_ld();
                    if (dbg_mode == 7) { tc_ld32(taddr + 32, vb); tc_ld32(taddr + 64, vc); tc_wait_ld(); tc_ld32(taddr + 96, va); tc_wait_ld(); }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_t_empty(tc & 1, rb));
                    if (t + 1 < I.tiles) {
                        wait_full(tc + 1);
                        tc_ld32(acc_addr(tc + 1), va);
                    }
                    continue;
                }
                if (tr) SFM_TRACE(1 + rb, tc, 0);
                tc_wait_ld();
                if (tr) SFM_TRACE(1 + rb, tc, 1);
                tc_ld32(taddr + 32, vb);
                tc_ld32(taddr + 64, vc);
                if (dbg_now) dump(va, 0);
                if (partial) mask_tail(va, 0, valid);
                submax4(va, c, 0);
                tc_wait_ld();
                tc_ld32(taddr + 96, va);
                if (dbg_now) { dump(vb, 32); dump(vc, 64); }
                if (partial) { mask_tail(vb, 32, valid); mask_tail(vc, 64, valid); }
                submax4(vb, c, 4);
                submax4(vc, c, 8);
                if (tr) SFM_TRACE(1 + rb, tc, 2);
                tc_wait_ld();
                if (tr) SFM_TRACE(1 + rb, tc, 3);
                // every TMEM read of this accumulator has landed: hand it back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_t_empty(tc & 1, rb));
                if (tr) SFM_TRACE(1 + rb, tc, 4);
                if (dbg_now) dump(va, 96);
                if (partial) mask_tail(va, 96, valid);
                submax4(va, c, 12);
                if (t + 1 < I.tiles) {                                   // prefetch the next tile's first chunk
                    wait_full(tc + 1);
                    tc_ld32(acc_addr(tc + 1), va);
                }
                if (tr) SFM_TRACE(1 + rb, tc, 5);
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
                        // explicit shared-space stores: through the generic `sub` pointer these compiled to ST.E (generic) instructions
                        // (0.2 % of the sweep, two same-box repetitions: 1.4565 against 1.4595 ms per 224 pairs)
                        const uint32_t dsts = sbase + TcSmem::kSub + (uint32_t)(((slot * 4) * 256 + eth) * 16);
                        sts128(dsts, c[0], c[1], c[2], c[3]);
                        sts128(dsts + 256 * 16, c[4], c[5], c[6], c[7]);
                        sts128(dsts + 512 * 16, c[8], c[9], c[10], c[11]);
                        sts128(dsts + 768 * 16, c[12], c[13], c[14], c[15]);
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
                if (tr) SFM_TRACE(1 + rb, tc, 6);
            }
            tcount += I.tiles;
            first_unit = false;
            // candidate record: for each kept tile the sub-groups whose maximum reaches M2
            const int q = I.qblk * kUnitRows + row_in_unit;
            if (q < I.nq) {
                const bool use3 = (M3 == M2) && (k3 & 0xFFFF) != kInvalidTile;
                int keys[3] = {k1, k2, k3};
                int rec[3];
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    const int tile = keys[e] & 0xFFFF, slot = keys[e] >> 16;
                    int mask = 0;
                    if (tile != kInvalidTile && (e < 2 || use3)) {
                        const int4* src = sub + (slot * 4) * 256 + eth;
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) {
                            const int4 w = src[qd * 256];
                            mask |= ((int)(w.x >= M2) << (4 * qd)) | ((int)(w.y >= M2) << (4 * qd + 1)) |
                                    ((int)(w.z >= M2) << (4 * qd + 2)) | ((int)(w.w >= M2) << (4 * qd + 3));
                        }
                    }
                    rec[e] = tile | (mask << 16);
                }
                int flags = (use3 && tie4) ? 2 : 0;
                // fourth word of the record: bit 0 = a second tile exists, bit 1 = brute-force the row, bits 3.. = M2 + kRecBias
                // (the second-largest tile maximum: everything outside the first tile has a distance >= C - 2 M2, which lets the
                //  fused refinement stop after the first tile for most rows)
                if ((k2 & 0xFFFF) != kInvalidTile) flags |= 1 | ((M2 + kRecBias) << 3);
                if (pf.mode != SFM_RATIO_NONE && (k2 & 0xFFFF) != kInvalidTile) {
                    // D = C - 2 acc + (|b|^2 & 1): the nearest distance is >= max(C - 2 M1, 0), the second nearest is
                    // <= C - 2 M2 + 1 (two different tiles hold elements that good).  A row whose bounds fail the
                    // (monotone) ratio test cannot pass it with the exact distances either: flag it, refine skips it.
                    const int cq = __ldg(norm + (long long)I.img_q * feat_stride + q) + 2 * kExtOffset;
                    if (!ratio_keep(max(cq - 2 * M1, 0), cq - 2 * M2 + 1, pf.mode, pf.ratio, pf.num2, pf.den2)) flags |= 4;
                }
                // a prefiltered row is final: it reads as "no neighbours" and the refinement leaves it alone
                *reinterpret_cast<int4*>(knn_out + ((long long)I.pair * feat_stride + q) * 4) =
                    (flags & 4) ? make_int4(-1, -1, -1, -1) : make_int4(rec[0], rec[1], rec[2], flags);
            }
        }
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------ exact refinement
// One thread per query row: recompute the recorded candidate sub-groups exactly (dp4a on the offset-int8 rows,
// D = |a|^2 + |b|^2 - 2 a.b), keep the lexicographic top-2 (distance, index) and overwrite the record.
// Rows flagged for brute force are collected per block and swept by whole warps afterwards.
__device__ unsigned long long g_refine_brute_rows = 0ull;
__device__ unsigned long long g_refine_candidates = 0ull;

// One block owns 256 consecutive query rows of one pair (feat_stride % 256 == 0).  Thread i first classifies row i
// from its record: prefiltered rows are blanked on the spot, rows to refine and rows to brute-force are compacted
// into shared-memory lists, so that the warps below always work on full groups whatever fraction of the rows the
// sweep's prefilter removed.  Refinement proper: 8 lanes per query row, lane j owning the 16-byte slice j of the
// query row and of every candidate row (one warp-wide LDG.128 touches 4 whole 128-byte lines); the 8 partial dot
// products of a candidate sub-group are transposed-and-summed with 7 shuffles so that lane j ends up with the
// distance of candidate j, keeps its own top-2, and the 8 lanes merge once per row.
constexpr int kRefineRows = 256;
#ifndef SFM_REFINE_MINB
#define SFM_REFINE_MINB 5
#endif

// kFuse: instead of the kNN row, apply the match filter of filter.cu on the spot (Lowe ratio, optional distance bound, optional
// mutual check against the REVERSE direction's finished kNN table) and leave the block's surviving rows, compacted in query
// order, as (queryIdx, trainIdx, D1) triples at the start of the block's own 4 KB slice of knn_out (the candidate records
// that lived there have been consumed), with their number in blk_count[block]: the 16-byte-per-row kNN table is never
// written or read again (round-1 VERDICT: the table's round trips were ~10x the bytes of the packed result).
struct FuseParams {
    int mode, mutual, max_d;
    double ratio;
    long long num2, den2;
    const int32_t* knn_rev;
    int32_t* blk_count;
    int32_t* pair_count;            // zeroed by the caller; every block adds its survivors
};

template <bool kFuse>
__global__ void __launch_bounds__(256, SFM_REFINE_MINB) refine_kernel(const int8_t* __restrict__ desc, const int32_t* __restrict__ norm,
                                                     const int32_t* __restrict__ count, const int32_t* __restrict__ pairs,
                                                     int n_pairs, int feat_stride, int32_t* __restrict__ knn_out, int stats,
                                                     const FuseParams fp)
{
    __shared__ int4 recs[kRefineRows];
    __shared__ int4 res[kFuse ? kRefineRows : 1];                       // fused: (idx1, D1, idx2, D2) of every row of the block
    __shared__ int wtot[8];
    __shared__ int act_rows[kRefineRows];
    __shared__ int brute_rows[kRefineRows];
    __shared__ int n_act, n_brute;
    if (threadIdx.x == 0) { n_act = 0; n_brute = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sl = lane & 7;
    const unsigned gmask = 0xFFu << (lane & 24);
    const long long grow0 = (long long)blockIdx.x * kRefineRows;
    const int p = (int)(grow0 / feat_stride), q0 = (int)(grow0 % feat_stride);
    const int img_q = __ldg(pairs + 2 * p), img_t = __ldg(pairs + 2 * p + 1);
    const int nq = __ldg(count + img_q), nt = __ldg(count + img_t);
    const long long trow0 = (long long)img_t * feat_stride;
    int4* out = reinterpret_cast<int4*>(knn_out) + grow0;
    if (kFuse) res[threadIdx.x] = make_int4(-1, -1, -1, -1);
    {
        const int q = q0 + threadIdx.x;
        if (q < nq && nt > 0) {
            const int4 rec = out[threadIdx.x];
            if (rec.w < 0) {
                // prefiltered by the sweep (provably fails the ratio test): already written as (-1,-1,-1,-1)
            } else if (rec.w & 2) {
                brute_rows[atomicAdd(&n_brute, 1)] = threadIdx.x;
            } else {
                const int slot = atomicAdd(&n_act, 1);
                act_rows[slot] = threadIdx.x;
                recs[slot] = rec;
            }
        } else if (!kFuse) {
            out[threadIdx.x] = make_int4(-1, -1, -1, -1);            // rows no sweep unit owns (padding, empty train image)
        }
    }
    __syncthreads();
    const int nact = n_act, nbr = n_brute;
    int ncand = 0;
    for (int base = warp * 4; base < nact; base += 32) {
        const int i = base + (lane >> 3);
        if (i < nact) {                                              // whole 8-lane groups take this branch together
            const int r = act_rows[i];
            const int4 rec = recs[i];
            const long long qrow = (long long)img_q * feat_stride + q0 + r;
            const int4 a = __ldg(reinterpret_cast<const int4*>(desc + qrow * kDescDim) + sl);
            const int na = __ldg(norm + qrow);
            Top2 best;
            best.clear();
            // every recorded sub-group of one record entry: exact distances of its eight candidates into the lanes' top-2
            auto eval_entry = [&](const int key) {
                const int c0 = (key & 0xFFFF) * kTileRows;
                unsigned m16 = ((unsigned)key >> 16);                 // 0 for an unused entry
                while (m16) {
                    const int g = __ffs(m16) - 1;
                    m16 &= m16 - 1;
                    const int t0 = c0 + g * 8;
                    // rows [nt, feat_stride) of an image are zero padding inside the bank: an 8-row group never leaves the
                    // image's slab, so the eight loads are one base address + immediate offsets (no per-row clamp)
                    const int4* bp = reinterpret_cast<const int4*>(desc + (trow0 + t0) * kDescDim) + sl;
                    int4 b[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) b[j] = __ldg(bp + j * (kDescDim / 16));
                    const int nb = __ldg(norm + trow0 + t0 + sl);
                    int d[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        int dot = __dp4a(a.x, b[j].x, 0);
                        dot = __dp4a(a.y, b[j].y, dot);
                        dot = __dp4a(a.z, b[j].z, dot);
                        d[j] = __dp4a(a.w, b[j].w, dot);
                    }
                    // transpose-and-sum over the 8 lanes of the group: lane j ends with the full dot product of candidate j
                    int v4[4], v2[2];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int send = (sl & 4) ? d[j] : d[j + 4], keep = (sl & 4) ? d[j + 4] : d[j];
                        v4[j] = keep + __shfl_xor_sync(gmask, send, 4);
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int send = (sl & 2) ? v4[j] : v4[j + 2], keep = (sl & 2) ? v4[j + 2] : v4[j];
                        v2[j] = keep + __shfl_xor_sync(gmask, send, 2);
                    }
                    const int send = (sl & 1) ? v2[0] : v2[1], keep = (sl & 1) ? v2[1] : v2[0];
                    const int dot = keep + __shfl_xor_sync(gmask, send, 1);
                    if (t0 + sl < nt) { best.push(na + nb - 2 * dot, t0 + sl); ++ncand; }
                }
            };
            auto group_merge = [&](Top2 t) {
#pragma unroll
                for (int o = 1; o <= 4; o <<= 1) {
                    Top2 u;
                    u.d1 = __shfl_xor_sync(gmask, t.d1, o);
                    u.i1 = __shfl_xor_sync(gmask, t.i1, o);
                    u.d2 = __shfl_xor_sync(gmask, t.d2, o);
                    u.i2 = __shfl_xor_sync(gmask, t.i2, o);
                    t.merge(u);
                }
                return t;
            };
            Top2 all;
            if (kFuse) {
                // The fused form only needs (nearest index, nearest distance) and the OUTCOME of the ratio test.  The first entry is the
                // tile that holds the largest accumulator; every element outside it has acc <= M2, i.e. D >= dlo = C - 2 M2, and one
                // of them has D <= dlo + 1.  After the first tile: if its nearest is below dlo it is THE nearest; the second nearest is
                // its own second if that is <= dlo, else dlo or dlo + 1 -- and only when the (monotone) ratio test answers differently
                // for those two values do the other tiles have to be read.  Halves the refinement's reads for the common row.
                eval_entry(rec.x);
                all = group_merge(best);
                bool more = false;
                if (rec.w & 1) {
                    const int dlo = na + 2 * kExtOffset - 2 * ((rec.w >> 3) - kRecBias);
                    if (all.d1 >= dlo) {
                        more = true;                                       // an element of another tile may beat or tie the nearest
                    } else if (all.d2 > dlo) {
                        const bool lo = ratio_keep(all.d1, dlo, fp.mode, fp.ratio, fp.num2, fp.den2);
                        const bool hi = ratio_keep(all.d1, dlo + 1, fp.mode, fp.ratio, fp.num2, fp.den2);
                        if (lo != hi) more = true;
                        else { all.d2 = dlo; all.i2 = 0; }                 // (only the test's outcome leaves this kernel)
                    }
                }
                if (more) {                                                // (uniform inside the 8-lane group)
                    eval_entry(rec.y);
                    eval_entry(rec.z);
                    all = group_merge(best);
                }
            } else {
                eval_entry(rec.x);
                eval_entry(rec.y);
                eval_entry(rec.z);
                all = group_merge(best);
            }
            best = all;
            if (sl == 0) {
                if (kFuse) res[r] = make_int4(best.i1, best.i1 >= 0 ? best.d1 : -1, best.i2, best.i2 >= 0 ? best.d2 : -1);
                else store_knn(reinterpret_cast<int32_t*>(out + r), best);
            }
        }
    }
    if (nbr > 0) {
        for (int i = warp; i < nbr; i += 8) {
            // whole-image sweep of one row by a warp: 4 candidates per step, 8 lanes per candidate
            const int qq = q0 + brute_rows[i];
            const long long qrow = (long long)img_q * feat_stride + qq;
            const int4 a = __ldg(reinterpret_cast<const int4*>(desc + qrow * kDescDim) + sl);
            const int na = __ldg(norm + qrow);
            Top2 best;
            best.clear();
            for (int t0 = 0; t0 < nt; t0 += 4) {
                const int t = t0 + (lane >> 3);
                const long long tr = trow0 + min(t, nt - 1);
                const int4 b = __ldg(reinterpret_cast<const int4*>(desc + tr * kDescDim) + sl);
                int dot = __dp4a(a.x, b.x, 0);
                dot = __dp4a(a.y, b.y, dot);
                dot = __dp4a(a.z, b.z, dot);
                dot = __dp4a(a.w, b.w, dot);
                dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                dot += __shfl_xor_sync(0xffffffffu, dot, 4);
                if (t < nt) best.push_ordered(na + __ldg(norm + tr) - 2 * dot, t);
            }
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
                Top2 u;
                u.d1 = __shfl_xor_sync(0xffffffffu, best.d1, o);
                u.i1 = __shfl_xor_sync(0xffffffffu, best.i1, o);
                u.d2 = __shfl_xor_sync(0xffffffffu, best.d2, o);
                u.i2 = __shfl_xor_sync(0xffffffffu, best.i2, o);
                best.merge(u);
            }
            if (lane == 0) {
                if (kFuse) res[brute_rows[i]] = make_int4(best.i1, best.i1 >= 0 ? best.d1 : -1, best.i2, best.i2 >= 0 ? best.d2 : -1);
                else store_knn(knn_out + ((long long)p * feat_stride + qq) * 4, best);
            }
        }
    }
    if (kFuse) {
        __syncthreads();
        const int4 k = res[threadIdx.x];
        const int q = q0 + threadIdx.x;
        bool keep = k.x >= 0 && ratio_keep(k.y, k.w, fp.mode, fp.ratio, fp.num2, fp.den2);
        if (keep && fp.max_d > 0) keep = k.y < fp.max_d;
        if (keep && fp.mutual) keep = __ldg(fp.knn_rev + ((long long)p * feat_stride + k.x) * 4) == q;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wtot[warp] = __popc(bal);
        __syncthreads();                                             // (also: every record of the slice has been read long ago)
        int rank = __popc(bal & ((1u << lane) - 1u)), total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (w < warp) rank += wtot[w];
            total += wtot[w];
        }
        if (keep) {
            int32_t* o = reinterpret_cast<int32_t*>(out) + 3 * rank;
            o[0] = q; o[1] = k.x; o[2] = k.y;
        }
        if (threadIdx.x == 0) {
            fp.blk_count[blockIdx.x] = total;
            if (total) atomicAdd(fp.pair_count + p, total);
        }
    }
    if (stats) {
        for (int o = 16; o; o >>= 1) ncand += __shfl_xor_sync(0xffffffffu, ncand, o);
        if (lane == 0 && ncand) atomicAdd(&g_refine_candidates, (unsigned long long)ncand);
        if (threadIdx.x == 0 && nbr) atomicAdd(&g_refine_brute_rows, (unsigned long long)nbr);
    }
}

static int g_refine_stats = 0;

int launch_refine(const sfm_bank* b, const int32_t* pairs, int n_pairs, int32_t* knn_out, cudaStream_t st)
{
    const long long rows = (long long)n_pairs * b->L.feat_stride;
    refine_kernel<false><<<(unsigned)(rows / kRefineRows), 256, 0, st>>>(b->desc, b->norm, b->count, pairs, n_pairs, (int)b->L.feat_stride,
                                                                knn_out, g_refine_stats, FuseParams{});
    SFM_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return SFM_OK;
}

int launch_refine_filter(const sfm_bank* b, const int32_t* pairs, int n_pairs, int32_t* knn_out, const sfm_filter_params* prm,
                         const int32_t* knn_rev, int32_t* blk_count, int32_t* pair_count, cudaStream_t st)
{
    SFM_CUDA_CHECK(cudaMemsetAsync(pair_count, 0, sizeof(int32_t) * (size_t)n_pairs, st));
    const long long rows = (long long)n_pairs * b->L.feat_stride;
    FuseParams fp;
    fp.mode = prm->ratio_mode;
    fp.mutual = prm->mutual;
    fp.max_d = prm->max_distance_sq;
    fp.ratio = prm->ratio;
    fp.num2 = prm->ratio_num * prm->ratio_num;
    fp.den2 = prm->ratio_den * prm->ratio_den;
    fp.knn_rev = knn_rev;
    fp.blk_count = blk_count;
    fp.pair_count = pair_count;
    refine_kernel<true><<<(unsigned)(rows / kRefineRows), 256, 0, st>>>(b->desc, b->norm, b->count, pairs, n_pairs, (int)b->L.feat_stride,
                                                               knn_out, g_refine_stats, fp);
    SFM_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return SFM_OK;
}

int launch_match_tc(const sfm_bank* b, const int32_t* pairs, int n_pairs, int grid_req, int32_t* knn_out, int32_t* dbg_acc,
                    int dbg_mode, const Prefilter& pf, cudaStream_t st)
{
    if (!b->tmap_ready) {
        set_error("bank has no descriptor tensor map (metric must be L2)");
        return SFM_ERR_STATE;
    }
    static SmemAttrTable attr_prod, attr_dbg;                // per device (one process may drive several GPUs)
    SFM_CUDA_CHECK(ensure_dyn_smem(match_tc_kernel<false>, kTcSmemBytes, b->device, attr_prod));
    SFM_CUDA_CHECK(ensure_dyn_smem(match_tc_kernel<true>, kTcSmemBytes, b->device, attr_dbg));
    const long long units = (long long)n_pairs * (b->L.feat_stride / kUnitRows);
    int grid = grid_req > 0 ? grid_req : b->sm_count;
    if (grid > units) grid = (int)units;
    if (grid < 1) grid = 1;
    if ((dbg_acc != nullptr && dbg_mode != 5) || (dbg_mode != 0 && dbg_mode != 3 && dbg_mode != 5))
        match_tc_kernel<true><<<grid, kTcThreads, kTcSmemBytes, st>>>(b->tmap_desc, b->ext, b->count, pairs, n_pairs,
                                                                     (int)b->L.feat_stride, knn_out, dbg_acc, dbg_mode, b->norm, pf);
    else
        match_tc_kernel<false><<<grid, kTcThreads, kTcSmemBytes, st>>>(b->tmap_desc, b->ext, b->count, pairs, n_pairs,
                                                                      (int)b->L.feat_stride, knn_out, nullptr, dbg_mode == 5 ? 5 : 0, b->norm, pf);
    SFM_CUDA_CHECK(cudaGetLastError());
    count_launch();
    if (dbg_mode == 0) return launch_refine(b, pairs, n_pairs, knn_out, st);
    return SFM_OK;                                                    // (dbg_mode 3: the caller runs its own refinement, e.g. the fused one)
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

// Diagnostics: (rows brute-forced, candidates evaluated) by refine_kernel since the counters were enabled.
extern "C" int sfm_debug_refine_stats(int enable, int64_t out[2])
{
    g_refine_stats = enable;
    if (out) {
        unsigned long long v[2] = {0, 0};
        SFM_CUDA_CHECK(cudaMemcpyFromSymbol(&v[0], g_refine_brute_rows, sizeof(unsigned long long)));
        SFM_CUDA_CHECK(cudaMemcpyFromSymbol(&v[1], g_refine_candidates, sizeof(unsigned long long)));
        out[0] = (int64_t)v[0];
        out[1] = (int64_t)v[1];
    }
    return SFM_OK;
}

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
