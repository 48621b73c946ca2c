// tc_ptx.cuh -- PTX wrappers (mbarrier, TMA, tcgen05), UMMA descriptors and small helpers shared by the tcgen05 matcher
// kernels (match_tc.cu: one CTA per SM; match_tc2.cu: 2-CTA clusters with multicast B tiles).
#pragma once
#include "match_common.cuh"

namespace sfm {

constexpr int kUnitRows = 256;
constexpr int kTileBytes = kTileRows * kDescDim;             // 16384
constexpr int kBStageBytes = kTileBytes + kExtTileBytes;     // 20480
constexpr int kTmemCols = 512;
constexpr int kMaskedAcc = -2147483647 - 1;                  // INT_MIN: below every real accumulator and every sentinel

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
// suspend-time hint: without it a failed try_wait returns after a few tens of cycles and the single-lane producer /
// issuer threads spend a third of the SM's issue slots re-polling (ncu source view of round 1); with it ptxas emits
// NANOSLEEP.SYNCS and the thread sleeps until the barrier phase flips (sweep -2.5 %)
#ifndef SFM_SUSPEND_NS
#define SFM_SUSPEND_NS 20000
#endif
constexpr uint32_t kSuspendHintNs = SFM_SUSPEND_NS;
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(kSuspendHintNs)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_nohint(uint32_t bar, uint32_t parity)
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
// latency-critical handshakes (accumulator full / empty): plain re-polling wakes faster than NANOSLEEP.SYNCS
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
#pragma unroll 1
    while (!mbar_try_wait_nohint(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
#ifndef SFM_SPIN_MASK
#define SFM_SPIN_MASK 0
#endif
// Bounded wait: a protocol bug traps (sticky error reported to the host) instead of hanging the GPU.  The loop is
// kept rolled on purpose: unrolled copies at every call site pushed the kernel past the instruction cache and
// every role switch of the single-lane issuer warps then paid an I-cache miss (measured with the clock64 trace).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
#pragma unroll 1
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 16)) __trap();      // each failed try sleeps up to kSuspendHintNs: ~1 s in total
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
__device__ __forceinline__ void sts128(uint32_t saddr, int a, int b, int c, int d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
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

// UMMA shared-memory descriptors, split into 32-bit halves so the issuing thread only adds to the low word.
//   K-major SWIZZLE_128B operand: rows of 128 B, 8-row groups 1024 B apart (SBO), descriptor version 1.
//   K-major no-swizzle operand [k-chunk][row][16 B]: LBO = 2048 B between the two K chunks, SBO = 128 B between 8-row groups.
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t kDescHiExt = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t desc_lo_sw128(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_lo_ext(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | ((uint32_t)((kTileRows * 16) >> 4) << 16); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) { return mk_desc(kDescHiSw128, desc_lo_sw128(saddr)); }
__device__ __forceinline__ uint64_t desc_ext(uint32_t saddr) { return mk_desc(kDescHiExt, desc_lo_ext(saddr)); }
// kind::i8 instruction descriptor: D = s32, M = 128, N = 128, both operands K-major.
__host__ __device__ constexpr uint32_t idesc_i8(uint32_t a_signed, uint32_t b_signed)
{
    return (2u << 4) | (a_signed << 7) | (b_signed << 10) | ((uint32_t)(kTileRows >> 3) << 17) | ((uint32_t)(kTileRows >> 4) << 24);
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

// Epilogue bookkeeping, one thread per query row.  A 128-column tile is reduced to 16 sub-maxima (8 columns
// each) and their maximum m.  The thread keeps the three largest tile maxima (M1 >= M2 >= M3) with
// key = tile | (slot << 16); the 16 sub-maxima of each kept tile are parked in the thread's shared-memory
// slot so that, at the end of the sweep, the sub-groups that can still hold a top-2 element
// (sub-maximum >= M2) are known exactly.  tie4 = some tile outside the top three has maximum == M3.
constexpr int kInvalidTile = 0xFFFF;
constexpr int kRecBias = 1 << 22;    // accumulators lie in (-2^21 - 1, 2^21 + 2^20]: M2 + kRecBias fits the 29 bits of a record's fourth word
constexpr int kTraceTiles = 64;      // dbg_mode 4: clock64 timeline of the first 64 tiles of CTA 0 (role, tile, event)
#define SFM_TRACE(role, tile, ev)                                                                         \
    do {                                                                                                  \
        if (kDbg && dbg_mode == 4 && blockIdx.x == 0 && (tile) < kTraceTiles)                             \
            reinterpret_cast<long long*>(dbg_acc)[((role) * kTraceTiles + (tile)) * 8 + (ev)] = clock64(); \
    } while (0)

__device__ __forceinline__ void submax4(const uint32_t (&u)[32], int (&c)[16], int base)
{
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        int m = __vimax3_s32((int)u[8 * g], (int)u[8 * g + 1], (int)u[8 * g + 2]);
        m = __vimax3_s32(m, (int)u[8 * g + 3], (int)u[8 * g + 4]);
        m = __vimax3_s32(m, (int)u[8 * g + 5], (int)u[8 * g + 6]);
        c[base + g] = max(m, (int)u[8 * g + 7]);
    }
}


}  // namespace sfm
