// match_common.cuh -- exact top-2 bookkeeping shared by the SIMT and tcgen05 matchers.
#pragma once
#include "common.cuh"

namespace sfm {

constexpr int kNoDist = 0x7fffffff;

// Lowe's ratio test on exact squared distances, bit for bit the knnMatch idiom `m.distance < ratio * n.distance`
// with cv2's float32 distances (SFM_RATIO_CV2_F32) or in exact integers (SFM_RATIO_EXACT_INT); SURVEY.md D8 / A.3.
// Monotone: if it holds for (d1, d2) it holds for every (d1' <= d1, d2' >= d2) -- the sweep's prefilter relies on that.
__device__ __forceinline__ bool ratio_keep(int d1, int d2, int mode, double ratio, long long num2, long long den2)
{
    if (mode == SFM_RATIO_NONE) return true;
    if (d2 < 0) return false;                       // no second neighbour
    if (mode == SFM_RATIO_CV2_F32) {
        const float s1 = __fsqrt_rn((float)d1), s2 = __fsqrt_rn((float)d2);
        return (double)s1 < __dmul_rn(ratio, (double)s2);
    }
    return (long long)d1 * den2 < (long long)d2 * num2;
}

// ratio test the sweep applies to distance BOUNDS (sfm_match_params.prefilter_*)
struct Prefilter {
    int mode;
    double ratio;
    long long num2, den2;
};

// (distance, index) ascending, lowest index wins a tie: the order knnMatch reports
// (SURVEY.md A.1).  d == kNoDist means "empty".
struct Top2 {
    int d1, i1, d2, i2;
    __device__ __forceinline__ void clear() { d1 = d2 = kNoDist; i1 = i2 = -1; }
    // candidates arriving in increasing index order only need the distance comparison
    __device__ __forceinline__ void push_ordered(int d, int i)
    {
        if (d < d2) {
            if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = i; }
            else { d2 = d; i2 = i; }
        }
    }
    __device__ __forceinline__ static bool less(int da, int ia, int db, int ib)
    {
        return da < db || (da == db && (unsigned)ia < (unsigned)ib);
    }
    // arbitrary arrival order: full lexicographic comparison (i = -1 sorts last via unsigned)
    __device__ __forceinline__ void push(int d, int i)
    {
        if (less(d, i, d2, i2)) {
            if (less(d, i, d1, i1)) { d2 = d1; i2 = i1; d1 = d; i1 = i; }
            else { d2 = d; i2 = i; }
        }
    }
    __device__ __forceinline__ void merge(const Top2& o) { push(o.d1, o.i1); push(o.d2, o.i2); }
};

__device__ __forceinline__ Top2 warp_merge(Top2 t)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        Top2 u;
        u.d1 = __shfl_xor_sync(0xffffffffu, t.d1, o);
        u.i1 = __shfl_xor_sync(0xffffffffu, t.i1, o);
        u.d2 = __shfl_xor_sync(0xffffffffu, t.d2, o);
        u.i2 = __shfl_xor_sync(0xffffffffu, t.i2, o);
        t.merge(u);
    }
    return t;
}

// exact squared distance of an offset-int8 query row held in registers (32 words) to bank row `grow`
__device__ __forceinline__ int exact_sqdist(const int (&a)[32], int na, const int8_t* __restrict__ desc,
                                            const int32_t* __restrict__ norm, long long grow)
{
    const int4* b = reinterpret_cast<const int4*>(desc + grow * kDescDim);
    int dot = 0;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const int4 w = __ldg(b + v);
        dot = __dp4a(a[4 * v + 0], w.x, dot);
        dot = __dp4a(a[4 * v + 1], w.y, dot);
        dot = __dp4a(a[4 * v + 2], w.z, dot);
        dot = __dp4a(a[4 * v + 3], w.w, dot);
    }
    return na + __ldg(norm + grow) - 2 * dot;
}

// One warp, one query row, every train row of the image: the always-exact path used for rows the
// tensor-core sweep could not resolve (chunk-maximum ties) and by the SIMT kernel's row mode.
__device__ __forceinline__ Top2 warp_bruteforce_row(const int (&a)[32], int na, const int8_t* __restrict__ desc,
                                                    const int32_t* __restrict__ norm, long long train_row0, int n_train,
                                                    int lane)
{
    Top2 t;
    t.clear();
    for (int c = lane; c < n_train; c += 32) t.push_ordered(exact_sqdist(a, na, desc, norm, train_row0 + c), c);
    return warp_merge(t);
}

__device__ __forceinline__ void load_query_row(int (&a)[32], const int8_t* __restrict__ desc, long long grow)
{
    const int4* p = reinterpret_cast<const int4*>(desc + grow * kDescDim);
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const int4 w = __ldg(p + v);
        a[4 * v + 0] = w.x; a[4 * v + 1] = w.y; a[4 * v + 2] = w.z; a[4 * v + 3] = w.w;
    }
}

__device__ __forceinline__ void store_knn(int32_t* __restrict__ knn_row, const Top2& t)
{
    int4 o;
    o.x = t.i1; o.y = (t.i1 >= 0) ? t.d1 : -1; o.z = t.i2; o.w = (t.i2 >= 0) ? t.d2 : -1;
    *reinterpret_cast<int4*>(knn_row) = o;
}

}  // namespace sfm
