// ransac_common.cuh -- pieces shared by the batched RANSAC kernels (ransac_f.cu, ransac_h.cu) and the pose kernel.
// Every translation unit that includes this is compiled with -fmad=false: all floating-point work is IEEE basic
// arithmetic in a fixed order so that results are bit-identical to the C oracles under oracle/.
#pragma once
#include <math.h>

#include "common.cuh"

namespace sfm {

constexpr int kRansacThreads = 256;

struct Norm2d { double s, cx, cy; };

// Correspondences of one pair: the first `cap` are staged in shared memory, the rest (very wide pairs only) are read
// through L1/L2.  Values are identical either way, so results do not depend on `cap`.
struct Pts {
    const float4* s;
    const float4* g;
    int cap;
    __device__ __forceinline__ float4 operator[](int i) const { return i < cap ? s[i] : __ldg(g + i); }
};

static __device__ __forceinline__ uint32_t rng_u32(uint64_t seed, uint32_t pair, uint32_t hyp, uint32_t ctr)
{
    uint64_t x = seed + 0x9E3779B97F4A7C15ULL * ((((uint64_t)pair) << 32) | (uint64_t)hyp);
    x ^= 0xD1B54A32D192ED03ULL * (uint64_t)(ctr + 1u);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return (uint32_t)(x >> 32);
}

static __device__ void draw_sample(uint64_t seed, uint32_t pair, uint32_t hyp, int m, int M, int* idx)
{
    for (int k = 0; k < m; ++k) {
        int v = 0;
        for (int attempt = 0; attempt < 16; ++attempt) {
            const uint32_t r = rng_u32(seed, pair, hyp, (uint32_t)(k * 16 + attempt));
            v = (int)(((uint64_t)r * (uint64_t)(uint32_t)M) >> 32);
            int dup = 0;
            for (int j = 0; j < k; ++j) dup |= (idx[j] == v);
            if (!dup) break;
        }
        idx[k] = v;
    }
}

// cyclic Jacobi on a symmetric n x n matrix; A diagonal -> eigenvalues, V columns -> eigenvectors
static __device__ void jacobi_eig(double* A, double* V, int n, int sweeps)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int s = 0; s < sweeps; ++s) {
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p * n + q];
                if (apq == 0.0) continue;
                const double app = A[p * n + p], aqq = A[q * n + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double at = fabs(theta);
                double t = 1.0 / (at + sqrt(theta * theta + 1.0));
                if (theta < 0.0) t = -t;
                const double c = 1.0 / sqrt(t * t + 1.0);
                const double sn = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - sn * akq;
                    A[k * n + q] = sn * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - sn * aqk;
                    A[q * n + k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - sn * vkq;
                    V[k * n + q] = sn * vkp + c * vkq;
                }
            }
        }
    }
}

static __device__ bool should_stop(int best, int M, int m, int done, double confidence)
{
    if (confidence >= 1.0 || best <= 0) return false;
    const double w = (double)best / (double)M;
    double wm = 1.0;
    for (int k = 0; k < m; ++k) wm *= w;
    const double q = 1.0 - wm;
    if (!(q > 0.0)) return true;
    double res = 1.0, base = q;
    int e = done;
    while (e) { if (e & 1) res *= base; base *= base; e >>= 1; }
    return res <= (1.0 - confidence);
}

// fixed-order block reduction: shfl_down tree per warp, warps summed in order by thread 0 (oracle: lane_tree)
static __device__ double block_tree_sum(double v, double* wsum)
{
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    double total = wsum[0];
    for (int w = 1; w < kRansacThreads / 32; ++w) total += wsum[w];
    return total;                       // every thread computes the same value in the same order
}

}  // namespace sfm
