// ransac_common.cuh -- pieces shared by the batched RANSAC kernels (ransac_f.cu, ransac_h.cu) and the pose kernel.
// Every translation unit that includes this is compiled with -fmad=false: all floating-point work is IEEE basic
// arithmetic in a fixed order so that results are bit-identical to the C oracles under oracle/.
#pragma once
#include <math.h>

#include "common.cuh"

namespace sfm {

constexpr int kRansacThreads = 256;
constexpr int kBailEvery = 16;          // scoring: bail-out test every 16 warp iterations (512 correspondences)

// Hypotheses evaluated before the next termination check: 32, 32, 64, then 128 at a time (checks after 32, 64, 128, 256,
// 384, ... hypotheses).  Clean pairs -- the common case after the ratio test -- stop after the first 32.
__host__ __device__ constexpr int ransac_batch(int done) { return done < 64 ? 32 : (done < 128 ? 64 : 128); }

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


// Same cyclic Jacobi with the dimension known at compile time: every index is static, so A and V live in registers
// (the generic version indexes them dynamically, i.e. through local memory).  Operation order is identical.
template <int N>
static __device__ __forceinline__ void jacobi_eig_n(double (&A)[N * N], double (&V)[N * N], int sweeps)
{
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) V[i * N + j] = (i == j) ? 1.0 : 0.0;
#pragma unroll 1
    for (int s = 0; s < sweeps; ++s) {
#pragma unroll
        for (int p = 0; p < N - 1; ++p) {
#pragma unroll
            for (int q = p + 1; q < N; ++q) {
                const double apq = A[p * N + q];
                if (apq == 0.0) continue;
                const double app = A[p * N + p], aqq = A[q * N + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double at = fabs(theta);
                double t = 1.0 / (at + sqrt(theta * theta + 1.0));
                if (theta < 0.0) t = -t;
                const double c = 1.0 / sqrt(t * t + 1.0);
                const double sn = t * c;
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double akp = A[k * N + p], akq = A[k * N + q];
                    A[k * N + p] = c * akp - sn * akq;
                    A[k * N + q] = sn * akp + c * akq;
                }
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double apk = A[p * N + k], aqk = A[q * N + k];
                    A[p * N + k] = c * apk - sn * aqk;
                    A[q * N + k] = sn * apk + c * aqk;
                }
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double vkp = V[k * N + p], vkq = V[k * N + q];
                    V[k * N + p] = c * vkp - sn * vkq;
                    V[k * N + q] = sn * vkp + c * vkq;
                }
            }
        }
    }
}

// N block sums at once with ONE barrier pair (the LO refit needs 5 + 2 + 45 of them per round; one at a time that was 104
// barriers).  Same arithmetic as N calls of block_tree_sum: shfl_down tree per warp, then the eight warp sums added in warp
// order.  part: N * 8 doubles, tot: N doubles (shared memory); every thread returns with v[e] = the block total.
template <int N>
static __device__ __forceinline__ void block_tree_sum_many(double (&v)[N], double* part, double* tot)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < N; ++e) {
        double x = v[e];
        for (int off = 16; off >= 1; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
        if (lane == 0) part[e * 8 + warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double t = part[threadIdx.x * 8];
        for (int w = 1; w < kRansacThreads / 32; ++w) t += part[threadIdx.x * 8 + w];
        tot[threadIdx.x] = t;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < N; ++e) v[e] = tot[e];
}

// Eigenvector of the smallest eigenvalue of a symmetric positive semi-definite 9 x 9 matrix (the normal matrix of a
// DLT / 8-point system): Cholesky factor of A + eps*I (eps = 1e-13 * trace keeps exact data factorisable and does not
// move the eigenvectors), then inverse iterations from a fixed start vector until the normalised iterate stops moving
// (at most 16; well-conditioned systems take 3-4).  Serial, a few thousand cycles --
// the 10-sweep cyclic Jacobi it replaces took ~900 K cycles on one thread and made the LO refit 13x the cost of the
// whole hypothesis loop.  `w` is scratch for 45 doubles (shared memory).  Returns 0 when A is not usable.
static __device__ int smallest_eigvec9(const double* __restrict__ A, double* __restrict__ w, double* __restrict__ xout)
{
    double* __restrict__ L = w;            // packed lower triangle in shared memory, row i at L[i*(i+1)/2]
    double invd[9], x[9], y[9];            // statically indexed (every loop below is fully unrolled): registers
    double tr = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) tr += A[i * 9 + i];
    if (!(tr > 0.0) || !(tr < 1e300)) return 0;
    const double eps = tr * 1e-13;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        double sj = A[j * 9 + j] + eps;
#pragma unroll
        for (int k = 0; k < j; ++k) sj -= L[j * (j + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
        if (!(sj > 0.0)) return 0;
        const double d = sqrt(sj);
        L[j * (j + 1) / 2 + j] = d;
        invd[j] = 1.0 / d;
#pragma unroll
        for (int i = j + 1; i < 9; ++i) {
            double v = A[i * 9 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
            L[i * (i + 1) / 2 + j] = v * invd[j];
        }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] = 1.0 + 0.125 * (double)i;
#pragma unroll 1
    for (int it = 0; it < 16; ++it) {
#pragma unroll
        for (int i = 0; i < 9; ++i) {                       // L y = x
            double v = x[i];
#pragma unroll
            for (int k = 0; k < i; ++k) v -= L[i * (i + 1) / 2 + k] * y[k];
            y[i] = v * invd[i];
        }
        double z[9];
#pragma unroll
        for (int i = 8; i >= 0; --i) {                      // L^T z = y
            double v = y[i];
#pragma unroll
            for (int k = i + 1; k < 9; ++k) v -= L[k * (k + 1) / 2 + i] * z[k];
            z[i] = v * invd[i];
        }
        double ss = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) ss += z[i] * z[i];
        if (!(ss > 0.0) || !(ss < 1e300)) return 0;
        const double inv = 1.0 / sqrt(ss);
        double diff = 0.0;                                  // converged when the normalised iterate stops moving
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const double xi = z[i] * inv;
            diff = fmax(diff, fabs(xi - x[i]));
            x[i] = xi;
        }
        if (it > 0 && diff < 1e-13) break;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) xout[i] = x[i];
    return 1;
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


// ---------------------------------------------------------------------------------------------------------------
// Cooperative Gauss-Jordan elimination with complete pivoting: 8 lanes (one aligned octet of a warp) reduce ONE
// m x 9 system, lane r holding row r in registers (a[0..8], statically indexed).  Value for value the serial
// elimination of oracle/ransac_f.c / ransac_h.c: rows and columns are never moved, their LOGICAL positions are
// tracked instead (lrow per lane; lc / pc = logical<->physical column maps, 4 bits per entry), the pivot search
// breaks ties by the smallest logical (row, column) exactly like the serial strict-greater scan, and every element
// sees the same operations (a *= 1/pivot on the pivot row, a -= f * pivot_row elsewhere) in the same order.
// The serial version keeps 8x9 doubles per thread in local memory, which thrashes L1 at 2 CTAs / SM; this one
// keeps 18 registers per lane and uses all 256 threads of the CTA.
struct CoopGJ {
    unsigned long long lc;   // lc[j]: logical position of physical column j
    unsigned long long pc;   // pc[k]: physical column at logical position k
    int lrow;                // logical position of this lane's row (15 = lane holds no row)
};

static __device__ __forceinline__ double sel9(const double (&a)[9], int i)
{
    double v = a[0];
#pragma unroll
    for (int j = 1; j < 9; ++j) v = (i == j) ? a[j] : v;
    return v;
}

static __device__ __forceinline__ unsigned long long set4(unsigned long long w, int pos, int val)
{
    return (w & ~(15ull << (4 * pos))) | ((unsigned long long)val << (4 * pos));
}

// Returns false (in every lane of the octet) when a pivot is <= tol.  On success st describes the final layout.
template <int M>
static __device__ __forceinline__ bool coop_gauss_jordan(double (&a)[9], int sl, unsigned gmask, double tol, CoopGJ& st)
{
    st.lc = 0x876543210ull;
    st.pc = 0x876543210ull;
    st.lrow = (sl < M) ? sl : 15;
    const int src0 = (threadIdx.x & 24);                     // first lane of this octet within the warp
#pragma unroll 1
    for (int k = 0; k < M; ++k) {
        double bv = -1.0;
        int bkey = 0x7fffffff;
        if (st.lrow != 15 && st.lrow >= k) {
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                const int lj = (int)(st.lc >> (4 * j)) & 15;
                const double v = fabs(a[j]);
                const int key = ((st.lrow * 9 + lj) << 8) | (sl << 4) | j;
                if (lj >= k && (v > bv || (v == bv && key < bkey))) { bv = v; bkey = key; }
            }
        }
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) {
            const double ov = __shfl_xor_sync(gmask, bv, o);
            const int ok = __shfl_xor_sync(gmask, bkey, o);
            if (ov > bv || (ov == bv && ok < bkey)) { bv = ov; bkey = ok; }
        }
        if (!(bv > tol)) return false;
        const int pl = (bkey >> 4) & 15, pcol = bkey & 15;
        const int lidx = bkey >> 8, pi = lidx / 9, pj = lidx - 9 * pi;
        if (st.lrow == k) st.lrow = pi;
        else if (sl == pl) st.lrow = k;
        const int c0 = (int)(st.pc >> (4 * k)) & 15;          // physical column now at logical k moves to logical pj
        st.lc = set4(set4(st.lc, c0, pj), pcol, k);
        st.pc = set4(set4(st.pc, pj, c0), k, pcol);
        const double own = sel9(a, pcol);                     // pivot lane: the pivot; other lanes: their factor f
        if (sl == pl) {
            const double inv = 1.0 / own;
#pragma unroll
            for (int j = 0; j < 9; ++j)
                if (((int)(st.lc >> (4 * j)) & 15) >= k) a[j] *= inv;
        }
        double r[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) r[j] = __shfl_sync(gmask, a[j], src0 + pl);
        if (sl != pl && sl < M) {
#pragma unroll
            for (int j = 0; j < 9; ++j)
                if (((int)(st.lc >> (4 * j)) & 15) >= k) a[j] -= own * r[j];
        }
    }
    return true;
}

// After a successful elimination: scatter null vector c (c < 9 - M) into dst[0..8] (shared memory, one writer per entry).
template <int M>
static __device__ __forceinline__ void coop_null_vector(const double (&a)[9], int sl, const CoopGJ& st, int c, double* dst)
{
    const int fc = (int)(st.pc >> (4 * (M + c))) & 15;        // physical free column of this null vector
    if (sl < M) dst[(int)(st.pc >> (4 * st.lrow)) & 15] = -sel9(a, fc);
    if (sl == 0) {
#pragma unroll
        for (int e = 0; e < 9 - M; ++e) dst[(int)(st.pc >> (4 * (M + e))) & 15] = (e == c) ? 1.0 : 0.0;
    }
}

}  // namespace sfm
