// pose.cu -- batched two-view initialisation from verified pairs.  Compile with -fmad=false.
//
// The step immediately downstream of the hot path, feeding the reference's empty code/3d_reconstruction.py
// (0 bytes; imported at code/pipeline.py:4): for every verified pair, E = K2^T F K1, the four (R, t) decompositions,
// the cheirality vote by DLT triangulation of the inlier correspondences, and the triangulated points of the winner.
// Conventions follow cv2.recoverPose(E, pts1, pts2, K, distanceThresh) + cv2.triangulatePoints: x2 ~ R x1 + t,
// |t| = 1, candidate order (R1,t) (R2,t) (R1,-t) (R2,-t) with ties going to the earlier one, a point counts iff its
// depth is in (0, distanceThresh) in BOTH cameras.
//
// One CTA (256 threads) per pair.  Thread 0 does the 3x3 algebra in fp64 (Jacobi eigen-decomposition of E^T E);
// every thread then triangulates its share of the correspondences for all four candidates (4x4 DLT, null vector by
// Jacobi on A^T A, fp64), the four vote counts are reduced in shared memory, and a second pass writes the points
// and the cheirality mask of the chosen candidate.  Integer votes and a fixed operation order make the result
// bit-identical to oracle/pose.c.
#include "ransac_common.cuh"

namespace sfm {

#ifndef SFM_POSE_MINB
#define SFM_POSE_MINB 2
#endif

#ifndef SFM_TRI_INLINE
#define SFM_TRI_INLINE __noinline__
#endif

struct PoseSmem {
    double R[2][9];
    double t[3];
    double tneg[3];
    double E[9];
    int votes[4];
    int ok, choice;
};

static __device__ void cross3(const double* a, const double* b, double* c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// E (unit Frobenius norm) -> R1, R2 (row-major), t.  Returns 0 when E has rank < 2.
static __device__ int decompose_essential(const double* E, double* R1, double* R2, double* t)
{
    double G[9], V[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) G[i * 3 + j] = E[0 + i] * E[0 + j] + E[3 + i] * E[3 + j] + E[6 + i] * E[6 + j];
    jacobi_eig_n<3>(G, V, 8);
    const double g[3] = {G[0], G[4], G[8]};
    int i0 = 0;
    if (g[1] > g[i0]) i0 = 1;
    if (g[2] > g[i0]) i0 = 2;
    int i2 = 0;
    if (g[1] < g[i2]) i2 = 1;
    if (g[2] < g[i2]) i2 = 2;
    if (i0 == i2) return 0;
    const int i1 = 3 - i0 - i2;
    double v0[3] = {V[0 + i0], V[3 + i0], V[6 + i0]};
    double v1[3] = {V[0 + i1], V[3 + i1], V[6 + i1]};
    double v2[3];
    cross3(v0, v1, v2);
    double u0[3], u1[3], u2[3];
    for (int r = 0; r < 3; ++r) {
        u0[r] = E[r * 3 + 0] * v0[0] + E[r * 3 + 1] * v0[1] + E[r * 3 + 2] * v0[2];
        u1[r] = E[r * 3 + 0] * v1[0] + E[r * 3 + 1] * v1[1] + E[r * 3 + 2] * v1[2];
    }
    const double n0 = sqrt(u0[0] * u0[0] + u0[1] * u0[1] + u0[2] * u0[2]);
    if (!(n0 > 1e-12)) return 0;
    for (int r = 0; r < 3; ++r) u0[r] = u0[r] / n0;
    const double d = u1[0] * u0[0] + u1[1] * u0[1] + u1[2] * u0[2];
    for (int r = 0; r < 3; ++r) u1[r] = u1[r] - d * u0[r];
    const double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    if (!(n1 > 1e-9)) return 0;
    for (int r = 0; r < 3; ++r) u1[r] = u1[r] / n1;
    cross3(u0, u1, u2);
    // R1 = U W V^T = u1 v0^T - u0 v1^T + u2 v2^T,  R2 = U W^T V^T = -u1 v0^T + u0 v1^T + u2 v2^T
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            const double a = u1[r] * v0[c] - u0[r] * v1[c];
            const double b = u2[r] * v2[c];
            R1[r * 3 + c] = b + a;
            R2[r * 3 + c] = b - a;
        }
    t[0] = u2[0]; t[1] = u2[1]; t[2] = u2[2];
    return 1;
}

// DLT triangulation of one correspondence in normalised camera coordinates, P0 = [I|0], P1 = [R|t]:
// X (camera-1 frame) and its depth z2 in camera 2; returns 0 when the null vector has no finite de-homogenisation.
//
// Mirror property used by the vote: replacing t by -t negates exactly the entries of A^T A (and of every Jacobi iterate
// and eigenvector) that carry the index 3 once -- IEEE multiplication, addition and division are sign-symmetric -- so
// the triangulated point and both depths of candidate (R, -t) are the exact negatives of those of (R, t): one
// triangulation serves two candidates, bit for bit (oracle/pose.c triangulates all four; tests compare).
static __device__ SFM_TRI_INLINE int triangulate(const double* R, const double* t, double x1, double y1, double x2, double y2, double* X, double* z2)
{
    double A[4][4];
    A[0][0] = -1.0; A[0][1] = 0.0;  A[0][2] = x1; A[0][3] = 0.0;
    A[1][0] = 0.0;  A[1][1] = -1.0; A[1][2] = y1; A[1][3] = 0.0;
    for (int c = 0; c < 3; ++c) {
        A[2][c] = x2 * R[6 + c] - R[0 + c];
        A[3][c] = y2 * R[6 + c] - R[3 + c];
    }
    A[2][3] = x2 * t[2] - t[0];
    A[3][3] = y2 * t[2] - t[1];
    double G[16], V[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) {
            const double s = A[0][i] * A[0][j] + A[1][i] * A[1][j] + A[2][i] * A[2][j] + A[3][i] * A[3][j];
            G[i * 4 + j] = s;
            G[j * 4 + i] = s;
        }
    jacobi_eig_n<4>(G, V, 6);
    // smallest eigenvalue -> its eigenvector, selected without dynamic indexing
    int k = 0;
    double gk = G[0];
    if (G[5] < gk) { k = 1; gk = G[5]; }
    if (G[10] < gk) { k = 2; gk = G[10]; }
    if (G[15] < gk) { k = 3; gk = G[15]; }
    const double e0 = k == 0 ? V[0] : k == 1 ? V[1] : k == 2 ? V[2] : V[3];
    const double e1 = k == 0 ? V[4] : k == 1 ? V[5] : k == 2 ? V[6] : V[7];
    const double e2 = k == 0 ? V[8] : k == 1 ? V[9] : k == 2 ? V[10] : V[11];
    const double w = k == 0 ? V[12] : k == 1 ? V[13] : k == 2 ? V[14] : V[15];
    X[0] = 0.0; X[1] = 0.0; X[2] = 0.0;
    *z2 = 0.0;
    if (!(fabs(w) > 1e-300)) return 0;
    X[0] = e0 / w;
    X[1] = e1 / w;
    X[2] = e2 / w;
    *z2 = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    return 1;
}

// both depths in (0, dist)
static __device__ __forceinline__ int in_front(double z1, double z2, double dist)
{
    return (z1 > 0.0 && z1 < dist && z2 > 0.0 && z2 < dist) ? 1 : 0;
}

__global__ void __launch_bounds__(kRansacThreads, SFM_POSE_MINB) pose_kernel(
    const float* __restrict__ corr, int corr_stride, const int32_t* __restrict__ count, const int32_t* __restrict__ offsets,
    const uint8_t* __restrict__ in_mask, const double* __restrict__ Fin, const double* __restrict__ cam, double dist,
    double* __restrict__ out_R, double* __restrict__ out_t, double* __restrict__ out_E, int32_t* __restrict__ out_ngood,
    uint8_t* __restrict__ out_mask, float* __restrict__ out_X, int stash_cap)
{
    __shared__ PoseSmem S;
    extern __shared__ float stash[];                   // [stash_cap][3]: candidate R2's point of every correspondence
    const int p = blockIdx.x, tid = threadIdx.x;
    const long long base = offsets ? (long long)offsets[p] : (long long)p * corr_stride;
    const int M = offsets ? (offsets[p + 1] - offsets[p]) : min(count[p], corr_stride);
    const float4* pts = reinterpret_cast<const float4*>(corr) + base;
    const uint8_t* imask = in_mask ? in_mask + base : nullptr;
    uint8_t* omask = out_mask + base;
    float* oX = out_X ? out_X + 3 * base : nullptr;
    const double fx1 = cam[p * 8 + 0], fy1 = cam[p * 8 + 1], cx1 = cam[p * 8 + 2], cy1 = cam[p * 8 + 3];
    const double fx2 = cam[p * 8 + 4], fy2 = cam[p * 8 + 5], cx2 = cam[p * 8 + 6], cy2 = cam[p * 8 + 7];

    for (int i = tid; i < (offsets ? M : corr_stride); i += kRansacThreads) {
        omask[i] = 0;
        if (oX) { oX[3 * i + 0] = 0.f; oX[3 * i + 1] = 0.f; oX[3 * i + 2] = 0.f; }
    }
    if (tid < 9) { out_R[(long long)p * 9 + tid] = 0.0; if (out_E) out_E[(long long)p * 9 + tid] = 0.0; }
    if (tid < 3) out_t[(long long)p * 3 + tid] = 0.0;
    if (tid < 4) S.votes[tid] = 0;
    if (tid == 0) {
        out_ngood[p] = 0;
        const double* F = Fin + (long long)p * 9;
        double G[9], E[9];
        for (int r = 0; r < 3; ++r) {
            G[r * 3 + 0] = F[r * 3 + 0] * fx1;
            G[r * 3 + 1] = F[r * 3 + 1] * fy1;
            G[r * 3 + 2] = F[r * 3 + 0] * cx1 + F[r * 3 + 1] * cy1 + F[r * 3 + 2];
        }
        for (int c = 0; c < 3; ++c) {
            E[0 + c] = fx2 * G[0 + c];
            E[3 + c] = fy2 * G[3 + c];
            E[6 + c] = cx2 * G[0 + c] + cy2 * G[3 + c] + G[6 + c];
        }
        double ss = 0.0;
        for (int i = 0; i < 9; ++i) ss += E[i] * E[i];
        int ok = (ss > 0.0) && (ss < 1e300) && M > 0;
        if (ok) {
            const double inv = 1.0 / sqrt(ss);
            for (int i = 0; i < 9; ++i) { E[i] *= inv; S.E[i] = E[i]; }
            ok = decompose_essential(E, S.R[0], S.R[1], S.t);
        }
        for (int i = 0; i < 3; ++i) S.tneg[i] = -S.t[i];
        S.ok = ok;
    }
    __syncthreads();
    if (!S.ok) return;

    // candidates are read from shared memory (broadcast loads) to keep the triangulation's 4x4 eigen-problem in registers
    const double* R1 = S.R[0];
    const double* R2 = S.R[1];
    const double* tp = S.t;
    const double* tn = S.tneg;
    // Pass 1 triangulates every used correspondence once per rotation and keeps what the selection needs: the four vote bits
    // in out_mask, candidate R1's point (float, the output format) in out_X and candidate R2's in shared memory.  Pass 2 is
    // then a selection (sign flip for the -t candidates); only correspondences beyond the stash are triangulated again.
    int v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    for (int i = tid; i < M; i += kRansacThreads) {
        if (imask && !imask[i]) continue;
        const float4 c = pts[i];
        const double x1 = ((double)c.x - cx1) / fx1, y1 = ((double)c.y - cy1) / fy1;
        const double x2 = ((double)c.z - cx2) / fx2, y2 = ((double)c.w - cy2) / fy2;
        double X[3], z2;
        int bits = 0;
        if (triangulate(R1, tp, x1, y1, x2, y2, X, &z2)) {
            bits |= in_front(X[2], z2, dist) | (in_front(-X[2], -z2, dist) << 2);      // (R1, t) and its mirror (R1, -t)
            if (oX) { oX[3 * i + 0] = (float)X[0]; oX[3 * i + 1] = (float)X[1]; oX[3 * i + 2] = (float)X[2]; }
        }
        if (triangulate(R2, tp, x1, y1, x2, y2, X, &z2)) {
            bits |= (in_front(X[2], z2, dist) << 1) | (in_front(-X[2], -z2, dist) << 3);
            if (oX && i < stash_cap) { stash[3 * i + 0] = (float)X[0]; stash[3 * i + 1] = (float)X[1]; stash[3 * i + 2] = (float)X[2]; }
        }
        omask[i] = (uint8_t)bits;
        v0 += bits & 1; v1 += (bits >> 1) & 1; v2 += (bits >> 2) & 1; v3 += (bits >> 3) & 1;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        v0 += __shfl_down_sync(0xffffffffu, v0, off);
        v1 += __shfl_down_sync(0xffffffffu, v1, off);
        v2 += __shfl_down_sync(0xffffffffu, v2, off);
        v3 += __shfl_down_sync(0xffffffffu, v3, off);
    }
    if ((tid & 31) == 0) {
        atomicAdd(&S.votes[0], v0); atomicAdd(&S.votes[1], v1); atomicAdd(&S.votes[2], v2); atomicAdd(&S.votes[3], v3);
    }
    __syncthreads();
    if (tid == 0) {
        int best = 0;
        for (int k = 1; k < 4; ++k)
            if (S.votes[k] > S.votes[best]) best = k;       // ties keep the earlier candidate (cv2's >= chain)
        S.choice = best;
    }
    __syncthreads();
    const int ch = S.choice;
    const double* R = (ch & 1) ? R2 : R1;
    const double* t = (ch & 2) ? tn : tp;
    const float sgn = (ch & 2) ? -1.f : 1.f;
    for (int i = tid; i < M; i += kRansacThreads) {
        if (imask && !imask[i]) continue;
        const int good = (omask[i] >> ch) & 1;
        omask[i] = (uint8_t)good;
        if (!oX) continue;
        float x = 0.f, y = 0.f, z = 0.f;
        if (good) {
            if (!(ch & 1)) {
                x = sgn * oX[3 * i + 0]; y = sgn * oX[3 * i + 1]; z = sgn * oX[3 * i + 2];
            } else if (i < stash_cap) {
                x = sgn * stash[3 * i + 0]; y = sgn * stash[3 * i + 1]; z = sgn * stash[3 * i + 2];
            } else {                                                   // beyond the stash: triangulate again
                const float4 c = pts[i];
                const double x1 = ((double)c.x - cx1) / fx1, y1 = ((double)c.y - cy1) / fy1;
                const double x2 = ((double)c.z - cx2) / fx2, y2 = ((double)c.w - cy2) / fy2;
                double X[3], z2;
                triangulate(R, t, x1, y1, x2, y2, X, &z2);
                x = (float)X[0]; y = (float)X[1]; z = (float)X[2];
            }
        }
        oX[3 * i + 0] = x; oX[3 * i + 1] = y; oX[3 * i + 2] = z;
    }
    if (tid < 9) {
        out_R[(long long)p * 9 + tid] = R[tid];
        if (out_E) out_E[(long long)p * 9 + tid] = S.E[tid];
    }
    if (tid < 3) out_t[(long long)p * 3 + tid] = t[tid];
    if (tid == 0) out_ngood[p] = S.votes[ch];
}

}  // namespace sfm

using namespace sfm;

static int launch_pose(const float* corr, int corr_stride, const int32_t* count, const int32_t* offsets, int n_pairs,
                       const uint8_t* in_mask, const double* F, const double* cam, double dist, double* out_R, double* out_t,
                       double* out_E, int32_t* out_ngood, uint8_t* out_mask, float* out_X, void* stream)
{
    SFM_REQUIRE(corr && (count || offsets) && F && cam && out_R && out_t && out_ngood && out_mask, "sfm_two_view_pose: NULL argument");
    SFM_REQUIRE(corr_stride > 0 && n_pairs >= 0, "bad sizes");
    SFM_REQUIRE(dist > 0.0, "distance_thresh must be positive");
    SFM_REQUIRE(((uintptr_t)corr & 15) == 0, "corr must be 16-byte aligned");
    if (n_pairs == 0) return SFM_OK;
    // shared-memory stash of candidate R2's points: 4,096 correspondences (48 KB; two CTAs per SM); wider pairs re-triangulate the rest
    constexpr int kStashCap = 4096;
    const size_t smem = out_X ? (size_t)kStashCap * 12 : 0;
    static SmemAttrTable attr;                               // per device; these entry points run on the caller's current device
    SFM_CUDA_CHECK(ensure_dyn_smem(pose_kernel, (size_t)kStashCap * 12, current_device(), attr));
    pose_kernel<<<n_pairs, kRansacThreads, smem, (cudaStream_t)stream>>>(corr, corr_stride, count, offsets, in_mask, F, cam, dist, out_R,
                                                                        out_t, out_E, out_ngood, out_mask, out_X, out_X ? kStashCap : 0);
    SFM_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return SFM_OK;
}

extern "C" int sfm_two_view_pose_batch(const float* corr, int corr_stride, const int32_t* count, int n_pairs, const uint8_t* in_mask,
                                       const double* F, const double* cam, double distance_thresh, double* out_R, double* out_t,
                                       double* out_E, int32_t* out_ngood, uint8_t* out_mask, float* out_X, void* stream)
{
    SFM_REQUIRE(count != nullptr, "sfm_two_view_pose_batch: NULL argument");
    return launch_pose(corr, corr_stride, count, nullptr, n_pairs, in_mask, F, cam, distance_thresh, out_R, out_t, out_E, out_ngood,
                       out_mask, out_X, stream);
}

extern "C" int sfm_two_view_pose_packed(const float* corr, const int32_t* offsets, int n_pairs, const uint8_t* in_mask,
                                        const double* F, const double* cam, double distance_thresh, double* out_R, double* out_t,
                                        double* out_E, int32_t* out_ngood, uint8_t* out_mask, float* out_X, void* stream)
{
    SFM_REQUIRE(offsets != nullptr, "sfm_two_view_pose_packed: NULL argument");
    return launch_pose(corr, 1, nullptr, offsets, n_pairs, in_mask, F, cam, distance_thresh, out_R, out_t, out_E, out_ngood, out_mask,
                       out_X, stream);
}
