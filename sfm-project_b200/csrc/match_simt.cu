// match_simt.cu -- CUDA-core (dp4a) exact L2 kNN(k=2) matcher.
//
// The correctness-first implementation of the matcher: it pins the epilogue
// semantics (exact integer distances, lowest-index tie-break) on the GPU and serves as an
// independent on-device cross-check of the tcgen05 kernel.  Selected with SFM_MATCH_SIMT.
// Replaces cv2.BFMatcher(NORM_L2).knnMatch(k=2) (north-star workload; the literal call site it
// displaces is bf.match, code/feature_matching.py:50).
#include "match_common.cuh"

namespace sfm {

constexpr int kSimtTile = 64;          // query rows / train rows per smem tile
constexpr int kSimtPitch = 33;         // words per smem row (32 + 1 pad)

// grid = (feat_stride / 64, n_pairs); block = 256 threads as 16 x 16; thread (ty,tx) owns query rows
// {ty + 16 i} and train columns {tx + 16 j} of each 64 x 64 tile.
__global__ void __launch_bounds__(256) match_simt_kernel(
    const int8_t* __restrict__ desc, const int32_t* __restrict__ norm, const int32_t* __restrict__ count,
    const int32_t* __restrict__ pairs, int feat_stride, int32_t* __restrict__ knn_out)
{
    __shared__ int sa[kSimtTile * kSimtPitch];
    __shared__ int sb[kSimtTile * kSimtPitch];
    __shared__ int snb[kSimtTile];
    __shared__ int4 smerge[kSimtTile * 16];

    const int p = blockIdx.y;
    const int img_q = pairs[2 * p], img_t = pairs[2 * p + 1];
    const int nq = count[img_q], nt = count[img_t];
    const int q0 = blockIdx.x * kSimtTile;
    if (q0 >= nq) return;
    const long long qrow0 = (long long)img_q * feat_stride + q0;
    const long long trow0 = (long long)img_t * feat_stride;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

    for (int e = tid; e < kSimtTile * 32; e += 256) {
        const int r = e >> 5, w = e & 31;
        sa[r * kSimtPitch + w] = reinterpret_cast<const int*>(desc + (qrow0 + r) * kDescDim)[w];
    }
    Top2 best[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) best[i].clear();

    for (int t0 = 0; t0 < nt; t0 += kSimtTile) {
        __syncthreads();
        for (int e = tid; e < kSimtTile * 32; e += 256) {
            const int r = e >> 5, w = e & 31;
            sb[r * kSimtPitch + w] = reinterpret_cast<const int*>(desc + (trow0 + t0 + r) * kDescDim)[w];
        }
        if (tid < kSimtTile) snb[tid] = norm[trow0 + t0 + tid];
        __syncthreads();
        int acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0;
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            int av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = sa[(ty + 16 * i) * kSimtPitch + k];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = sb[(tx + 16 * j) * kSimtPitch + k];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __dp4a(av[i], bv[j], acc[i][j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = t0 + tx + 16 * j;
            if (c < nt) {
                const int nb = snb[tx + 16 * j];
#pragma unroll
                for (int i = 0; i < 4; ++i) best[i].push_ordered(nb - 2 * acc[i][j], c);   // + |a|^2 added at the end
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        smerge[(ty + 16 * i) * 16 + tx] = make_int4(best[i].d1, best[i].i1, best[i].d2, best[i].i2);
    __syncthreads();
    if (tid < kSimtTile) {
        const int r = tid;
        Top2 t;
        t.clear();
        for (int x = 0; x < 16; ++x) {
            const int4 v = smerge[r * 16 + x];
            Top2 u;
            u.d1 = v.x; u.i1 = v.y; u.d2 = v.z; u.i2 = v.w;
            t.merge(u);
        }
        if (q0 + r < nq) {
            const int na = norm[qrow0 + r];
            if (t.i1 >= 0) t.d1 += na;
            if (t.i2 >= 0) t.d2 += na;
            store_knn(knn_out + ((long long)p * feat_stride + q0 + r) * 4, t);
        }
    }
}

int launch_match_simt(const sfm_bank* b, const int32_t* pairs, int n_pairs, int32_t* knn_out, cudaStream_t st)
{
    // the pair index is gridDim.y (limit 65535): long pair lists go in chunks
    constexpr int kChunk = 65535;
    for (int p0 = 0; p0 < n_pairs; p0 += kChunk) {
        const int np = n_pairs - p0 < kChunk ? n_pairs - p0 : kChunk;
        dim3 grid((unsigned)(b->L.feat_stride / kSimtTile), (unsigned)np);
        match_simt_kernel<<<grid, 256, 0, st>>>(b->desc, b->norm, b->count, pairs + 2 * (size_t)p0, (int)b->L.feat_stride,
                                                knn_out + (size_t)p0 * b->L.feat_stride * 4);
        SFM_CUDA_CHECK(cudaGetLastError());
        count_launch();
    }
    return SFM_OK;
}

}  // namespace sfm
