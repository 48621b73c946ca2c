// hamming.cu -- the reference's literal matcher on the GPU (K3).
//
// Replaces, per image pair, code/feature_matching.py:48-58:
//     bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True); matches = bf.match(des1, des2)
//     matches = sorted(matches, key=lambda x: x.distance);  keep the prefix with distance < 26
// Semantics: row argmin and column argmin take the lowest index on ties, a match
// survives iff it is mutual, output order is (distance asc, queryIdx asc).
//
// Kernel A (grid: row tiles x pairs x 2 directions): one thread per query row keeps its 256-bit
// descriptor in 8 registers and scans the other image through shared memory (xor + popc).
// Kernel B (one CTA per pair): mutual check, threshold, bitonic sort of (distance << 16 | queryIdx).
#include "common.cuh"

namespace sfm {

constexpr int kHamTile = 256;

__global__ void __launch_bounds__(256) hamming_nn_kernel(const int8_t* __restrict__ desc, const int32_t* __restrict__ count,
                                                         const int32_t* __restrict__ pairs, int feat_stride,
                                                         int2* __restrict__ nn /* [2][P][feat_stride] */, int n_pairs, int pair0)
{
    __shared__ uint32_t tile[kHamTile * 8];
    const int p = pair0 + blockIdx.y, dir = blockIdx.z;
    const int img_q = pairs[2 * p + dir], img_t = pairs[2 * p + 1 - dir];
    const int nq = count[img_q], nt = count[img_t];
    const int q0 = blockIdx.x * kHamTile;
    if (q0 >= nq) return;
    const int q = q0 + threadIdx.x;
    uint32_t a[8];
    {
        const uint4* src = reinterpret_cast<const uint4*>(desc + ((long long)img_q * feat_stride + min(q, feat_stride - 1)) * kHammingDim);
        const uint4 lo = src[0], hi = src[1];
        a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w; a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
    }
    int best = 0x7fffffff, arg = -1;
    const uint32_t* tsrc = reinterpret_cast<const uint32_t*>(desc + (long long)img_t * feat_stride * kHammingDim);
    for (int t0 = 0; t0 < nt; t0 += kHamTile) {
        __syncthreads();
        for (int e = threadIdx.x; e < kHamTile * 8; e += 256) tile[e] = tsrc[(long long)t0 * 8 + e];
        __syncthreads();
        const int lim = min(kHamTile, nt - t0);
        for (int j = 0; j < lim; ++j) {
            int d = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) d += __popc(a[w] ^ tile[j * 8 + w]);
            if (d < best) { best = d; arg = t0 + j; }        // strict: lowest index wins ties
        }
    }
    if (q < nq) nn[((long long)dir * n_pairs + p) * feat_stride + q] = make_int2(arg, best);
}

__global__ void __launch_bounds__(256) hamming_select_kernel(const int32_t* __restrict__ count, const int32_t* __restrict__ pairs,
                                                             int feat_stride, const int2* __restrict__ nn, int n_pairs,
                                                             int max_distance, int sort_n, int32_t* __restrict__ out_count,
                                                             int32_t* __restrict__ out_match)
{
    extern __shared__ uint32_t keys[];
    __shared__ int kept;
    const int p = blockIdx.x;
    const int nq = count[pairs[2 * p]];
    const int2* nn12 = nn + (long long)p * feat_stride;
    const int2* nn21 = nn + ((long long)n_pairs + p) * feat_stride;
    if (threadIdx.x == 0) kept = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < sort_n; i += 256) {
        uint32_t key = 0xFFFFFFFFu;
        if (i < nq) {
            const int2 f = nn12[i];
            if (f.x >= 0 && f.y < max_distance && nn21[f.x].x == i) { key = ((uint32_t)f.y << 16) | (uint32_t)i; ++mine; }
        }
        keys[i] = key;
    }
    atomicAdd(&kept, mine);
    __syncthreads();
    for (int k = 2; k <= sort_n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < sort_n; i += 256) {
                const int l = i ^ j;
                if (l > i) {
                    const uint32_t x = keys[i], y = keys[l];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { keys[i] = y; keys[l] = x; }
                }
            }
            __syncthreads();
        }
    const int n = kept;
    for (int i = threadIdx.x; i < n; i += 256) {
        const uint32_t key = keys[i];
        const int q = (int)(key & 0xFFFFu);
        const long long o = ((long long)p * feat_stride + i) * 3;
        out_match[o + 0] = q;
        out_match[o + 1] = nn12[q].x;
        out_match[o + 2] = (int)(key >> 16);
    }
    if (threadIdx.x == 0) out_count[p] = n;
}

}  // namespace sfm

using namespace sfm;

extern "C" int sfm_match_hamming(const sfm_bank_t* bank, const int32_t* pairs_dev, int n_pairs, int max_distance,
                                 int32_t* out_count, int32_t* out_match, void* workspace, size_t workspace_bytes, void* stream)
{
    SFM_REQUIRE(bank && pairs_dev && out_count && out_match && workspace, "sfm_match_hamming: NULL argument");
    SFM_REQUIRE(bank->metric == SFM_METRIC_HAMMING, "sfm_match_hamming: bank metric is not Hamming");
    SFM_REQUIRE(n_pairs >= 0 && max_distance >= 0 && max_distance <= 257, "sfm_match_hamming: bad argument");
    SFM_REQUIRE(bank->L.feat_stride <= 32768, "sfm_match_hamming: at most 32768 features per image");
    const size_t need = (size_t)2 * n_pairs * bank->L.feat_stride * sizeof(int2);
    SFM_REQUIRE(workspace_bytes >= need, "sfm_match_hamming: workspace too small (%zu < %zu)", workspace_bytes, need);
    if (n_pairs == 0) return SFM_OK;
    SFM_ON_DEVICE(bank->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int fs = (int)bank->L.feat_stride;
    int2* nn = (int2*)workspace;
    SFM_CUDA_CHECK(cudaMemsetAsync(nn, 0xFF, need, st));
    int sort_n = 256;
    while (sort_n < bank->max_feats) sort_n <<= 1;
    const size_t smem = (size_t)sort_n * 4;
    static SmemAttrTable attr;
    SFM_CUDA_CHECK(ensure_dyn_smem(hamming_select_kernel, smem, bank->device, attr));
    // the pair index is gridDim.y of the distance kernel (limit 65535): long pair lists go in chunks
    constexpr int kChunk = 65535;
    for (int p0 = 0; p0 < n_pairs; p0 += kChunk) {
        const int np = n_pairs - p0 < kChunk ? n_pairs - p0 : kChunk;
        // (both directions of pair p live at nn[(dir * n_pairs + p) * fs]: the chunk passes the TOTAL pair count and its offset)
        dim3 grid((unsigned)(fs / kHamTile), (unsigned)np, 2);
        hamming_nn_kernel<<<grid, 256, 0, st>>>(bank->desc, bank->count, pairs_dev, fs, nn, n_pairs, p0);
        SFM_CUDA_CHECK(cudaGetLastError());
        count_launch();
    }
    hamming_select_kernel<<<n_pairs, 256, smem, st>>>(bank->count, pairs_dev, fs, nn, n_pairs, max_distance, sort_n, out_count, out_match);
    SFM_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return SFM_OK;
}
