// orb.cu -- ORB's descriptor stage on the GPU (stage 1).
//
// Replaces, for keypoints that cv2 detected, the descriptor half of `orb.detectAndCompute(gray, None)`
// (code/feature_matching.py:42-45): the image pyramid (resize with INTER_LINEAR_EXACT, each level from the previous one),
// the 7x7 sigma-2 Gaussian blur of every level, and the 256 rotated rBRIEF intensity tests per keypoint.  Every stage is
// bit-exact against cv2 4.13 (oracle/orb_oracle.py restates the arithmetic; tests pin it to cv2.resize, cv2.sepFilter2D and
// cv2.ORB itself):
//   resize : 8.8 fixed-point coefficients (tables from the host), horizontal then vertical, (v + 2^15) >> 16
//   blur   : float32, rows  s = k[-3] x[-3]; s = fma(k[d], x[d], s), d = -2..3; columns  s = k[0] x[0];
//            s = fma(k[d], x[d] + x[-d], s), d = 1..3; round-half-even -- the association of cv2's sepFilter2D on a host with FMA
//   tests  : sample i at (round(x a - y b), round(x b + y a)) around the keypoint's rounded level position, bit = I(a_i) < I(b_i)
// The sampling pattern is measured from cv2 by tools/recover_orb_pattern.py (orb_pattern.inc).  Built with -fmad=false: every
// fused operation below is explicit.
#include "common.cuh"

namespace sfm {

__constant__ int8_t c_orb_pattern[256][4] = {
#include "orb_pattern.inc"
};

// cv2.getGaussianKernel(7, 2, CV_32F): k[0..3] = taps at distance 3, 2, 1, 0 (bit patterns; tests compare with cv2)
__device__ __forceinline__ float orb_tap(int d)
{
    const uint32_t bits[4] = {0x3d8fafb1u, 0x3e06387eu, 0x3e434a39u, 0x3e5d4ae0u};
    return __uint_as_float(bits[3 - (d < 0 ? -d : d)]);
}

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

// ---- INTER_LINEAR_EXACT: one thread per destination pixel; tab = (source index, weight of the NEXT source pixel in 1/256)
__global__ void __launch_bounds__(256) orb_resize_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch, uint8_t* __restrict__ dst,
                                                         int dw, int dh, int dpitch, const int2* __restrict__ xtab, const int2* __restrict__ ytab)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dw || y >= dh) return;
    const int2 tx = __ldg(xtab + x), ty = __ldg(ytab + y);
    const int x0 = tx.x, x1 = min(tx.x + 1, sw - 1), y0 = ty.x, y1 = min(ty.x + 1, sh - 1);
    const uint8_t* r0 = src + (size_t)y0 * spitch;
    const uint8_t* r1 = src + (size_t)y1 * spitch;
    const int h0 = (256 - tx.y) * r0[x0] + tx.y * r0[x1];
    const int h1 = (256 - tx.y) * r1[x0] + tx.y * r1[x1];
    const int v = (256 - ty.y) * h0 + ty.y * h1;
    dst[(size_t)y * dpitch + x] = (uint8_t)((v + (1 << 15)) >> 16);
}

// ---- GaussianBlur(7 x 7, sigma 2, BORDER_REFLECT_101): 32 x 16 output tile per block, rows then columns through shared memory
constexpr int kBlurW = 32, kBlurH = 16;
__global__ void __launch_bounds__(kBlurW* kBlurH) orb_blur_kernel(const uint8_t* __restrict__ src, int w, int h, int spitch,
                                                                   uint8_t* __restrict__ dst, int dpitch)
{
    __shared__ uint8_t tile[kBlurH + 6][kBlurW + 6 + 2];
    __shared__ float rows[kBlurH + 6][kBlurW];
    const int tx = threadIdx.x & (kBlurW - 1), ty = threadIdx.x / kBlurW;
    const int x0 = blockIdx.x * kBlurW, y0 = blockIdx.y * kBlurH;
    for (int e = threadIdx.x; e < (kBlurH + 6) * (kBlurW + 6); e += kBlurW * kBlurH) {
        const int r = e / (kBlurW + 6), c = e - r * (kBlurW + 6);
        tile[r][c] = src[(size_t)reflect101(y0 + r - 3, h) * spitch + reflect101(x0 + c - 3, w)];
    }
    __syncthreads();
    for (int r = ty; r < kBlurH + 6; r += kBlurH) {
        float s = __fmul_rn(orb_tap(-3), (float)tile[r][tx]);
#pragma unroll
        for (int d = -2; d <= 3; ++d) s = __fmaf_rn(orb_tap(d), (float)tile[r][tx + 3 + d], s);
        rows[r][tx] = s;
    }
    __syncthreads();
    const int x = x0 + tx, y = y0 + ty;
    if (x < w && y < h) {
        float s = __fmul_rn(orb_tap(0), rows[ty + 3][tx]);
#pragma unroll
        for (int d = 1; d <= 3; ++d) s = __fmaf_rn(orb_tap(d), __fadd_rn(rows[ty + 3 + d][tx], rows[ty + 3 - d][tx]), s);
        const int v = __float2int_rn(s);
        dst[(size_t)y * dpitch + x] = (uint8_t)min(max(v, 0), 255);
    }
}

struct OrbLevels {
    const uint8_t* ptr[16];
    int pitch[16];
};

// ---- rBRIEF: one warp per keypoint, lane = descriptor byte.  kp = (cx, cy, level, 0) integers, rot = (cos, sin) float32
__global__ void __launch_bounds__(256) orb_describe_kernel(const OrbLevels L, const int4* __restrict__ kp, const float2* __restrict__ rot,
                                                           int n, uint8_t* __restrict__ out, int out_stride)
{
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= n) return;
    const int4 c = __ldg(kp + k);
    const float2 ab = __ldg(rot + k);
    const uint8_t* img = L.ptr[c.z] + (size_t)c.y * L.pitch[c.z] + c.x;
    const int pitch = L.pitch[c.z];
    auto sample = [&](int px, int py) {
        const float x = __fsub_rn(__fmul_rn((float)px, ab.x), __fmul_rn((float)py, ab.y));
        const float y = __fadd_rn(__fmul_rn((float)px, ab.y), __fmul_rn((float)py, ab.x));
        return (int)__ldg(img + (ptrdiff_t)__float2int_rn(y) * pitch + __float2int_rn(x));
    };
    unsigned byte = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int i = lane * 8 + j;
        const int va = sample(c_orb_pattern[i][0], c_orb_pattern[i][1]);
        const int vb = sample(c_orb_pattern[i][2], c_orb_pattern[i][3]);
        byte |= (unsigned)(va < vb) << j;
    }
    out[(size_t)k * out_stride + lane] = (uint8_t)byte;
}

// ================================================================================================ detection (stage 2)
// cv2.ORB's keypoint half, per pyramid level: FAST-9/16 (threshold 20) with its corner score, 3 x 3 non-maximum suppression,
// the 31-pixel border rule, row-major compaction; then, for the candidates the host's retainBest keeps, the Harris response and
// the intensity-centroid orientation.  All integer except the two float32 expressions, whose operations are explicit.

// score = (largest t for which the pixel is a FAST corner) = max over the 16 arcs of 9 contiguous circle pixels of the arc's
// smallest |difference| of one sign, minus 1; 0 below the detection threshold.
__global__ void __launch_bounds__(256) orb_fast_score_kernel(const uint8_t* __restrict__ img, int w, int h, int pitch,
                                                             uint8_t* __restrict__ score, int spitch, int threshold)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    int out = 0;
    if (x >= 3 && y >= 3 && x < w - 3 && y < h - 3) {
        const uint8_t* p = img + (size_t)y * pitch + x;
        const int v = p[0];
        const int o1 = pitch, o2 = 2 * pitch, o3 = 3 * pitch;
        int d[16];
        d[0] = v - p[-o3];      d[1] = v - p[-o3 + 1];  d[2] = v - p[-o2 + 2];  d[3] = v - p[-o1 + 3];
        d[4] = v - p[3];        d[5] = v - p[o1 + 3];   d[6] = v - p[o2 + 2];   d[7] = v - p[o3 + 1];
        d[8] = v - p[o3];       d[9] = v - p[o3 - 1];   d[10] = v - p[o2 - 2];  d[11] = v - p[o1 - 3];
        d[12] = v - p[-3];      d[13] = v - p[-o1 - 3]; d[14] = v - p[-o2 - 2]; d[15] = v - p[-o3 - 1];
        // sliding minimum / maximum over 9 consecutive entries of the circular sequence: 2, 4, 8, then one more
        int lo2[16], hi2[16], lo4[16], hi4[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { lo2[k] = min(d[k], d[(k + 1) & 15]); hi2[k] = max(d[k], d[(k + 1) & 15]); }
#pragma unroll
        for (int k = 0; k < 16; ++k) { lo4[k] = min(lo2[k], lo2[(k + 2) & 15]); hi4[k] = max(hi2[k], hi2[(k + 2) & 15]); }
        int best_pos = -1000, best_neg = 1000;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int lo9 = min(min(lo4[k], lo4[(k + 4) & 15]), d[(k + 8) & 15]);
            const int hi9 = max(max(hi4[k], hi4[(k + 4) & 15]), d[(k + 8) & 15]);
            best_pos = max(best_pos, lo9);                        // all nine darker than the centre by at least lo9
            best_neg = min(best_neg, hi9);                        // all nine brighter by at least -hi9
        }
        const int sc = max(best_pos, -best_neg) - 1;
        out = sc >= threshold ? sc : 0;
    }
    score[(size_t)y * spitch + x] = (uint8_t)out;
}

__device__ __forceinline__ bool orb_is_peak(const uint8_t* __restrict__ s, int pitch, int x, int y)
{
    const int v = s[(size_t)y * pitch + x];
    if (v == 0) return false;
    const uint8_t* a = s + (size_t)(y - 1) * pitch + x;
    const uint8_t* b = a + pitch;
    const uint8_t* c = b + pitch;
    const int m = max(max(max((int)a[-1], (int)a[0]), max((int)a[1], (int)b[-1])), max(max((int)b[1], (int)c[-1]), max((int)c[0], (int)c[1])));
    return v > m;
}

// one block per row of [border, h - border): kWrite == false counts the row's keypoints; kWrite == true writes them in x order at
// (number of keypoints in earlier rows) + rank
template <bool kWrite>
__global__ void __launch_bounds__(256) orb_nms_kernel(const uint8_t* __restrict__ score, int w, int h, int pitch, int border,
                                                      int32_t* __restrict__ row_count, int32_t* __restrict__ total, int32_t* __restrict__ out_xy,
                                                      float* __restrict__ out_resp)
{
    __shared__ int wsum[8];
    __shared__ int base_s;
    const int y = border + blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int base = 0;
    if (kWrite) {
        int v = 0;
        for (int r = threadIdx.x; r < (int)blockIdx.x; r += 256) v += row_count[r];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) wsum[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int k = 0; k < 8; ++k) t += wsum[k];
            base_s = t;
        }
        __syncthreads();
        base = base_s;
        __syncthreads();
    }
    int running = 0;
    for (int x0 = border; x0 < w - border; x0 += 256) {
        const int x = x0 + threadIdx.x;
        const bool keep = x < w - border && orb_is_peak(score, pitch, x, y);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int before = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k < warp) before += wsum[k];
            tot += wsum[k];
        }
        if (kWrite && keep) {
            const int o = base + running + before + __popc(bal & ((1u << lane) - 1u));
            out_xy[2 * o] = x;
            out_xy[2 * o + 1] = y;
            out_resp[o] = (float)score[(size_t)y * pitch + x];
        }
        running += tot;
        __syncthreads();
    }
    if (!kWrite && threadIdx.x == 0) {
        row_count[blockIdx.x] = running;
        if (running) atomicAdd(total, running);
    }
}

__constant__ int c_orb_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// cv::fastAtan2 (degrees): 7th-order odd polynomial on min/max, float32, every operation rounded on its own
__device__ __forceinline__ float orb_fast_atan2(float y, float x)
{
    const float p1 = __fmul_rn(0.9997878412794807f, 57.29577951308232f), p3 = __fmul_rn(-0.3258083974640975f, 57.29577951308232f);
    const float p5 = __fmul_rn(0.1555786518463281f, 57.29577951308232f), p7 = __fmul_rn(-0.04432655554792128f, 57.29577951308232f);
    const float eps = 2.220446049250313e-16f;
    const float ax = fabsf(x), ay = fabsf(y);
    float a;
    if (ax >= ay) {
        const float c = __fdiv_rn(ay, __fadd_rn(ax, eps)), c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        const float c = __fdiv_rn(ax, __fadd_rn(ay, eps)), c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

// one warp per selected candidate: Harris response of the 7 x 7 block and orientation of the radius-15 disc, on the UNBLURRED level
__global__ void __launch_bounds__(256) orb_harris_angle_kernel(const uint8_t* __restrict__ img, int pitch, const int32_t* __restrict__ xy,
                                                               const int32_t* __restrict__ sel, int n, int32_t* __restrict__ out_xy,
                                                               float* __restrict__ out_resp, float* __restrict__ out_angle)
{
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= n) return;
    const int src = sel ? sel[k] : k;
    const int x0 = xy[2 * src], y0 = xy[2 * src + 1];
    const uint8_t* c0 = img + (size_t)y0 * pitch + x0;
    // ---- Harris: 49 pixels over the lanes
    long long a = 0, b = 0, c = 0;
    for (int e = lane; e < 49; e += 32) {
        const uint8_t* p = c0 + (e / 7 - 3) * pitch + (e % 7 - 3);
        const int Ix = ((int)p[1] - (int)p[-1]) * 2 + ((int)p[-pitch + 1] - (int)p[-pitch - 1]) + ((int)p[pitch + 1] - (int)p[pitch - 1]);
        const int Iy = ((int)p[pitch] - (int)p[-pitch]) * 2 + ((int)p[pitch - 1] - (int)p[-pitch - 1]) + ((int)p[pitch + 1] - (int)p[-pitch + 1]);
        a += Ix * Ix;
        b += Iy * Iy;
        c += Ix * Iy;
    }
    // ---- moments: rows v = -15..15 over the lanes (31 rows), each lane sums its row
    int m01 = 0, m10 = 0;
    if (lane < 31) {
        const int v = lane - 15, d = c_orb_umax[v < 0 ? -v : v];
        const uint8_t* row = c0 + v * pitch;
        int s = 0, su = 0;
        for (int u = -d; u <= d; ++u) {
            const int val = row[u];
            s += val;
            su += u * val;
        }
        m01 = v * s;
        m10 = su;
    }
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    }
    if (lane == 0) {
        const float fa = (float)(int)a, fb = (float)(int)b, fc = (float)(int)c;
        const float scale = __fdiv_rn(1.f, __fmul_rn(28.f, 255.f));
        const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
        const float t = __fadd_rn(fa, fb);
        const float r = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(0.04f, t), t)), s4);
        out_xy[2 * k] = x0;
        out_xy[2 * k + 1] = y0;
        out_resp[k] = r;
        out_angle[k] = orb_fast_atan2((float)m01, (float)m10);
    }
}

}  // namespace sfm

using namespace sfm;

extern "C" {

int sfm_orb_resize(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dw, int dh, int dpitch, const int32_t* xtab,
                   const int32_t* ytab, void* stream)
{
    SFM_REQUIRE(src && dst && xtab && ytab, "sfm_orb_resize: NULL argument");
    SFM_REQUIRE(sw > 0 && sh > 0 && dw > 0 && dh > 0 && spitch >= sw && dpitch >= dw, "sfm_orb_resize: bad sizes");
    SFM_REQUIRE(((uintptr_t)xtab & 7) == 0 && ((uintptr_t)ytab & 7) == 0, "sfm_orb_resize: tables must be 8-byte aligned");
    dim3 grid((unsigned)((dw + 31) / 32), (unsigned)((dh + 7) / 8));
    orb_resize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, sw, sh, spitch, dst, dw, dh, dpitch, (const int2*)xtab, (const int2*)ytab);
    SFM_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return SFM_OK;
}

int sfm_orb_blur(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch, void* stream)
{
    SFM_REQUIRE(src && dst && src != dst, "sfm_orb_blur: NULL or aliased argument");
    SFM_REQUIRE(w > 0 && h > 0 && spitch >= w && dpitch >= w, "sfm_orb_blur: bad sizes");
    dim3 grid((unsigned)((w + kBlurW - 1) / kBlurW), (unsigned)((h + kBlurH - 1) / kBlurH));
    orb_blur_kernel<<<grid, kBlurW * kBlurH, 0, (cudaStream_t)stream>>>(src, w, h, spitch, dst, dpitch);
    SFM_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return SFM_OK;
}

int sfm_orb_describe(const uint8_t* const* level_ptr, const int32_t* level_pitch, int n_levels, const int32_t* kp, const float* rot,
                     int n_keypoints, uint8_t* out_desc, int out_stride, void* stream)
{
    SFM_REQUIRE(level_ptr && level_pitch && kp && rot && out_desc, "sfm_orb_describe: NULL argument");
    SFM_REQUIRE(n_levels > 0 && n_levels <= 16 && n_keypoints >= 0 && out_stride >= 32, "sfm_orb_describe: bad sizes");
    SFM_REQUIRE(((uintptr_t)kp & 15) == 0 && ((uintptr_t)rot & 7) == 0, "sfm_orb_describe: keypoint arrays must be 16 / 8-byte aligned");
    if (n_keypoints == 0) return SFM_OK;
    OrbLevels L;
    for (int i = 0; i < 16; ++i) {
        L.ptr[i] = i < n_levels ? level_ptr[i] : nullptr;
        L.pitch[i] = i < n_levels ? level_pitch[i] : 0;
    }
    const int blocks = (n_keypoints * 32 + 255) / 256;
    orb_describe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(L, (const int4*)kp, (const float2*)rot, n_keypoints, out_desc, out_stride);
    SFM_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return SFM_OK;
}

int sfm_orb_fast_detect(const uint8_t* img, int w, int h, int pitch, int threshold, int border, uint8_t* score, int32_t* row_count,
                        int32_t* total, int32_t* out_xy, float* out_resp, void* stream)
{
    SFM_REQUIRE(img && score && row_count && total && out_xy && out_resp, "sfm_orb_fast_detect: NULL argument");
    SFM_REQUIRE(w > 0 && h > 0 && pitch >= w && threshold > 0 && threshold < 255 && border >= 4, "sfm_orb_fast_detect: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    SFM_CUDA_CHECK(cudaMemsetAsync(total, 0, sizeof(int32_t), st));
    if (w <= 2 * border || h <= 2 * border) return SFM_OK;      // no pixel is far enough from the border: no keypoints
    dim3 grid((unsigned)((w + 31) / 32), (unsigned)((h + 7) / 8));
    orb_fast_score_kernel<<<grid, 256, 0, st>>>(img, w, h, pitch, score, w, threshold);
    const int rows = h - 2 * border;
    orb_nms_kernel<false><<<rows, 256, 0, st>>>(score, w, h, w, border, row_count, total, nullptr, nullptr);
    orb_nms_kernel<true><<<rows, 256, 0, st>>>(score, w, h, w, border, row_count, total, out_xy, out_resp);
    SFM_CUDA_CHECK(cudaGetLastError());
    count_launch(3);
    return SFM_OK;
}

int sfm_orb_harris_angle(const uint8_t* img, int w, int h, int pitch, const int32_t* xy, const int32_t* sel, int n, int32_t* out_xy,
                         float* out_resp, float* out_angle, void* stream)
{
    SFM_REQUIRE(img && xy && out_xy && out_resp && out_angle, "sfm_orb_harris_angle: NULL argument");
    SFM_REQUIRE(w > 0 && h > 0 && pitch >= w && n >= 0, "sfm_orb_harris_angle: bad sizes");
    if (n == 0) return SFM_OK;
    orb_harris_angle_kernel<<<(n * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(img, pitch, xy, sel, n, out_xy, out_resp, out_angle);
    SFM_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return SFM_OK;
}

}  // extern "C"
