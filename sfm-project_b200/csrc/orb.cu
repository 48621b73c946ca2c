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

}  // extern "C"
