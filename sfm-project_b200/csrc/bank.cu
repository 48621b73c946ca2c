// bank.cu -- descriptor bank (K1): pack uint8 descriptors once, keep them resident in HBM.
//
// Replaces the implicit hand-off "orb.detectAndCompute output -> bf.match input" of the reference
// (code/feature_matching.py:44-50).  Layout and the K-extension encoding are described in DESIGN.md.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace sfm {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches += n; }

// One warp per bank row.
//   L2:      s8 = u8 ^ 0x80, norm = sum s8^2, ext bytes encode v = H0 - (norm >> 1) as
//            255 * sum_{k<24} e_k + e_24 (all u8), laid out as the no-swizzle K-major UMMA tile
//            [k-chunk 0..1][row 0..127][16 B] per 128 rows.
//   Hamming: raw 32-byte copy.
__global__ void __launch_bounds__(256) bank_pack_kernel(
    const uint8_t* __restrict__ src, int src_stride, const int32_t* __restrict__ counts,
    const float* __restrict__ src_xy, int first_image, int n_images, int feat_stride, int metric,
    int8_t* __restrict__ desc, int8_t* __restrict__ ext, int32_t* __restrict__ norm,
    float* __restrict__ xy, int32_t* __restrict__ count)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long total = (long long)n_images * feat_stride;
    if (warp >= total) return;
    const int img = (int)(warp / feat_stride);
    const int r = (int)(warp % feat_stride);
    int n = counts ? counts[img] : src_stride;
    n = min(min(n, src_stride), feat_stride);
    if (n < 0) n = 0;
    const long long grow = (long long)(first_image + img) * feat_stride + r;   // bank row
    const bool valid = r < n;
    if (r == 0 && lane == 0) count[first_image + img] = n;

    if (metric == SFM_METRIC_L2) {
        uint32_t w = 0;
        if (valid) w = reinterpret_cast<const uint32_t*>(src + ((long long)img * src_stride + r) * kDescDim)[lane] ^ 0x80808080u;
        reinterpret_cast<uint32_t*>(desc + grow * kDescDim)[lane] = w;
        int ss = __dp4a((int)w, (int)w, 0);
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        int e = 0;
        if (valid) {
            const int v = kExtOffset - (ss >> 1);
            const int hi = v / 255, rem = v - hi * 255;
            if (lane < kExtHi) e = max(0, min(255, hi - 255 * lane));
            else if (lane == kExtHi) e = rem;
        }
        const long long tile = grow / kTileRows;
        const int rin = (int)(grow % kTileRows);
        ext[tile * kExtTileBytes + (lane >> 4) * (kTileRows * 16) + rin * 16 + (lane & 15)] = (int8_t)e;
        if (lane == 0) norm[grow] = valid ? ss : 0;
    } else {
        if (lane < kHammingDim / 4) {
            uint32_t w = 0;
            if (valid) w = reinterpret_cast<const uint32_t*>(src + ((long long)img * src_stride + r) * kHammingDim)[lane];
            reinterpret_cast<uint32_t*>(desc + grow * kHammingDim)[lane] = w;
        }
        if (lane == 0) norm[grow] = 0;
    }
    if (lane < 2) {
        float v = 0.f;
        if (valid && src_xy) v = src_xy[((long long)img * src_stride + r) * 2 + lane];
        xy[grow * 2 + lane] = v;
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_desc_tmap(sfm_bank* b)
{
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    SFM_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return SFM_ERR_CUDA;
    }
    const cuuint64_t rows = (cuuint64_t)b->max_images * (cuuint64_t)b->L.feat_stride;
    cuuint64_t dims[2] = {(cuuint64_t)kDescDim, rows};
    cuuint64_t strides[1] = {(cuuint64_t)kDescDim};
    cuuint32_t box[2] = {(cuuint32_t)kDescDim, (cuuint32_t)kTileRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((PFN_encodeTiled)fn)(&b->tmap_desc, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)b->desc, dims, strides,
                                       box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return SFM_ERR_CUDA;
    }
    b->tmap_ready = true;
    return SFM_OK;
}

}  // namespace sfm

using namespace sfm;

extern "C" {

const char* sfm_last_error(void) { return g_err; }
int sfm_abi_version(void) { return SFM_B200_ABI_VERSION; }
int64_t sfm_launch_count(void) { return (int64_t)g_launches.load(); }

int sfm_device_info(int device, int32_t out[4])
{
    SFM_REQUIRE(out != nullptr, "sfm_device_info: out is NULL");
    cudaDeviceProp p;
    SFM_CUDA_CHECK(cudaGetDeviceProperties(&p, device));
    out[0] = p.multiProcessorCount;
    out[1] = p.major;
    out[2] = p.minor;
    out[3] = (int32_t)p.sharedMemPerBlockOptin;
    return SFM_OK;
}

int sfm_bank_storage_bytes(int max_images, int max_feats, int metric, size_t* out_bytes)
{
    SFM_REQUIRE(out_bytes != nullptr, "sfm_bank_storage_bytes: out is NULL");
    SFM_REQUIRE(max_images > 0 && max_feats > 0, "sfm_bank_storage_bytes: sizes must be positive");
    SFM_REQUIRE(metric == SFM_METRIC_L2 || metric == SFM_METRIC_HAMMING, "unknown metric %d", metric);
    SFM_REQUIRE((int64_t)max_images * align_up(max_feats, kFeatAlign) < (1ll << 31), "bank too large: rows must fit int32");
    *out_bytes = (size_t)bank_layout(max_images, max_feats, metric).total;
    return SFM_OK;
}

int sfm_bank_create(int device, int max_images, int max_feats, int metric, void* storage, size_t storage_bytes,
                    sfm_bank_t** out)
{
    SFM_REQUIRE(out != nullptr && storage != nullptr, "sfm_bank_create: NULL argument");
    size_t need = 0;
    int rc = sfm_bank_storage_bytes(max_images, max_feats, metric, &need);
    if (rc) return rc;
    SFM_REQUIRE(storage_bytes >= need, "sfm_bank_create: storage too small (%zu < %zu)", storage_bytes, need);
    SFM_REQUIRE(((uintptr_t)storage & 1023) == 0, "sfm_bank_create: storage must be 1024-byte aligned");
    cudaDeviceProp p;
    SFM_CUDA_CHECK(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) {
        set_error("device %d is sm_%d%d; this library contains sm_100a code only (no fallback)", device, p.major, p.minor);
        return SFM_ERR_DEVICE;
    }
    SFM_ON_DEVICE(device);                                   // the caller's current device is restored on return
    sfm_bank* b = new sfm_bank();
    memset(b, 0, sizeof *b);
    b->device = device;
    b->max_images = max_images;
    b->max_feats = max_feats;
    b->metric = metric;
    b->L = bank_layout(max_images, max_feats, metric);
    b->base = (uint8_t*)storage;
    b->desc = (int8_t*)(b->base + b->L.off_desc);
    b->ext = (int8_t*)(b->base + b->L.off_ext);
    b->norm = (int32_t*)(b->base + b->L.off_norm);
    b->xy = (float*)(b->base + b->L.off_xy);
    b->count = (int32_t*)(b->base + b->L.off_count);
    b->sm_count = p.multiProcessorCount;
    if (metric == SFM_METRIC_L2) {
        rc = make_desc_tmap(b);
        if (rc) { delete b; return rc; }
    }
    *out = b;
    return SFM_OK;
}

int sfm_bank_destroy(sfm_bank_t* bank)
{
    delete bank;
    return SFM_OK;
}

int sfm_bank_layout(const sfm_bank_t* bank, int64_t out[6])
{
    SFM_REQUIRE(bank && out, "sfm_bank_layout: NULL argument");
    out[0] = bank->L.feat_stride;
    out[1] = bank->L.off_desc;
    out[2] = bank->L.off_ext;
    out[3] = bank->L.off_norm;
    out[4] = bank->L.off_xy;
    out[5] = bank->L.off_count;
    return SFM_OK;
}

int sfm_bank_put_batch(sfm_bank_t* bank, int first_image, int n_images, const uint8_t* desc_u8, int src_stride,
                       const int32_t* counts, const float* xy, void* stream)
{
    SFM_REQUIRE(bank && desc_u8, "sfm_bank_put_batch: NULL argument");
    SFM_REQUIRE(n_images > 0 && first_image >= 0 && first_image + n_images <= bank->max_images,
                "sfm_bank_put_batch: images [%d,%d) outside bank of %d", first_image, first_image + n_images, bank->max_images);
    SFM_REQUIRE(src_stride > 0 && src_stride <= bank->L.feat_stride, "sfm_bank_put_batch: src_stride %d exceeds feat_stride %lld",
                src_stride, (long long)bank->L.feat_stride);
    SFM_REQUIRE(((uintptr_t)desc_u8 & 3) == 0, "sfm_bank_put_batch: descriptors must be 4-byte aligned");
    SFM_ON_DEVICE(bank->device);
    const long long warps = (long long)n_images * bank->L.feat_stride;
    const int block = 256;
    const long long grid = (warps * 32 + block - 1) / block;
    bank_pack_kernel<<<(unsigned)grid, block, 0, (cudaStream_t)stream>>>(
        desc_u8, src_stride, counts, xy, first_image, n_images, (int)bank->L.feat_stride, bank->metric, bank->desc,
        bank->ext, bank->norm, bank->xy, bank->count);
    SFM_CUDA_CHECK(cudaGetLastError());
    count_launch();
    if (first_image + n_images > bank->n_filled) bank->n_filled = first_image + n_images;
    return SFM_OK;
}

int sfm_bank_mark_filled(sfm_bank_t* bank, int n_images)
{
    SFM_REQUIRE(bank && n_images >= 0 && n_images <= bank->max_images, "sfm_bank_mark_filled: bad argument");
    bank->n_filled = n_images;
    return SFM_OK;
}

}  // extern "C"
