// common.cuh -- shared host-side plumbing of libsfm_b200 (error text, launch counter, bank handle).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "../../include/sfm_b200.h"

namespace sfm {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define SFM_CUDA_CHECK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            sfm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                           __LINE__);                                                          \
            return SFM_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define SFM_REQUIRE(cond, ...)                                                                 \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            sfm::set_error(__VA_ARGS__);                                                       \
            return SFM_ERR_ARG;                                                                \
        }                                                                                      \
    } while (0)

// Entry points that take a bank run on the bank's device whatever the caller's current device is, and leave the
// caller's (and torch's) current device as they found it.
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) {
            err = cudaSetDevice(device);
            changed = (err == cudaSuccess);
        }
    }
    ~DeviceGuard()
    {
        if (changed) cudaSetDevice(prev);
    }
};
#define SFM_ON_DEVICE(dev)                                                                     \
    sfm::DeviceGuard _guard(dev);                                                              \
    SFM_CUDA_CHECK(_guard.err)

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device); the "already raised to" table of each launch
// site is shared by every host thread of the process, hence the lock (uncontended in normal use).
struct SmemAttrTable {
    std::mutex mu;
    size_t raised[64] = {};
};
template <typename Kernel>
inline cudaError_t ensure_dyn_smem(Kernel kernel, size_t bytes, int device, SmemAttrTable& t)
{
    std::lock_guard<std::mutex> lock(t.mu);
    size_t& have = t.raised[device & 63];
    if (bytes <= have) return cudaSuccess;           // (no 48 KB shortcut: static shared memory counts against the default limit)
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}
inline int current_device()
{
    int d = 0;
    cudaGetDevice(&d);
    return d;
}

constexpr int kDescDim = 128;        // bytes per L2 descriptor row
constexpr int kHammingDim = 32;      // bytes per binary descriptor row
constexpr int kTileRows = 128;       // rows per MMA operand tile
constexpr int kFeatAlign = 256;      // feat_stride granularity (one matcher unit = 256 query rows)
constexpr int kExtK = 32;            // bytes of K-extension per row
constexpr int kExtTileBytes = 2 * kTileRows * 16;   // [2][128][16]
constexpr int kExtHi = 24;           // slots weighted 255
constexpr int kExtOffset = 1 << 20;  // H0: ext encodes H0 - floor(|b_s|^2 / 2) >= 0

struct BankLayout {
    int64_t feat_stride;
    int64_t off_desc, off_ext, off_norm, off_xy, off_count, total;
};

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

inline BankLayout bank_layout(int max_images, int max_feats, int metric)
{
    BankLayout L;
    L.feat_stride = align_up(max_feats, kFeatAlign);
    int64_t rows = (int64_t)max_images * L.feat_stride;
    int64_t row_bytes = (metric == SFM_METRIC_L2) ? kDescDim : kHammingDim;
    int64_t o = 0;
    L.off_desc = o;  o = align_up(o + rows * row_bytes, 1024);
    L.off_ext = o;   o = align_up(o + (metric == SFM_METRIC_L2 ? rows / kTileRows * kExtTileBytes : 0), 1024);
    L.off_norm = o;  o = align_up(o + rows * 4, 1024);
    L.off_xy = o;    o = align_up(o + rows * 8, 1024);
    L.off_count = o; o = align_up(o + (int64_t)max_images * 4, 1024);
    L.total = o;
    return L;
}

}  // namespace sfm

struct sfm_bank {
    int device;
    int max_images, max_feats, metric;
    int n_filled;
    sfm::BankLayout L;
    uint8_t* base;
    int8_t* desc;
    int8_t* ext;
    int32_t* norm;
    float* xy;
    int32_t* count;
    int sm_count;
    bool tmap_ready;
    alignas(64) CUtensorMap tmap_desc;   // 2-D [rows][128 B], box 128x128, SWIZZLE_128B
};
