// match_api.cu -- C-ABI entry points of the L2 matcher (dispatch between the tcgen05 and SIMT kernels).
#include "match_common.cuh"

namespace sfm {
int launch_match_simt(const sfm_bank* b, const int32_t* pairs, int n_pairs, int32_t* knn_out, cudaStream_t st);
int launch_match_tc(const sfm_bank* b, const int32_t* pairs, int n_pairs, int grid_req, int32_t* knn_out, int32_t* dbg_acc,
                    int dbg_mode, const Prefilter& pf, cudaStream_t st);
}  // namespace sfm

using namespace sfm;

extern "C" {

int sfm_match_knn2(const sfm_bank_t* bank, const int32_t* pairs_dev, int n_pairs, const sfm_match_params* params,
                   int32_t* knn_out, void* stream)
{
    SFM_REQUIRE(bank && pairs_dev && knn_out, "sfm_match_knn2: NULL argument");
    SFM_REQUIRE(bank->metric == SFM_METRIC_L2, "sfm_match_knn2: bank metric is not L2");
    SFM_REQUIRE(n_pairs >= 0, "sfm_match_knn2: negative pair count");
    SFM_REQUIRE(((uintptr_t)knn_out & 15) == 0, "sfm_match_knn2: knn_out must be 16-byte aligned");
    if (bank->n_filled <= 0) {
        set_error("sfm_match_knn2: bank is empty");
        return SFM_ERR_STATE;
    }
    if (n_pairs == 0) return SFM_OK;
    SFM_ON_DEVICE(bank->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int impl = params ? params->impl : SFM_MATCH_AUTO;
    const int grid = params ? params->grid : 0;
    // rows that no unit owns (>= count of the query image) read as (-1,-1,-1,-1): the tcgen05 paths leave that to the refinement
    // kernel, which visits every row anyway; the SIMT kernel and the sweep-only diagnostics start from a filled buffer
    const bool refine_fills = impl != SFM_MATCH_SIMT && !(params && params->sweep_only);
    const bool timing_only = params && params->sweep_only >= 4 && impl != SFM_MATCH_SIMT;
    if (!refine_fills && !timing_only) SFM_CUDA_CHECK(cudaMemsetAsync(knn_out, 0xFF, (size_t)n_pairs * bank->L.feat_stride * 16, st));
    if (impl == SFM_MATCH_SIMT) return launch_match_simt(bank, pairs_dev, n_pairs, knn_out, st);
    SFM_REQUIRE(impl == SFM_MATCH_AUTO || impl == SFM_MATCH_TCGEN05, "unknown matcher impl %d", impl);
    const int dbg = (params && params->sweep_only) ? 3 : 0;      // diagnostics: leave the candidate records in knn_out
    Prefilter pf{SFM_RATIO_NONE, 1.0, 1, 1};
    if (params && params->prefilter_mode != SFM_RATIO_NONE) {
        SFM_REQUIRE(params->prefilter_mode == SFM_RATIO_CV2_F32 || params->prefilter_mode == SFM_RATIO_EXACT_INT,
                    "unknown prefilter mode %d", params->prefilter_mode);
        if (params->prefilter_mode == SFM_RATIO_EXACT_INT)
            SFM_REQUIRE(params->prefilter_num > 0 && params->prefilter_den > 0 && params->prefilter_num < 4096 &&
                            params->prefilter_den < 4096, "exact_int prefilter needs 0 < num, den < 4096");
        else
            SFM_REQUIRE(params->prefilter_ratio > 0.0, "prefilter ratio must be positive");
        pf.mode = params->prefilter_mode;
        pf.ratio = params->prefilter_ratio;
        pf.num2 = (long long)params->prefilter_num * params->prefilter_num;
        pf.den2 = (long long)params->prefilter_den * params->prefilter_den;
    }
    return launch_match_tc(bank, pairs_dev, n_pairs, grid, knn_out, nullptr, dbg, pf, st);
}

// Bring-up aid (tests only): run the tcgen05 kernel on ONE pair and dump the raw accumulators of its first
// unit / first train tile (256 x 128 int32).  mode 0 = main + K-extension, 1 = main only, 2 = extension only.
int sfm_debug_tc_tile(const sfm_bank_t* bank, const int32_t* pairs_dev, int mode, int32_t* knn_out, int32_t* acc_out,
                      void* stream)
{
    SFM_REQUIRE(bank && pairs_dev && knn_out && acc_out, "sfm_debug_tc_tile: NULL argument");
    SFM_REQUIRE(bank->metric == SFM_METRIC_L2, "sfm_debug_tc_tile: bank metric is not L2");
    SFM_ON_DEVICE(bank->device);
    cudaStream_t st = (cudaStream_t)stream;
    SFM_CUDA_CHECK(cudaMemsetAsync(knn_out, 0xFF, (size_t)(mode >= 4 ? (mode >> 8) : 1) * bank->L.feat_stride * 16, st));
    // mode 4 (timeline trace) may run a whole pair list on the full grid: n_pairs is passed in the high bits
    const int n_pairs = mode >= 4 ? (mode >> 8) : 1;
    return launch_match_tc(bank, pairs_dev, n_pairs > 0 ? n_pairs : 1, mode >= 4 ? 0 : 1, knn_out, acc_out, mode & 0xFF,
                           Prefilter{SFM_RATIO_NONE, 1.0, 1, 1}, st);
}

}  // extern "C"
