// orb_select.cu -- KeyPointsFilter::retainBest for the GPU ORB detector (host code; no device work).
//
// cv2.ORB keeps, per pyramid level, the keypoints whose response reaches the n-th largest one (ties included) with
//     std::nth_element(first, first + n - 1, last, greater response);  threshold = (first + n - 1)->response;
//     new_end = std::partition(first + n, last, response >= threshold)
// and the ORDER those two algorithms leave the survivors in becomes the order of the descriptors, i.e. the queryIdx / trainIdx
// of the reference's match list (code/feature_matching.py:50-58).  They only ever compare responses, so running libstdc++'s
// own nth_element / partition on (response, index) records reproduces cv2's permutation (pinned by tests/test_gpu_orb.py
// against cv2.ORB's keypoint order, and against oracle/stl_select.cpp, which keeps records of cv::KeyPoint's 28 bytes): every
// decision inside introselect / partition -- pivot choice, the switch to insertion sort at 3 elements, the heap fallback --
// counts ELEMENTS and compares responses, never bytes, so 8-byte records take the same path with a third of the memory traffic
// (the FAST stage of a textured 1080p image hands ~10^5 keypoints of level 0 to this function).
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {
struct Rec {
    float response;
    int32_t index;
};
}  // namespace

extern "C" int sfm_orb_retain_best(const float* response_host, int n, int n_points, int32_t* out_index)
{
    SFM_REQUIRE((response_host || n == 0) && out_index && n >= 0, "sfm_orb_retain_best: bad argument");
    static thread_local std::vector<Rec> v;                  // one call per pyramid level and image: keep the storage
    v.resize((size_t)n);
    for (int i = 0; i < n; ++i) {
        v[i].response = response_host[i];
        v[i].index = i;
    }
    if (n_points >= 0 && n > n_points) {
        if (n_points == 0) return 0;
        std::nth_element(v.begin(), v.begin() + n_points - 1, v.end(), [](const Rec& a, const Rec& b) { return a.response > b.response; });
        const float ambiguous = v[n_points - 1].response;
        auto new_end = std::partition(v.begin() + n_points, v.end(), [ambiguous](const Rec& a) { return a.response >= ambiguous; });
        v.resize((size_t)(new_end - v.begin()));
    }
    for (size_t i = 0; i < v.size(); ++i) out_index[i] = v[i].index;
    return (int)v.size();
}
