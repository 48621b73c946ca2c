// stl_select.cpp -- KeyPointsFilter::retainBest as cv2 runs it inside ORB (TEST INFRASTRUCTURE; built into oracle/_build).
//
// cv2.ORB keeps, per pyramid level, the keypoints whose response reaches the n-th largest one (ties included):
//     std::nth_element(first, first + n - 1, last, response greater);  threshold = (first + n - 1)->response;
//     new_end = std::partition(first + n, last, response >= threshold)
// The SET is well defined; the ORDER the survivors are left in is whatever libstdc++'s introselect and partition produce, and
// that order becomes the order of the descriptors (queryIdx / trainIdx of code/feature_matching.py:50-58).  The algorithms
// only ever compare responses, so running them on (response, original index) records reproduces cv2's permutation.
#include <algorithm>
#include <cstdint>
#include <vector>

struct Rec {
    float response;
    int32_t index;
    int32_t pad[5];      // same size as cv::KeyPoint (28 bytes): insertion-sort thresholds inside libstdc++ count elements, not bytes,
};                       // but keeping the size equal rules the question out

extern "C" int sfm_oracle_retain_best(const float* response, int n, int n_points, int32_t* out_index)
{
    std::vector<Rec> v((size_t)n);
    for (int i = 0; i < n; ++i) { v[i].response = response[i]; v[i].index = i; }
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
