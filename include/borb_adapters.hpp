// borb_adapters.hpp — header-only C++ adapters that keep the reference's class signatures and forward
// to the C ABI in borb.h.  Include AFTER the OpenCV core headers (cv::Mat, cv::KeyPoint, cv::InputArray).
//
//   ORB_SLAM2::ORBextractor    replaces include/ORBextractor.h:45-111 / src/ORBextractor.cc
//   borb::ComputeStereoMatches replaces the body of Frame::ComputeStereoMatches (src/Frame.cc:466-640)
//
// Drop-in recipe: see INTEGRATION.md.  Errors: the reference's extractor never throws on its own;
// these adapters throw std::runtime_error only when the CUDA library reports a failure (there is no
// CPU fallback to fall back to), never call exit().
#pragma once
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "borb.h"

namespace borb {
inline void check(borb_status s, const char* where) {
    if (s != BORB_OK) throw std::runtime_error(std::string(where) + ": " + borb_status_str(s) + ": " + borb_last_error());
}
}  // namespace borb

#ifndef BORB_ADAPTER_NO_EXTRACTOR   /* borb_matcher_adapters.hpp alone: define it to skip the cv::InputArray-based extractor class */
namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device = 0)
        : nfeatures_(nfeatures), nlevels_(nlevels), scaleFactor_(scaleFactor) {
        static_assert(sizeof(cv::KeyPoint) == sizeof(borb_keypoint), "cv::KeyPoint must be the 28-byte POD layout");
        borb_extractor_cfg cfg = {nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST};
        borb::check(borb_extractor_create(&cfg, device, &h_), "borb_extractor_create");
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels);
        mvInvLevelSigma2.resize(nlevels); mnFeaturesPerLevel.resize(nlevels);
        borb::check(borb_extractor_tables(h_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                                          mvInvLevelSigma2.data(), mnFeaturesPerLevel.data()), "borb_extractor_tables");
        mvImagePyramid.resize(nlevels);
    }
    ~ORBextractor() { borb_extractor_destroy(h_); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image (mask is ignored, as in the reference).
    void operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints,
                    cv::OutputArray descriptors) {
        if (_image.empty()) return;                                   // ORBextractor.cc:1046-1047
        cv::Mat image = _image.getMat();
        int cap = 0;
        borb::check(borb_extractor_capacity(h_, image.cols, image.rows, &cap), "borb_extractor_capacity");
        keypoints.resize(cap);
        scratch_.resize((size_t)cap * 32);
        int n = 0;
        borb::check(borb_extract(h_, image.data, image.cols, image.rows, (int)image.step,
                                 reinterpret_cast<borb_keypoint*>(keypoints.data()), scratch_.data(), cap, &n), "borb_extract");
        keypoints.resize(n);
        if (n == 0) { descriptors.release(); return; }                // :1064-1065
        descriptors.create(n, 32, CV_8U);
        cv::Mat d = descriptors.getMat();
        for (int i = 0; i < n; i++) std::memcpy(d.ptr(i), scratch_.data() + (size_t)i * 32, 32);
        pyramid_valid_ = false;
    }

    // mvImagePyramid is device-resident; call this only if host code still reads it (the stereo
    // association replacement below does not).
    void SyncPyramid() {
        if (pyramid_valid_) return;
        for (int l = 0; l < nlevels_; l++) {
            int w = 0, h = 0;
            borb::check(borb_extractor_pyramid(h_, 0, l, nullptr, &w, &h), "borb_extractor_pyramid");
            mvImagePyramid[l].create(h, w, CV_8UC1);
            borb::check(borb_extractor_pyramid(h_, 0, l, mvImagePyramid[l].data, &w, &h), "borb_extractor_pyramid");
        }
        pyramid_valid_ = true;
    }

    int inline GetLevels() { return nlevels_; }
    float inline GetScaleFactor() { return scaleFactor_; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    std::vector<cv::Mat> mvImagePyramid;      // filled by SyncPyramid()
    borb_extractor* handle() const { return h_; }

protected:
    borb_extractor* h_ = nullptr;
    int nfeatures_, nlevels_;
    float scaleFactor_;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<unsigned char> scratch_;
    bool pyramid_valid_ = false;
};

}  // namespace ORB_SLAM2

namespace borb {
// Body of Frame::ComputeStereoMatches (src/Frame.cc:466-640) for the two extractors that just ran on
// this frame's left/right images (src/Frame.cc:78-81).  mbf, mb as in Frame.h; N = mvKeys.size().
inline void ComputeStereoMatches(ORB_SLAM2::ORBextractor& left, ORB_SLAM2::ORBextractor& right, float mbf, float mb, int N,
                                 std::vector<float>& mvuRight, std::vector<float>& mvDepth) {
    mvuRight.assign(N, -1.0f);
    mvDepth.assign(N, -1.0f);
    if (N == 0) return;
    check(borb_stereo_match2(left.handle(), right.handle(), mbf, mb, mvuRight.data(), mvDepth.data(), N), "borb_stereo_match2");
}
}  // namespace borb
#endif  /* BORB_ADAPTER_NO_EXTRACTOR */
