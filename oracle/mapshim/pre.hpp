// TEST INFRASTRUCTURE ONLY (oracle/).  Force-included ahead of the reference's src/MapPoint.cc so that it compiles VERBATIM against
// the real include/MapPoint.h (oracle/Makefile target `ref`, _ref/libmapref.so): KeyFrame.h / Frame.h / Map.h / ORBmatcher.h are
// replaced by plain-data stand-ins with the members MapPoint.cc touches.  Pins MapPoint::PredictScale (:385-417), the
// *DistanceInvariance getters (:373-383), ComputeDistinctiveDescriptors (:242-307) and UpdateNormalAndDepth (:330-371).
#pragma once
#define KEYFRAME_H
#define FRAME_H
#define MAP_H
#define ORBMATCHER_H

#include <algorithm>
#include <climits>
#include <cmath>
#include <map>
#include <mutex>
#include <set>
#include <vector>

#include <opencv2/core/core.hpp>

using namespace std;        // the reference's headers rely on it (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:36)

namespace ORB_SLAM2 {

class MapPoint;

class KeyFrame {
public:
    long unsigned int mnId = 0, mnFrameId = 0;
    bool bad = false;
    cv::Mat mDescriptors, Ow;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight, mvScaleFactors;
    int mnScaleLevels = 0;
    float mfLogScaleFactor = 1.f;
    bool isBad() { return bad; }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    void EraseMapPointMatch(const size_t&) {}
    void ReplaceMapPointMatch(const size_t&, MapPoint*) {}
};

class Frame {
public:
    long unsigned int mnId = 0;
    cv::Mat mDescriptors, Ow;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvScaleFactors;
    int mnScaleLevels = 0;
    float mfLogScaleFactor = 1.f;
    cv::Mat GetCameraCenter() { return Ow.clone(); }
};

class Map {
public:
    std::mutex mMutexPointCreation;
    void EraseMapPoint(MapPoint*) {}
};

class ORBmatcher {
public:
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {   // a10, pinned on its own in libmatchref
        const uint32_t* pa = a.ptr<uint32_t>();
        const uint32_t* pb = b.ptr<uint32_t>();
        int d = 0;
        for (int i = 0; i < 8; i++) d += __builtin_popcount(pa[i] ^ pb[i]);
        return d;
    }
};

}  // namespace ORB_SLAM2

#define protected public     // the wrapper fills mfMaxDistance / mObservations of the reference's MapPoint directly
