// TEST INFRASTRUCTURE ONLY (oracle/).  C entry points around the reference's src/MapPoint.cc compiled VERBATIM against its real
// include/MapPoint.h (oracle/Makefile target `ref`, _ref/libmapref.so; oracle/mapshim/pre.hpp).  Pins MapPoint::PredictScale,
// the *DistanceInvariance getters and ComputeDistinctiveDescriptors of the restatements to the reference source.
#include <cstdint>
#include <cstring>
#include <vector>

#include "MapPoint.h"

using namespace ORB_SLAM2;

namespace {
cv::Mat origin() { return cv::Mat(3, 1, CV_32F); }
}

extern "C" {

// PredictScale(currentDist, KeyFrame*) (use_frame = 0, :385-400) or (currentDist, Frame*) (use_frame = 1, :402-417), element-wise
void mapref_predict_scale(const float* max_distance, const float* dist, int n, float log_scale, int n_levels, int use_frame, int32_t* out) {
    Map map;
    KeyFrame kf; Frame fr;
    kf.mfLogScaleFactor = fr.mfLogScaleFactor = log_scale;
    kf.mnScaleLevels = fr.mnScaleLevels = n_levels;
    MapPoint mp(origin(), &kf, &map);
    for (int i = 0; i < n; i++) {
        mp.mfMaxDistance = max_distance[i];
        out[i] = use_frame ? mp.PredictScale(dist[i], &fr) : mp.PredictScale(dist[i], &kf);
    }
}

void mapref_distance_invariance(const float* max_distance, const float* min_distance, int n, float* out_max, float* out_min) {
    Map map; KeyFrame kf;
    MapPoint mp(origin(), &kf, &map);
    for (int i = 0; i < n; i++) {
        mp.mfMaxDistance = max_distance[i]; mp.mfMinDistance = min_distance[i];
        out_max[i] = mp.GetMaxDistanceInvariance(); out_min[i] = mp.GetMinDistanceInvariance();
    }
}

// ComputeDistinctiveDescriptors over N observations (descriptor i lives in keyframe i; the keyframes sit in one array, so the
// std::map<KeyFrame*, size_t> walks them in index order).  bad[i] != 0 marks keyframe i as bad.  Writes the chosen descriptor;
// returns 1 if one was chosen.
int mapref_distinctive_descriptor(const uint8_t* desc, const uint8_t* bad, int n, uint8_t* chosen) {
    Map map;
    std::vector<KeyFrame> kfs(n > 0 ? n : 1);
    MapPoint mp(origin(), &kfs[0], &map);
    for (int i = 0; i < n; i++) {
        kfs[i].mnId = i;
        kfs[i].bad = bad && bad[i];
        kfs[i].mDescriptors = cv::Mat(1, 32, CV_8U);
        std::memcpy(kfs[i].mDescriptors.data, desc + (size_t)i * 32, 32);
        kfs[i].mvuRight.assign(1, -1.f);
        mp.AddObservation(&kfs[i], 0);
    }
    mp.ComputeDistinctiveDescriptors();
    const cv::Mat d = mp.GetDescriptor();
    if (d.empty()) return 0;
    std::memcpy(chosen, d.data, 32);
    return 1;
}

}  // extern "C"
