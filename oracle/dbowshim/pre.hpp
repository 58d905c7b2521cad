// TEST INFRASTRUCTURE ONLY (oracle/).  Force-included ahead of the reference's src/KeyFrameDatabase.cc and the DBoW2 sources
// (Thirdparty/DBoW2/DBoW2/{FORB,ScoringObject,BowVector,FeatureVector}.cpp, TemplatedVocabulary.h, DUtils/{Random,Timestamp}.cpp)
// so that they compile VERBATIM (oracle/Makefile target `ref`, _ref/libdbowref.so): the real include/KeyFrameDatabase.h and
// include/ORBVocabulary.h are used as they are; KeyFrame.h / Frame.h are replaced by the two plain-data stand-ins below, which
// carry exactly the members KeyFrameDatabase.cc touches.
#pragma once
#define KEYFRAME_H
#define FRAME_H

#include <list>
#include <map>
#include <mutex>
#include <set>
#include <vector>

#include <opencv2/core/core.hpp>

#include "Thirdparty/DBoW2/DBoW2/BowVector.h"

namespace ORB_SLAM2 {

class KeyFrame {
public:
    long unsigned int mnId = 0;
    DBoW2::BowVector mBowVec;
    // loop / relocalisation query bookkeeping (include/KeyFrame.h:141-149); nothing has queried a fresh keyframe
    long unsigned int mnLoopQuery = ~0ul, mnRelocQuery = ~0ul;
    int mnLoopWords = 0, mnRelocWords = 0;
    float mLoopScore = 0.f, mRelocScore = 0.f;
    std::set<KeyFrame*> connected;
    std::vector<KeyFrame*> covisible;               // best-first
    std::set<KeyFrame*> GetConnectedKeyFrames() { return connected; }
    std::vector<KeyFrame*> GetBestCovisibilityKeyFrames(const int& N) {
        return (int)covisible.size() < N ? covisible : std::vector<KeyFrame*>(covisible.begin(), covisible.begin() + N);
    }
};

class Frame {
public:
    long unsigned int mnId = 0;
    DBoW2::BowVector mBowVec;
};

}  // namespace ORB_SLAM2
