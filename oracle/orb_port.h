// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/orb_prims.h header).
//
// Independent CPU restatement ("port") of the ORB front-end hot path, on POD arrays:
//   extractor  : src/ORBextractor.cc:410-470 (ctor tables), :1107-1132 (pyramid), :765-853 (cells +
//                FAST + quadtree + orientation), :1043-1105 (blur + rBRIEF + concatenation)
//   stereo     : src/Frame.cc:466-640
//   matchers   : src/ORBmatcher.cc (SearchByProjection :45-137, SearchByBoW :159-288 / :522-655,
//                SearchForTriangulation :657-823, ComputeThreeMaxima :1601-1642,
//                DescriptorDistance :1647-1663), grid src/Frame.cc:230-245,327-392
//   BoW feeder : Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1259
// It is validated against the verbatim-compiled reference (oracle/_ref/liborbref.so) and against cv2
// 4.13 in tests/.  The quadtree uses the canonical tie-break "equal sizes: later-created node first"
//  in a list-order/prefix-sum formulation that the CUDA kernel mirrors.
#pragma once
#include <cstdint>

extern "C" {

// cv::KeyPoint-compatible 28-byte record (ORBextractor.cc:1103 output element)
typedef struct orbport_kp {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orbport_kp;

void* orbport_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
void orbport_destroy(void* h);
void orbport_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* per_level, int* umax16);
int orbport_extract(void* h, const uint8_t* img, int w, int hgt, int stride, orbport_kp* kps, uint8_t* desc, int cap);
// state of the last orbport_extract call
int orbport_level_size(void* h, int level, int* w, int* hgt);
const uint8_t* orbport_level_ptr(void* h, int level);   // unblurred, stride == w
const uint8_t* orbport_blur_ptr(void* h, int level);    // blurred (ORBextractor.cc:1086), stride == w
int orbport_candidates(void* h, int level, int32_t* xys, int cap);  // (x,y,score) absolute level px, reference order
int orbport_level_count(void* h, int level);            // keypoints kept on that level

// standalone pieces (for kernel-level parity tests)
int orbport_distribute(const int32_t* xys, int n, int width, int height, int N, int32_t* out_xys, int cap);

// Frame::ComputeStereoMatches (Frame.cc:466-640). pyrL/pyrR: arrays of nlevels pointers (stride == lw[l]).
int orbport_stereo(const orbport_kp* kL, const uint8_t* dL, int nL, const orbport_kp* kR, const uint8_t* dR, int nR,
                   const uint8_t* const* pyrL, const uint8_t* const* pyrR, const int* lw, const int* lh, int nlevels,
                   const float* scale, const float* inv_scale, float bf, float b, float* uRight, float* depth,
                   int32_t* best_dist_dbg);

int orbport_hamming(const uint8_t* a, const uint8_t* b);
}
