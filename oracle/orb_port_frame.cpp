// ORACLE — TEST INFRASTRUCTURE ONLY.  Restatement of the RGB-D leg of the Frame constructor (reference src/Frame.cc:119-178):
//   UndistortKeyPoints (:404-434) -> cv::undistortPoints (OpenCV calib3d, un-vendored; algorithm restated, pinned to cv2 4.13)
//   ComputeStereoFromRGBD (:643-664), ComputeImageBounds (:436-464)
//   depth map conversion of Tracking::GrabImageRGBD (src/Tracking.cc:227-228): imDepth.convertTo(CV_32F, mDepthMapFactor)
// Validated against the verbatim-compiled Frame.cc (oracle/_ref/libframeref.so) and cv2 in tests/test_oracle_frame_ref.py.
#include <cmath>
#include <cstring>

#include "orb_port.h"

namespace {
// one point through cv::undistortPoints(src, dst, K, D, Mat(), K): double arithmetic, 5 iterations, result narrowed to float
void undistort_point(float u_in, float v_in, const float* K4, const float* dist, int n_dist, float* xo, float* yo) {
    double k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n_dist && i < 8; i++) k[i] = (double)dist[i];
    const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
    const double ifx = 1. / fx, ify = 1. / fy;
    const double u = u_in, v = v_in;
    double x = (u - cx) * ifx, y = (v - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
        const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
        const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
        x = (x0 - deltaX) * icdist; y = (y0 - deltaY) * icdist;
    }
    // new camera matrix P = K: rows (fx 0 cx), (0 fy cy), (0 0 1)
    const double xx = fx * x + 0.0 * y + cx, yy = 0.0 * x + fy * y + cy, ww = 1. / (0.0 * x + 0.0 * y + 1.0);
    *xo = (float)(xx * ww); *yo = (float)(yy * ww);
}
}  // namespace

extern "C" {

// Frame::UndistortKeyPoints (:404-434): dist[0] == 0 -> copy
void orbport_undistort_keypoints(const orbport_kp* keys, int n, const float* K4, const float* dist, int n_dist, orbport_kp* keys_un) {
    for (int i = 0; i < n; i++) {
        keys_un[i] = keys[i];
        if (n_dist > 0 && dist[0] != 0.0f) undistort_point(keys[i].x, keys[i].y, K4, dist, n_dist, &keys_un[i].x, &keys_un[i].y);
    }
}

// Frame::ComputeImageBounds (:436-464)
void orbport_image_bounds(int w, int h, const float* K4, const float* dist, int n_dist, float* b4) {
    if (n_dist > 0 && dist[0] != 0.0f) {
        float x[4], y[4];
        const float cu[4] = {0.f, (float)w, 0.f, (float)w}, cv[4] = {0.f, 0.f, (float)h, (float)h};
        for (int i = 0; i < 4; i++) undistort_point(cu[i], cv[i], K4, dist, n_dist, &x[i], &y[i]);
        b4[0] = std::fmin(x[0], x[2]); b4[2] = std::fmax(x[1], x[3]);
        b4[1] = std::fmin(y[0], y[1]); b4[3] = std::fmax(y[2], y[3]);
    } else { b4[0] = 0.f; b4[1] = 0.f; b4[2] = (float)w; b4[3] = (float)h; }
}

// Frame::ComputeStereoFromRGBD (:643-664); depth: h x w floats (metres)
int orbport_stereo_from_rgbd(const orbport_kp* keys, const orbport_kp* keys_un, int n, const float* depth, int w, int h, float bf,
                             float* u_right, float* depth_out) {
    int cnt = 0;
    for (int i = 0; i < n; i++) {
        u_right[i] = -1.f; depth_out[i] = -1.f;
        const int v = (int)keys[i].y, u = (int)keys[i].x;                  // imDepth.at<float>(v,u): float -> int truncation
        if (u < 0 || v < 0 || u >= w || v >= h) continue;                  // never happens for extractor output (EDGE_THRESHOLD border)
        const float d = depth[(size_t)v * w + u];
        if (d > 0) { depth_out[i] = d; u_right[i] = keys_un[i].x - bf / d; cnt++; }
    }
    return cnt;
}

// imDepth.convertTo(imDepth, CV_32F, mDepthMapFactor) for a CV_16U depth map (OpenCV cvtScale 16u->32f: float multiply-add)
void orbport_depth_to_float(const uint16_t* raw, int n, float factor, float* out) {
    for (int i = 0; i < n; i++) out[i] = (float)raw[i] * factor;
}

}  // extern "C"
