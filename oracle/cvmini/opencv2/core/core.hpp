// TEST INFRASTRUCTURE ONLY (oracle/): a minimal stand-in for the parts of OpenCV's core module that the reference's
// src/ORBmatcher.cc touches, so that file can be compiled VERBATIM where it lies (oracle/Makefile, target `matchref`).
// Written from scratch; only the semantics the matcher code relies on:
//   cv::Mat  CV_8U / CV_32F, row-major, views by row()/col()/rowRange()/colRange(), at<T>(), ptr<T>(), t(), dot(),
//            operator* (float gemm: sum_k a_ik*b_kj accumulated left to right in float32, no FMA — what OpenCV's gemm
//            does for the 3x3 / 3x1 products of this file), +, -, unary -, scalar * and /.
//   cv::norm (L2: squares accumulated in double, index order), cv::KeyPoint, cv::Point2f.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5

namespace cv {
typedef unsigned char uchar;

struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float a, float b) : x(a), y(b) {} };
struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};

class Mat {
public:
    int rows, cols, type_;
    size_t step;
    uchar* data;
    std::shared_ptr<std::vector<uchar>> buf;
    Mat() : rows(0), cols(0), type_(CV_8U), step(0), data(nullptr) {}
    Mat(int r, int c, int t) : rows(r), cols(c), type_(t), step((size_t)c * esz(t)), data(nullptr) {
        buf = std::make_shared<std::vector<uchar>>((size_t)r * step, 0);
        data = buf->data();
    }
    static int esz(int t) { return t == CV_32F ? 4 : 1; }
    void create(int r, int c, int t) { *this = Mat(r, c, t); }
    void release() { *this = Mat(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> T& at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
    template <typename T> const T& at(int i) const { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    Mat view(int r0, int r1, int c0, int c1) const {
        Mat m; m.rows = r1 - r0; m.cols = c1 - c0; m.type_ = type_; m.step = step; m.buf = buf;
        m.data = data + (size_t)r0 * step + (size_t)c0 * esz(type_);
        return m;
    }
    Mat row(int r) const { return view(r, r + 1, 0, cols); }
    Mat col(int c) const { return view(0, rows, c, c + 1); }
    Mat rowRange(int a, int b) const { return view(a, b, 0, cols); }
    Mat colRange(int a, int b) const { return view(0, rows, a, b); }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * esz(type_));
        return m;
    }
    void copyTo(Mat& dst) const { dst = clone(); }
    Mat t() const {
        Mat m(cols, rows, CV_32F);
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++) m.at<float>(c, r) = at<float>(r, c);
        return m;
    }
    // u8 -> f32 (or a plain copy); `dst` may be *this (a view becomes an independent float matrix, as in OpenCV)
    void convertTo(Mat& dst, int t) const {
        Mat m(rows, cols, t);
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++) {
                const float v = type_ == CV_32F ? at<float>(r, c) : (float)at<uchar>(r, c);
                if (t == CV_32F) m.at<float>(r, c) = v; else m.at<uchar>(r, c) = (uchar)v;
            }
        dst = m;
    }
    static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
    static Mat ones(int r, int c, int t) {
        Mat m(r, c, t);
        for (int i = 0; i < r; i++)
            for (int j = 0; j < c; j++) { if (t == CV_32F) m.at<float>(i, j) = 1.f; else m.at<uchar>(i, j) = 1; }
        return m;
    }
    Mat reshape(int) const { return *this; }        // N x 2 CV_32F <-> N x 1 CV_32FC2: the same memory; undistortPoints below reads N x 2
    double dot(const Mat& o) const {                // float inputs, products and sum in double, index order
        double s = 0;
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++) s += (double)at<float>(r, c) * o.at<float>(r, c);
        return s;
    }
};

inline Mat operator*(const Mat& a, const Mat& b) {
    Mat m(a.rows, b.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < b.cols; j++) {
            float s = a.at<float>(i, 0) * b.at<float>(0, j);
            for (int k = 1; k < a.cols; k++) s = s + a.at<float>(i, k) * b.at<float>(k, j);
            m.at<float>(i, j) = s;
        }
    return m;
}
inline Mat operator+(const Mat& a, const Mat& b) {
    Mat m(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) m.at<float>(i, j) = a.at<float>(i, j) + b.at<float>(i, j);
    return m;
}
inline Mat operator-(const Mat& a, const Mat& b) {
    Mat m(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) m.at<float>(i, j) = a.at<float>(i, j) - b.at<float>(i, j);
    return m;
}
inline Mat operator-(const Mat& a) {
    Mat m(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) m.at<float>(i, j) = -a.at<float>(i, j);
    return m;
}
// Scalar scaling as cv::MatExpr evaluates it for CV_32F: convertTo with alpha narrowed to float.
inline Mat operator*(double s, const Mat& a) {
    Mat m(a.rows, a.cols, CV_32F);
    const float f = (float)s;
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) m.at<float>(i, j) = a.at<float>(i, j) * f;
    return m;
}
inline Mat operator*(const Mat& a, double s) { return s * a; }
inline Mat operator/(const Mat& a, double s) { return (1.0 / s) * a; }

enum { NORM_L1 = 2 };
inline double norm(const Mat& a, const Mat& b, int /*NORM_L1*/) {      // sum |a-b| in double (exact for the integer-valued SAD patches)
    double s = 0;
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) s += std::fabs((double)a.at<float>(i, j) - (double)b.at<float>(i, j));
    return s;
}
// cv::FileStorage / cv::FileNode: only so that the YAML save/load members of DBoW2's TemplatedVocabulary (virtual, hence always
// instantiated) compile; the oracle loads vocabularies through loadFromTextFile and never calls them.
class FileNode {
public:
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0.0; }
    operator float() const { return 0.f; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const char*, int) {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    template <typename T> FileStorage& operator<<(const T&) { return *this; }
};

// cv::undistortPoints(src, dst, cameraMatrix, distCoeffs, R = empty, P) as OpenCV 4.x evaluates it for CV_32FC2 points
// (calib3d/undistort.dispatch.cpp, cvUndistortPointsInternal with the default TermCriteria(MAX_ITER, 5, 0.01)): camera
// matrix and coefficients widened to double, x = (u - cx) * (1/fx), FIVE fixed-point iterations of the Brown model
// (k1 k2 p1 p2 k3 [k4 k5 k6]; thin-prism and tilt terms are zero here, the tilt step multiplies by the identity),
// re-projection with P, result narrowed to float.  Restated from the published algorithm; pinned against cv2 4.13 in
// tests/test_oracle_frame_ref.py.  src: N x 2 CV_32F (the reference's reshape(2) is a no-op in this Mat); dst may be src.
inline void undistortPoints(const Mat& src, Mat& dst, const Mat& Kc, const Mat& D, const Mat& /*R*/, const Mat& P) {
    double k[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const int nd = D.empty() ? 0 : D.rows * D.cols;
    for (int i = 0; i < nd && i < 12; i++) k[i] = (double)D.at<float>(i);
    const double fx = Kc.at<float>(0, 0), fy = Kc.at<float>(1, 1), cx = Kc.at<float>(0, 2), cy = Kc.at<float>(1, 2);
    const double ifx = 1. / fx, ify = 1. / fy;
    double RR[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    if (!P.empty())
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) RR[r][c] = (double)P.at<float>(r, c);
    Mat out = (dst.data == src.data) ? dst : Mat(src.rows, src.cols, CV_32F);
    for (int i = 0; i < src.rows; i++) {
        double x = src.at<float>(i, 0), y = src.at<float>(i, 1);
        const double u = x, v = y;
        x = (x - cx) * ifx; y = (y - cy) * ify;
        if (nd > 0) {
            const double x0 = x, y0 = y;
            for (int j = 0; j < 5; j++) {
                const double r2 = x * x + y * y;
                const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
                if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
                const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
                const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
                x = (x0 - deltaX) * icdist; y = (y0 - deltaY) * icdist;
            }
        }
        const double xx = RR[0][0] * x + RR[0][1] * y + RR[0][2];
        const double yy = RR[1][0] * x + RR[1][1] * y + RR[1][2];
        const double ww = 1. / (RR[2][0] * x + RR[2][1] * y + RR[2][2]);
        out.at<float>(i, 0) = (float)(xx * ww); out.at<float>(i, 1) = (float)(yy * ww);
    }
    dst = out;
}

// cv::Mat_<float>(3,1) << x, y, z   (Frame::UnprojectStereo)
template <typename T> class Mat_ : public Mat {
public:
    Mat_(int r, int c) : Mat(r, c, CV_32F) {}
    struct Init {
        Mat_* m; int i;
        Init& operator,(T v) { m->template at<T>(i / m->cols, i % m->cols) = v; i++; return *this; }
        operator Mat() const { return *m; }
    };
    Init operator<<(T v) { this->template at<T>(0, 0) = v; return Init{this, 1}; }
};

inline double norm(const Mat& a) {
    double s = 0;
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) { const double v = a.at<float>(i, j); s += v * v; }
    return std::sqrt(s);
}
}  // namespace cv
