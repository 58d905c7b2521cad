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
    Mat reshape(int) const { return *this; }        // only on paths the oracle never runs (UndistortKeyPoints)
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

inline void undistortPoints(const Mat&, Mat&, const Mat&, const Mat&, const Mat&, const Mat&) {}   // never run by the oracle

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
