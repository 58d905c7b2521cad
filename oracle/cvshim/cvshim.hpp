// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/orb_prims.h header).
//
// Minimal `cv` compatibility shim: just enough of the OpenCV C++ surface for
// /root/reference/src/ORBextractor.cc to compile VERBATIM (in place, never copied) without an
// OpenCV install.  Image primitives forward to oracle/orb_prims.h (pinned to cv2 4.13).
// Nothing here is OpenCV source; it is a from-scratch stand-in with the same call signatures.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "../orb_prims.h"

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0

static inline int cvRound(double v) { return orbprims::cv_round(v); }
static inline int cvRound(float v) { return orbprims::cv_round(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { return orbprims::cv_floor(v); }
static inline int cvCeil(double v) { return orbprims::cv_ceil(v); }

namespace cv {

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3,
       BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0,
             int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
static_assert(sizeof(KeyPoint) == 28, "layout must match cv::KeyPoint");

struct MatZeros { int rows, cols, type; };

class Mat {
public:
    int rows, cols;
    size_t step;
    uchar* data;
    std::shared_ptr<std::vector<uchar>> buf;

    Mat() : rows(0), cols(0), step(0), data(nullptr) {}
    Mat(Size sz, int type) : rows(0), cols(0), step(0), data(nullptr) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(nullptr) { create(r, c, type); }
    // non-owning view over caller memory
    Mat(int r, int c, int /*type*/, void* p, size_t _step) : rows(r), cols(c), step(_step), data((uchar*)p) {}

    void create(int r, int c, int /*type*/) {
        if (data && r == rows && c == cols) return;
        buf = std::make_shared<std::vector<uchar>>((size_t)r * c);
        rows = r; cols = c; step = (size_t)c; data = buf->data();
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    size_t step1() const { return step; }

    Mat operator()(const Rect& r) const {
        Mat m; m.rows = r.height; m.cols = r.width; m.step = step;
        m.data = data + (size_t)r.y * step + r.x; m.buf = buf; return m;
    }
    Mat rowRange(int a, int b) const {
        assert(0 <= a && a <= b && b <= rows);
        Mat m; m.rows = b - a; m.cols = cols; m.step = step; m.data = data + (size_t)a * step; m.buf = buf; return m;
    }
    Mat colRange(int a, int b) const {
        assert(0 <= a && a <= b && b <= cols);
        Mat m; m.rows = rows; m.cols = b - a; m.step = step; m.data = data + a; m.buf = buf; return m;
    }
    Mat clone() const {
        Mat m; m.create(rows, cols, CV_8UC1);
        for (int y = 0; y < rows; y++) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, cols);
        return m;
    }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + x * sizeof(T)); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }

    static MatZeros zeros(int r, int c, int type) { return MatZeros{r, c, type}; }
    // OpenCV semantics: assigning a MatExpr to an existing Mat of equal shape writes IN PLACE
    // (ORBextractor.cc:1037 relies on this to zero the caller's descriptor rows).
    Mat& operator=(const MatZeros& z) {
        create(z.rows, z.cols, z.type);
        for (int y = 0; y < rows; y++) std::memset(data + (size_t)y * step, 0, cols);
        return *this;
    }
};

class _InputArray {
public:
    Mat* m;
    _InputArray() : m(nullptr) {}
    _InputArray(const Mat& mat) : m(const_cast<Mat*>(&mat)) {}
    bool empty() const { return m == nullptr || m->empty(); }
    Mat getMat() const { return m ? *m : Mat(); }
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& mat) { m = &mat; }
    void create(int r, int c, int type) const { m->create(r, c, type); }
    void release() const { if (m) m->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

static inline float fastAtan2(float y, float x) { return orbprims::fast_atan2(y, x); }

static inline void FAST(const Mat& img, std::vector<KeyPoint>& kps, int threshold, bool nms) {
    std::vector<orbprims::FastPt> pts;
    orbprims::fast9_16(img.data, img.cols, img.rows, img.step, threshold, nms, pts);
    kps.clear();
    kps.reserve(pts.size());
    for (const auto& p : pts) kps.push_back(KeyPoint((float)p.x, (float)p.y, 7.f, -1.f, (float)p.score));
}

static inline void resize(const Mat& src, Mat& dst, Size dsize, double, double, int interp) {
    assert(interp == INTER_LINEAR);
    (void)interp;
    dst.create(dsize, CV_8UC1);
    orbprims::resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

static inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType) {
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    (void)borderType;
    const int w = src.cols, h = src.rows;
    Mat s = src;  // keep src alive / stable if dst is re-created
    dst.create(h + top + bottom, w + left + right, CV_8UC1);
    // interior first (memmove: src may alias dst's interior exactly, ORBextractor.cc:1122)
    for (int y = 0; y < h; y++) {
        uchar* d = dst.data + (size_t)(y + top) * dst.step + left;
        const uchar* sp = s.data + (size_t)y * s.step;
        if (d != sp) std::memmove(d, sp, w);
    }
    for (int y = 0; y < h; y++) {
        uchar* row = dst.data + (size_t)(y + top) * dst.step;
        for (int x = 0; x < left; x++) row[x] = row[left + orbprims::reflect101(x - left, w)];
        for (int x = 0; x < right; x++) row[left + w + x] = row[left + orbprims::reflect101(w + x, w)];
    }
    const int W = w + left + right;
    for (int y = 0; y < top; y++)
        std::memcpy(dst.data + (size_t)y * dst.step, dst.data + (size_t)(top + orbprims::reflect101(y - top, h)) * dst.step, W);
    for (int y = 0; y < bottom; y++)
        std::memcpy(dst.data + (size_t)(top + h + y) * dst.step, dst.data + (size_t)(top + orbprims::reflect101(h + y, h)) * dst.step, W);
}

static inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sx, double sy, int borderType) {
    assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && borderType == BORDER_REFLECT_101);
    (void)ksize; (void)sx; (void)sy; (void)borderType;
    Mat s = src;
    dst.create(s.rows, s.cols, CV_8UC1);
    orbprims::gaussian_blur7_u8(s.data, s.cols, s.rows, s.step, dst.data, dst.step);
}

struct KeyPointsFilter {
    // Only referenced by the dead ComputeKeyPointsOld (ORBextractor.cc:1006,1024).
    static void retainBest(std::vector<KeyPoint>&, int) { std::abort(); }
};

}  // namespace cv
