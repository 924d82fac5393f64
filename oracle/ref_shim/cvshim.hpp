// cvshim.hpp — TEST INFRASTRUCTURE.  The slice of the OpenCV C++ API that /root/reference/src/ORBextractor.cc uses, so that the
// reference's OWN, UNMODIFIED translation unit can be compiled in an image that has no OpenCV C++ headers or libraries
// (oracle/Makefile target `_ref`).  The data types are written here; the five image primitives the reference delegates to OpenCV
// (cv::FAST, cv::resize INTER_LINEAR, cv::GaussianBlur 7x7 sigma 2, cv::copyMakeBorder REFLECT_101, cv::fastAtan2) are the oracle's
// restatements (liborb_oracle.so), each pinned bit for bit against cv2 4.13 golden vectors (tests/golden/prim_*.npz).
// What this buys: the reference's control flow — ComputePyramid, ComputeKeyPointsOctTree, DistributeOctTree with its std::list and
// its real pointer-valued tie-break, IC_Angle, computeOrbDescriptor, operator() — runs as the authors wrote it.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_PI 3.1415926535897932384626433832795

extern "C" {
struct orc_xyr { int x, y, r; };
int orc_fast(const uchar* img, int w, int h, int stride, int threshold, int nms, orc_xyr* out, int cap);
void orc_resize(const uchar* src, int sw, int sh, int sstride, uchar* dst, int dw, int dh, int dstride);
void orc_blur(const uchar* src, int w, int h, int sstride, uchar* dst, int dstride);
float orc_fast_atan2(float y, float x);
}

// cvRound: round half to even (SSE2 cvtsd2si / lrint under the default rounding mode); cvFloor / cvCeil as in OpenCV's fast_math.hpp
static inline int cvRound(double v) { return (int)std::nearbyint(v); }
static inline int cvRound(float v) { return (int)std::nearbyintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

struct ZerosExpr { int rows, cols, type; };

class Mat {
public:
    int rows, cols;
    size_t step;                       // bytes per row (uchar matrices only)
    uchar* data;
    Mat() : rows(0), cols(0), step(0), data(nullptr), type_(CV_8U) {}
    Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(nullptr), type_(CV_8U) { create(r, c, type); }
    Mat(Size s, int type) : rows(0), cols(0), step(0), data(nullptr), type_(CV_8U) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, void* ext, size_t step_) : rows(r), cols(c), step(step_), data((uchar*)ext), type_(type) {}
    static ZerosExpr zeros(int r, int c, int type) { ZerosExpr z = {r, c, type}; return z; }
    // like cv::Mat::operator=(const MatExpr&): evaluated INTO this matrix when size and type already match (a row-range view of the
    // descriptor matrix in computeDescriptors), otherwise into fresh storage
    Mat& operator=(const ZerosExpr& z) {
        create(z.rows, z.cols, z.type);
        for (int r = 0; r < rows; r++) std::memset(data + (size_t)r * step, 0, (size_t)cols);
        return *this;
    }
    void create(int r, int c, int type) {          // no-op when the geometry already matches (cv::Mat::create)
        assert(type == CV_8U);
        if (data && r == rows && c == cols && type == type_) return;
        rows = r; cols = c; type_ = type; step = (size_t)c;
        buf_.reset(new std::vector<uchar>((size_t)r * step));
        data = buf_->data();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { rows = cols = 0; step = 0; data = nullptr; buf_.reset(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    size_t step1() const { return step; }
    Mat rowRange(int a, int b) const { Mat m = *this; m.rows = b - a; m.data = data + (size_t)a * step; return m; }
    Mat colRange(int a, int b) const { Mat m = *this; m.cols = b - a; m.data = data + a; return m; }
    Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols);
        return m;
    }
    // cv::Mat::copyTo(dst, mask): a newly allocated destination is zero-filled first, then the pixels with mask != 0 are copied
    void copyTo(Mat& dst, const Mat& mask) const {
        if (mask.empty()) { dst = clone(); return; }
        assert(mask.rows == rows && mask.cols == cols);
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++) m.data[(size_t)r * m.step + c] = mask.data[(size_t)r * mask.step + c] ? data[(size_t)r * step + c] : 0;
        dst = m;
    }
    template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    Mat getMat() const { return *this; }           // Mat doubles as InputArray / OutputArray
private:
    int type_;
    std::shared_ptr<std::vector<uchar> > buf_;      // shared by all views, like cv::Mat's reference count
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;

// cv::FAST(image, keypoints, threshold, nonmaxSuppression): FAST-9/16, KeyPoint(x, y, 7.f, -1, score)
inline void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nms = true) {
    keypoints.clear();
    if (image.rows < 7 || image.cols < 7) return;
    std::vector<orc_xyr> tmp((size_t)image.rows * image.cols / 4 + 16);
    int n = orc_fast(image.data, image.cols, image.rows, (int)image.step, threshold, nms ? 1 : 0, tmp.data(), (int)tmp.size());
    if (n > (int)tmp.size()) {
        tmp.resize(n);
        n = orc_fast(image.data, image.cols, image.rows, (int)image.step, threshold, nms ? 1 : 0, tmp.data(), (int)tmp.size());
    }
    keypoints.reserve(n);
    for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)tmp[i].x, (float)tmp[i].y, 7.f, -1, (float)tmp[i].r));
}

inline void resize(const Mat& src, Mat& dst, Size dsize, double, double, int interpolation) {
    assert(interpolation == INTER_LINEAR);
    (void)interpolation;
    dst.create(dsize, src.type());                 // a view of the right size is written in place (ComputePyramid relies on it)
    orc_resize(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sx, double sy, int borderType) {
    assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && borderType == BORDER_REFLECT_101);
    (void)ksize; (void)sx; (void)sy; (void)borderType;
    Mat in = src.clone();                          // the reference blurs in place
    dst.create(src.rows, src.cols, src.type());
    orc_blur(in.data, in.cols, in.rows, (int)in.step, dst.data, (int)dst.step);
}

static inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}
// cv::copyMakeBorder with BORDER_REFLECT_101 (with or without BORDER_ISOLATED: the sources used by ComputePyramid are either whole
// matrices or treated as isolated).  The source may be a view INSIDE the destination (level >= 1), so rows are moved with memmove
// and every border pixel is computed from source pixels only, which no border write can touch.
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType) {
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    (void)borderType;
    dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
    for (int y = 0; y < src.rows; y++) {           // interior rows first (they may alias the source exactly)
        uchar* d = dst.data + (size_t)(y + top) * dst.step;
        const uchar* s = src.data + (size_t)y * src.step;
        if (d + left != s) std::memmove(d + left, s, (size_t)src.cols);
        for (int x = 0; x < left; x++) d[x] = s[reflect101(x - left, src.cols)];
        for (int x = 0; x < right; x++) d[left + src.cols + x] = s[reflect101(src.cols + x, src.cols)];
    }
    for (int y = 0; y < top; y++)
        std::memcpy(dst.data + (size_t)y * dst.step, dst.data + (size_t)(top + reflect101(y - top, src.rows)) * dst.step, (size_t)dst.cols);
    for (int y = 0; y < bottom; y++)
        std::memcpy(dst.data + (size_t)(top + src.rows + y) * dst.step,
                    dst.data + (size_t)(top + reflect101(src.rows + y, src.rows)) * dst.step, (size_t)dst.cols);
}

inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

struct KeyPointsFilter {                           // only the dead ComputeKeyPointsOld uses it
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints) {
        if (npoints >= 0 && (int)keypoints.size() > npoints) {
            std::stable_sort(keypoints.begin(), keypoints.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
            keypoints.resize(npoints);
        }
    }
};

}  // namespace cv
