// TEST INFRASTRUCTURE.  A header with the SHAPE of the real <opencv2/core/core.hpp> (OpenCV 3 / 4) for the slice the drop-in classes
// touch: cv::Mat with OpenCV's member order (flags, dims, rows, cols, data ... MatSize size; MatStep step), cv::_InputArray /
// cv::_OutputArray proxies behind the InputArray / OutputArray typedefs, Point_ / Rect_ / KeyPoint.  Putting this directory on the
// include path makes orbslam_mapsave_b200/host/cv_compat.h take its `__has_include(<opencv2/core/core.hpp>)` branch, i.e. the way the
// host sources are compiled inside the reference's tree; tests/test_capi_load.py compiles them that way (OpenCV itself is not in this image).
#pragma once
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5
#define CV_MAT_DEPTH(t) ((t) & 7)

namespace cv {
typedef unsigned char uchar;

template <typename T> class Point_ {
public:
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
template <typename T> class Rect_ {
public:
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

class KeyPoint {
public:
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

struct MatSize {
    int* p;
    explicit MatSize(int* p_) : p(p_) {}
};
struct MatStep {
    size_t* p;
    size_t buf[2];
    MatStep() : p(buf) { buf[0] = buf[1] = 0; }
    operator size_t() const { return buf[0]; }
    MatStep& operator=(size_t s) { buf[0] = s; return *this; }
    size_t operator[](int i) const { return p[i]; }
};

class Mat {
public:
    enum { AUTO_STEP = 0 };
    int flags;                       // OpenCV's member order
    int dims;
    int rows, cols;
    uchar* data;
    const uchar* datastart;
    const uchar* dataend;
    const uchar* datalimit;
    void* allocator;
    void* u;
    MatSize size;
    MatStep step;

    Mat() : flags(0), dims(0), rows(0), cols(0), data(0), datastart(0), dataend(0), datalimit(0), allocator(0), u(0), size(&rows) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t step_ = AUTO_STEP) : Mat() {
        flags = type; dims = 2; rows = r; cols = c; data = (uchar*)ext;
        step = step_ ? step_ : (size_t)c * elemSize();
    }
    Mat(const Mat& m) : Mat() { *this = m; }
    Mat& operator=(const Mat& m) {
        flags = m.flags; dims = m.dims; rows = m.rows; cols = m.cols; data = m.data; step.buf[0] = m.step.buf[0]; step.buf[1] = m.step.buf[1];
        buf_ = m.buf_;
        return *this;
    }
    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == this->type()) return;
        flags = type; dims = 2; rows = r; cols = c;
        step = (size_t)c * elemSize(); step.buf[1] = elemSize();
        buf_.reset(new std::vector<uchar>((size_t)r * step.buf[0]));
        data = buf_->data();
    }
    void release() { rows = cols = 0; dims = 0; data = 0; buf_.reset(); }
    bool empty() const { return data == 0 || rows * cols == 0 || dims == 0; }
    int type() const { return flags & 0xFFF; }
    size_t elemSize() const { return dims > 0 ? (CV_MAT_DEPTH(flags) == CV_32F ? 4 : 1) : 0; }
    bool isContinuous() const { return step.buf[0] == (size_t)cols * elemSize(); }
    Mat row(int y) const { Mat m(*this); m.rows = 1; m.data = data + (size_t)y * step.buf[0]; return m; }
    Mat rowRange(int a, int b) const { Mat m(*this); m.rows = b - a; m.data = data + (size_t)a * step.buf[0]; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.cols = b - a; m.data = data + (size_t)a * elemSize(); return m; }
    Mat col(int x) const { return colRange(x, x + 1); }
    Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
    Mat clone() const {
        Mat m(rows, cols, type());
        for (int y = 0; y < rows; y++) std::memcpy(m.data + (size_t)y * m.step.buf[0], data + (size_t)y * step.buf[0], (size_t)cols * elemSize());
        return m;
    }
    Mat t() const {
        Mat m(cols, rows, type());
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++)
                std::memcpy(m.data + (size_t)x * m.step.buf[0] + (size_t)y * elemSize(), data + (size_t)y * step.buf[0] + (size_t)x * elemSize(), elemSize());
        return m;
    }
    double dot(const Mat& o) const {
        double s = 0;
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++) s += (double)at<float>(y, x) * (double)o.at<float>(y, x);
        return s;
    }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step.buf[0]); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step.buf[0]); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step.buf[0]; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step.buf[0]; }
    template <typename T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step.buf[0]))[x]; }
    template <typename T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step.buf[0]))[x]; }
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
private:
    std::shared_ptr<std::vector<uchar> > buf_;      // stands for OpenCV's reference-counted UMatData
};

// the proxies every OpenCV function signature is written with
class _InputArray {
public:
    _InputArray() : obj(0) {}
    _InputArray(const Mat& m) : obj((void*)&m) {}
    Mat getMat(int = -1) const { return obj ? *(const Mat*)obj : Mat(); }
    bool empty() const { return !obj || ((const Mat*)obj)->empty(); }
protected:
    void* obj;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) { obj = &m; }
    void create(int r, int c, int type) const { ((Mat*)obj)->create(r, c, type); }
    void release() const { ((Mat*)obj)->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
}  // namespace cv
