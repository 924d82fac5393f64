// cv_compat.h — the handful of OpenCV types ORBextractor.h / ORBmatcher.h name in their signatures.
// With real OpenCV available (the reference's build) the real headers are used; in this repository's container OpenCV's
// C++ headers are absent, so a minimal stand-in with the same names, member names and memory layout is provided
// (cv::KeyPoint is the same 28-byte POD; cv::Mat is a ref-counted 2-D byte/float matrix with ROI views).
#pragma once
#if !defined(ORB_B200_FORCE_CV_SHIM) && __has_include(<opencv2/core/core.hpp>)
#include <opencv2/core/core.hpp>
#else
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5

namespace cv {

struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct Point { int x = 0, y = 0; Point() {} Point(int x_, int y_) : x(x_), y(y_) {} };
typedef Point Point2i;

struct KeyPoint {                       // same field order and size (28 bytes) as cv::KeyPoint
    Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
    KeyPoint() {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;                    // bytes per row
    unsigned char* data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t step_ = 0) : rows(r), cols(c), data((unsigned char*)ext), type_(type) {
        step = step_ ? step_ : (size_t)c * elemSize();
    }
    void create(int r, int c, int type) {
        if (r == rows && c == cols && type == type_ && data && isContinuous()) return;
        rows = r; cols = c; type_ = type; step = (size_t)c * elemSize();
        buf_ = std::shared_ptr<std::vector<unsigned char>>(new std::vector<unsigned char>((size_t)r * step));
        data = buf_->data();
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }   // create() value-initialises the storage
    void release() { rows = cols = 0; step = 0; data = nullptr; buf_.reset(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    size_t elemSize() const { return (!data && rows == 0 && cols == 0) ? 0 : (type_ == CV_32F ? 4 : 1); }   // 0 for an empty Mat, like cv::Mat
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    Mat row(int r) const { Mat m = *this; m.rows = 1; m.data = data + (size_t)r * step; return m; }
    Mat rowRange(int a, int b) const { Mat m = *this; m.rows = b - a; m.data = data + (size_t)a * step; return m; }
    Mat colRange(int a, int b) const { Mat m = *this; m.cols = b - a; m.data = data + (size_t)a * elemSize(); return m; }
    Mat roi(int x, int y, int w, int h) const { return rowRange(y, y + h).colRange(x, x + w); }
    Mat col(int c) const { return colRange(c, c + 1); }
    Mat t() const {                                       // transposed copy (CV_32F or CV_8U)
        Mat m(cols, rows, type_);
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++)
                std::memcpy(m.data + (size_t)c * m.step + (size_t)r * elemSize(), data + (size_t)r * step + (size_t)c * elemSize(), elemSize());
        return m;
    }
    double dot(const Mat& o) const {                      // CV_32F, accumulated in double like cv::Mat::dot
        double s = 0;
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++) s += (double)at<float>(r, c) * (double)o.at<float>(r, c);
        return s;
    }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat& dst) const { dst = clone(); }
    Mat reshape(int /*cn*/) const { return *this; }          // channels are not represented in this single-channel stand-in
    static Mat ones(int r, int c, int type) {
        Mat m(r, c, type);
        for (int y = 0; y < r; y++)
            for (int x = 0; x < c; x++) { if (type == CV_32F) m.at<float>(y, x) = 1.f; else m.at<unsigned char>(y, x) = 1; }
        return m;
    }
    void convertTo(Mat& dst, int rtype) const {           // CV_8U -> CV_32F (exact) or a same-type copy; dst may be *this (a view is left alone)
        Mat out(rows, cols, rtype);
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++) {
                const float v = type_ == CV_32F ? at<float>(y, x) : (float)at<unsigned char>(y, x);
                if (rtype == CV_32F) out.at<float>(y, x) = v; else out.at<unsigned char>(y, x) = (unsigned char)v;
            }
        dst = out;
    }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    unsigned char* ptr(int r = 0) { return data + (size_t)r * step; }
    const unsigned char* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    Mat getMat() const { return *this; }                 // so that Mat doubles as InputArray / OutputArray
private:
    int type_ = CV_8U;
    std::shared_ptr<std::vector<unsigned char>> buf_;
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;

}  // namespace cv
#endif
