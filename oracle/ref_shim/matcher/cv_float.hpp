// cv_float.hpp — TEST INFRASTRUCTURE (oracle/_ref matcher build): the float-matrix expressions src/ORBmatcher.cc writes with
// cv::Mat, on top of the host's stand-in cv::Mat (orbslam_mapsave_b200/host/cv_compat.h).
// `A*B + C` is ONE cv::gemm call that multiplies and adds in float32, left to right (characterised against cv2 4.13 in
// tests/golden/prim_gemm3.npz); that is the only product the PINNED functions use (SearchForTriangulation's epipole, :667-673).
// `M / s` and `s * M` scale by (float)(1.0 / s) resp. (float)s (convertTo); dot() and norm() accumulate in double.
// `-R.t()*t` (the only product without `+ C`) follows gemm's general path: double accumulation (same golden file).  The similarity
// scalings of SearchBySim3 / Fuse(Scw) (`M / s`, `s * M.t()`) are written as convertTo does them but are not exercised by the tests.
#pragma once
#define ORB_B200_FORCE_CV_SHIM 1
#include <cassert>
#include <cmath>
#include "../../../orbslam_mapsave_b200/host/cv_compat.h"

namespace cv {

inline Mat gemm_small(const Mat& A, const Mat& B, const Mat* C) {
    Mat out(A.rows, B.cols, CV_32F);
    for (int r = 0; r < A.rows; r++)
        for (int c = 0; c < B.cols; c++) {
            float acc = A.at<float>(r, 0) * B.at<float>(0, c);
            for (int k = 1; k < A.cols; k++) acc = acc + A.at<float>(r, k) * B.at<float>(k, c);
            if (C) acc = acc + C->at<float>(r, c);
            out.at<float>(r, c) = acc;
        }
    return out;
}
// A * B WITHOUT `+ C`: the only such products in src/ORBmatcher.cc are `-R.t()*t` (:306, :993, :1344, :1481), which cv::gemm runs on
// its general path (transposed operand): products accumulated in DOUBLE, rounded once.  The negation of the materialised transpose is
// exact, so (-R^T)*t accumulated in double equals gemm(R, t, alpha = -1, GEMM_1_T).
inline Mat gemm_general(const Mat& A, const Mat& B) {
    Mat out(A.rows, B.cols, CV_32F);
    for (int r = 0; r < A.rows; r++)
        for (int c = 0; c < B.cols; c++) {
            double acc = (double)A.at<float>(r, 0) * (double)B.at<float>(0, c);
            for (int k = 1; k < A.cols; k++) acc = acc + (double)A.at<float>(r, k) * (double)B.at<float>(k, c);
            out.at<float>(r, c) = (float)acc;
        }
    return out;
}
struct MatProd {                                          // A * B, evaluated when it meets `+ C` (one small gemm) or a Mat
    Mat a, b;
    operator Mat() const { return gemm_general(a, b); }
};
inline MatProd operator*(const Mat& a, const Mat& b) { MatProd p = {a, b}; return p; }
inline Mat operator+(const MatProd& p, const Mat& c) { return gemm_small(p.a, p.b, &c); }

template <class F> inline Mat map1(const Mat& a, F f) {
    Mat o(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < a.cols; c++) o.at<float>(r, c) = f(a.at<float>(r, c));
    return o;
}
template <class F> inline Mat map2(const Mat& a, const Mat& b, F f) {
    Mat o(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < a.cols; c++) o.at<float>(r, c) = f(a.at<float>(r, c), b.at<float>(r, c));
    return o;
}
inline Mat operator-(const Mat& a) { return map1(a, [](float v) { return -v; }); }
inline Mat operator-(const Mat& a, const Mat& b) { return map2(a, b, [](float x, float y) { return x - y; }); }
inline Mat operator+(const Mat& a, const Mat& b) { return map2(a, b, [](float x, float y) { return x + y; }); }
inline Mat operator/(const Mat& a, double s) { const float k = (float)(1.0 / s); return map1(a, [k](float v) { return v * k; }); }
inline Mat operator*(double s, const Mat& a) { const float k = (float)s; return map1(a, [k](float v) { return v * k; }); }
inline double norm(const Mat& a) {
    double s = 0;
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < a.cols; c++) s += (double)a.at<float>(r, c) * (double)a.at<float>(r, c);
    return std::sqrt(s);
}

}  // namespace cv
