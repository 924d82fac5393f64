// TEST INFRASTRUCTURE (oracle/_ref Frame build).  Stand-in for <opencv2/core/core.hpp> / <opencv2/opencv.hpp> as far as the reference's
// src/Frame.cc, src/MapPoint.cc and the headers they include need it: the host's cv::Mat stand-in with the float expressions of
// ../../matcher/cv_float.hpp, plus the few extra names Frame.cc mentions.  What ComputeStereoMatches computes with them — the 11x11
// patches converted to float, centre-subtracted, compared with an L1 norm (src/Frame.cc:679-700) — is exact integer arithmetic in
// float32 / double for 8-bit images, so the stand-ins have no rounding freedom.
#pragma once
#include "../../../dbow2/opencv2/core/core.hpp"     // cv::Mat stand-in + FileStorage stubs (ORBVocabulary.h pulls in TemplatedVocabulary.h)
#include "../../../matcher/cv_float.hpp"
#include <initializer_list>

namespace cv {
enum { NORM_L1 = 2 };
inline double norm(const Mat& a, const Mat& b, int type) {
    if (type != NORM_L1) std::abort();
    double s = 0;
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < a.cols; c++) s += std::fabs((double)a.at<float>(r, c) - (double)b.at<float>(r, c));
    return s;
}
inline void undistortPoints(const Mat&, Mat&, const Mat&, const Mat&, const Mat&, const Mat&) { std::abort(); }   // only reached with distortion
struct MatReshape { };
// cv::Mat_<float>(3,1) << x, y, z   (src/Frame.cc:791)
template <class T> struct MatCommaInit {
    Mat m; int i;
    MatCommaInit& operator,(T v) { m.at<T>(i / m.cols, i % m.cols) = v; i++; return *this; }
    operator Mat() const { return m; }
};
template <class T> struct Mat_ {
    Mat m;
    Mat_(int r, int c) : m(r, c, CV_32F) {}
    MatCommaInit<T> operator<<(T v) { MatCommaInit<T> ci = {m, 0}; ci, v; return ci; }
};
}  // namespace cv
