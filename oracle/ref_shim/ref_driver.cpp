// ref_driver.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's own ORB_SLAM2::ORBextractor (compiled unmodified
// from /root/reference/src/ORBextractor.cc against cvshim.hpp) for tests/ and bench.py's CPU arm.
#include <vector>
#include "ORBextractor.h"

extern "C" {

void* refx_create(int nfeatures, float scale, int nlevels, int ini, int min) {
    return new ORB_SLAM2::ORBextractor(nfeatures, scale, nlevels, ini, min);
}
void refx_destroy(void* h) { delete (ORB_SLAM2::ORBextractor*)h; }

// kp_out: 7 floats per keypoint (x, y, size, angle, response, octave, class_id); returns the number of keypoints (may exceed cap:
// then only `cap` are written)
int refx_extract(void* h, const unsigned char* img, int w, int hgt, int stride, const unsigned char* mask, int mask_stride,
                 float* kp_out, unsigned char* desc_out, int cap) {
    ORB_SLAM2::ORBextractor* ex = (ORB_SLAM2::ORBextractor*)h;
    cv::Mat image(hgt, w, CV_8UC1, (void*)img, (size_t)stride), m;
    if (mask) m = cv::Mat(hgt, w, CV_8UC1, (void*)mask, (size_t)mask_stride);
    std::vector<cv::KeyPoint> kps;
    cv::Mat desc;
    (*ex)(image, m, kps, desc);
    const int n = (int)kps.size();
    for (int i = 0; i < n && i < cap; i++) {
        float* o = kp_out + 7 * (size_t)i;
        o[0] = kps[i].pt.x; o[1] = kps[i].pt.y; o[2] = kps[i].size; o[3] = kps[i].angle; o[4] = kps[i].response;
        o[5] = (float)kps[i].octave; o[6] = (float)kps[i].class_id;
        for (int c = 0; c < 32; c++) desc_out[32 * (size_t)i + c] = desc.ptr(i)[c];
    }
    return n;
}

int refx_level(void* h, int level, int* w, int* hgt, unsigned char* dst, int dst_stride) {
    ORB_SLAM2::ORBextractor* ex = (ORB_SLAM2::ORBextractor*)h;
    if (level < 0 || level >= (int)ex->mvImagePyramid.size()) return -1;
    const cv::Mat& m = ex->mvImagePyramid[level];
    if (w) *w = m.cols;
    if (hgt) *hgt = m.rows;
    if (dst)
        for (int r = 0; r < m.rows; r++)
            for (int c = 0; c < m.cols; c++) dst[(size_t)r * dst_stride + c] = m.ptr(r)[c];
    return 0;
}

}  // extern "C"
