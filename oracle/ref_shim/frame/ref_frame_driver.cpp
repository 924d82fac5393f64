// ref_frame_driver.cpp — TEST INFRASTRUCTURE.  C entry point around the reference's own ORB_SLAM2::Frame::ComputeStereoMatches
// (/root/reference/src/Frame.cc:584-756 compiled unmodified together with include/Frame.h, src/MapPoint.cc, include/MapPoint.h,
// ORBVocabulary.h / DBoW2; see prelude.hpp for what is a stand-in).  The two extractors are inputs of that function (it reads their
// public mvImagePyramid); DescriptorDistance is the reference's own, forwarded to oracle/_ref/libref_orbmatcher.so.
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <string>
#define private public          // Frame::AssignFeaturesToGrid is private; access only, the layout is that of src/Frame.cc's own build
#include "Frame.h"
#undef private

namespace ORB_SLAM2 {
int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    typedef int (*fn_t)(const unsigned char*, const unsigned char*);
    static fn_t fn = [] {
        Dl_info info;
        dladdr((void*)&ORBmatcher::DescriptorDistance, &info);
        std::string dir(info.dli_fname);
        dir = dir.substr(0, dir.find_last_of('/'));
        void* h = dlopen((dir + "/libref_orbmatcher.so").c_str(), RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND);
        fn_t f = h ? (fn_t)dlsym(h, "refm_descriptor_distance") : nullptr;
        if (!f) { fprintf(stderr, "ref_frame: cannot bind refm_descriptor_distance: %s\n", dlerror()); abort(); }
        return f;
    }();
    return fn(a.ptr(), b.ptr());
}
}  // namespace ORB_SLAM2

using namespace ORB_SLAM2;

struct reff_kp { float x, y, size, angle, response; int octave, class_id; };      // = cv::KeyPoint = orbx_keypoint

static void fill_side(std::vector<cv::KeyPoint>& keys, cv::Mat& desc, const reff_kp* kp, const unsigned char* d, int n) {
    keys.resize(n);
    desc.create(n > 0 ? n : 1, 32, CV_8U);
    desc.rows = n;
    for (int i = 0; i < n; i++) {
        keys[i] = cv::KeyPoint(kp[i].x, kp[i].y, kp[i].size, kp[i].angle, kp[i].response, kp[i].octave, kp[i].class_id);
        memcpy(desc.ptr(i), d + 32 * (size_t)i, 32);
    }
}

extern "C" void reff_stereo_matches(int nlevels, const unsigned char* const* left, const unsigned char* const* right, const int* w, const int* h,
                                    const float* sf, const float* isf, const reff_kp* kpL, const unsigned char* dL, int nL, const reff_kp* kpR,
                                    const unsigned char* dR, int nR, float mbf, float mb, float* uright, float* depth) {
    ORBextractor exL, exR;
    for (int l = 0; l < nlevels; l++) {
        exL.mvImagePyramid.push_back(cv::Mat(h[l], w[l], CV_8U, (void*)left[l], (size_t)w[l]));
        exR.mvImagePyramid.push_back(cv::Mat(h[l], w[l], CV_8U, (void*)right[l], (size_t)w[l]));
    }
    Frame F;
    F.mpORBextractorLeft = &exL;
    F.mpORBextractorRight = &exR;
    F.N = nL;
    fill_side(F.mvKeys, F.mDescriptors, kpL, dL, nL);
    fill_side(F.mvKeysRight, F.mDescriptorsRight, kpR, dR, nR);
    F.mvScaleFactors.assign(sf, sf + nlevels);
    F.mvInvScaleFactors.assign(isf, isf + nlevels);
    F.mnScaleLevels = nlevels;
    F.mbf = mbf;
    F.mb = mb;
    F.ComputeStereoMatches();
    for (int i = 0; i < nL; i++) { uright[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
}

// Frame::AssignFeaturesToGrid / PosInGrid (src/Frame.cc:341-356, 500-510) and Frame::GetFeaturesInArea (:445-498) of the reference on n
// undistorted keypoints: the grid as CSR (cell = ix * 48 + iy, entries in push order) and, for nq queries (x, y, r, minLevel, maxLevel),
// the returned index lists back to back (q_off[nq + 1]).
extern "C" int reff_grid_and_areas(int n, const float* x, const float* y, const int* octave, const float* bounds /*minx miny maxx maxy*/,
                                   int* cell_offsets, int* cell_features, int nq, const float* q /*nq x 3*/, const int* qlev /*nq x 2*/,
                                   int* q_off, int* q_idx, int cap) {
    Frame F;
    F.N = n;
    F.mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) F.mvKeysUn[i] = cv::KeyPoint(x[i], y[i], 31.f, 0.f, 0.f, octave[i], -1);
    Frame::mnMinX = bounds[0]; Frame::mnMinY = bounds[1]; Frame::mnMaxX = bounds[2]; Frame::mnMaxY = bounds[3];
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);      // :101-102
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
    F.AssignFeaturesToGrid();
    int e = 0;
    for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
        for (int iy = 0; iy < FRAME_GRID_ROWS; iy++) {
            cell_offsets[ix * FRAME_GRID_ROWS + iy] = e;
            for (size_t j = 0; j < F.mGrid[ix][iy].size(); j++) cell_features[e++] = (int)F.mGrid[ix][iy][j];
        }
    cell_offsets[FRAME_GRID_COLS * FRAME_GRID_ROWS] = e;
    int o = 0;
    for (int k = 0; k < nq; k++) {
        q_off[k] = o;
        const std::vector<size_t> v = F.GetFeaturesInArea(q[3 * k], q[3 * k + 1], q[3 * k + 2], qlev[2 * k], qlev[2 * k + 1]);
        for (size_t j = 0; j < v.size() && o < cap; j++) q_idx[o++] = (int)v[j];
    }
    q_off[nq] = o;
    return e;
}
