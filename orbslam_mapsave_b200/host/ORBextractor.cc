// ORBextractor.cc — host side of the drop-in class: marshals ORB_SLAM2::ORBextractor onto the C ABI (include/orb_b200.h).
// Replaces src/ORBextractor.cc of the reference; every computation (tables included) happens in liborb_b200.so.
#include "ORBextractor.h"

#include <cassert>
#include <cstdio>
#include <cstring>

#include "../../include/orb_b200.h"

namespace ORB_SLAM2 {

static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint must be the 28-byte POD the C ABI writes");

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
    mvImagePyramid.resize(nlevels);
    // The tables of the reference's constructor (ORBextractor.cc:414-445) depend on the parameters only: no device memory, no plan.
    // The workspace is planned by the first operator() call, for the size of the image it is given.
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    mLastStatus = orbx_compute_tables(nfeatures, (float)scaleFactor, nlevels, mvScaleFactor.data(), mvInvScaleFactor.data(),
                                      mvLevelSigma2.data(), mvInvLevelSigma2.data(), mnFeaturesPerLevel.data());
    if (mLastStatus != ORB_OK) std::fprintf(stderr, "ORBextractor: %s\n", orb_last_error());
}

ORBextractor::~ORBextractor() { orbx_destroy(mHandle); orbx_destroy(mBatchHandle); }

void ORBextractor::Plan(int width, int height) {
    if (mHandle && width == mPlanW && height == mPlanH) return;
    orbx_destroy(mHandle);
    mHandle = nullptr;
    mLastStatus = orbx_create(&mHandle, nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST, width, height, 1, mDevice);
    if (mLastStatus != ORB_OK) {
        std::fprintf(stderr, "ORBextractor: %s\n", orb_last_error());
        return;
    }
    mPlanW = width; mPlanH = height;
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    orbx_tables(mHandle, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(),
                mnFeaturesPerLevel.data());
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors) {
    if (_image.empty()) return;                                  // reference :1045-1046: outputs untouched
    cv::Mat image = _image.getMat();
    cv::Mat mask = _mask.getMat();
    assert(image.type() == CV_8UC1);                             // reference :1051
    Plan(image.cols, image.rows);
    if (!mHandle) return;

    const int cap = orbx_max_keypoints(mHandle);
    std::vector<cv::KeyPoint> kps(cap);
    cv::Mat desc(cap, 32, CV_8U);
    int n = 0;
    mLastStatus = orbx_extract(mHandle, image.data, image.cols, image.rows, (int)image.step,
                               mask.empty() ? nullptr : mask.data, mask.empty() ? 0 : (int)mask.step,
                               reinterpret_cast<orbx_keypoint*>(kps.data()), desc.data, cap, &n);
    if (mLastStatus != ORB_OK) {
        std::fprintf(stderr, "ORBextractor: %s\n", orb_last_error());
        return;
    }
    if (n == 0) _descriptors.release();                          // reference :1067-1073
    else {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat out = _descriptors.getMat();
        for (int i = 0; i < n; i++) std::memcpy(out.ptr(i), desc.ptr(i), 32);
    }
    _keypoints.assign(kps.begin(), kps.begin() + n);             // reference :1075-1106

    if (mbDownloadPyramid) {                                     // reference :1110-1135: ROI views inside (w+38)x(h+38) buffers
        mvImagePyramid.resize(nlevels);
        std::vector<cv::Mat> whole(nlevels);
        std::vector<uint8_t*> ptrs(nlevels);
        std::vector<int> strides(nlevels), ws(nlevels), hs(nlevels);
        for (int l = 0; l < nlevels; l++) {
            orbx_level_size(mHandle, l, &ws[l], &hs[l]);
            whole[l] = cv::Mat(hs[l] + 38, ws[l] + 38, CV_8U);
            ptrs[l] = whole[l].data;
            strides[l] = (int)whole[l].step;
        }
        if (orbx_get_pyramid(mHandle, 0, 1, ptrs.data(), strides.data()) == ORB_OK)
            for (int l = 0; l < nlevels; l++) {
                mvImagePyramid[l] = whole[l].rowRange(19, 19 + hs[l]).colRange(19, 19 + ws[l]);      // the ROI inside the bordered buffer
            }
    }
}


void ORBextractor::operator()(const std::vector<cv::Mat>& images, std::vector<std::vector<cv::KeyPoint> >& keypoints,
                              std::vector<cv::Mat>& descriptors) {
    const int N = (int)images.size();
    keypoints.assign(N, std::vector<cv::KeyPoint>());
    descriptors.assign(N, cv::Mat());
    if (N == 0) return;
    const int w = images[0].cols, h = images[0].rows;
    const int step = (int)images[0].step;
    std::vector<const uint8_t*> ptrs(N);
    for (int i = 0; i < N; i++) {
        assert(images[i].type() == CV_8UC1 && images[i].cols == w && images[i].rows == h && (int)images[i].step == step);
        ptrs[i] = images[i].data;
    }
    if (!mBatchHandle || w != mBatchW || h != mBatchH) {
        orbx_destroy(mBatchHandle);
        mBatchHandle = nullptr;
        mLastStatus = orbx_create(&mBatchHandle, nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST, w, h, 128, mDevice);
        if (mLastStatus != ORB_OK) { std::fprintf(stderr, "ORBextractor: %s\n", orb_last_error()); return; }
        mBatchW = w; mBatchH = h;
    }
    const int cap = orbx_max_keypoints(mBatchHandle);
    std::vector<cv::KeyPoint> kps((size_t)N * cap);
    std::vector<uint8_t> desc((size_t)N * cap * 32);
    std::vector<int> n(N, 0);
    mLastStatus = orbx_extract_batch_ptrs(mBatchHandle, ptrs.data(), N, w, h, step, reinterpret_cast<orbx_keypoint*>(kps.data()), desc.data(), cap,
                                          n.data());
    if (mLastStatus != ORB_OK) { std::fprintf(stderr, "ORBextractor: %s\n", orb_last_error()); return; }
    for (int i = 0; i < N; i++) {
        keypoints[i].assign(kps.begin() + (size_t)i * cap, kps.begin() + (size_t)i * cap + n[i]);
        if (n[i] > 0) {
            descriptors[i] = cv::Mat(n[i], 32, CV_8U);
            for (int r = 0; r < n[i]; r++) std::memcpy(descriptors[i].ptr(r), &desc[((size_t)i * cap + r) * 32], 32);
        }
    }
}

void ORBextractor::ComputeStereoMatches(ORBextractor* left, ORBextractor* right, const std::vector<cv::KeyPoint>& mvKeys,
                                        const cv::Mat& mDescriptors, const std::vector<cv::KeyPoint>& mvKeysRight,
                                        const cv::Mat& mDescriptorsRight, float mbf, float mb, std::vector<float>& mvuRight,
                                        std::vector<float>& mvDepth) {
    const int N = (int)mvKeys.size(), Nr = (int)mvKeysRight.size();
    mvuRight = std::vector<float>(N, -1.0f);                                              // src/Frame.cc:586-587
    mvDepth = std::vector<float>(N, -1.0f);
    if (N == 0 || !left->mHandle || !right->mHandle) return;
    std::vector<unsigned char> dl((size_t)N * 32), dr((size_t)(Nr > 0 ? Nr : 1) * 32);
    for (int i = 0; i < N; i++) std::memcpy(&dl[(size_t)i * 32], mDescriptors.ptr(i), 32);
    for (int i = 0; i < Nr; i++) std::memcpy(&dr[(size_t)i * 32], mDescriptorsRight.ptr(i), 32);
    left->mLastStatus = orbx_stereo_matches(left->mHandle, right->mHandle, 0, 0, reinterpret_cast<const orbx_keypoint*>(mvKeys.data()),
                                            dl.data(), N, reinterpret_cast<const orbx_keypoint*>(mvKeysRight.data()), dr.data(), Nr,
                                            mbf, mb, mvuRight.data(), mvDepth.data());
    if (left->mLastStatus != ORB_OK) std::fprintf(stderr, "ORBextractor::ComputeStereoMatches: %s\n", orb_last_error());
}

}  // namespace ORB_SLAM2
