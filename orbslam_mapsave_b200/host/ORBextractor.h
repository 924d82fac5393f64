// ORBextractor.h — drop-in for the reference's include/ORBextractor.h:50-116 (class ORB_SLAM2::ORBextractor):
// same constructor, operator()(image, mask, keypoints, descriptors), getters and public mvImagePyramid.
// The body marshals to the C ABI of include/orb_b200.h (CUDA, sm_100a); there is no CPU path.
#ifndef ORB_B200_ORBEXTRACTOR_H
#define ORB_B200_ORBEXTRACTOR_H

#include <vector>
#include "cv_compat.h"

struct orbx_extractor;

namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };                                           // reference :54

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);   // reference :56-57
    ~ORBextractor();
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image; the mask (this fork) zeroes masked-out pixels first.
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
                    cv::OutputArray descriptors);                                        // reference :64-66

    // Batch form (not in the reference): N images of one size, each its own cv::Mat, in ONE pipelined call (orbx_extract_batch_ptrs:
    // copies overlapped with the kernels, 128-frame device passes).  keypoints[i] / descriptors[i] equal what operator() returns for
    // images[i] without a mask; mvImagePyramid is left alone.
    void operator()(const std::vector<cv::Mat>& images, std::vector<std::vector<cv::KeyPoint> >& keypoints, std::vector<cv::Mat>& descriptors);

    int inline GetLevels() { return nlevels; }                                            // :68-69
    float inline GetScaleFactor() { return scaleFactor; }                                 // :71-72
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }                 // :74-76
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }       // :78-80
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }            // :82-84
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }  // :86-88

    std::vector<cv::Mat> mvImagePyramid;                                                  // :90 (ROI views into bordered buffers)

    // --- additions (not in the reference) ---
    void SetDevice(int device) { mDevice = device; }          // CUDA ordinal, before the first call
    void SetPyramidDownload(bool on) { mbDownloadPyramid = on; }   // mvImagePyramid costs 8 device->host copies per frame;
                                                                    // only Frame::ComputeStereoMatches reads it
    int LastStatus() const { return mLastStatus; }             // orb_status of the last call (the reference API has no error channel)
    // The body of Frame::ComputeStereoMatches (src/Frame.cc:584-756) on the device-resident pyramids of the two extractors'
    // last operator() calls: fills mvuRight / mvDepth exactly like the reference.  With this, a stereo Frame can call
    // SetPyramidDownload(false) on both extractors (nothing else reads mvImagePyramid).
    static void ComputeStereoMatches(ORBextractor* left, ORBextractor* right, const std::vector<cv::KeyPoint>& mvKeys,
                                     const cv::Mat& mDescriptors, const std::vector<cv::KeyPoint>& mvKeysRight,
                                     const cv::Mat& mDescriptorsRight, float mbf, float mb, std::vector<float>& mvuRight,
                                     std::vector<float>& mvDepth);

protected:
    void Plan(int width, int height);

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    orbx_extractor* mHandle = nullptr;
    orbx_extractor* mBatchHandle = nullptr;   // 128-frame workspace of the batch form, created on its first call
    int mBatchW = 0, mBatchH = 0;
    int mPlanW = 0, mPlanH = 0, mDevice = 0, mLastStatus = 0;
    bool mbDownloadPyramid = true;
};

}  // namespace ORB_SLAM2
#endif
