// prelude.hpp — TEST INFRASTRUCTURE (oracle/_ref Frame build), force-included before the reference's src/Frame.cc and src/MapPoint.cc.
// Compiled UNMODIFIED: src/Frame.cc, src/MapPoint.cc, include/Frame.h, include/MapPoint.h, include/ORBVocabulary.h and DBoW2.
// Switched off by their include guards and replaced by plain-data stand-ins: KeyFrame.h, Map.h, ORBextractor.h (the extractor is an
// INPUT of Frame::ComputeStereoMatches: it only reads the public mvImagePyramid), ORBmatcher.h (DescriptorDistance is forwarded to the
// reference's own in libref_orbmatcher.so), Converter.h (Eigen / g2o; Frame.cc uses toDescriptorVector only).
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <iostream>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <thread>
#include <vector>
#include <opencv2/core/core.hpp>

using namespace std;          // the reference's sources rely on it

#define KEYFRAME_H
#define MAP_H
#define ORBMATCHER_H
#define ORBEXTRACTOR_H
#define CONVERTER_H

namespace ORB_SLAM2 {

class MapPoint;

class KeyFrame {
public:
    long unsigned int mnId = 0, mnFrameId = 0;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight, mvScaleFactors;
    int mnScaleLevels = 0;
    cv::Mat mDescriptors, Ow;
    bool mbBad = false;
    std::vector<MapPoint*> mvpMapPoints;
    bool isBad() { return mbBad; }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    void EraseMapPointMatch(const size_t& idx) { mvpMapPoints[idx] = static_cast<MapPoint*>(NULL); }
    void ReplaceMapPointMatch(const size_t& idx, MapPoint* pMP) { mvpMapPoints[idx] = pMP; }
};

class Map {
public:
    std::mutex mMutexPointCreation;
    void EraseMapPoint(MapPoint*) {}
};

class ORBextractor {           // include/ORBextractor.h:50-116, as far as Frame.cc uses it
public:
    std::vector<cv::Mat> mvImagePyramid;
    int nlevels = 0;
    float scaleFactor = 1.f;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    void operator()(cv::InputArray, cv::InputArray, std::vector<cv::KeyPoint>&, cv::OutputArray) { std::abort(); }
    int GetLevels() { return nlevels; }
    float GetScaleFactor() { return scaleFactor; }
    std::vector<float> GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }
};

class ORBmatcher {
public:
    static const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
};

class Converter {
public:
    static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& Descriptors) {
        std::vector<cv::Mat> v;
        for (int j = 0; j < Descriptors.rows; j++) v.push_back(Descriptors.row(j));
        return v;
    }
};

}  // namespace ORB_SLAM2
