// prelude.hpp — TEST INFRASTRUCTURE (oracle/_ref MapPoint build), force-included before the reference's src/MapPoint.cc.
// The reference's own include/MapPoint.h and src/MapPoint.cc are compiled UNMODIFIED; what they pull in besides — KeyFrame.h,
// Frame.h, Map.h (Boost, DBoW2 vocabulary, Eigen, g2o ...) and ORBmatcher.h — is switched off by pre-defining the include guards, and
// plain-data stand-ins with the members src/MapPoint.cc touches take their place.  Boost's headers are the recording stand-ins next
// to this file; cv::Mat is the host's stand-in with the float expressions of ../matcher/cv_float.hpp.
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <iostream>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <vector>
#include <opencv2/core/core.hpp>

using namespace std;          // the reference's sources rely on it

#define KEYFRAME_H
#define FRAME_H
#define MAP_H
#define ORBMATCHER_H

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

class Frame {
public:
    long unsigned int mnId = 0;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvScaleFactors;
    int mnScaleLevels = 0;
    cv::Mat mDescriptors, mOw;
    cv::Mat GetCameraCenter() { return mOw.clone(); }
};

class Map {
public:
    std::mutex mMutexPointCreation;
    std::set<MapPoint*> erased;
    void EraseMapPoint(MapPoint* pMP) { erased.insert(pMP); }
};

class ORBmatcher {             // the real one lives in oracle/_ref/libref_orbmatcher.so; the driver forwards to it
public:
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
};

}  // namespace ORB_SLAM2
