// ref_types.hpp — TEST INFRASTRUCTURE (oracle/_ref matcher build): stand-ins for the reference's KeyFrame / Frame / MapPoint with
// the members src/ORBmatcher.cc touches, as plain data.  The reference's own headers of those classes need Boost.Serialization,
// DBoW2's vocabulary, Eigen ... and are switched off by their include guards (ORBmatcher.h next to this file).  DBoW2's
// FeatureVector / BowVector are the reference's own (Thirdparty/DBoW2, compiled from where they lie).
#pragma once
#include <cassert>
#include <cmath>
#include <list>
#include <map>
#include <set>
#include <vector>
#include "cv_float.hpp"
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

using namespace std;          // the reference's headers rely on it (include/MapPoint.h etc. say `using namespace std`-style names)

namespace ORB_SLAM2 {

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

class KeyFrame;
class Frame;

class MapPoint {
public:
    long unsigned int mnId = 0;
    bool isBad() { return mbBad; }
    int Observations() { return nObs; }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    cv::Mat GetNormal() { return mNormalVector.clone(); }
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    int PredictScale(const float& currentDist, const float& logScaleFactor) {        // src/MapPoint.cc:633-642 (this fork: no clamp)
        const float ratio = mfMaxDistance / currentDist;
        return (int)std::ceil(std::log(ratio) / logScaleFactor);
    }
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
    int GetIndexInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) ? (int)mObservations[pKF] : -1; }
    void AddObservation(KeyFrame* pKF, size_t idx);
    void Replace(MapPoint* pMP);
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0;
    bool mbTrackInView = false;
    int mnTrackScaleLevel = 0;
    float mTrackViewCos = 0;
    long unsigned int mnLastFrameSeen = 0, mnFuseCandidateForKF = 0;
    std::map<KeyFrame*, size_t> mObservations;
    bool mbBad = false;
    int nObs = 0;
    float mfMinDistance = 0, mfMaxDistance = 0;
    cv::Mat mDescriptor, mWorldPos, mNormalVector;
};

class Frame {
public:
    long unsigned int mnId = 0;
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    cv::Mat mTcw;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mb = 0;
    int mnScaleLevels = 0;
    std::vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    float mfLogScaleFactor = 0;
    float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    // restated from src/Frame.cc:445-498 (that translation unit cannot be compiled here); not used by the pinned functions
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const;
};

class KeyFrame {
public:
    long unsigned int mnId = 0;
    int N = 0;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    int mnScaleLevels = 0;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    float mfLogScaleFactor = 0;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0, mnGridCols = 64, mnGridRows = 48;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    std::vector<std::vector<std::vector<size_t> > > mGrid;
    bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }
    std::vector<MapPoint*> mvpMapPoints;
    cv::Mat Ow, Rcw, tcw;
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
    void ReplaceMapPointMatch(const size_t& idx, MapPoint* pMP) { mvpMapPoints[idx] = pMP; }
    void EraseMapPointMatch(const size_t& idx) { mvpMapPoints[idx] = static_cast<MapPoint*>(NULL); }
    std::set<MapPoint*> GetMapPoints() {
        std::set<MapPoint*> s;
        for (size_t i = 0; i < mvpMapPoints.size(); i++)
            if (mvpMapPoints[i] && !mvpMapPoints[i]->isBad()) s.insert(mvpMapPoints[i]);
        return s;
    }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    cv::Mat GetRotation() { return Rcw.clone(); }
    cv::Mat GetTranslation() { return tcw.clone(); }
    // restated from src/KeyFrame.cc:1311-1350; not used by the pinned functions
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const;
};

}  // namespace ORB_SLAM2
