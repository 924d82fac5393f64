// orbslam_types_min.h — stand-ins for the reference's KeyFrame / Frame / MapPoint / DBoW2::FeatureVector exposing exactly the
// members the hot-path ORBmatcher functions touch (include/KeyFrame.h:66-69,103-105,179-204; include/Frame.h:141-189;
// include/MapPoint.h:73; Thirdparty/DBoW2/DBoW2/FeatureVector.h:21-22).  When ORBmatcher.cc is built inside the reference tree
// (KeyFrame.h on the include path) the real classes are used instead.
#pragma once
#if !defined(ORB_B200_FORCE_MIN_TYPES) && __has_include("KeyFrame.h") && __has_include("Frame.h")
#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"
#else
#include <map>
#include <set>
#include <vector>
#include "cv_compat.h"

#include <cmath>
namespace DBoW2 {
typedef unsigned int NodeId;
typedef unsigned int WordId;
typedef double WordValue;
enum LNorm { L1, L2 };
enum WeightingType { TF_IDF, TF, IDF, BINARY };
enum ScoringType { L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT };
// Thirdparty/DBoW2/DBoW2/BowVector.{h,cpp}
class BowVector : public std::map<WordId, WordValue> {
public:
    void addWeight(WordId id, WordValue v) {
        iterator vit = this->lower_bound(id);
        if (vit != this->end() && !(this->key_comp()(id, vit->first))) vit->second += v;
        else this->insert(vit, value_type(id, v));
    }
    void addIfNotExist(WordId id, WordValue v) {
        iterator vit = this->lower_bound(id);
        if (vit == this->end() || (this->key_comp()(id, vit->first))) this->insert(vit, value_type(id, v));
    }
    void normalize(LNorm norm_type) {
        double norm = 0.0;
        if (norm_type == L1) { for (iterator it = begin(); it != end(); ++it) norm += std::fabs(it->second); }
        else { for (iterator it = begin(); it != end(); ++it) norm += it->second * it->second; norm = std::sqrt(norm); }
        if (norm > 0.0) for (iterator it = begin(); it != end(); ++it) it->second /= norm;
    }
};
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {
public:
    void addFeature(NodeId id, unsigned int i_feature) { (*this)[id].push_back(i_feature); }
};
}  // namespace DBoW2

namespace ORB_SLAM2 {

class KeyFrame;
class MapPoint {
public:
    bool isBad() { return mbBad; }
    int Observations() { return nObs; }                                // include/MapPoint.h:64
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }            // :87
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }                // :58
    // variables used by the tracking, written by Frame::isInFrustum (include/MapPoint.h:106-111)
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0;
    bool mbTrackInView = false;
    int mnTrackScaleLevel = 0;
    float mTrackViewCos = 0;
    cv::Mat GetNormal() { return mNormalVector.clone(); }              // :60
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }  // src/MapPoint.cc:621-631
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    int PredictScale(const float& currentDist, const float& logScaleFactor) {   // src/MapPoint.cc:633-642 (this fork: no clamp)
        const float ratio = mfMaxDistance / currentDist;
        return (int)std::ceil(std::log(ratio) / logScaleFactor);
    }
    // observations (include/MapPoint.h:61-70; bodies after KeyFrame below, as in src/MapPoint.cc:93-116, 226-283)
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
    int GetIndexInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) ? (int)mObservations[pKF] : -1; }
    inline void AddObservation(KeyFrame* pKF, size_t idx);
    inline void Replace(MapPoint* pMP);
    std::map<KeyFrame*, size_t> mObservations;
    bool mbBad = false;
    int nObs = 0;
    float mfMinDistance = 0, mfMaxDistance = 0;
    cv::Mat mDescriptor, mWorldPos, mNormalVector;                     // 1x32 CV_8U, 3x1 CV_32F, 3x1 CV_32F
};

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
class Frame {
public:
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    cv::Mat mTcw;                                                      // 4x4 CV_32F
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mb = 0;
    std::vector<float> mvScaleFactors;
    float mfLogScaleFactor = 0;
    float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;              // static members in the reference (include/Frame.h:192-195)
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;       // static in the reference (:168-169)
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];  // :170
};

class KeyFrame {
public:
    int N = 0;
    float fx = 0, fy = 0, cx = 0, cy = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    std::vector<float> mvScaleFactors, mvLevelSigma2;
    float mfLogScaleFactor = 0;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;                // include/KeyFrame.h:196-199 (ints here, floats in Frame)
    int mnGridCols = 64, mnGridRows = 48;                              // :152-155
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }
    std::vector<MapPoint*> mvpMapPoints;
    cv::Mat Ow, Rcw, tcw;                               // 3x1, 3x3, 3x1 CV_32F
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }                   // src/KeyFrame.cc:620-624
    void ReplaceMapPointMatch(const size_t& idx, MapPoint* pMP) { mvpMapPoints[idx] = pMP; }          // :640-643
    void EraseMapPointMatch(const size_t& idx) { mvpMapPoints[idx] = static_cast<MapPoint*>(NULL); }  // :626-630
    std::set<MapPoint*> GetMapPoints() {                                                              // :645-659
        std::set<MapPoint*> s;
        for (size_t i = 0; i < mvpMapPoints.size(); i++)
            if (mvpMapPoints[i] && !mvpMapPoints[i]->isBad()) s.insert(mvpMapPoints[i]);
        return s;
    }
    std::vector<float> mvInvLevelSigma2;
    float mbf = 0;
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    cv::Mat GetRotation() { return Rcw.clone(); }
    cv::Mat GetTranslation() { return tcw.clone(); }
};

inline void MapPoint::AddObservation(KeyFrame* pKF, size_t idx) {
    if (mObservations.count(pKF)) return;
    mObservations[pKF] = idx;
    if (pKF->mvuRight[idx] >= 0) nObs += 2; else nObs++;
}
inline void MapPoint::Replace(MapPoint* pMP) {
    if (pMP == this) return;
    std::map<KeyFrame*, size_t> obs = mObservations;
    mObservations.clear();
    mbBad = true;
    for (std::map<KeyFrame*, size_t>::iterator mit = obs.begin(); mit != obs.end(); ++mit) {
        KeyFrame* pKF = mit->first;
        if (!pMP->IsInKeyFrame(pKF)) { pKF->ReplaceMapPointMatch(mit->second, pMP); pMP->AddObservation(pKF, mit->second); }
        else pKF->EraseMapPointMatch(mit->second);
    }
}

}  // namespace ORB_SLAM2
#endif
