// MapArchive.h — the fork's saved map (System::SaveMap / LoadMap, src/System.cc:552-574) as plain C++ data, without Boost or
// the SLAM object graph: what Map::load (src/Map.cc:76-134), KeyFrame::load (src/KeyFrame.cc:308-510) and MapPoint::load
// (src/MapPoint.cc:142-213) read, handed out with the reference's member names.  Parsing is done by the orbmap_* entry points
// of include/orb_b200.h (byte layout: orbslam_mapsave_b200/csrc/orb_map.cpp; unpinned against a real Boost build).  The
// descriptor tables it returns are the inputs of the GPU matcher entry points (orbm_hamming_top2, orbm_allpairs_device,
// orbm_distinctive_descriptors, orbv_transform).
#ifndef ORB_B200_MAPARCHIVE_H
#define ORB_B200_MAPARCHIVE_H

#include <string>
#include <utility>
#include <vector>
#include "cv_compat.h"

struct orbmap_archive;

namespace ORB_SLAM2 {

class MapArchiveB200 {
public:
    struct KeyFrameData {                        // names as in include/KeyFrame.h
        long unsigned int mnId = 0, mnFrameId = 0;
        double mTimeStamp = 0;
        int N = 0;
        std::vector<cv::KeyPoint> mvKeys, mvKeysUn;          // `size` is 0: the fork's serializer never stores it
        std::vector<float> mvuRight, mvDepth;
        cv::Mat mDescriptors;                                // N x 32, CV_8U
        std::vector<long> mvpMapPointIds;                    // MapPoint::mnId per feature, -1 = none (mmMapPoints_nId)
        int mnScaleLevels = 0;
        float mfScaleFactor = 0, mfLogScaleFactor = 0;
        std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
        float fx = 0, fy = 0, cx = 0, cy = 0, invfx = 0, invfy = 0, mbf = 0, mb = 0, mThDepth = 0;
        int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0, mnGridCols = 0, mnGridRows = 0;
        float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
        cv::Mat Tcw, mK;                                     // 4x4 / 3x3, CV_32F
        bool hasParent = false, mbBad = false;
        long unsigned int parentId = 0;                      // mparent_KfId_map
        std::vector<std::pair<long, int> > mConnectedKeyFrameWeights;   // (mnId or -1, weight)
        std::vector<long> mvpOrderedConnectedKeyFrames, mspChildrens, mspLoopEdges;
        std::vector<int> mvOrderedWeights;
        std::vector<int> gridOffsets, gridFeatures;          // mGrid as CSR, cells column-major like mGrid[col][row]
    };
    struct MapPointData {                        // names as in include/MapPoint.h
        long unsigned int mnId = 0;
        cv::Mat mWorldPos, mNormalVector;                    // 3x1, CV_32F
        cv::Mat mDescriptor;                                 // 1 x 32, CV_8U
        long refKFId = -1;                                   // mref_KfId_pair
        int nObs = 0, mnVisible = 0, mnFound = 0;
        bool mbBad = false;
        float mfMinDistance = 0, mfMaxDistance = 0;
        std::vector<std::pair<long, long> > mObservations;   // (KeyFrame::mnId, feature index); -1 = entry stored without an id
    };

    MapArchiveB200() {}
    ~MapArchiveB200();
    MapArchiveB200(const MapArchiveB200&) = delete;
    MapArchiveB200& operator=(const MapArchiveB200&) = delete;

    bool Load(const std::string& filename);                  // System::LoadMap; false + LastError() on failure
    bool Save(const std::string& filename) const;            // System::SaveMap (byte-identical for a loaded file)
    const std::string& LastError() const { return mError; }

    long unsigned int KeyFramesInMap() const;                // Map::KeyFramesInMap
    long unsigned int MapPointsInMap() const;                // Map::MapPointsInMap
    long unsigned int GetMaxKFid() const;                    // Map::GetMaxKFid
    bool LoadValidated() const;                              // the 0xdeadbeef check of Map::load (src/Map.cc:127-131)

    bool GetKeyFrame(size_t i, KeyFrameData& out) const;     // i = position in the file (std::set order at save time)
    std::vector<MapPointData> GetAllMapPoints() const;

    // The gather loop of MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:495-510) for all map points: descriptors
    // (total x 32) and offsets (MapPointsInMap() + 1), the ragged batch orbm_distinctive_descriptors takes.
    bool ObservedDescriptors(cv::Mat& descriptors, std::vector<int>& offsets) const;

    const orbmap_archive* Handle() const { return mHandle; }

private:
    orbmap_archive* mHandle = nullptr;
    mutable std::string mError;
};

}  // namespace ORB_SLAM2
#endif
