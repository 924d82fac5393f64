// ORBmatcher.h — drop-in for the hot-path members of the reference's include/ORBmatcher.h:37-101:
// DescriptorDistance, both SearchByBoW overloads, SearchForTriangulation (+ the constants), and — first "next" row of
// SURVEY.md §8f — the window searches SearchByProjection(Frame, MapPoints), SearchByProjection(CurrentFrame, LastFrame) and
// SearchForInitialization, the two projection searches of relocalisation / loop closing, SearchBySim3 and Fuse x2.  Same signatures; bodies marshal
// to the C ABI of include/orb_b200.h.  SearchBySim3 and Fuse (x2) search on the device and apply their map updates on the host.
#ifndef ORB_B200_ORBMATCHER_H
#define ORB_B200_ORBMATCHER_H

#include <set>
#include <utility>
#include <vector>
#include "cv_compat.h"
#include "orbslam_types_min.h"

namespace ORB_SLAM2 {

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);                                // reference :41

    // Hamming distance between two ORB descriptors (reference :44, src/ORBmatcher.cc:1650-1666)
    // src/ORBmatcher.cc:1650-1666.  NOTE: one call is one device round trip (two 32-byte uploads, a kernel, a 4-byte download): right for
    // the odd stray call, wrong inside a loop — loops over descriptor pairs belong in DescriptorDistances below (one round trip for all
    // pairs) or in the searches of this class.  Returns 257 (more than any distance) if the device call fails; see LastStatus().
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    // Batched form (not in the reference): out[i] = distance(A.row(i), B.row(i)) for i < min(A.rows, B.rows); returns that count.
    static int DescriptorDistances(const cv::Mat& A, const cv::Mat& B, int* out);

    // Brute force constrained to ORB that belong to the same vocabulary node (reference :65-66)
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);
    // Batched overloads (not in the reference): the candidate loops of Tracking::Relocalization (src/Tracking.cc:1621-1643) and
    // LoopClosing::ComputeSim3 (src/LoopClosing.cc:240-266) as one device call; element i equals the single overload on candidate i.
    std::vector<int> SearchByBoW(const std::vector<KeyFrame*>& vpKFs, Frame& F, std::vector<std::vector<MapPoint*> >& vvpMapPointMatches);
    std::vector<int> SearchByBoW(KeyFrame* pKF1, const std::vector<KeyFrame*>& vpKF2s, std::vector<std::vector<MapPoint*> >& vvpMatches12);

    // Search matches between Frame keypoints and projected MapPoints. Returns number of matches.  Used to track the local map
    // (Tracking) (reference :48, src/ORBmatcher.cc:45-129)
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3);
    // Project MapPoints tracked in last frame into the current frame and search matches.  Used to track from previous frame
    // (Tracking) (reference :52, src/ORBmatcher.cc:1331-1463)
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
    // Batched overload (not in the reference): independent (CurrentFrame, LastFrame) pairs — several cameras or sessions tracked by one
    // process — as one device call; element i equals the single overload on pair i.
    std::vector<int> SearchByProjection(const std::vector<Frame*>& vpCurrentFrames, const std::vector<const Frame*>& vpLastFrames,
                                        const float th, const bool bMono);
    // Project MapPoints seen in KeyFrame into the Frame and search matches.  Used in relocalisation (Tracking)
    // (reference :56, src/ORBmatcher.cc:1465-1602)
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist);
    // Project MapPoints using a Similarity Transformation and search matches.  Used in loop detection (Loop Closing)
    // (reference :60, src/ORBmatcher.cc:293-406)
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th);
    // Search matches between MapPoints seen in KF1 and KF2 transforming by a Sim3 [s12*R12|t12]
    // (reference :77, src/ORBmatcher.cc:1105-1329)
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                     const cv::Mat& t12, const float th);
    // Project MapPoints into KeyFrame and search for duplicated MapPoints (reference :80, src/ORBmatcher.cc:828-972)
    int Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th = 3.0);
    // Project MapPoints into KeyFrame using a given Sim3 and search for duplicated MapPoints (reference :83, :974-1103)
    int Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint);
    // Matching for the Map Initialization (only used in the monocular case) (reference :69, src/ORBmatcher.cc:408-523)
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                                int windowSize = 10);

    // Matching to triangulate new MapPoints. Check Epipolar Constraint. (reference :72-73)
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs,
                               const bool bOnlyStereo);

    // Batched overload (not in the reference): the loop over the neighbours of LocalMapping::CreateNewMapPoints
    // (src/LocalMapping.cc:215-268) as one device call; element i equals the single overload on (pKF1, vpKF2s[i], vF12[i]).
    std::vector<int> SearchForTriangulation(KeyFrame* pKF1, const std::vector<KeyFrame*>& vpKF2s, const std::vector<cv::Mat>& vF12,
                                            std::vector<std::vector<std::pair<size_t, size_t> > >& vvMatchedPairs, const bool bOnlyStereo);

    static const int TH_LOW;                                                              // reference :87-89
    static const int TH_HIGH;
    static const int HISTO_LENGTH;

    static void SetDevice(int device);           // addition: CUDA ordinal used by all matcher calls of this process
    static int LastStatus();                     // addition: orb_status of this thread's last call

protected:
    float mfNNratio;
    bool mbCheckOrientation;
};

}  // namespace ORB_SLAM2
#endif
