// ORBmatcher.h — drop-in for the hot-path members of the reference's include/ORBmatcher.h:37-101:
// DescriptorDistance, both SearchByBoW overloads, SearchForTriangulation (+ the constants).  Same signatures; bodies marshal
// to the C ABI of include/orb_b200.h.  The projection / fuse / Sim3 searches of the reference class are outside this path
//  and stay in the reference's own ORBmatcher.cc — see INTEGRATION.md.
#ifndef ORB_B200_ORBMATCHER_H
#define ORB_B200_ORBMATCHER_H

#include <utility>
#include <vector>
#include "cv_compat.h"
#include "orbslam_types_min.h"

namespace ORB_SLAM2 {

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);                                // reference :41

    // Hamming distance between two ORB descriptors (reference :44, src/ORBmatcher.cc:1650-1666)
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);

    // Brute force constrained to ORB that belong to the same vocabulary node (reference :65-66)
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);

    // Matching to triangulate new MapPoints. Check Epipolar Constraint. (reference :72-73)
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs,
                               const bool bOnlyStereo);

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
