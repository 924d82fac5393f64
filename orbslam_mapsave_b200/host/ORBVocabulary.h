// ORBVocabulary.h — the slice of the reference's ORBVocabulary (include/ORBVocabulary.h: a
// DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>) that feeds the matcher: loading ORBvoc.txt / ORBvoc.bin and
// transform(features, BowVector&, FeatureVector&, levelsup) as called by Frame::ComputeBoW (src/Frame.cc:513-520) and
// KeyFrame::ComputeBoW (src/KeyFrame.cc:781-790).  The tree descent (60 Hamming distances per descriptor) runs on the GPU
// through orbv_transform; the two std::map containers are filled on the host in feature order, exactly like DBoW2
// (TemplatedVocabulary.h:1140-1207), so BowVector sums and the L1/L2 normalisation are bit-identical.
#ifndef ORB_B200_ORBVOCABULARY_H
#define ORB_B200_ORBVOCABULARY_H

#include <string>
#include <vector>
#include "cv_compat.h"
#include "orbslam_types_min.h"

struct orbv_vocabulary;

namespace ORB_SLAM2 {

class ORBVocabularyB200 {
public:
    ORBVocabularyB200() {}
    ~ORBVocabularyB200();
    ORBVocabularyB200(const ORBVocabularyB200&) = delete;
    ORBVocabularyB200& operator=(const ORBVocabularyB200&) = delete;

    bool loadFromTextFile(const std::string& filename);                       // TemplatedVocabulary.h:1351-1440
    bool loadFromBinaryFile(const std::string& filename);                     // :1467-1512
    void saveToBinaryFile(const std::string& filename) const;                 // :1515-1536
    bool empty() const { return mHandle == nullptr; }
    unsigned int size() const;                                                // number of words

    // transform(features, v, fv, levelsup)                                   // :1140-1207
    void transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const;

    void SetDevice(int device) { mDevice = device; }
    int LastStatus() const { return mLastStatus; }

private:
    orbv_vocabulary* mHandle = nullptr;
    int mDevice = 0;
    mutable int mLastStatus = 0;
};

}  // namespace ORB_SLAM2
#endif
