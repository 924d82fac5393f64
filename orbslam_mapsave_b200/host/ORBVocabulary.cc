// ORBVocabulary.cc — host side of the vocabulary drop-in (see ORBVocabulary.h).
#include "ORBVocabulary.h"

#include <cstdio>
#include <cstring>

#include "../../include/orb_b200.h"

namespace ORB_SLAM2 {

ORBVocabularyB200::~ORBVocabularyB200() { orbv_destroy(mHandle); }

static bool report(int rc, int& last) {
    last = rc;
    if (rc != ORB_OK) std::fprintf(stderr, "ORBVocabulary: %s\n", orb_last_error());
    return rc == ORB_OK;
}

bool ORBVocabularyB200::loadFromTextFile(const std::string& filename) {
    orbv_destroy(mHandle);
    mHandle = nullptr;
    return report(orbv_load_text(&mHandle, filename.c_str(), mDevice), mLastStatus);
}
bool ORBVocabularyB200::loadFromBinaryFile(const std::string& filename) {
    orbv_destroy(mHandle);
    mHandle = nullptr;
    return report(orbv_load_binary(&mHandle, filename.c_str(), mDevice), mLastStatus);
}
void ORBVocabularyB200::saveToBinaryFile(const std::string& filename) const {
    if (mHandle) report(orbv_save_binary(mHandle, filename.c_str()), mLastStatus);
}
unsigned int ORBVocabularyB200::size() const {
    int nw = 0;
    if (mHandle) orbv_info(mHandle, nullptr, nullptr, nullptr, &nw, nullptr, nullptr);
    return (unsigned)nw;
}

void ORBVocabularyB200::transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v, DBoW2::FeatureVector& fv,
                                  int levelsup) const {
    v.clear();
    fv.clear();
    if (empty()) return;                                           // TemplatedVocabulary.h:1147-1150
    const int n = (int)features.size();
    std::vector<unsigned char> desc((size_t)n * 32);
    for (int i = 0; i < n; i++) std::memcpy(&desc[(size_t)i * 32], features[i].ptr<unsigned char>(), 32);
    std::vector<int> word(n), node(n);
    std::vector<double> weight(n);
    if (!report(orbv_transform(mHandle, desc.data(), n, levelsup, word.data(), weight.data(), node.data()), mLastStatus)) return;
    int scoring = 0, weighting = 0;
    orbv_info(mHandle, nullptr, nullptr, nullptr, nullptr, &scoring, &weighting);
    // mustNormalize of the scoring classes (ScoringObject.h): L1 for L1/CHI_SQUARE/KL/BHATTACHARYYA, L2 for L2_NORM, none for DOT_PRODUCT
    const bool must = scoring != DBoW2::DOT_PRODUCT;
    const DBoW2::LNorm norm = scoring == DBoW2::L2_NORM ? DBoW2::L2 : DBoW2::L1;
    if (weighting == DBoW2::TF || weighting == DBoW2::TF_IDF) {
        for (int i = 0; i < n; i++)
            if (weight[i] > 0) { v.addWeight((DBoW2::WordId)word[i], weight[i]); fv.addFeature((DBoW2::NodeId)node[i], (unsigned)i); }
        if (!v.empty() && !must) {
            const double nd = (double)v.size();
            for (DBoW2::BowVector::iterator vit = v.begin(); vit != v.end(); ++vit) vit->second /= nd;
        }
    } else {
        for (int i = 0; i < n; i++)
            if (weight[i] > 0) { v.addIfNotExist((DBoW2::WordId)word[i], weight[i]); fv.addFeature((DBoW2::NodeId)node[i], (unsigned)i); }
    }
    if (must) v.normalize(norm);
}

}  // namespace ORB_SLAM2
