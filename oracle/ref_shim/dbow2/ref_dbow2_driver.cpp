// ref_dbow2_driver.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's own DBoW2 vocabulary
// (/root/reference/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h instantiated as ORBVocabulary = TemplatedVocabulary<FORB::TDescriptor, FORB>,
// include/ORBVocabulary.h:30-31, with FORB.cpp / ScoringObject.cpp / BowVector.cpp / FeatureVector.cpp compiled unmodified from where
// they lie; cv::Mat is the stand-in of opencv2/core/core.hpp next to this file).  Pinned through these: loadFromTextFile,
// loadFromBinaryFile, saveToBinaryFile, the per-feature transform (:1231-1272, FORB::distance FORB.cpp:81-101) and
// transform(features, BowVector&, FeatureVector&, levelsup) (:1140-1207).
#include <cstring>
#include <string>
#include <vector>
#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"

namespace {
typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> Base;
struct Voc : public Base {
    using Base::transform;                       // the protected per-feature overloads
    void one(const cv::Mat& f, DBoW2::WordId& id, DBoW2::WordValue& w, DBoW2::NodeId* nid, int levelsup) const { Base::transform(f, id, w, nid, levelsup); }
    size_t nodes() const { return m_nodes.size(); }
};
std::vector<cv::Mat> rows(const unsigned char* desc, int n) {
    std::vector<cv::Mat> v(n);
    for (int i = 0; i < n; i++) {
        v[i].create(1, 32, CV_8U);
        memcpy(v[i].data, desc + 32 * (size_t)i, 32);
    }
    return v;
}
}  // namespace

extern "C" {

void* refv_load(const char* path, int binary) {
    Voc* v = new Voc();
    const bool ok = binary ? v->loadFromBinaryFile(path) : v->loadFromTextFile(path);
    if (!ok) { delete v; return nullptr; }
    return v;
}
void refv_destroy(void* h) { delete (Voc*)h; }
void refv_save_binary(void* h, const char* path) { ((Voc*)h)->saveToBinaryFile(path); }
void refv_save_text(void* h, const char* path) { ((Voc*)h)->saveToTextFile(path); }
void refv_info(void* h, int* k, int* L, int* n_nodes, int* n_words) {
    Voc* v = (Voc*)h;
    *k = v->getBranchingFactor(); *L = v->getDepthLevels(); *n_nodes = (int)v->nodes(); *n_words = (int)v->size();
}
// per feature: word id, word weight, node id `levelsup` levels above the word (TemplatedVocabulary.h:1231-1272)
void refv_transform_raw(void* h, const unsigned char* desc, int n, int levelsup, int* word, double* weight, int* nid) {
    Voc* v = (Voc*)h;
    std::vector<cv::Mat> f = rows(desc, n);
    for (int i = 0; i < n; i++) {
        DBoW2::WordId id = 0; DBoW2::WordValue w = 0; DBoW2::NodeId nd = 0;
        v->one(f[i], id, w, &nd, levelsup);
        word[i] = (int)id; weight[i] = w; nid[i] = (int)nd;
    }
}
// transform(features, BowVector&, FeatureVector&, levelsup) (:1140-1207): the two maps flattened in key order.  Returns the number of
// BowVector entries; *n_fv_nodes the number of FeatureVector nodes; fv_off has n_fv_nodes + 1 entries.
int refv_transform(void* h, const unsigned char* desc, int n, int levelsup, int* bow_id, double* bow_val, int* fv_node, int* fv_off,
                   int* fv_feat, int* n_fv_nodes) {
    Voc* v = (Voc*)h;
    std::vector<cv::Mat> f = rows(desc, n);
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    v->transform(f, bv, fv, levelsup);
    int i = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++i) { bow_id[i] = (int)it->first; bow_val[i] = it->second; }
    int a = 0, e = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++a) {
        fv_node[a] = (int)it->first;
        fv_off[a] = e;
        for (size_t j = 0; j < it->second.size(); j++) fv_feat[e++] = (int)it->second[j];
    }
    fv_off[a] = e;
    *n_fv_nodes = a;
    return i;
}

}  // extern "C"
