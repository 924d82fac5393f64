// ORBmatcher.cc — host side of the drop-in ORBmatcher members: snapshots the KeyFrame / Frame fields the reference reads
// (under the reference's own accessors, so its locking discipline is kept), flattens mFeatVec to CSR and calls the C ABI.
// Replaces src/ORBmatcher.cc:159-291, 525-658, 660-826, 1650-1666 of the reference.
#include "ORBmatcher.h"

#include <cstdio>
#include <cstring>

#include "../../include/orb_b200.h"

namespace ORB_SLAM2 {

const int ORBmatcher::TH_HIGH = ORBM_TH_HIGH;
const int ORBmatcher::TH_LOW = ORBM_TH_LOW;
const int ORBmatcher::HISTO_LENGTH = ORBM_HISTO_LENGTH;

static int g_device = 0;
static thread_local int g_status = 0;
void ORBmatcher::SetDevice(int device) { g_device = device; }
int ORBmatcher::LastStatus() { return g_status; }

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

static void report(int rc) {
    g_status = rc;
    if (rc != ORB_OK) std::fprintf(stderr, "ORBmatcher: %s\n", orb_last_error());
}

int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    int d = -1;
    report(orbm_descriptor_distance(a.ptr<unsigned char>(), b.ptr<unsigned char>(), 1, &d, g_device));
    return d;
}

namespace {
// flat copies of one side of a search
struct Side {
    std::vector<unsigned char> desc, flag;
    std::vector<float> angle, x, y, uright;
    std::vector<int> octave, ids, off, feat;
    orbm_view view;
};

void flatten_featvec(const DBoW2::FeatureVector& fv, Side& s) {
    s.ids.clear(); s.off.assign(1, 0); s.feat.clear();
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
        s.ids.push_back((int)it->first);
        for (size_t k = 0; k < it->second.size(); k++) s.feat.push_back((int)it->second[k]);
        s.off.push_back((int)s.feat.size());
    }
}

void copy_desc(const cv::Mat& d, int n, Side& s) {
    s.desc.resize((size_t)n * 32);
    for (int i = 0; i < n; i++) std::memcpy(&s.desc[(size_t)i * 32], d.ptr(i), 32);
}

void finish(Side& s, int n, bool tri) {
    std::memset(&s.view, 0, sizeof(s.view));
    s.view.n = n;
    s.view.desc = s.desc.data();
    s.view.flag = s.flag.empty() ? nullptr : s.flag.data();
    s.view.angle = s.angle.data();
    if (tri) { s.view.x = s.x.data(); s.view.y = s.y.data(); s.view.octave = s.octave.data(); s.view.uright = s.uright.data(); }
    s.view.fv.n_nodes = (int)s.ids.size();
    s.view.fv.node_ids = s.ids.data();
    s.view.fv.offsets = s.off.data();
    s.view.fv.features = s.feat.data();
}

// KeyFrame side of SearchByBoW: flag = has a MapPoint that is not bad (ORBmatcher.cc:196-203, 565-569)
void snapshot_kf_bow(KeyFrame* kf, const std::vector<MapPoint*>& mps, Side& s) {
    const int n = (int)mps.size();
    copy_desc(kf->mDescriptors, n, s);
    s.flag.resize(n); s.angle.resize(n);
    for (int i = 0; i < n; i++) {
        MapPoint* p = mps[i];
        s.flag[i] = (p && !p->isBad()) ? 1 : 0;
        s.angle[i] = kf->mvKeysUn[i].angle;
    }
    flatten_featvec(kf->mFeatVec, s);
    finish(s, n, false);
}

void snapshot_kf_tri(KeyFrame* kf, Side& s) {
    const int n = kf->N;
    copy_desc(kf->mDescriptors, n, s);
    s.flag.resize(n); s.angle.resize(n); s.x.resize(n); s.y.resize(n); s.octave.resize(n); s.uright.resize(n);
    for (int i = 0; i < n; i++) {
        s.flag[i] = kf->GetMapPoint(i) ? 1 : 0;                       // ORBmatcher.cc:703-708, 731-735
        const cv::KeyPoint& kp = kf->mvKeysUn[i];
        s.angle[i] = kp.angle; s.x[i] = kp.pt.x; s.y[i] = kp.pt.y; s.octave[i] = kp.octave;
        s.uright[i] = kf->mvuRight[i];
    }
    flatten_featvec(kf->mFeatVec, s);
    finish(s, n, true);
}
}  // namespace

int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
    Side a, b;
    snapshot_kf_bow(pKF, vpMapPointsKF, a);
    copy_desc(F.mDescriptors, F.N, b);
    b.angle.resize(F.N);
    for (int i = 0; i < F.N; i++) b.angle[i] = F.mvKeys[i].angle;        // ORBmatcher.cc:241 uses F.mvKeys
    flatten_featvec(F.mFeatVec, b);
    finish(b, F.N, false);
    std::vector<int> m21(F.N > 0 ? F.N : 1, -1);
    int nmatches = 0;
    report(orbm_search_by_bow_kf_frame(&a.view, &b.view, mfNNratio, mbCheckOrientation ? 1 : 0, m21.data(), &nmatches, g_device));
    if (g_status != ORB_OK) return 0;
    for (int j = 0; j < F.N; j++)
        if (m21[j] >= 0) vpMapPointMatches[j] = vpMapPointsKF[m21[j]];
    return nmatches;
}

int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const std::vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    vpMatches12 = std::vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));
    Side a, b;
    snapshot_kf_bow(pKF1, vpMapPoints1, a);
    snapshot_kf_bow(pKF2, vpMapPoints2, b);
    std::vector<int> m12(vpMapPoints1.empty() ? 1 : vpMapPoints1.size(), -1);
    int nmatches = 0;
    report(orbm_search_by_bow_kf_kf(&a.view, &b.view, mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(), &nmatches, g_device));
    if (g_status != ORB_OK) return 0;
    for (size_t i = 0; i < vpMapPoints1.size(); i++)
        if (m12[i] >= 0) vpMatches12[i] = vpMapPoints2[m12[i]];
    return nmatches;
}

int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12,
                                       std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo) {
    // Epipole in the second image (ORBmatcher.cc:667-673).  cv::Mat float products accumulate in double (cv::gemm), then round.
    cv::Mat Cw = pKF1->GetCameraCenter();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    float C2[3];
    for (int i = 0; i < 3; i++)
        C2[i] = (float)((double)R2w.at<float>(i, 0) * Cw.at<float>(0) + (double)R2w.at<float>(i, 1) * Cw.at<float>(1) +
                        (double)R2w.at<float>(i, 2) * Cw.at<float>(2) + (double)t2w.at<float>(i));
    const float invz = 1.0f / C2[2];
    const float ex = pKF2->fx * C2[0] * invz + pKF2->cx;
    const float ey = pKF2->fy * C2[1] * invz + pKF2->cy;

    Side a, b;
    snapshot_kf_tri(pKF1, a);
    snapshot_kf_tri(pKF2, b);
    float F[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) F[3 * r + c] = F12.at<float>(r, c);
    std::vector<int> pairs(2 * (pKF1->N > 0 ? pKF1->N : 1));
    int npairs = 0, nmatches = 0;
    report(orbm_search_for_triangulation(&a.view, &b.view, F, ex, ey, pKF2->mvScaleFactors.data(), pKF2->mvLevelSigma2.data(),
                                         (int)pKF2->mvScaleFactors.size(), bOnlyStereo ? 1 : 0, mbCheckOrientation ? 1 : 0,
                                         pairs.data(), &npairs, &nmatches, g_device));
    vMatchedPairs.clear();
    if (g_status != ORB_OK) return 0;
    vMatchedPairs.reserve(npairs);
    for (int i = 0; i < npairs; i++) vMatchedPairs.push_back(std::make_pair((size_t)pairs[2 * i], (size_t)pairs[2 * i + 1]));
    return nmatches;
}

}  // namespace ORB_SLAM2
