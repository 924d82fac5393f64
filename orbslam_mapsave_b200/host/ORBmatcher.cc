// ORBmatcher.cc — host side of the drop-in ORBmatcher members: snapshots the KeyFrame / Frame fields the reference reads
// (under the reference's own accessors, so its locking discipline is kept), flattens mFeatVec to CSR and calls the C ABI.
// Replaces src/ORBmatcher.cc:159-291, 525-658, 660-826, 1650-1666 of the reference.
#include "ORBmatcher.h"

#include <cmath>
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
    int d = 257;                 // on a CUDA error: larger than any Hamming distance, so that no caller reads the failure as a match
    report(orbm_descriptor_distance(a.ptr<unsigned char>(), b.ptr<unsigned char>(), 1, &d, g_device));
    return g_status == ORB_OK ? d : 257;
}

int ORBmatcher::DescriptorDistances(const cv::Mat& A, const cv::Mat& B, int* out) {
    const int n = A.rows < B.rows ? A.rows : B.rows;
    if (n <= 0) return 0;
    if (A.isContinuous() && B.isContinuous()) {
        report(orbm_descriptor_distance(A.ptr<unsigned char>(), B.ptr<unsigned char>(), n, out, g_device));
    } else {                     // row views with a stride: gather
        std::vector<unsigned char> a((size_t)n * 32), b((size_t)n * 32);
        for (int i = 0; i < n; i++) { std::memcpy(&a[32 * (size_t)i], A.ptr<unsigned char>(i), 32); std::memcpy(&b[32 * (size_t)i], B.ptr<unsigned char>(i), 32); }
        report(orbm_descriptor_distance(a.data(), b.data(), n, out, g_device));
    }
    if (g_status != ORB_OK) for (int i = 0; i < n; i++) out[i] = 257;
    return n;
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

// Batched overloads (not in the reference): the candidate loops of Tracking::Relocalization (src/Tracking.cc:1621-1643) and
// LoopClosing::ComputeSim3 (src/LoopClosing.cc:240-266) as ONE device call; results equal calling the overloads above per candidate.
std::vector<int> ORBmatcher::SearchByBoW(const std::vector<KeyFrame*>& vpKFs, Frame& F, std::vector<std::vector<MapPoint*> >& vvpMapPointMatches) {
    const size_t k = vpKFs.size();
    std::vector<int> counts(k, 0);
    vvpMapPointMatches.assign(k, std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL)));
    if (k == 0) return counts;
    std::vector<std::vector<MapPoint*> > mps(k);
    std::vector<Side> sides(k);
    std::vector<orbm_view> views(k);
    for (size_t i = 0; i < k; i++) {
        mps[i] = vpKFs[i]->GetMapPointMatches();
        snapshot_kf_bow(vpKFs[i], mps[i], sides[i]);
        views[i] = sides[i].view;
    }
    Side b;
    copy_desc(F.mDescriptors, F.N, b);
    b.angle.resize(F.N);
    for (int i = 0; i < F.N; i++) b.angle[i] = F.mvKeys[i].angle;
    flatten_featvec(F.mFeatVec, b);
    finish(b, F.N, false);
    std::vector<int> m21(k * (size_t)(F.N > 0 ? F.N : 1), -1);
    report(orbm_search_by_bow_batch(&b.view, views.data(), (int)k, 0, mfNNratio, mbCheckOrientation ? 1 : 0, m21.data(), counts.data(), g_device));
    if (g_status != ORB_OK) return std::vector<int>(k, 0);
    for (size_t i = 0; i < k; i++)
        for (int j = 0; j < F.N; j++) {
            const int idx1 = m21[i * (size_t)F.N + j];
            if (idx1 >= 0) vvpMapPointMatches[i][j] = mps[i][idx1];
        }
    return counts;
}

std::vector<int> ORBmatcher::SearchByBoW(KeyFrame* pKF1, const std::vector<KeyFrame*>& vpKF2s, std::vector<std::vector<MapPoint*> >& vvpMatches12) {
    const size_t k = vpKF2s.size();
    std::vector<int> counts(k, 0);
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const size_t n1 = vpMapPoints1.size();
    vvpMatches12.assign(k, std::vector<MapPoint*>(n1, static_cast<MapPoint*>(NULL)));
    if (k == 0) return counts;
    Side a;
    snapshot_kf_bow(pKF1, vpMapPoints1, a);
    std::vector<std::vector<MapPoint*> > mps(k);
    std::vector<Side> sides(k);
    std::vector<orbm_view> views(k);
    for (size_t i = 0; i < k; i++) {
        mps[i] = vpKF2s[i]->GetMapPointMatches();
        snapshot_kf_bow(vpKF2s[i], mps[i], sides[i]);
        views[i] = sides[i].view;
    }
    std::vector<int> m12(k * (n1 ? n1 : 1), -1);
    report(orbm_search_by_bow_batch(&a.view, views.data(), (int)k, 1, mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(), counts.data(), g_device));
    if (g_status != ORB_OK) return std::vector<int>(k, 0);
    for (size_t i = 0; i < k; i++)
        for (size_t j = 0; j < n1; j++) {
            const int idx2 = m12[i * n1 + j];
            if (idx2 >= 0) vvpMatches12[i][j] = mps[i][idx2];
        }
    return counts;
}

namespace {
// Epipole of pKF1's centre in the second image (ORBmatcher.cc:667-673).  `R2w*Cw+t2w` is one cv::gemm(A, B, 1, C, 1) on 3x3 / 3x1 floats:
// float32, left to right (cv2 4.13, tests/golden/prim_gemm3.npz).
void epipole_in(KeyFrame* pKF1, KeyFrame* pKF2, float& ex, float& ey) {
    cv::Mat Cw = pKF1->GetCameraCenter();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    float C2[3];
    for (int i = 0; i < 3; i++)
        C2[i] = ((R2w.at<float>(i, 0) * Cw.at<float>(0) + R2w.at<float>(i, 1) * Cw.at<float>(1)) + R2w.at<float>(i, 2) * Cw.at<float>(2)) +
                t2w.at<float>(i);
    const float invz = 1.0f / C2[2];
    ex = pKF2->fx * C2[0] * invz + pKF2->cx;
    ey = pKF2->fy * C2[1] * invz + pKF2->cy;
}
}  // namespace

std::vector<int> ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, const std::vector<KeyFrame*>& vpKF2s, const std::vector<cv::Mat>& vF12,
                                                    std::vector<std::vector<std::pair<size_t, size_t> > >& vvMatchedPairs,
                                                    const bool bOnlyStereo) {
    const size_t k = std::min(vpKF2s.size(), vF12.size());
    std::vector<int> counts(k, 0);
    vvMatchedPairs.assign(k, std::vector<std::pair<size_t, size_t> >());
    if (k == 0) return counts;
    // the C entry takes one scale table width for all neighbours: keyframes of one map share it (mnScaleLevels is a system constant)
    const size_t nlv = vpKF2s[0]->mvScaleFactors.size();
    bool uniform = true;
    for (size_t i = 0; i < k; i++) uniform = uniform && vpKF2s[i]->mvScaleFactors.size() == nlv && vpKF2s[i]->mvLevelSigma2.size() == nlv;
    if (!uniform) {                                                   // (never in the reference; keep the semantics anyway)
        for (size_t i = 0; i < k; i++) counts[i] = SearchForTriangulation(pKF1, vpKF2s[i], vF12[i], vvMatchedPairs[i], bOnlyStereo);
        return counts;
    }
    Side a;
    snapshot_kf_tri(pKF1, a);
    std::vector<Side> sides(k);
    std::vector<orbm_view> views(k);
    std::vector<float> F(9 * k), ep(2 * k), sf(nlv * k), s2(nlv * k);
    for (size_t i = 0; i < k; i++) {
        snapshot_kf_tri(vpKF2s[i], sides[i]);
        views[i] = sides[i].view;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) F[9 * i + 3 * r + c] = vF12[i].at<float>(r, c);
        epipole_in(pKF1, vpKF2s[i], ep[2 * i], ep[2 * i + 1]);
        for (size_t l = 0; l < nlv; l++) { sf[i * nlv + l] = vpKF2s[i]->mvScaleFactors[l]; s2[i * nlv + l] = vpKF2s[i]->mvLevelSigma2[l]; }
    }
    const size_t n1 = (size_t)(pKF1->N > 0 ? pKF1->N : 1);
    std::vector<int> pairs(2 * n1 * k), npairs(k, 0);
    report(orbm_search_for_triangulation_batch(&a.view, views.data(), (int)k, F.data(), ep.data(), sf.data(), s2.data(), (int)nlv,
                                               bOnlyStereo ? 1 : 0, mbCheckOrientation ? 1 : 0, pairs.data(), npairs.data(), counts.data(),
                                               g_device));
    if (g_status != ORB_OK) return std::vector<int>(k, 0);
    for (size_t i = 0; i < k; i++) {
        vvMatchedPairs[i].reserve(npairs[i]);
        const int* p = &pairs[2 * (size_t)pKF1->N * i];
        for (int j = 0; j < npairs[i]; j++) vvMatchedPairs[i].push_back(std::make_pair((size_t)p[2 * j], (size_t)p[2 * j + 1]));
    }
    return counts;
}

int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12,
                                       std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo) {
    float ex, ey;
    epipole_in(pKF1, pKF2, ex, ey);

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


// ---- window searches ------------------------------------------------------------------------------------------
namespace {
// flat copy of the Frame fields GetFeaturesInArea and the candidate loops read (src/Frame.cc:445-498)
struct GridSide {
    std::vector<unsigned char> desc, blocked;
    std::vector<float> x, y, angle, uright;
    std::vector<int> octave, off, feat;
    orbm_grid_view view;
};
void snapshot_frame_grid(const Frame& F, bool blockObserved, GridSide& g) {
    const int n = F.N;
    g.desc.resize((size_t)n * 32);
    g.x.resize(n); g.y.resize(n); g.angle.resize(n); g.octave.resize(n); g.uright.resize(n); g.blocked.assign(n, 0);
    for (int i = 0; i < n; i++) {
        std::memcpy(&g.desc[(size_t)i * 32], F.mDescriptors.ptr(i), 32);
        const cv::KeyPoint& kp = F.mvKeysUn[i];
        g.x[i] = kp.pt.x; g.y[i] = kp.pt.y; g.angle[i] = kp.angle; g.octave[i] = kp.octave;
        g.uright[i] = F.mvuRight.empty() ? -1.0f : F.mvuRight[i];
        if (blockObserved && F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0) g.blocked[i] = 1;   // ORBmatcher.cc:87-89
    }
    g.off.assign(1, 0); g.feat.clear();
    for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
        for (int iy = 0; iy < FRAME_GRID_ROWS; iy++) {
            const std::vector<std::size_t>& cell = F.mGrid[ix][iy];
            for (size_t k = 0; k < cell.size(); k++) g.feat.push_back((int)cell[k]);
            g.off.push_back((int)g.feat.size());
        }
    std::memset(&g.view, 0, sizeof(g.view));
    g.view.n = n;
    g.view.desc = g.desc.data(); g.view.x = g.x.data(); g.view.y = g.y.data(); g.view.octave = g.octave.data();
    g.view.angle = g.angle.data(); g.view.uright = g.uright.data(); g.view.blocked = g.blocked.data();
    g.view.grid_cols = FRAME_GRID_COLS; g.view.grid_rows = FRAME_GRID_ROWS;
    g.view.min_x = F.mnMinX; g.view.min_y = F.mnMinY; g.view.max_x = F.mnMaxX; g.view.max_y = F.mnMaxY;
    g.view.inv_w = F.mfGridElementWidthInv; g.view.inv_h = F.mfGridElementHeightInv;
    g.view.cell_offsets = g.off.data(); g.view.cell_features = g.feat.data();
    g.view.scale_factors = F.mvScaleFactors.data(); g.view.n_levels = (int)F.mvScaleFactors.size();
}
}  // namespace

int ORBmatcher::SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    GridSide g;
    snapshot_frame_grid(F, true, g);
    const int np = (int)vpMapPoints.size();
    std::vector<unsigned char> in_view(np), claims(np), desc((size_t)np * 32);
    std::vector<float> px(np), py(np), pxr(np), vc(np);
    std::vector<int> level(np);
    for (int i = 0; i < np; i++) {
        MapPoint* pMP = vpMapPoints[i];
        in_view[i] = (pMP->mbTrackInView && !pMP->isBad()) ? 1 : 0;                      // ORBmatcher.cc:53-58
        if (!in_view[i]) continue;
        px[i] = pMP->mTrackProjX; py[i] = pMP->mTrackProjY; pxr[i] = pMP->mTrackProjXR;
        level[i] = pMP->mnTrackScaleLevel; vc[i] = pMP->mTrackViewCos;
        claims[i] = pMP->Observations() > 0 ? 1 : 0;
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&desc[(size_t)i * 32], d.ptr(0), 32);
    }
    std::vector<int> owner(F.N > 0 ? F.N : 1, -1);
    int nmatches = 0;
    report(orbm_search_by_projection_map(&g.view, np, in_view.data(), px.data(), py.data(), pxr.data(), level.data(), vc.data(),
                                         desc.data(), claims.data(), th, mfNNratio, owner.data(), &nmatches, g_device));
    if (g_status != ORB_OK) return 0;
    for (int j = 0; j < F.N; j++)
        if (owner[j] >= 0) F.mvpMapPoints[j] = vpMapPoints[owner[j]];
    return nmatches;
}

namespace {
// flat copy of what SearchByProjection(CurrentFrame, LastFrame) reads of the two frames (src/ORBmatcher.cc:1344-1396)
struct FramePairSide {
    GridSide g;
    float Tc[12], Tl[12];
    std::vector<unsigned char> has, claims, desc;
    std::vector<float> world, angle;
    std::vector<int> octave, owner;
};
void snapshot_frame_pair(Frame& CurrentFrame, const Frame& LastFrame, FramePairSide& s) {
    snapshot_frame_grid(CurrentFrame, true, s.g);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) { s.Tc[4 * r + c] = CurrentFrame.mTcw.at<float>(r, c); s.Tl[4 * r + c] = LastFrame.mTcw.at<float>(r, c); }
    const int nl = LastFrame.N;
    s.has.assign(nl, 0); s.claims.assign(nl, 0); s.desc.assign((size_t)nl * 32, 0);
    s.world.assign((size_t)nl * 3, 0.f); s.angle.assign(nl, 0.f); s.octave.assign(nl, 0);
    for (int i = 0; i < nl; i++) {
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;                                   // ORBmatcher.cc:1359-1363
        s.has[i] = 1;
        const cv::Mat x3Dw = pMP->GetWorldPos();
        for (int k = 0; k < 3; k++) s.world[3 * (size_t)i + k] = x3Dw.at<float>(k);
        s.octave[i] = LastFrame.mvKeys[i].octave;
        s.angle[i] = LastFrame.mvKeysUn[i].angle;
        s.claims[i] = pMP->Observations() > 0 ? 1 : 0;
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&s.desc[(size_t)i * 32], d.ptr(0), 32);
    }
    s.owner.assign(CurrentFrame.N > 0 ? CurrentFrame.N : 1, -1);
}
void apply_frame_owner(Frame& CurrentFrame, const Frame& LastFrame, const std::vector<int>& owner) {
    for (int j = 0; j < CurrentFrame.N; j++) {
        if (owner[j] >= 0) CurrentFrame.mvpMapPoints[j] = LastFrame.mvpMapPoints[owner[j]];
        else if (owner[j] == -2) CurrentFrame.mvpMapPoints[j] = static_cast<MapPoint*>(NULL);   // rotation cull, :1452-1456
    }
}
}  // namespace

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
    FramePairSide s;
    snapshot_frame_pair(CurrentFrame, LastFrame, s);
    int nmatches = 0;
    report(orbm_search_by_projection_frame(&s.g.view, s.Tc, s.Tl, CurrentFrame.fx, CurrentFrame.fy, CurrentFrame.cx, CurrentFrame.cy,
                                           CurrentFrame.mbf, CurrentFrame.mb, LastFrame.N, s.has.data(), s.world.data(), s.octave.data(),
                                           s.angle.data(), s.desc.data(), s.claims.data(), th, bMono ? 1 : 0, mbCheckOrientation ? 1 : 0,
                                           s.owner.data(), &nmatches, g_device));
    if (g_status != ORB_OK) return 0;
    apply_frame_owner(CurrentFrame, LastFrame, s.owner);
    return nmatches;
}

std::vector<int> ORBmatcher::SearchByProjection(const std::vector<Frame*>& vpCurrentFrames, const std::vector<const Frame*>& vpLastFrames,
                                                const float th, const bool bMono) {
    const size_t n = std::min(vpCurrentFrames.size(), vpLastFrames.size());
    std::vector<FramePairSide> sides(n);
    std::vector<orbm_frame_search_job> jobs(n);
    for (size_t k = 0; k < n; k++) {
        Frame& C = *vpCurrentFrames[k];
        FramePairSide& s = sides[k];
        snapshot_frame_pair(C, *vpLastFrames[k], s);
        orbm_frame_search_job& j = jobs[k];
        std::memset(&j, 0, sizeof(j));
        j.cur = &s.g.view; j.Tcw_cur = s.Tc; j.Tcw_last = s.Tl;
        j.fx = C.fx; j.fy = C.fy; j.cx = C.cx; j.cy = C.cy; j.mbf = C.mbf; j.mb = C.mb;
        j.n_last = vpLastFrames[k]->N;
        j.has_point = s.has.data(); j.world = s.world.data(); j.octave = s.octave.data(); j.angle = s.angle.data();
        j.desc = s.desc.data(); j.claims = s.claims.data(); j.th = th; j.mono = bMono ? 1 : 0; j.owner = s.owner.data();
    }
    std::vector<int> nmatches(n, 0);
    report(orbm_search_by_projection_frame_batch(jobs.data(), (int)n, mbCheckOrientation ? 1 : 0, g_device));
    if (g_status != ORB_OK) return nmatches;
    for (size_t k = 0; k < n; k++) {
        apply_frame_owner(*vpCurrentFrames[k], *vpLastFrames[k], sides[k].owner);
        nmatches[k] = jobs[k].n_matches;
    }
    return nmatches;
}

int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                                        int windowSize) {
    GridSide g;
    snapshot_frame_grid(F2, false, g);
    g.view.uright = NULL; g.view.blocked = NULL;
    const int n1 = (int)F1.mvKeysUn.size();
    vnMatches12 = std::vector<int>(n1, -1);
    std::vector<unsigned char> desc((size_t)n1 * 32);
    std::vector<int> octave(n1);
    std::vector<float> angle(n1), prev((size_t)2 * n1);
    for (int i = 0; i < n1; i++) {
        std::memcpy(&desc[(size_t)i * 32], F1.mDescriptors.ptr(i), 32);
        octave[i] = F1.mvKeysUn[i].octave; angle[i] = F1.mvKeysUn[i].angle;
        prev[2 * (size_t)i] = vbPrevMatched[i].x; prev[2 * (size_t)i + 1] = vbPrevMatched[i].y;
    }
    std::vector<int> m12(n1 > 0 ? n1 : 1, -1);
    int nmatches = 0;
    report(orbm_search_for_initialization(&g.view, n1, desc.data(), octave.data(), angle.data(), prev.data(), windowSize, mfNNratio,
                                          mbCheckOrientation ? 1 : 0, m12.data(), &nmatches, g_device));
    if (g_status != ORB_OK) return 0;
    for (int i = 0; i < n1; i++) {
        vnMatches12[i] = m12[i];
        vbPrevMatched[i] = cv::Point2f(prev[2 * (size_t)i], prev[2 * (size_t)i + 1]);
    }
    return nmatches;
}


// ---- relocalisation / loop-closing projection searches: the per-point projection runs here (where the MapPoint objects live),
// the window search, Hamming distances and the serial claim order on the GPU (orbm_search_windows) --------------------------------
namespace {
// cv::gemm(A, x, 1, b, 1) on 3x3 / 3x1 floats: float32, left to right; transposed products and cv::norm / Mat::dot accumulate in
// double (cv2 4.13, tests/golden/prim_gemm3.npz)
inline void rt_apply(const float* T, const float* p, float* out) {
    for (int r = 0; r < 3; r++) out[r] = ((T[4 * r] * p[0] + T[4 * r + 1] * p[1]) + T[4 * r + 2] * p[2]) + T[4 * r + 3];
}
inline void minus_rt_t(const float* T, float* out) {
    for (int r = 0; r < 3; r++) out[r] = (float)(-(((double)T[r] * T[3] + (double)T[4 + r] * T[7]) + (double)T[8 + r] * T[11]));
}
inline float norm3(const float* p) { return (float)std::sqrt((double)p[0] * p[0] + (double)p[1] * p[1] + (double)p[2] * p[2]); }
inline int clamp_level(int l, int n) { return l < 0 ? 0 : (l >= n ? n - 1 : l); }   // the fork indexes mvScaleFactors unclamped (UB)

// the KeyFrame side: mGrid is not public, but it is Frame::AssignFeaturesToGrid of mvKeysUn (src/KeyFrame.cc:48-53), rebuilt here
void snapshot_keyframe_grid(KeyFrame* pKF, GridSide& g) {
    const int n = pKF->N;
    g.desc.resize((size_t)n * 32);
    g.x.resize(n); g.y.resize(n); g.angle.resize(n); g.octave.resize(n); g.uright.resize(n); g.blocked.assign(n, 0);
    std::vector<std::vector<int> > grid((size_t)pKF->mnGridCols * pKF->mnGridRows);
    for (int i = 0; i < n; i++) {
        std::memcpy(&g.desc[(size_t)i * 32], pKF->mDescriptors.ptr(i), 32);
        const cv::KeyPoint& kp = pKF->mvKeysUn[i];
        g.x[i] = kp.pt.x; g.y[i] = kp.pt.y; g.angle[i] = kp.angle; g.octave[i] = kp.octave;
        g.uright[i] = pKF->mvuRight.empty() ? -1.0f : pKF->mvuRight[i];
        const int posX = (int)round((kp.pt.x - pKF->mnMinX) * pKF->mfGridElementWidthInv);
        const int posY = (int)round((kp.pt.y - pKF->mnMinY) * pKF->mfGridElementHeightInv);
        if (posX < 0 || posX >= pKF->mnGridCols || posY < 0 || posY >= pKF->mnGridRows) continue;
        grid[(size_t)posX * pKF->mnGridRows + posY].push_back(i);
    }
    g.off.assign(1, 0); g.feat.clear();
    for (size_t c = 0; c < grid.size(); c++) { g.feat.insert(g.feat.end(), grid[c].begin(), grid[c].end()); g.off.push_back((int)g.feat.size()); }
    std::memset(&g.view, 0, sizeof(g.view));
    g.view.n = n; g.view.desc = g.desc.data(); g.view.x = g.x.data(); g.view.y = g.y.data(); g.view.octave = g.octave.data();
    g.view.angle = g.angle.data(); g.view.uright = g.uright.data(); g.view.blocked = g.blocked.data();
    g.view.grid_cols = pKF->mnGridCols; g.view.grid_rows = pKF->mnGridRows;
    g.view.min_x = (float)pKF->mnMinX; g.view.min_y = (float)pKF->mnMinY; g.view.max_x = (float)pKF->mnMaxX; g.view.max_y = (float)pKF->mnMaxY;
    g.view.inv_w = pKF->mfGridElementWidthInv; g.view.inv_h = pKF->mfGridElementHeightInv;
    g.view.cell_offsets = g.off.data(); g.view.cell_features = g.feat.data();
    g.view.scale_factors = pKF->mvScaleFactors.data(); g.view.n_levels = (int)pKF->mvScaleFactors.size();
}

struct Windows {
    std::vector<unsigned char> active, desc;
    std::vector<float> u, v, r, angle;
    std::vector<int> minL, maxL;
    explicit Windows(int n) : active(n, 0), desc((size_t)n * 32), u(n), v(n), r(n), angle(n), minL(n, -1), maxL(n, -1) {}
};
}  // namespace

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th,
                                   const int ORBdist) {
    GridSide g;
    snapshot_frame_grid(CurrentFrame, false, g);
    for (int i = 0; i < CurrentFrame.N; i++) g.blocked[i] = CurrentFrame.mvpMapPoints[i] ? 1 : 0;          // ORBmatcher.cc:1548-1549
    float T[12], Ow[3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) T[4 * r + c] = CurrentFrame.mTcw.at<float>(r, c);
    minus_rt_t(T, Ow);                                                                                      // :1471
    const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
    const int n = (int)vpMPs.size(), nlev = (int)CurrentFrame.mvScaleFactors.size();
    Windows w(n);
    for (int i = 0; i < n; i++) {
        MapPoint* pMP = vpMPs[i];
        if (!pMP) continue;
        if (pMP->isBad() || sAlreadyFound.count(pMP)) continue;
        const cv::Mat x3Dw = pMP->GetWorldPos();
        const float X[3] = {x3Dw.at<float>(0), x3Dw.at<float>(1), x3Dw.at<float>(2)};
        float x3Dc[3];
        rt_apply(T, X, x3Dc);
        const float xc = x3Dc[0], yc = x3Dc[1];
        const float invzc = 1.0 / x3Dc[2];
        const float u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
        const float v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
        if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX) continue;
        if (v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
        const float PO[3] = {X[0] - Ow[0], X[1] - Ow[1], X[2] - Ow[2]};
        float dist3D = norm3(PO);
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const int nPredictedLevel = clamp_level(pMP->PredictScale(dist3D, CurrentFrame.mfLogScaleFactor), nlev);
        w.active[i] = 1; w.u[i] = u; w.v[i] = v;
        w.r[i] = th * CurrentFrame.mvScaleFactors[nPredictedLevel];
        w.minL[i] = nPredictedLevel - 1; w.maxL[i] = nPredictedLevel + 1;
        w.angle[i] = pKF->mvKeysUn[i].angle;
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&w.desc[(size_t)i * 32], d.ptr(0), 32);
    }
    std::vector<int> owner(CurrentFrame.N > 0 ? CurrentFrame.N : 1, -1);
    int nmatches = 0;
    report(orbm_search_windows(&g.view, n, w.active.data(), w.u.data(), w.v.data(), w.r.data(), w.minL.data(), w.maxL.data(),
                               w.desc.data(), w.angle.data(), ORBdist, mbCheckOrientation ? 1 : 0, owner.data(), &nmatches, g_device));
    if (g_status != ORB_OK) return 0;
    for (int j = 0; j < CurrentFrame.N; j++) {
        if (owner[j] >= 0) CurrentFrame.mvpMapPoints[j] = vpMPs[owner[j]];
        else if (owner[j] == -2) CurrentFrame.mvpMapPoints[j] = NULL;
    }
    return nmatches;
}

int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched,
                                   int th) {
    const float &fx = pKF->fx, &fy = pKF->fy, &cx = pKF->cx, &cy = pKF->cy;
    // Decompose Scw (ORBmatcher.cc:301-306): scw = sqrt(row0 . row0) (Mat::dot: double), Rcw = sRcw / scw, tcw = Scw.col(3) / scw
    const float s0 = Scw.at<float>(0, 0), s1 = Scw.at<float>(0, 1), s2 = Scw.at<float>(0, 2);
    const float scw = (float)std::sqrt((double)s0 * s0 + (double)s1 * s1 + (double)s2 * s2);
    const float inv = (float)(1.0 / (double)scw);
    float T[12], Ow[3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) T[4 * r + c] = Scw.at<float>(r, c) * inv;
    minus_rt_t(T, Ow);
    const int n = pKF->N;
    GridSide g;
    snapshot_keyframe_grid(pKF, g);
    for (int i = 0; i < n; i++) g.blocked[i] = vpMatched[i] ? 1 : 0;                                          // :375-376

    std::set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint*>(NULL));
    const int np = (int)vpPoints.size(), nlev = (int)pKF->mvScaleFactors.size();
    Windows w(np);
    for (int iMP = 0; iMP < np; iMP++) {
        MapPoint* pMP = vpPoints[iMP];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        const cv::Mat p3Dw = pMP->GetWorldPos();
        const float X[3] = {p3Dw.at<float>(0), p3Dw.at<float>(1), p3Dw.at<float>(2)};
        float p3Dc[3];
        rt_apply(T, X, p3Dc);
        if (p3Dc[2] < 0.0) continue;
        const float invz = 1 / p3Dc[2];
        const float x = p3Dc[0] * invz, y = p3Dc[1] * invz;
        const float u = fx * x + cx, v = fy * y + cy;
        if (!pKF->IsInImage(u, v)) continue;
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        const float PO[3] = {X[0] - Ow[0], X[1] - Ow[1], X[2] - Ow[2]};
        const float dist = norm3(PO);
        if (dist < minDistance || dist > maxDistance) continue;
        const cv::Mat Pn = pMP->GetNormal();
        const double dot = (double)PO[0] * Pn.at<float>(0) + (double)PO[1] * Pn.at<float>(1) + (double)PO[2] * Pn.at<float>(2);
        if (dot < 0.5 * dist) continue;
        const int nPredictedLevel = clamp_level(pMP->PredictScale(dist, pKF->mfLogScaleFactor), nlev);
        w.active[iMP] = 1; w.u[iMP] = u; w.v[iMP] = v;
        w.r[iMP] = th * pKF->mvScaleFactors[nPredictedLevel];
        w.minL[iMP] = nPredictedLevel - 1; w.maxL[iMP] = nPredictedLevel;                                   // :380-381
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&w.desc[(size_t)iMP * 32], d.ptr(0), 32);
    }
    std::vector<int> owner(n > 0 ? n : 1, -1);
    int nmatches = 0;
    report(orbm_search_windows(&g.view, np, w.active.data(), w.u.data(), w.v.data(), w.r.data(), w.minL.data(), w.maxL.data(),
                               w.desc.data(), NULL, TH_LOW, 0, owner.data(), &nmatches, g_device));
    if (g_status != ORB_OK) return 0;
    for (int j = 0; j < n; j++)
        if (owner[j] >= 0) vpMatched[j] = vpPoints[owner[j]];
    return nmatches;
}


// ---- SearchBySim3 and Fuse: the candidate loops do not depend on earlier iterations, so all points are searched at once on the
// device (orbm_search_windows_best) and the reference's bookkeeping / map updates run afterwards on the host, in order -------------
namespace {
// scale a 3x3 by a scalar the way `s*R` / `(1.0/s)*R.t()` do: cv::Mat::convertTo scales CV_32F data by (float)alpha
// (cv2 4.13, probed through cv2.normalize; same convention as sRcw/scw above)
inline float scaled(float v, double alpha) { return v * (float)alpha; }

// one direction of SearchBySim3 (ORBmatcher.cc:1150-1219 / 1222-1299): points of `from` into `to`
void sim3_direction(KeyFrame* to, const std::vector<MapPoint*>& pts, const std::vector<bool>& already, const float* Rfw, const float* tfw,
                    const float* sR, const float* t, float th, std::vector<int>& vnMatch, int device, int* status) {
    const int n = (int)pts.size(), nlev = (int)to->mvScaleFactors.size();
    Windows w(n);
    for (int i = 0; i < n; i++) {
        MapPoint* pMP = pts[i];
        if (!pMP || already[i]) continue;
        if (pMP->isBad()) continue;
        const cv::Mat p3Dw = pMP->GetWorldPos();
        const float X[3] = {p3Dw.at<float>(0), p3Dw.at<float>(1), p3Dw.at<float>(2)};
        float T1[12], T2[12], c1[3], c2[3];
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) { T1[4 * r + c] = Rfw[3 * r + c]; T2[4 * r + c] = sR[3 * r + c]; } T1[4 * r + 3] = tfw[r]; T2[4 * r + 3] = t[r]; }
        rt_apply(T1, X, c1);
        rt_apply(T2, c1, c2);
        if (c2[2] < 0.0) continue;
        const float invz = 1.0 / c2[2];
        const float x = c2[0] * invz, y = c2[1] * invz;
        const float u = to->fx * x + to->cx, v = to->fy * y + to->cy;
        if (!to->IsInImage(u, v)) continue;
        const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
        const float dist3D = norm3(c2);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const int nPredictedLevel = clamp_level(pMP->PredictScale(dist3D, to->mfLogScaleFactor), nlev);
        w.active[i] = 1; w.u[i] = u; w.v[i] = v;
        w.r[i] = th * to->mvScaleFactors[nPredictedLevel];
        w.minL[i] = nPredictedLevel - 1; w.maxL[i] = nPredictedLevel;
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&w.desc[(size_t)i * 32], d.ptr(0), 32);
    }
    GridSide g;
    snapshot_keyframe_grid(to, g);
    vnMatch.assign(n > 0 ? n : 1, -1);
    *status = orbm_search_windows_best(&g.view, n, w.active.data(), w.u.data(), w.v.data(), w.r.data(), w.minL.data(), w.maxL.data(),
                                       w.desc.data(), NULL, NULL, ORBM_TH_HIGH, vnMatch.data(), device);
    vnMatch.resize(n);
}
}  // namespace

int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                             const cv::Mat& t12, const float th) {
    cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
    // sR12 = s12*R12 ; sR21 = (1.0/s12)*R12.t() ; t21 = -sR21*t12 (:1123-1126)
    float r1w[9], r2w[9], T1w[3], T2w[3], sR12[9], sR21[9], T12[3], T21[3];
    for (int r = 0; r < 3; r++) {
        T1w[r] = t1w.at<float>(r); T2w[r] = t2w.at<float>(r); T12[r] = t12.at<float>(r);
        for (int c = 0; c < 3; c++) {
            r1w[3 * r + c] = R1w.at<float>(r, c); r2w[3 * r + c] = R2w.at<float>(r, c);
            sR12[3 * r + c] = scaled(R12.at<float>(r, c), (double)s12);
            sR21[3 * r + c] = scaled(R12.at<float>(c, r), 1.0 / s12);
        }
    }
    for (int r = 0; r < 3; r++) T21[r] = -((sR21[3 * r] * T12[0] + sR21[3 * r + 1] * T12[1]) + sR21[3 * r + 2] * T12[2]);
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const int N1 = (int)vpMapPoints1.size();
    const std::vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N2 = (int)vpMapPoints2.size();
    std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
    for (int i = 0; i < N1; i++) {
        MapPoint* pMP = vpMatches12[i];
        if (pMP) {
            vbAlreadyMatched1[i] = true;
            int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
        }
    }
    std::vector<int> vnMatch1, vnMatch2;
    int st = ORB_OK;
    sim3_direction(pKF2, vpMapPoints1, vbAlreadyMatched1, r1w, T1w, sR21, T21, th, vnMatch1, g_device, &st);
    report(st);
    if (st != ORB_OK) return 0;
    sim3_direction(pKF1, vpMapPoints2, vbAlreadyMatched2, r2w, T2w, sR12, T12, th, vnMatch2, g_device, &st);
    report(st);
    if (st != ORB_OK) return 0;
    int nFound = 0;                                                    // check agreement (:1301-1326)
    for (int i1 = 0; i1 < N1; i1++) {
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0) {
            const int idx1 = vnMatch2[idx2];
            if (idx1 == i1) { vpMatches12[i1] = vpMapPoints2[idx2]; nFound++; }
        }
    }
    return nFound;
}

namespace {
// the projection part shared by both Fuse overloads (:846-893, 1012-1051); T = [Rcw|tcw] row-major 3x4
void fuse_windows(KeyFrame* pKF, const float* T, const float* Ow, const std::vector<MapPoint*>& pts, const std::vector<unsigned char>& skip,
                  float th, bool stereoGate, Windows& w, std::vector<float>& ur) {
    const int n = (int)pts.size(), nlev = (int)pKF->mvScaleFactors.size();
    ur.assign(n > 0 ? n : 1, 0.f);
    for (int i = 0; i < n; i++) {
        MapPoint* pMP = pts[i];
        if (!pMP || skip[i]) continue;
        const cv::Mat p3Dw = pMP->GetWorldPos();
        const float X[3] = {p3Dw.at<float>(0), p3Dw.at<float>(1), p3Dw.at<float>(2)};
        float p3Dc[3];
        rt_apply(T, X, p3Dc);
        if (p3Dc[2] < 0.0f) continue;
        const float invz = stereoGate ? 1 / p3Dc[2] : (float)(1.0 / p3Dc[2]);       // `1/z` (:855) vs `1.0/z` (:1023)
        const float x = p3Dc[0] * invz, y = p3Dc[1] * invz;
        const float u = pKF->fx * x + pKF->cx, v = pKF->fy * y + pKF->cy;
        if (!pKF->IsInImage(u, v)) continue;
        ur[i] = u - pKF->mbf * invz;
        const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
        const float PO[3] = {X[0] - Ow[0], X[1] - Ow[1], X[2] - Ow[2]};
        const float dist3D = norm3(PO);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const cv::Mat Pn = pMP->GetNormal();
        const double dot = (double)PO[0] * Pn.at<float>(0) + (double)PO[1] * Pn.at<float>(1) + (double)PO[2] * Pn.at<float>(2);
        if (dot < 0.5 * dist3D) continue;
        const int nPredictedLevel = clamp_level(pMP->PredictScale(dist3D, pKF->mfLogScaleFactor), nlev);
        w.active[i] = 1; w.u[i] = u; w.v[i] = v;
        w.r[i] = th * pKF->mvScaleFactors[nPredictedLevel];
        w.minL[i] = nPredictedLevel - 1; w.maxL[i] = nPredictedLevel;
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&w.desc[(size_t)i * 32], d.ptr(0), 32);
    }
}
}  // namespace

// Precondition inherited from the callers (LocalMapping::SearchInNeighbors de-duplicates with mnFuseCandidateForKF,
// LoopClosing::SearchAndFuse passes a set): a MapPoint appears once in vpMapPoints, so the state snapshotted before the device
// search (position, descriptor, normal) is the state the reference would read at that point's turn; isBad() / IsInKeyFrame(),
// which earlier replacements do change, are re-evaluated at the point's turn below.
int ORBmatcher::Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    cv::Mat Rcw = pKF->GetRotation(), tcw = pKF->GetTranslation(), Owm = pKF->GetCameraCenter();
    float T[12], Ow[3];
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) T[4 * r + c] = Rcw.at<float>(r, c); T[4 * r + 3] = tcw.at<float>(r); Ow[r] = Owm.at<float>(r); }
    const int nMPs = (int)vpMapPoints.size();
    std::vector<unsigned char> skip(nMPs > 0 ? nMPs : 1, 0);
    for (int i = 0; i < nMPs; i++) {
        MapPoint* pMP = vpMapPoints[i];
        if (pMP && (pMP->isBad() || pMP->IsInKeyFrame(pKF))) skip[i] = 1;           // both only ever turn true during Fuse
    }
    Windows w(nMPs);
    std::vector<float> ur;
    fuse_windows(pKF, T, Ow, vpMapPoints, skip, th, true, w, ur);
    GridSide g;
    snapshot_keyframe_grid(pKF, g);
    std::vector<int> best(nMPs > 0 ? nMPs : 1, -1);
    report(orbm_search_windows_best(&g.view, nMPs, w.active.data(), w.u.data(), w.v.data(), w.r.data(), w.minL.data(), w.maxL.data(),
                                    w.desc.data(), ur.data(), pKF->mvInvLevelSigma2.data(), TH_LOW, best.data(), g_device));
    if (g_status != ORB_OK) return 0;
    int nFused = 0;
    for (int i = 0; i < nMPs; i++) {                                                 // :952-971, in the reference's order
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP) continue;
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        if (best[i] < 0) continue;
        const int bestIdx = best[i];
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, bestIdx);
            pKF->AddMapPoint(pMP, bestIdx);
        }
        nFused++;
    }
    return nFused;
}

int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint) {
    const float s0 = Scw.at<float>(0, 0), s1 = Scw.at<float>(0, 1), s2 = Scw.at<float>(0, 2);
    const float scw = (float)std::sqrt((double)s0 * s0 + (double)s1 * s1 + (double)s2 * s2);       // :983-987
    const float inv = (float)(1.0 / (double)scw);
    float T[12], Ow[3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) T[4 * r + c] = Scw.at<float>(r, c) * inv;
    minus_rt_t(T, Ow);
    const std::set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();
    const int nPoints = (int)vpPoints.size();
    std::vector<unsigned char> skip(nPoints > 0 ? nPoints : 1, 0);
    for (int i = 0; i < nPoints; i++)
        if (vpPoints[i]->isBad() || spAlreadyFound.count(vpPoints[i])) skip[i] = 1;               // :1002-1003 (this overload mutates neither)
    Windows w(nPoints);
    std::vector<float> ur;
    fuse_windows(pKF, T, Ow, vpPoints, skip, th, false, w, ur);
    GridSide g;
    snapshot_keyframe_grid(pKF, g);
    std::vector<int> best(nPoints > 0 ? nPoints : 1, -1);
    report(orbm_search_windows_best(&g.view, nPoints, w.active.data(), w.u.data(), w.v.data(), w.r.data(), w.minL.data(), w.maxL.data(),
                                    w.desc.data(), NULL, NULL, TH_LOW, best.data(), g_device));
    if (g_status != ORB_OK) return 0;
    int nFused = 0;
    for (int iMP = 0; iMP < nPoints; iMP++) {                                        // :1083-1099
        if (best[iMP] < 0) continue;
        MapPoint* pMP = vpPoints[iMP];
        MapPoint* pMPinKF = pKF->GetMapPoint(best[iMP]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) vpReplacePoint[iMP] = pMPinKF;
        } else {
            pMP->AddObservation(pKF, best[iMP]);
            pKF->AddMapPoint(pMP, best[iMP]);
        }
        nFused++;
    }
    return nFused;
}

}  // namespace ORB_SLAM2
