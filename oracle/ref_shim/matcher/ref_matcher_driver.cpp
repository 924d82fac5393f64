// ref_matcher_driver.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's own ORB_SLAM2::ORBmatcher
// (/root/reference/src/ORBmatcher.cc compiled unmodified; KeyFrame / Frame / MapPoint are the plain-data stand-ins of
// ref_types.hpp, DBoW2::FeatureVector is the reference's own).  Pinned through these: DescriptorDistance, SearchByBoW x2,
// SearchForTriangulation (+ CheckDistEpipolarLine), ComputeThreeMaxima.
#include <vector>
#include <ORBmatcher.h>      // the shadow next to this file first (angle brackets: via -I, so that its #include_next works)

namespace ORB_SLAM2 {

// members of the stand-ins that the reference defines in translation units which cannot be compiled here
void MapPoint::AddObservation(KeyFrame* pKF, size_t idx) {                 // src/MapPoint.cc:93-104
    if (mObservations.count(pKF)) return;
    mObservations[pKF] = idx;
    if (pKF->mvuRight[idx] >= 0) nObs += 2; else nObs++;
}
void MapPoint::Replace(MapPoint* pMP) {                                     // src/MapPoint.cc:247-283 (without the map bookkeeping)
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
std::vector<size_t> Frame::GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel, const int maxLevel) const {
    std::vector<size_t> vIndices;                                           // src/Frame.cc:445-498
    const int nMinCellX = std::max(0, (int)floor((x - mnMinX - r) * mfGridElementWidthInv));
    if (nMinCellX >= FRAME_GRID_COLS) return vIndices;
    const int nMaxCellX = std::min((int)FRAME_GRID_COLS - 1, (int)ceil((x - mnMinX + r) * mfGridElementWidthInv));
    if (nMaxCellX < 0) return vIndices;
    const int nMinCellY = std::max(0, (int)floor((y - mnMinY - r) * mfGridElementHeightInv));
    if (nMinCellY >= FRAME_GRID_ROWS) return vIndices;
    const int nMaxCellY = std::min((int)FRAME_GRID_ROWS - 1, (int)ceil((y - mnMinY + r) * mfGridElementHeightInv));
    if (nMaxCellY < 0) return vIndices;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const std::vector<size_t>& vCell = mGrid[ix][iy];
            for (size_t j = 0; j < vCell.size(); j++) {
                const cv::KeyPoint& kpUn = mvKeysUn[vCell[j]];
                if (bCheckLevels) {
                    if (kpUn.octave < minLevel) continue;
                    if (maxLevel >= 0 && kpUn.octave > maxLevel) continue;
                }
                const float distx = kpUn.pt.x - x, disty = kpUn.pt.y - y;
                if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(vCell[j]);
            }
        }
    return vIndices;
}
std::vector<size_t> KeyFrame::GetFeaturesInArea(const float& x, const float& y, const float& r) const {
    std::vector<size_t> vIndices;                                           // src/KeyFrame.cc:1311-1350
    const int nMinCellX = std::max(0, (int)floor((x - mnMinX - r) * mfGridElementWidthInv));
    if (nMinCellX >= mnGridCols) return vIndices;
    const int nMaxCellX = std::min((int)mnGridCols - 1, (int)ceil((x - mnMinX + r) * mfGridElementWidthInv));
    if (nMaxCellX < 0) return vIndices;
    const int nMinCellY = std::max(0, (int)floor((y - mnMinY - r) * mfGridElementHeightInv));
    if (nMinCellY >= mnGridRows) return vIndices;
    const int nMaxCellY = std::min((int)mnGridRows - 1, (int)ceil((y - mnMinY + r) * mfGridElementHeightInv));
    if (nMaxCellY < 0) return vIndices;
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const std::vector<size_t>& vCell = mGrid[ix][iy];
            for (size_t j = 0; j < vCell.size(); j++) {
                const cv::KeyPoint& kpUn = mvKeysUn[vCell[j]];
                const float distx = kpUn.pt.x - x, disty = kpUn.pt.y - y;
                if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(vCell[j]);
            }
        }
    return vIndices;
}

struct MatcherAccess : public ORBmatcher {                                  // ComputeThreeMaxima is protected
    MatcherAccess(float r, bool o) : ORBmatcher(r, o) {}
    using ORBmatcher::ComputeThreeMaxima;
};

}  // namespace ORB_SLAM2

using namespace ORB_SLAM2;

extern "C" {

struct refm_side {
    int n;
    const unsigned char* desc;      // n x 32
    const unsigned char* flag;      // n: the feature holds a good MapPoint (may be NULL = none)
    const float *angle, *x, *y;     // keypoint angle / undistorted position (x, y may be NULL)
    const int* octave;              // may be NULL
    const float* uright;            // may be NULL (= -1)
    int n_nodes;
    const int *node_ids, *off, *feat;
};

}

namespace {

template <class T> void fill_common(T& k, const refm_side& s, std::vector<MapPoint>& pool) {
    k.N = s.n;
    k.mDescriptors.create(s.n > 0 ? s.n : 1, 32, CV_8U);
    k.mDescriptors.rows = s.n;
    k.mvKeysUn.resize(s.n);
    k.mvKeys.resize(s.n);
    k.mvuRight.assign(s.n, -1.f);
    k.mvpMapPoints.assign(s.n, static_cast<MapPoint*>(NULL));
    pool.resize(s.n);
    for (int i = 0; i < s.n; i++) {
        memcpy(k.mDescriptors.ptr(i), s.desc + 32 * (size_t)i, 32);
        cv::KeyPoint kp(s.x ? s.x[i] : 0.f, s.y ? s.y[i] : 0.f, 31.f, s.angle ? s.angle[i] : 0.f, 0.f, s.octave ? s.octave[i] : 0, -1);
        k.mvKeysUn[i] = kp;
        k.mvKeys[i] = kp;
        if (s.uright) k.mvuRight[i] = s.uright[i];
        if (s.flag && s.flag[i]) k.mvpMapPoints[i] = &pool[i];
    }
    for (int a = 0; a < s.n_nodes; a++)
        for (int e = s.off[a]; e < s.off[a + 1]; e++) k.mFeatVec[(DBoW2::NodeId)s.node_ids[a]].push_back((unsigned int)s.feat[e]);
}

}  // namespace

extern "C" {

int refm_descriptor_distance(const unsigned char* a, const unsigned char* b) {
    cv::Mat A(1, 32, CV_8U, (void*)a), B(1, 32, CV_8U, (void*)b);
    return ORBmatcher::DescriptorDistance(A, B);
}

void refm_three_maxima(const int* counts, int L, int* out3) {
    std::vector<std::vector<int> > histo(L);
    for (int i = 0; i < L; i++) histo[i].assign(counts[i], 0);
    MatcherAccess m(0.6f, true);
    int i1 = -1, i2 = -1, i3 = -1;
    m.ComputeThreeMaxima(histo.data(), L, i1, i2, i3);
    out3[0] = i1; out3[1] = i2; out3[2] = i3;
}

int refm_search_bow_kf_frame(const refm_side* kf, const refm_side* fr, float nnratio, int check_ori, int* m21) {
    KeyFrame K;
    Frame F;
    std::vector<MapPoint> p1, p2;
    fill_common(K, *kf, p1);
    fill_common(F, *fr, p2);
    std::fill(F.mvpMapPoints.begin(), F.mvpMapPoints.end(), static_cast<MapPoint*>(NULL));
    ORBmatcher m(nnratio, check_ori != 0);
    std::vector<MapPoint*> out;
    const int n = m.SearchByBoW(&K, F, out);
    for (int j = 0; j < fr->n; j++) m21[j] = out[j] ? (int)(out[j] - p1.data()) : -1;
    return n;
}

int refm_search_bow_kf_kf(const refm_side* k1, const refm_side* k2, float nnratio, int check_ori, int* m12) {
    KeyFrame A, B;
    std::vector<MapPoint> p1, p2;
    fill_common(A, *k1, p1);
    fill_common(B, *k2, p2);
    ORBmatcher m(nnratio, check_ori != 0);
    std::vector<MapPoint*> out;
    const int n = m.SearchByBoW(&A, &B, out);
    for (int i = 0; i < k1->n; i++) m12[i] = out[i] ? (int)(out[i] - p2.data()) : -1;
    return n;
}

int refm_search_triangulation(const refm_side* k1, const refm_side* k2, const float* F12, float ex, float ey, const float* sf2,
                              const float* sigma2_2, int nlev, int only_stereo, float nnratio, int check_ori, int* pairs, int cap) {
    KeyFrame A, B;
    std::vector<MapPoint> p1, p2;
    fill_common(A, *k1, p1);
    fill_common(B, *k2, p2);
    // epipole (:667-673): C2 = R2w * Cw + t2w with R2w = I, t2w = 0, Cw = (ex, ey, 1) and unit intrinsics gives exactly (ex, ey)
    A.Ow.create(3, 1, CV_32F);
    A.Ow.at<float>(0) = ex; A.Ow.at<float>(1) = ey; A.Ow.at<float>(2) = 1.f;
    B.Rcw.create(3, 3, CV_32F);
    B.tcw.create(3, 1, CV_32F);
    for (int r = 0; r < 3; r++) {
        B.tcw.at<float>(r) = 0.f;
        for (int c = 0; c < 3; c++) B.Rcw.at<float>(r, c) = r == c ? 1.f : 0.f;
    }
    B.fx = B.fy = 1.f; B.cx = B.cy = 0.f;
    B.mvScaleFactors.assign(sf2, sf2 + nlev);
    B.mvLevelSigma2.assign(sigma2_2, sigma2_2 + nlev);
    cv::Mat F(3, 3, CV_32F);
    for (int i = 0; i < 9; i++) F.at<float>(i / 3, i % 3) = F12[i];
    ORBmatcher m(nnratio, check_ori != 0);
    std::vector<std::pair<size_t, size_t> > out;
    const int n = m.SearchForTriangulation(&A, &B, F, out, only_stereo != 0);
    for (size_t i = 0; i < out.size() && (int)i < cap; i++) { pairs[2 * i] = (int)out[i].first; pairs[2 * i + 1] = (int)out[i].second; }
    return n >= 0 ? (int)out.size() : n;
}

// ---- window searches of the tracker: Frame stand-ins filled from the oracle's grid view (same struct layout) ----
struct refm_grid {
    int n;
    const unsigned char* desc;
    const float *x, *y;
    const int* octave;
    const float *angle, *uright;
    const unsigned char* blocked;       // the feature already holds a MapPoint with Observations() > 0
    int grid_cols, grid_rows;
    float min_x, min_y, max_x, max_y, inv_w, inv_h;
    const int *cell_offsets, *cell_features;    // CSR of mGrid, cell = ix * grid_rows + iy, entries in push order
    const float* scale_factors;
    int n_levels;
};

static void fill_frame(Frame& F, const refm_grid& g, std::vector<MapPoint>& holders) {
    F.N = g.n;
    F.mDescriptors.create(g.n > 0 ? g.n : 1, 32, CV_8U);
    F.mDescriptors.rows = g.n;
    F.mvKeysUn.resize(g.n);
    F.mvuRight.assign(g.n, -1.f);
    F.mvpMapPoints.assign(g.n, static_cast<MapPoint*>(NULL));
    F.mvbOutlier.assign(g.n, false);
    holders.resize(g.n);
    for (int i = 0; i < g.n; i++) {
        memcpy(F.mDescriptors.ptr(i), g.desc + 32 * (size_t)i, 32);
        F.mvKeysUn[i] = cv::KeyPoint(g.x[i], g.y[i], 31.f, g.angle ? g.angle[i] : 0.f, 0.f, g.octave[i], -1);
        if (g.uright) F.mvuRight[i] = g.uright[i];
        if (g.blocked && g.blocked[i]) { holders[i].nObs = 1; F.mvpMapPoints[i] = &holders[i]; }
    }
    F.mvKeys = F.mvKeysUn;
    F.mvScaleFactors.assign(g.scale_factors, g.scale_factors + g.n_levels);
    F.mnScaleLevels = g.n_levels;
    F.mnMinX = g.min_x; F.mnMinY = g.min_y; F.mnMaxX = g.max_x; F.mnMaxY = g.max_y;
    F.mfGridElementWidthInv = g.inv_w; F.mfGridElementHeightInv = g.inv_h;
    assert(g.grid_cols == FRAME_GRID_COLS && g.grid_rows == FRAME_GRID_ROWS);
    for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
        for (int iy = 0; iy < FRAME_GRID_ROWS; iy++) {
            const int c = ix * FRAME_GRID_ROWS + iy;
            for (int e = g.cell_offsets[c]; e < g.cell_offsets[c + 1]; e++) F.mGrid[ix][iy].push_back((size_t)g.cell_features[e]);
        }
}

static void owners_out(const Frame& F, const std::vector<MapPoint>& pool, int* owner) {
    for (int i = 0; i < F.N; i++) {
        const MapPoint* p = F.mvpMapPoints[i];
        owner[i] = (p && !pool.empty() && p >= pool.data() && p < pool.data() + pool.size()) ? (int)(p - pool.data()) : -1;
    }
}

// SearchByProjection(Frame&, const vector<MapPoint*>&, th)  (:45-129)
int refm_search_projection_map(const refm_grid* g, int npts, const unsigned char* in_view, const float* proj_x, const float* proj_y,
                               const float* proj_xr, const int* level, const float* view_cos, const unsigned char* desc,
                               const unsigned char* claims, float th, float nnratio, int* owner) {
    Frame F;
    std::vector<MapPoint> holders, pool(npts);
    fill_frame(F, *g, holders);
    std::vector<MapPoint*> pts(npts);
    for (int i = 0; i < npts; i++) {
        MapPoint& p = pool[i];
        p.mbTrackInView = in_view[i] != 0;
        p.mTrackProjX = proj_x[i]; p.mTrackProjY = proj_y[i]; p.mTrackProjXR = proj_xr[i];
        p.mnTrackScaleLevel = level[i]; p.mTrackViewCos = view_cos[i];
        p.nObs = claims[i] ? 1 : 0;
        p.mDescriptor.create(1, 32, CV_8U);
        memcpy(p.mDescriptor.ptr(), desc + 32 * (size_t)i, 32);
        pts[i] = &p;
    }
    ORBmatcher m(nnratio, true);
    const int n = m.SearchByProjection(F, pts, th);
    owners_out(F, pool, owner);
    return n;
}

// SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono)  (:1331-1463); owner: -1 also for entries the rotation
// cull set back to NULL
int refm_search_projection_frame(const refm_grid* g, const float* Tcw, const float* Tlw, float fx, float fy, float cx, float cy, float mbf,
                                 float mb, int n_last, const unsigned char* has_point, const float* world, const int* octave,
                                 const float* angle, const unsigned char* desc, const unsigned char* claims, float th, int mono,
                                 int check_ori, float nnratio, int* owner) {
    Frame C, L;
    std::vector<MapPoint> holders, pool(n_last);
    fill_frame(C, *g, holders);
    auto pose = [](Frame& F, const float* T) {
        F.mTcw.create(4, 4, CV_32F);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) F.mTcw.at<float>(r, c) = T[4 * r + c];
        F.mTcw.at<float>(3, 0) = F.mTcw.at<float>(3, 1) = F.mTcw.at<float>(3, 2) = 0.f;
        F.mTcw.at<float>(3, 3) = 1.f;
    };
    pose(C, Tcw);
    pose(L, Tlw);
    C.fx = fx; C.fy = fy; C.cx = cx; C.cy = cy; C.mbf = mbf; C.mb = mb;
    L.N = n_last;
    L.mvKeys.resize(n_last);
    L.mvKeysUn.resize(n_last);
    L.mvpMapPoints.assign(n_last, static_cast<MapPoint*>(NULL));
    L.mvbOutlier.assign(n_last, false);
    for (int i = 0; i < n_last; i++) {
        L.mvKeys[i] = L.mvKeysUn[i] = cv::KeyPoint(0.f, 0.f, 31.f, angle[i], 0.f, octave[i], -1);
        if (!has_point[i]) continue;
        MapPoint& p = pool[i];
        p.nObs = claims[i] ? 1 : 0;
        p.mWorldPos.create(3, 1, CV_32F);
        for (int c = 0; c < 3; c++) p.mWorldPos.at<float>(c) = world[3 * (size_t)i + c];
        p.mDescriptor.create(1, 32, CV_8U);
        memcpy(p.mDescriptor.ptr(), desc + 32 * (size_t)i, 32);
        L.mvpMapPoints[i] = &p;
    }
    ORBmatcher m(nnratio, check_ori != 0);
    const int n = m.SearchByProjection(C, L, th, mono != 0);
    owners_out(C, pool, owner);
    return n;
}

// SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)  (:408-523); prev_matched: n1 x 2 floats, updated in place
int refm_search_initialization(const refm_grid* g2, int n1, const unsigned char* desc1, const int* octave1, const float* angle1,
                               float* prev_matched, int window_size, float nnratio, int check_ori, int* m12) {
    Frame F1, F2;
    std::vector<MapPoint> holders;
    fill_frame(F2, *g2, holders);
    F1.N = n1;
    F1.mvKeysUn.resize(n1);
    F1.mDescriptors.create(n1 > 0 ? n1 : 1, 32, CV_8U);
    F1.mDescriptors.rows = n1;
    std::vector<cv::Point2f> prev(n1);
    for (int i = 0; i < n1; i++) {
        F1.mvKeysUn[i] = cv::KeyPoint(prev_matched[2 * i], prev_matched[2 * i + 1], 31.f, angle1[i], 0.f, octave1[i], -1);
        memcpy(F1.mDescriptors.ptr(i), desc1 + 32 * (size_t)i, 32);
        prev[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
    }
    ORBmatcher m(nnratio, check_ori != 0);
    std::vector<int> out;
    const int n = m.SearchForInitialization(F1, F2, prev, out, window_size);
    for (int i = 0; i < n1; i++) {
        m12[i] = out[i];
        prev_matched[2 * i] = prev[i].x;
        prev_matched[2 * i + 1] = prev[i].y;
    }
    return n;
}

static void fill_points(std::vector<MapPoint>& pool, int npts, const unsigned char* state, const float* world, const float* normal,
                        const float* mf_max, const float* mf_min, const unsigned char* desc) {
    pool.resize(npts);
    for (int i = 0; i < npts; i++) {
        MapPoint& p = pool[i];
        p.mbBad = state[i] == 2;
        p.nObs = 1;
        p.mfMaxDistance = mf_max[i];
        p.mfMinDistance = mf_min[i];
        p.mWorldPos.create(3, 1, CV_32F);
        p.mNormalVector.create(3, 1, CV_32F);
        for (int c = 0; c < 3; c++) {
            p.mWorldPos.at<float>(c) = world[3 * (size_t)i + c];
            p.mNormalVector.at<float>(c) = normal ? normal[3 * (size_t)i + c] : 0.f;
        }
        p.mDescriptor.create(1, 32, CV_8U);
        memcpy(p.mDescriptor.ptr(), desc + 32 * (size_t)i, 32);
    }
}

// SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist)  (:1465-1602, relocalisation)
// state[i]: 0 = the keyframe feature holds no MapPoint, 1 = good, 2 = bad, 3 = already found
int refm_search_projection_kf(const refm_grid* g, const float* Tcw, float fx, float fy, float cx, float cy, float log_sf, int npts,
                              const unsigned char* state, const float* world, const float* mf_max, const float* mf_min, const float* angle,
                              const unsigned char* desc, float th, int orb_dist, int check_ori, int* owner) {
    Frame C;
    std::vector<MapPoint> holders, pool;
    fill_frame(C, *g, holders);
    C.mTcw.create(4, 4, CV_32F);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) C.mTcw.at<float>(r, c) = Tcw[4 * r + c];
    C.mTcw.at<float>(3, 0) = C.mTcw.at<float>(3, 1) = C.mTcw.at<float>(3, 2) = 0.f;
    C.mTcw.at<float>(3, 3) = 1.f;
    C.fx = fx; C.fy = fy; C.cx = cx; C.cy = cy; C.mfLogScaleFactor = log_sf;
    fill_points(pool, npts, state, world, nullptr, mf_max, mf_min, desc);
    KeyFrame K;
    K.N = npts;
    K.mvKeysUn.resize(npts);
    K.mvpMapPoints.assign(npts, static_cast<MapPoint*>(NULL));
    std::set<MapPoint*> found;
    for (int i = 0; i < npts; i++) {
        K.mvKeysUn[i] = cv::KeyPoint(0.f, 0.f, 31.f, angle[i], 0.f, 0, -1);
        if (state[i] != 0) K.mvpMapPoints[i] = &pool[i];
        if (state[i] == 3) found.insert(&pool[i]);
    }
    ORBmatcher m(0.9f, check_ori != 0);
    const int n = m.SearchByProjection(C, &K, found, th, orb_dist);
    owners_out(C, pool, owner);
    return n;
}

// SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th)  (:293-406, loop
// closing); points whose state is not 1 are handed over as bad (skipped either way)
int refm_search_projection_sim3(const refm_grid* g, const float* Scw, float fx, float fy, float cx, float cy, float log_sf, int npts,
                                const unsigned char* state, const float* world, const float* mf_max, const float* mf_min, const float* normal,
                                const unsigned char* desc, int th, int* owner) {
    KeyFrame K;
    std::vector<MapPoint> holders(g->n), pool;
    K.N = g->n;
    K.fx = fx; K.fy = fy; K.cx = cx; K.cy = cy; K.mfLogScaleFactor = log_sf;
    K.mvKeysUn.resize(g->n);
    K.mDescriptors.create(g->n > 0 ? g->n : 1, 32, CV_8U);
    K.mDescriptors.rows = g->n;
    std::vector<MapPoint*> matched(g->n, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < g->n; i++) {
        memcpy(K.mDescriptors.ptr(i), g->desc + 32 * (size_t)i, 32);
        K.mvKeysUn[i] = cv::KeyPoint(g->x[i], g->y[i], 31.f, g->angle ? g->angle[i] : 0.f, 0.f, g->octave[i], -1);
        if (g->blocked && g->blocked[i]) matched[i] = &holders[i];
    }
    K.mvScaleFactors.assign(g->scale_factors, g->scale_factors + g->n_levels);
    K.mnMinX = (int)g->min_x; K.mnMinY = (int)g->min_y; K.mnMaxX = (int)g->max_x; K.mnMaxY = (int)g->max_y;
    K.mnGridCols = g->grid_cols; K.mnGridRows = g->grid_rows;
    K.mfGridElementWidthInv = g->inv_w; K.mfGridElementHeightInv = g->inv_h;
    K.mGrid.assign(g->grid_cols, std::vector<std::vector<size_t> >(g->grid_rows));
    for (int ix = 0; ix < g->grid_cols; ix++)
        for (int iy = 0; iy < g->grid_rows; iy++) {
            const int c = ix * g->grid_rows + iy;
            for (int e = g->cell_offsets[c]; e < g->cell_offsets[c + 1]; e++) K.mGrid[ix][iy].push_back((size_t)g->cell_features[e]);
        }
    fill_points(pool, npts, state, world, normal, mf_max, mf_min, desc);
    std::vector<MapPoint*> pts(npts);
    for (int i = 0; i < npts; i++) {
        if (state[i] != 1) pool[i].mbBad = true;
        pts[i] = &pool[i];
    }
    cv::Mat S(4, 4, CV_32F);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) S.at<float>(r, c) = Scw[4 * r + c];
    S.at<float>(3, 0) = S.at<float>(3, 1) = S.at<float>(3, 2) = 0.f;
    S.at<float>(3, 3) = 1.f;
    ORBmatcher m(0.75f, true);
    const int n = m.SearchByProjection(&K, S, pts, matched, th);
    for (int i = 0; i < g->n; i++) {
        const MapPoint* p = matched[i];
        owner[i] = (p && !pool.empty() && p >= pool.data() && p < pool.data() + pool.size()) ? (int)(p - pool.data()) : -1;
    }
    return n;
}

// ---- Fuse x2 and SearchBySim3 (src/ORBmatcher.cc:828-972, 974-1103, 1105-1329): the KeyFrame stand-in filled from the grid view, the
// MapPoint object graph from flat arrays.  Point state: 0 = no point (NULL), 1 = good, 2 = bad, 3 = (candidates only) already observed by
// the keyframe at feature in_at[i].  Results are reported as the graph the reference leaves behind.
static void fill_keyframe(KeyFrame& K, const refm_grid& g, const float* K6, const float* inv_sigma2, const float* T /*3x4*/, const float* Ow) {
    K.N = g.n;
    K.fx = K6[0]; K.fy = K6[1]; K.cx = K6[2]; K.cy = K6[3]; K.mbf = K6[4]; K.mfLogScaleFactor = K6[5];
    K.mvKeysUn.resize(g.n);
    K.mvuRight.assign(g.n, -1.f);
    K.mDescriptors.create(g.n > 0 ? g.n : 1, 32, CV_8U);
    K.mDescriptors.rows = g.n;
    K.mvpMapPoints.assign(g.n, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < g.n; i++) {
        memcpy(K.mDescriptors.ptr(i), g.desc + 32 * (size_t)i, 32);
        K.mvKeysUn[i] = cv::KeyPoint(g.x[i], g.y[i], 31.f, g.angle ? g.angle[i] : 0.f, 0.f, g.octave[i], -1);
        if (g.uright) K.mvuRight[i] = g.uright[i];
    }
    K.mvKeys = K.mvKeysUn;
    K.mnScaleLevels = g.n_levels;
    K.mvScaleFactors.assign(g.scale_factors, g.scale_factors + g.n_levels);
    if (inv_sigma2) K.mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + g.n_levels);
    K.mnMinX = (int)g.min_x; K.mnMinY = (int)g.min_y; K.mnMaxX = (int)g.max_x; K.mnMaxY = (int)g.max_y;
    K.mnGridCols = g.grid_cols; K.mnGridRows = g.grid_rows;
    K.mfGridElementWidthInv = g.inv_w; K.mfGridElementHeightInv = g.inv_h;
    K.mGrid.assign(g.grid_cols, std::vector<std::vector<size_t> >(g.grid_rows));
    for (int ix = 0; ix < g.grid_cols; ix++)
        for (int iy = 0; iy < g.grid_rows; iy++) {
            const int c = ix * g.grid_rows + iy;
            for (int e = g.cell_offsets[c]; e < g.cell_offsets[c + 1]; e++) K.mGrid[ix][iy].push_back((size_t)g.cell_features[e]);
        }
    if (T) {
        K.Rcw.create(3, 3, CV_32F);
        K.tcw.create(3, 1, CV_32F);
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 3; c++) K.Rcw.at<float>(r, c) = T[4 * r + c];
            K.tcw.at<float>(r) = T[4 * r + 3];
        }
    }
    if (Ow) {
        K.Ow.create(3, 1, CV_32F);
        for (int r = 0; r < 3; r++) K.Ow.at<float>(r) = Ow[r];
    }
}

struct refm_points {
    int n;
    const unsigned char* state;
    const int* nobs;
    const unsigned char* desc;
    const float *world, *normal, *mf_max, *mf_min;
};

static void fill_pool(std::vector<MapPoint>& pool, const refm_points& p, long unsigned int id0) {
    fill_points(pool, p.n, p.state, p.world, p.normal, p.mf_max, p.mf_min, p.desc);
    for (int i = 0; i < p.n; i++) { pool[i].nObs = p.nobs ? p.nobs[i] : 1; pool[i].mnId = id0 + i; }
}

static int encode_point(const MapPoint* p, const std::vector<MapPoint>& kfp, const std::vector<MapPoint>& cand) {
    if (!p) return -1;
    if (!kfp.empty() && p >= kfp.data() && p < kfp.data() + kfp.size()) return (int)(p - kfp.data());
    if (!cand.empty() && p >= cand.data() && p < cand.data() + cand.size()) return 100000 + (int)(p - cand.data());
    return -2;
}

// variant 0: Fuse(pKF, vpMapPoints, th); variant 1: Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) with Scw 4x4 row-major.
// kf_ptr_out[g->n]: which point every keyframe feature holds afterwards (index into the keyframe's pool, 100000 + candidate index, -1);
// pts_out[(g->n + cand->n) x 2]: isBad, Observations of every point; replace_out[cand->n] (variant 1): vpReplacePoint.
int refm_fuse(int variant, const refm_grid* g, const float* K6, const float* inv_sigma2, const float* T, const float* Ow, const float* Scw,
              const refm_points* kfpts, const refm_points* cand, const int* in_at, float th, int* kf_ptr_out, int* pts_out, int* replace_out) {
    KeyFrame K;
    fill_keyframe(K, *g, K6, inv_sigma2, T, Ow);
    std::vector<MapPoint> kfp, cp;
    fill_pool(kfp, *kfpts, 0);
    fill_pool(cp, *cand, 1000000);
    for (int j = 0; j < g->n; j++)
        if (kfpts->state[j]) { K.mvpMapPoints[j] = &kfp[j]; kfp[j].mObservations[&K] = (size_t)j; }
    std::vector<MapPoint*> pts(cand->n, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < cand->n; i++) {
        if (cand->state[i] == 0) continue;
        pts[i] = &cp[i];
        if (cand->state[i] == 3) { cp[i].mObservations[&K] = (size_t)in_at[i]; K.mvpMapPoints[in_at[i]] = &cp[i]; }
    }
    ORBmatcher m(0.8f, true);
    int n;
    if (variant == 0) n = m.Fuse(&K, pts, th);
    else {
        cv::Mat S(4, 4, CV_32F);
        for (int i = 0; i < 16; i++) S.at<float>(i / 4, i % 4) = Scw[i];
        std::vector<MapPoint*> rep(cand->n, static_cast<MapPoint*>(NULL));
        n = m.Fuse(&K, S, pts, th, rep);
        for (int i = 0; i < cand->n; i++) replace_out[i] = encode_point(rep[i], kfp, cp);
    }
    for (int j = 0; j < g->n; j++) kf_ptr_out[j] = encode_point(K.mvpMapPoints[j], kfp, cp);
    for (int j = 0; j < g->n; j++) { pts_out[2 * j] = kfp[j].isBad(); pts_out[2 * j + 1] = kfp[j].Observations(); }
    for (int i = 0; i < cand->n; i++) { pts_out[2 * (g->n + i)] = cp[i].isBad(); pts_out[2 * (g->n + i) + 1] = cp[i].Observations(); }
    return n;
}

// SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th)  (:1105-1329).  pre12[n1]: index of the KF2 feature whose point vpMatches12[i]
// holds on entry (-1 = NULL); matches12_out[n1]: the same view of vpMatches12 afterwards.
int refm_search_by_sim3(const refm_grid* g1, const refm_grid* g2, const float* K6, const float* T1, const float* T2, const refm_points* p1,
                        const refm_points* p2, float s12, const float* R12, const float* t12, float th, const int* pre12, int* matches12_out) {
    KeyFrame A, B;
    fill_keyframe(A, *g1, K6, nullptr, T1, nullptr);
    fill_keyframe(B, *g2, K6, nullptr, T2, nullptr);
    std::vector<MapPoint> a, b, none;
    fill_pool(a, *p1, 0);
    fill_pool(b, *p2, 1000000);
    for (int i = 0; i < g1->n; i++) if (p1->state[i]) { A.mvpMapPoints[i] = &a[i]; a[i].mObservations[&A] = (size_t)i; }
    for (int i = 0; i < g2->n; i++) if (p2->state[i]) { B.mvpMapPoints[i] = &b[i]; b[i].mObservations[&B] = (size_t)i; }
    std::vector<MapPoint*> m12(g1->n, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < g1->n; i++) if (pre12[i] >= 0) m12[i] = &b[pre12[i]];
    cv::Mat R(3, 3, CV_32F), t(3, 1, CV_32F);
    for (int r = 0; r < 3; r++) { t.at<float>(r) = t12[r]; for (int c = 0; c < 3; c++) R.at<float>(r, c) = R12[3 * r + c]; }
    ORBmatcher m(0.75f, true);
    const int n = m.SearchBySim3(&A, &B, m12, s12, R, t, th);
    for (int i = 0; i < g1->n; i++) {
        const MapPoint* p = m12[i];
        matches12_out[i] = (p && p >= b.data() && p < b.data() + b.size()) ? (int)(p - b.data()) : -1;
    }
    return n;
}

}  // extern "C"