// host_api_driver.cc — exercises the C++ drop-in classes (ORB_SLAM2::ORBextractor / ORBmatcher) the way the reference's
// Frame / LocalMapping code does, on inputs written by tests/test_gpu_host_cpp.py; writes results back as raw binaries.
//   driver extract <in.raw> <w> <h> <nfeat> <scale> <nlevels> <ini> <min> <out_prefix> [mask.raw]
//   driver match   <in.bin> <out.bin>
//   driver project <in.bin> <out.bin>
//   driver project2 <in.bin> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "ORBextractor.h"
#include "ORBmatcher.h"
#include "ORBVocabulary.h"
#include "MapArchive.h"

using namespace ORB_SLAM2;

static std::vector<unsigned char> slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
template <typename T> static void put(std::ofstream& f, const T* p, size_t n) { f.write((const char*)p, sizeof(T) * n); }

static int run_extract(int argc, char** argv) {
    if (argc < 11) return 2;
    const int w = atoi(argv[3]), h = atoi(argv[4]);
    std::vector<unsigned char> img = slurp(argv[2]);
    if ((int)img.size() != w * h) return 3;
    ORBextractor ex(atoi(argv[5]), (float)atof(argv[6]), atoi(argv[7]), atoi(argv[8]), atoi(argv[9]));
    cv::Mat image(h, w, CV_8U, img.data());
    std::vector<unsigned char> mk;
    cv::Mat mask;
    if (argc > 11) { mk = slurp(argv[11]); mask = cv::Mat(h, w, CV_8U, mk.data()); }
    std::vector<cv::KeyPoint> kps(3, cv::KeyPoint(1, 2, 3));       // must be cleared and refilled
    cv::Mat desc;
    ex(image, mask, kps, desc);
    if (ex.LastStatus() != 0) return 4;
    std::string pre(argv[10]);
    { std::ofstream f(pre + ".kp", std::ios::binary); put(f, kps.data(), kps.size()); }
    { std::ofstream f(pre + ".desc", std::ios::binary); for (int i = 0; i < desc.rows; i++) put(f, desc.ptr(i), 32); }
    {   // getters + the public pyramid (ROI views into bordered buffers, like the reference)
        std::ofstream f(pre + ".meta", std::ios::binary);
        const int nl = ex.GetLevels();
        put(f, &nl, 1);
        const float sf = ex.GetScaleFactor();
        put(f, &sf, 1);
        std::vector<float> a = ex.GetScaleFactors(), b = ex.GetInverseScaleFactors(), c = ex.GetScaleSigmaSquares(), d = ex.GetInverseScaleSigmaSquares();
        put(f, a.data(), nl); put(f, b.data(), nl); put(f, c.data(), nl); put(f, d.data(), nl);
        for (int l = 0; l < nl; l++) {
            const cv::Mat& m = ex.mvImagePyramid[l];
            const int dims[2] = {m.cols, m.rows};
            put(f, dims, 2);
            for (int y = -19; y < m.rows + 19; y++) put(f, m.data + (ptrdiff_t)y * (ptrdiff_t)m.step - 19, m.cols + 38);   // the border is addressable around the ROI
        }
    }
    if (mask.empty()) {   // batch overload: separately allocated frames; every copy of `image` must come back exactly like the single call
        const int nb = 150;                                        // > 128: two chunks through the pipelined path
        std::vector<std::vector<unsigned char> > store(nb, img);
        std::vector<cv::Mat> frames(nb);
        for (int i = 0; i < nb; i++) {
            if (i % 3 == 1) for (int y = 0; y < h; y++) for (int x = 0; x + 1 < w; x++) store[i][(size_t)y * w + x] = img[(size_t)y * w + x + 1];   // a different frame
            frames[i] = cv::Mat(h, w, CV_8U, store[i].data());
        }
        std::vector<std::vector<cv::KeyPoint> > bk;
        std::vector<cv::Mat> bd;
        ex(frames, bk, bd);
        if (ex.LastStatus() != 0 || (int)bk.size() != nb) return 6;
        std::vector<cv::KeyPoint> k1; cv::Mat d1;
        ex(frames[1], cv::Mat(), k1, d1);
        for (int i = 0; i < nb; i++) {
            const std::vector<cv::KeyPoint>& want = (i % 3 == 1) ? k1 : kps;
            const cv::Mat& wd = (i % 3 == 1) ? d1 : desc;
            if (bk[i].size() != want.size() || bd[i].rows != wd.rows) return 7;
            if (memcmp(bk[i].data(), want.data(), want.size() * sizeof(cv::KeyPoint))) return 8;
            for (int r = 0; r < wd.rows; r++) if (memcmp(bd[i].ptr(r), wd.ptr(r), 32)) return 9;
        }
    }
    // empty image: silent no-op, outputs untouched (reference :1045-1046)
    cv::Mat empty;
    std::vector<cv::KeyPoint> keep(kps);
    ex(empty, empty, keep, desc);
    return keep.size() == kps.size() ? 0 : 5;
}

struct Reader {
    std::vector<unsigned char> buf; size_t pos = 0;
    template <typename T> T get() { T v; memcpy(&v, &buf[pos], sizeof(T)); pos += sizeof(T); return v; }
    template <typename T> std::vector<T> arr(size_t n) { std::vector<T> v(n); if (n) memcpy(v.data(), &buf[pos], n * sizeof(T)); pos += n * sizeof(T); return v; }
};

struct World {
    KeyFrame kf[2];
    Frame fr;
    std::vector<MapPoint> pool[2];
};

static void load_side(Reader& r, KeyFrame& kf, std::vector<MapPoint>& pool, Frame* alsoFrame) {
    const int n = r.get<int>();
    kf.N = n;
    std::vector<unsigned char> desc = r.arr<unsigned char>((size_t)n * 32);
    kf.mDescriptors = cv::Mat(n, 32, CV_8U);
    for (int i = 0; i < n; i++) memcpy(kf.mDescriptors.ptr(i), &desc[(size_t)i * 32], 32);
    std::vector<int> node = r.arr<int>(n);
    std::vector<unsigned char> state = r.arr<unsigned char>(n);    // 0 = no MapPoint, 1 = good MapPoint, 2 = bad MapPoint
    std::vector<float> ang = r.arr<float>(n), x = r.arr<float>(n), y = r.arr<float>(n), ur = r.arr<float>(n);
    std::vector<int> oct = r.arr<int>(n);
    pool.resize(n);
    kf.mvpMapPoints.assign(n, (MapPoint*)NULL);
    kf.mvKeysUn.resize(n); kf.mvKeys.resize(n); kf.mvuRight = ur;
    for (int i = 0; i < n; i++) {
        if (state[i]) { pool[i].mbBad = state[i] == 2; kf.mvpMapPoints[i] = &pool[i]; }
        kf.mvKeysUn[i] = cv::KeyPoint(x[i], y[i], 31, ang[i], 10, oct[i]);
        kf.mvKeys[i] = kf.mvKeysUn[i];
        kf.mFeatVec.addFeature((DBoW2::NodeId)node[i], (unsigned)i);
    }
    if (alsoFrame) {
        alsoFrame->N = n; alsoFrame->mDescriptors = kf.mDescriptors; alsoFrame->mvKeys = kf.mvKeys; alsoFrame->mvKeysUn = kf.mvKeysUn;
        alsoFrame->mFeatVec = kf.mFeatVec; alsoFrame->mvuRight = ur;
    }
}

static int run_match(int argc, char** argv) {
    if (argc < 4) return 2;
    Reader r; r.buf = slurp(argv[2]);
    World w;
    load_side(r, w.kf[0], w.pool[0], NULL);
    load_side(r, w.kf[1], w.pool[1], &w.fr);
    const float ratio = r.get<float>();
    const int ori = r.get<int>(), onlyStereo = r.get<int>();
    std::vector<float> F = r.arr<float>(9), Ow = r.arr<float>(3), R = r.arr<float>(9), t = r.arr<float>(3), K = r.arr<float>(4);
    const int nlev = r.get<int>();
    w.kf[1].mvScaleFactors = r.arr<float>(nlev);
    w.kf[1].mvLevelSigma2 = r.arr<float>(nlev);
    w.kf[0].Ow = cv::Mat(3, 1, CV_32F); w.kf[1].Rcw = cv::Mat(3, 3, CV_32F); w.kf[1].tcw = cv::Mat(3, 1, CV_32F);
    for (int i = 0; i < 3; i++) { w.kf[0].Ow.at<float>(i) = Ow[i]; w.kf[1].tcw.at<float>(i) = t[i]; for (int j = 0; j < 3; j++) w.kf[1].Rcw.at<float>(i, j) = R[3 * i + j]; }
    w.kf[1].fx = K[0]; w.kf[1].fy = K[1]; w.kf[1].cx = K[2]; w.kf[1].cy = K[3];
    cv::Mat F12(3, 3, CV_32F);
    for (int i = 0; i < 9; i++) F12.at<float>(i / 3, i % 3) = F[i];

    std::ofstream out(argv[3], std::ios::binary);
    ORBmatcher m(ratio, ori != 0);
    // SearchByBoW(KF, Frame): report, per Frame feature, the index of the KF feature whose MapPoint was assigned
    {
        std::vector<MapPoint*> res;
        const int n = m.SearchByBoW(&w.kf[0], w.fr, res);
        put(out, &n, 1);
        std::vector<int> idx(res.size(), -1);
        for (size_t j = 0; j < res.size(); j++) if (res[j]) idx[j] = (int)(res[j] - &w.pool[0][0]);
        const int sz = (int)idx.size(); put(out, &sz, 1); put(out, idx.data(), idx.size());
    }
    {
        std::vector<MapPoint*> res;
        const int n = m.SearchByBoW(&w.kf[0], &w.kf[1], res);
        put(out, &n, 1);
        std::vector<int> idx(res.size(), -1);
        for (size_t j = 0; j < res.size(); j++) if (res[j]) idx[j] = (int)(res[j] - &w.pool[1][0]);
        const int sz = (int)idx.size(); put(out, &sz, 1); put(out, idx.data(), idx.size());
    }
    {
        std::vector<std::pair<size_t, size_t> > pairs(1, std::make_pair((size_t)7, (size_t)7));
        const int n = m.SearchForTriangulation(&w.kf[0], &w.kf[1], F12, pairs, onlyStereo != 0);
        put(out, &n, 1);
        const int sz = (int)pairs.size(); put(out, &sz, 1);
        for (size_t i = 0; i < pairs.size(); i++) { const int p[2] = {(int)pairs[i].first, (int)pairs[i].second}; put(out, p, 2); }
    }
    {
        const int d = ORBmatcher::DescriptorDistance(w.kf[0].mDescriptors.row(0), w.kf[1].mDescriptors.row(0));
        put(out, &d, 1);
        const int c[3] = {ORBmatcher::TH_LOW, ORBmatcher::TH_HIGH, ORBmatcher::HISTO_LENGTH};
        put(out, c, 3);
    }
    {   // batched overloads: {kf0, kf0} against the frame and kf0 against {kf1, kf1} — every element must equal the single call
        std::vector<KeyFrame*> two0(2, &w.kf[0]), two1(2, &w.kf[1]);
        std::vector<std::vector<MapPoint*> > res;
        std::vector<int> cnt = m.SearchByBoW(two0, w.fr, res);
        for (int c = 0; c < 2; c++) {
            put(out, &cnt[c], 1);
            std::vector<int> idx(res[c].size(), -1);
            for (size_t j = 0; j < res[c].size(); j++) if (res[c][j]) idx[j] = (int)(res[c][j] - &w.pool[0][0]);
            const int sz = (int)idx.size(); put(out, &sz, 1); put(out, idx.data(), idx.size());
        }
        cnt = m.SearchByBoW(&w.kf[0], two1, res);
        for (int c = 0; c < 2; c++) {
            put(out, &cnt[c], 1);
            std::vector<int> idx(res[c].size(), -1);
            for (size_t j = 0; j < res[c].size(); j++) if (res[c][j]) idx[j] = (int)(res[c][j] - &w.pool[1][0]);
            const int sz = (int)idx.size(); put(out, &sz, 1); put(out, idx.data(), idx.size());
        }
        const int nd = std::min(w.kf[0].mDescriptors.rows, w.kf[1].mDescriptors.rows);
        std::vector<int> dist(nd > 0 ? nd : 1, -1);
        const int got = ORBmatcher::DescriptorDistances(w.kf[0].mDescriptors, w.kf[1].mDescriptors, dist.data());
        put(out, &got, 1); put(out, dist.data(), (size_t)got);
    }
    {   // batched SearchForTriangulation: kf0 against {kf1, kf1, kf1} — every element must equal the single call
        std::vector<std::pair<size_t, size_t> > single;
        const int n1 = m.SearchForTriangulation(&w.kf[0], &w.kf[1], F12, single, onlyStereo != 0);
        std::vector<std::vector<std::pair<size_t, size_t> > > many;
        const std::vector<int> nb = m.SearchForTriangulation(&w.kf[0], std::vector<KeyFrame*>(3, &w.kf[1]), std::vector<cv::Mat>(3, F12), many,
                                                             onlyStereo != 0);
        if (nb.size() != 3 || many.size() != 3) return 10;
        for (int c = 0; c < 3; c++)
            if (nb[c] != n1 || many[c] != single) return 11;
    }
    return ORBmatcher::LastStatus() == 0 ? 0 : 4;
}

// driver bow <voc.bin|voc.txt> <desc.raw> <levelsup> <out.bin>: Frame::ComputeBoW the way the reference does it
static int run_bow(int argc, char** argv) {
    if (argc < 6) return 2;
    ORBVocabularyB200 voc;
    const std::string path(argv[2]);
    const bool ok = path.size() > 4 && path.substr(path.size() - 4) == ".txt" ? voc.loadFromTextFile(path) : voc.loadFromBinaryFile(path);
    if (!ok) return 3;
    std::vector<unsigned char> d = slurp(argv[3]);
    const int n = (int)(d.size() / 32);
    std::vector<cv::Mat> feats(n);
    for (int i = 0; i < n; i++) feats[i] = cv::Mat(1, 32, CV_8U, &d[(size_t)i * 32]);    // Converter::toDescriptorVector
    DBoW2::BowVector bow;
    DBoW2::FeatureVector fv;
    voc.transform(feats, bow, fv, atoi(argv[4]));
    if (voc.LastStatus() != 0) return 4;
    std::ofstream out(argv[5], std::ios::binary);
    const int nb = (int)bow.size(), nn = (int)fv.size(), nw = (int)voc.size();
    put(out, &nw, 1); put(out, &nb, 1);
    for (DBoW2::BowVector::const_iterator it = bow.begin(); it != bow.end(); ++it) { const int w = (int)it->first; put(out, &w, 1); put(out, &it->second, 1); }
    put(out, &nn, 1);
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
        const int id = (int)it->first, c = (int)it->second.size();
        put(out, &id, 1); put(out, &c, 1);
        for (int j = 0; j < c; j++) { const int f = (int)it->second[j]; put(out, &f, 1); }
    }
    return 0;
}


// driver project <in.bin> <out.bin>: the window searches the way Tracking calls them (SearchLocalPoints, TrackWithMotionModel,
// MonocularInitialization).  The Frame is assembled like Frame::Frame does: AssignFeaturesToGrid with PosInGrid's round().
#include <algorithm>
#include <cmath>
static void build_frame(Reader& r, Frame& F) {
    const int n = r.get<int>();
    F.N = n;
    std::vector<unsigned char> desc = r.arr<unsigned char>((size_t)n * 32);
    F.mDescriptors = cv::Mat(n, 32, CV_8U);
    for (int i = 0; i < n; i++) memcpy(F.mDescriptors.ptr(i), &desc[(size_t)i * 32], 32);
    std::vector<float> x = r.arr<float>(n), y = r.arr<float>(n), ang = r.arr<float>(n), ur = r.arr<float>(n);
    std::vector<int> oct = r.arr<int>(n);
    std::vector<float> b = r.arr<float>(4);
    const int nlev = r.get<int>();
    F.mvScaleFactors = r.arr<float>(nlev);
    F.mnMinX = b[0]; F.mnMinY = b[1]; F.mnMaxX = b[2]; F.mnMaxY = b[3];
    F.mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (F.mnMaxX - F.mnMinX);
    F.mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (F.mnMaxY - F.mnMinY);
    F.mvKeysUn.resize(n); F.mvKeys.resize(n); F.mvuRight = ur;
    F.mvpMapPoints.assign(n, (MapPoint*)NULL);
    F.mvbOutlier.assign(n, false);
    for (int i = 0; i < n; i++) {
        F.mvKeysUn[i] = cv::KeyPoint(x[i], y[i], 31, ang[i], 10, oct[i]);
        F.mvKeys[i] = F.mvKeysUn[i];
        const int px = (int)round((x[i] - F.mnMinX) * F.mfGridElementWidthInv), py = (int)round((y[i] - F.mnMinY) * F.mfGridElementHeightInv);
        if (px < 0 || px >= FRAME_GRID_COLS || py < 0 || py >= FRAME_GRID_ROWS) continue;
        F.mGrid[px][py].push_back(i);
    }
}
static cv::Mat mat_from(const std::vector<float>& v, int rows, int cols) {
    cv::Mat m(rows, cols, CV_32F);
    for (int i = 0; i < rows * cols; i++) m.at<float>(i / cols, i % cols) = v[i];
    return m;
}
static int run_project(int argc, char** argv) {
    if (argc < 4) return 2;
    Reader r; r.buf = slurp(argv[2]);
    std::ofstream out(argv[3], std::ios::binary);
    const float ratio = r.get<float>();
    const int ori = r.get<int>();
    ORBmatcher m(ratio, ori != 0);
    {   // SearchByProjection(F, vpMapPoints, th)
        Frame* F = new Frame();
        build_frame(r, *F);
        std::vector<unsigned char> blocked = r.arr<unsigned char>(F->N);
        std::vector<MapPoint> own(F->N);                 // pre-existing points of the frame: observed ones block their feature
        for (int i = 0; i < F->N; i++) { own[i].nObs = blocked[i] ? 2 : 0; F->mvpMapPoints[i] = (i % 3 == 0 || blocked[i]) ? &own[i] : NULL; }
        const int np = r.get<int>();
        const float th = r.get<float>();
        std::vector<unsigned char> in_view = r.arr<unsigned char>(np), claims = r.arr<unsigned char>(np), d = r.arr<unsigned char>((size_t)np * 32);
        std::vector<float> px = r.arr<float>(np), py = r.arr<float>(np), pxr = r.arr<float>(np), vc = r.arr<float>(np);
        std::vector<int> lvl = r.arr<int>(np);
        std::vector<MapPoint> pool(np);
        std::vector<MapPoint*> vp(np);
        for (int i = 0; i < np; i++) {
            pool[i].mbTrackInView = in_view[i] != 0; pool[i].mTrackProjX = px[i]; pool[i].mTrackProjY = py[i]; pool[i].mTrackProjXR = pxr[i];
            pool[i].mnTrackScaleLevel = lvl[i]; pool[i].mTrackViewCos = vc[i]; pool[i].nObs = claims[i] ? 3 : 0;
            pool[i].mDescriptor = cv::Mat(1, 32, CV_8U); memcpy(pool[i].mDescriptor.ptr(0), &d[(size_t)i * 32], 32);
            vp[i] = &pool[i];
        }
        const int n = m.SearchByProjection(*F, vp, th);
        put(out, &n, 1);
        std::vector<int> owner(F->N, -1);
        for (int j = 0; j < F->N; j++)
            if (F->mvpMapPoints[j] && np && F->mvpMapPoints[j] >= &pool[0] && F->mvpMapPoints[j] <= &pool[np - 1]) owner[j] = (int)(F->mvpMapPoints[j] - &pool[0]);
        put(out, owner.data(), owner.size());
        delete F;
    }
    {   // SearchByProjection(CurrentFrame, LastFrame, th, bMono)
        Frame* C = new Frame(); Frame* L = new Frame();
        build_frame(r, *C);
        std::vector<unsigned char> blocked = r.arr<unsigned char>(C->N);
        std::vector<MapPoint> own(C->N);
        for (int i = 0; i < C->N; i++) { own[i].nObs = blocked[i] ? 2 : 0; C->mvpMapPoints[i] = (i % 3 == 0 || blocked[i]) ? &own[i] : NULL; }
        C->mTcw = mat_from(r.arr<float>(16), 4, 4); L->mTcw = mat_from(r.arr<float>(16), 4, 4);
        std::vector<float> K = r.arr<float>(6);
        C->fx = K[0]; C->fy = K[1]; C->cx = K[2]; C->cy = K[3]; C->mbf = K[4]; C->mb = K[5];
        const float th = r.get<float>();
        const int mono = r.get<int>(), nl = r.get<int>();
        std::vector<unsigned char> has = r.arr<unsigned char>(nl), claims = r.arr<unsigned char>(nl), d = r.arr<unsigned char>((size_t)nl * 32);
        std::vector<float> world = r.arr<float>((size_t)nl * 3), ang = r.arr<float>(nl);
        std::vector<int> oct = r.arr<int>(nl);
        L->N = nl; L->mvKeys.resize(nl); L->mvKeysUn.resize(nl); L->mvpMapPoints.assign(nl, (MapPoint*)NULL); L->mvbOutlier.assign(nl, false);
        std::vector<MapPoint> pool(nl);
        for (int i = 0; i < nl; i++) {
            L->mvKeys[i] = cv::KeyPoint(0, 0, 31, ang[i], 10, oct[i]); L->mvKeysUn[i] = L->mvKeys[i];
            pool[i].nObs = claims[i] ? 3 : 0;
            pool[i].mDescriptor = cv::Mat(1, 32, CV_8U); memcpy(pool[i].mDescriptor.ptr(0), &d[(size_t)i * 32], 32);
            pool[i].mWorldPos = cv::Mat(3, 1, CV_32F);
            for (int k = 0; k < 3; k++) pool[i].mWorldPos.at<float>(k) = world[3 * (size_t)i + k];
            if (has[i]) L->mvpMapPoints[i] = &pool[i];
            else if (i % 2) { L->mvpMapPoints[i] = &pool[i]; L->mvbOutlier[i] = true; }      // outliers are skipped like missing points
        }
        const std::vector<MapPoint*> before = C->mvpMapPoints;
        const int n = m.SearchByProjection(*C, *L, th, mono != 0);
        put(out, &n, 1);
        std::vector<int> owner(C->N, -1);
        for (int j = 0; j < C->N; j++) {
            MapPoint* p = C->mvpMapPoints[j];
            if (p && nl && p >= &pool[0] && p <= &pool[nl - 1]) owner[j] = (int)(p - &pool[0]);
            else if (!p && (j % 3 == 0 || blocked[j])) owner[j] = -2;                        // a pre-existing point was set to NULL
        }
        put(out, owner.data(), owner.size());
        // the batched overload on three copies of the pair must leave every copy exactly like the single call left C
        Frame* Cs[3];
        for (int k = 0; k < 3; k++) { Cs[k] = new Frame(*C); Cs[k]->mvpMapPoints = before; }
        const std::vector<int> nb = m.SearchByProjection(std::vector<Frame*>(Cs, Cs + 3), std::vector<const Frame*>(3, L), th, mono != 0);
        int batch_ok = nb.size() == 3 ? 1 : 0;
        for (int k = 0; k < 3 && batch_ok; k++) batch_ok = (nb[k] == n && Cs[k]->mvpMapPoints == C->mvpMapPoints) ? 1 : 0;
        put(out, &batch_ok, 1);
        for (int k = 0; k < 3; k++) delete Cs[k];
        delete C; delete L;
    }
    {   // SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)
        Frame* F2 = new Frame(); Frame* F1 = new Frame();
        build_frame(r, *F2);
        const int n1 = r.get<int>(), window = r.get<int>();
        std::vector<unsigned char> d = r.arr<unsigned char>((size_t)n1 * 32);
        std::vector<int> oct = r.arr<int>(n1);
        std::vector<float> ang = r.arr<float>(n1), prev = r.arr<float>((size_t)n1 * 2);
        F1->N = n1; F1->mvKeysUn.resize(n1);
        F1->mDescriptors = cv::Mat(n1, 32, CV_8U);
        std::vector<cv::Point2f> vbPrev(n1);
        for (int i = 0; i < n1; i++) {
            memcpy(F1->mDescriptors.ptr(i), &d[(size_t)i * 32], 32);
            F1->mvKeysUn[i] = cv::KeyPoint(prev[2 * i], prev[2 * i + 1], 31, ang[i], 10, oct[i]);
            vbPrev[i] = cv::Point2f(prev[2 * i], prev[2 * i + 1]);
        }
        std::vector<int> m12(3, 7);
        const int n = m.SearchForInitialization(*F1, *F2, vbPrev, m12, window);
        put(out, &n, 1);
        const int sz = (int)m12.size(); put(out, &sz, 1);
        put(out, m12.data(), m12.size());
        for (int i = 0; i < n1; i++) { const float p[2] = {vbPrev[i].x, vbPrev[i].y}; put(out, p, 2); }
        delete F1; delete F2;
    }
    return ORBmatcher::LastStatus() == 0 ? 0 : 4;
}

// driver project2 <in.bin> <out.bin>: the relocalisation and loop-closing projection searches (Tracking::Relocalization,
// LoopClosing::ComputeSim3).  Point states: 0 = NULL entry, 1 = good, 2 = bad, 3 = already found.
static void fill_point(MapPoint& p, const float* world, float mfmax, float mfmin, const float* normal, const unsigned char* d) {
    p.mWorldPos = cv::Mat(3, 1, CV_32F); p.mNormalVector = cv::Mat(3, 1, CV_32F);
    for (int k = 0; k < 3; k++) { p.mWorldPos.at<float>(k) = world[k]; p.mNormalVector.at<float>(k) = normal[k]; }
    p.mfMaxDistance = mfmax; p.mfMinDistance = mfmin;
    p.mDescriptor = cv::Mat(1, 32, CV_8U); memcpy(p.mDescriptor.ptr(0), d, 32);
}
static int run_project2(int argc, char** argv) {
    if (argc < 4) return 2;
    Reader r; r.buf = slurp(argv[2]);
    std::ofstream out(argv[3], std::ios::binary);
    const int ori = r.get<int>();
    ORBmatcher m(0.9f, ori != 0);
    {   // SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist)
        Frame* C = new Frame();
        build_frame(r, *C);
        std::vector<unsigned char> blocked = r.arr<unsigned char>(C->N);
        std::vector<MapPoint> own(C->N);
        for (int i = 0; i < C->N; i++) C->mvpMapPoints[i] = blocked[i] ? &own[i] : NULL;
        C->mTcw = mat_from(r.arr<float>(16), 4, 4);
        std::vector<float> K = r.arr<float>(5);
        C->fx = K[0]; C->fy = K[1]; C->cx = K[2]; C->cy = K[3]; C->mfLogScaleFactor = K[4];
        const float th = r.get<float>();
        const int orbdist = r.get<int>(), n = r.get<int>();
        std::vector<unsigned char> state = r.arr<unsigned char>(n), d = r.arr<unsigned char>((size_t)n * 32);
        std::vector<float> world = r.arr<float>((size_t)n * 3), normal = r.arr<float>((size_t)n * 3), mfmax = r.arr<float>(n), mfmin = r.arr<float>(n), ang = r.arr<float>(n);
        KeyFrame kf; kf.N = n; kf.mvKeysUn.resize(n); kf.mvpMapPoints.assign(n, (MapPoint*)NULL);
        std::vector<MapPoint> pool(n);
        std::set<MapPoint*> found;
        for (int i = 0; i < n; i++) {
            kf.mvKeysUn[i] = cv::KeyPoint(0, 0, 31, ang[i], 10, 0);
            fill_point(pool[i], &world[3 * (size_t)i], mfmax[i], mfmin[i], &normal[3 * (size_t)i], &d[(size_t)i * 32]);
            pool[i].mbBad = state[i] == 2;
            if (state[i]) kf.mvpMapPoints[i] = &pool[i];
            if (state[i] == 3) found.insert(&pool[i]);
        }
        const int nm = m.SearchByProjection(*C, &kf, found, th, orbdist);
        put(out, &nm, 1);
        std::vector<int> owner(C->N, -1);
        for (int j = 0; j < C->N; j++) {
            MapPoint* p = C->mvpMapPoints[j];
            if (p && n && p >= &pool[0] && p <= &pool[n - 1]) owner[j] = (int)(p - &pool[0]);
        }
        put(out, owner.data(), owner.size());
        delete C;
    }
    {   // SearchByProjection(pKF, Scw, vpPoints, vpMatched, th)
        Frame* F = new Frame();
        build_frame(r, *F);
        KeyFrame kf;
        kf.N = F->N; kf.mDescriptors = F->mDescriptors; kf.mvKeysUn = F->mvKeysUn; kf.mvScaleFactors = F->mvScaleFactors;
        kf.mnMinX = (int)F->mnMinX; kf.mnMinY = (int)F->mnMinY; kf.mnMaxX = (int)F->mnMaxX; kf.mnMaxY = (int)F->mnMaxY;
        kf.mfGridElementWidthInv = F->mfGridElementWidthInv; kf.mfGridElementHeightInv = F->mfGridElementHeightInv;
        std::vector<unsigned char> blocked = r.arr<unsigned char>(kf.N);
        cv::Mat Scw = mat_from(r.arr<float>(16), 4, 4);
        std::vector<float> K = r.arr<float>(5);
        kf.fx = K[0]; kf.fy = K[1]; kf.cx = K[2]; kf.cy = K[3]; kf.mfLogScaleFactor = K[4];
        const int th = r.get<int>(), n = r.get<int>();
        std::vector<unsigned char> state = r.arr<unsigned char>(n), d = r.arr<unsigned char>((size_t)n * 32);
        std::vector<float> world = r.arr<float>((size_t)n * 3), normal = r.arr<float>((size_t)n * 3), mfmax = r.arr<float>(n), mfmin = r.arr<float>(n);
        std::vector<MapPoint> pool(n), dummy(kf.N);
        std::vector<MapPoint*> vpPoints(n), vpMatched(kf.N, (MapPoint*)NULL);
        std::vector<int> foundIdx;
        for (int i = 0; i < n; i++) {
            fill_point(pool[i], &world[3 * (size_t)i], mfmax[i], mfmin[i], &normal[3 * (size_t)i], &d[(size_t)i * 32]);
            pool[i].mbBad = state[i] == 2 || state[i] == 0;             // vpPoints holds no NULLs in the reference; 0 -> bad here
            vpPoints[i] = &pool[i];
            if (state[i] == 3) foundIdx.push_back(i);
        }
        size_t k = 0;
        for (int j = 0; j < kf.N; j++)
            if (blocked[j]) { vpMatched[j] = k < foundIdx.size() ? &pool[foundIdx[k]] : &dummy[j]; k++; }
        if (k < foundIdx.size()) return 6;                               // the generator guarantees enough blocked features
        const int nm = m.SearchByProjection(&kf, Scw, vpPoints, vpMatched, th);
        put(out, &nm, 1);
        std::vector<int> owner(kf.N, -1);
        for (int j = 0; j < kf.N; j++) {
            MapPoint* p = vpMatched[j];
            if (!blocked[j] && p && n && p >= &pool[0] && p <= &pool[n - 1]) owner[j] = (int)(p - &pool[0]);
        }
        put(out, owner.data(), owner.size());
        delete F;
    }
    return ORBmatcher::LastStatus() == 0 ? 0 : 4;
}

// driver stereo <left.raw> <right.raw> <w> <h> <nfeat> <nlevels> <mbf> <mb> <out.bin>: the stereo Frame constructor's path
// (two extractors, then ComputeStereoMatches) without ever downloading a pyramid
static int run_stereo(int argc, char** argv) {
    if (argc < 11) return 2;
    const int w = atoi(argv[4]), h = atoi(argv[5]);
    std::vector<unsigned char> l = slurp(argv[2]), r = slurp(argv[3]);
    if ((int)l.size() != w * h || (int)r.size() != w * h) return 3;
    ORBextractor exL(atoi(argv[6]), 1.2f, atoi(argv[7]), 20, 7), exR(atoi(argv[6]), 1.2f, atoi(argv[7]), 20, 7);
    exL.SetPyramidDownload(false); exR.SetPyramidDownload(false);
    cv::Mat imL(h, w, CV_8U, l.data()), imR(h, w, CV_8U, r.data()), nomask;
    std::vector<cv::KeyPoint> kL, kR;
    cv::Mat dL, dR;
    exL(imL, nomask, kL, dL);
    exR(imR, nomask, kR, dR);
    std::vector<float> uR(2, 5.f), depth;
    ORBextractor::ComputeStereoMatches(&exL, &exR, kL, dL, kR, dR, (float)atof(argv[8]), (float)atof(argv[9]), uR, depth);
    if (exL.LastStatus() != 0) return 4;
    std::ofstream out(argv[10], std::ios::binary);
    const int n = (int)uR.size();
    put(out, &n, 1); put(out, uR.data(), uR.size()); put(out, depth.data(), depth.size());
    return 0;
}

// driver fuse <in.bin> <out.bin>: Fuse (LocalMapping::SearchInNeighbors), Fuse with a similarity (LoopClosing::SearchAndFuse) and
// SearchBySim3 (LoopClosing::ComputeSim3) on stand-in KeyFrame / MapPoint objects; the final object graph is written back.
static void build_keyframe(Reader& r, KeyFrame& kf) {
    Frame* F = new Frame();
    build_frame(r, *F);
    kf.N = F->N; kf.mDescriptors = F->mDescriptors; kf.mvKeysUn = F->mvKeysUn; kf.mvKeys = F->mvKeys; kf.mvuRight = F->mvuRight;
    kf.mvScaleFactors = F->mvScaleFactors;
    kf.mnMinX = (int)F->mnMinX; kf.mnMinY = (int)F->mnMinY; kf.mnMaxX = (int)F->mnMaxX; kf.mnMaxY = (int)F->mnMaxY;
    kf.mfGridElementWidthInv = F->mfGridElementWidthInv; kf.mfGridElementHeightInv = F->mfGridElementHeightInv;
    kf.mvpMapPoints.assign(kf.N, (MapPoint*)NULL);
    std::vector<float> K = r.arr<float>(6);
    kf.fx = K[0]; kf.fy = K[1]; kf.cx = K[2]; kf.cy = K[3]; kf.mbf = K[4]; kf.mfLogScaleFactor = K[5];
    kf.mvInvLevelSigma2 = r.arr<float>(kf.mvScaleFactors.size());
    std::vector<float> T = r.arr<float>(12), Ow = r.arr<float>(3);
    kf.Rcw = cv::Mat(3, 3, CV_32F); kf.tcw = cv::Mat(3, 1, CV_32F); kf.Ow = cv::Mat(3, 1, CV_32F);
    for (int i = 0; i < 3; i++) { kf.tcw.at<float>(i) = T[4 * i + 3]; kf.Ow.at<float>(i) = Ow[i]; for (int j = 0; j < 3; j++) kf.Rcw.at<float>(i, j) = T[4 * i + j]; }
    delete F;
}
struct PointSet {
    std::vector<MapPoint> pool;
    std::vector<unsigned char> state;
    void load(Reader& r) {
        const int n = r.get<int>();
        state = r.arr<unsigned char>(n);
        std::vector<int> nobs = r.arr<int>(n);
        std::vector<unsigned char> d = r.arr<unsigned char>((size_t)n * 32);
        std::vector<float> world = r.arr<float>((size_t)n * 3), normal = r.arr<float>((size_t)n * 3), mfmax = r.arr<float>(n), mfmin = r.arr<float>(n);
        pool.resize(n);
        for (int i = 0; i < n; i++) {
            fill_point(pool[i], &world[3 * (size_t)i], mfmax[i], mfmin[i], &normal[3 * (size_t)i], &d[(size_t)i * 32]);
            pool[i].mbBad = state[i] == 2; pool[i].nObs = nobs[i];
        }
    }
};
static int encode(MapPoint* p, PointSet& kfp, PointSet& cand) {
    if (!p) return -1;
    if (!kfp.pool.empty() && p >= &kfp.pool[0] && p <= &kfp.pool.back()) return (int)(p - &kfp.pool[0]);
    if (!cand.pool.empty() && p >= &cand.pool[0] && p <= &cand.pool.back()) return 100000 + (int)(p - &cand.pool[0]);
    return -2;
}
static void attach_kf_points(KeyFrame& kf, PointSet& kfp) {           // state 1/2: the KeyFrame observes the point at that feature
    for (int j = 0; j < kf.N; j++)
        if (kfp.state[j]) { kf.mvpMapPoints[j] = &kfp.pool[j]; kfp.pool[j].mObservations[&kf] = j; }
}
static void dump_state(std::ofstream& out, KeyFrame& kf, PointSet& kfp, PointSet& cand) {
    std::vector<int> ptr(kf.N);
    for (int j = 0; j < kf.N; j++) ptr[j] = encode(kf.mvpMapPoints[j], kfp, cand);
    put(out, ptr.data(), ptr.size());
    for (PointSet* ps : {&kfp, &cand})
        for (size_t i = 0; i < ps->pool.size(); i++) { const int v[2] = {ps->pool[i].mbBad ? 1 : 0, ps->pool[i].nObs}; put(out, v, 2); }
}
static int run_fuse(int argc, char** argv) {
    if (argc < 4) return 2;
    Reader r; r.buf = slurp(argv[2]);
    std::ofstream out(argv[3], std::ios::binary);
    ORBmatcher m(0.8f, true);
    for (int variant = 0; variant < 2; variant++) {
        KeyFrame kf;
        build_keyframe(r, kf);
        PointSet kfp, cand;
        kfp.load(r); cand.load(r);
        attach_kf_points(kf, kfp);
        const float th = r.get<float>();
        std::vector<int> inKfAt = r.arr<int>(cand.pool.size());       // state 3: the candidate already sits in the KeyFrame at this feature
        std::vector<MapPoint*> vp(cand.pool.size());
        for (size_t i = 0; i < cand.pool.size(); i++) {
            vp[i] = (variant == 0 && cand.state[i] == 0) ? NULL : &cand.pool[i];
            if (cand.state[i] == 3) { cand.pool[i].mObservations[&kf] = inKfAt[i]; kf.mvpMapPoints[inKfAt[i]] = &cand.pool[i]; }
        }
        int n;
        if (variant == 0) n = m.Fuse(&kf, vp, th);
        else {
            cv::Mat Scw = mat_from(r.arr<float>(16), 4, 4);
            std::vector<MapPoint*> rep(vp.size(), (MapPoint*)NULL);
            n = m.Fuse(&kf, Scw, vp, th, rep);
            std::vector<int> e(rep.size());
            for (size_t i = 0; i < rep.size(); i++) e[i] = encode(rep[i], kfp, cand);
            put(out, e.data(), e.size());
        }
        put(out, &n, 1);
        dump_state(out, kf, kfp, cand);
    }
    {   // SearchBySim3
        KeyFrame kf1, kf2;
        build_keyframe(r, kf1); build_keyframe(r, kf2);
        PointSet p1, p2;
        p1.load(r); p2.load(r);
        attach_kf_points(kf1, p1); attach_kf_points(kf2, p2);
        const float s12 = r.get<float>(), th = r.get<float>();
        cv::Mat R12 = mat_from(r.arr<float>(9), 3, 3), t12 = mat_from(r.arr<float>(3), 3, 1);
        std::vector<int> pre = r.arr<int>(kf1.N);                      // vpMatches12 on entry: index of a KF2 feature's point, or -1
        std::vector<MapPoint*> vpMatches12(kf1.N, (MapPoint*)NULL);
        for (int i = 0; i < kf1.N; i++) if (pre[i] >= 0) vpMatches12[i] = &p2.pool[pre[i]];
        const int n = m.SearchBySim3(&kf1, &kf2, vpMatches12, s12, R12, t12, th);
        put(out, &n, 1);
        std::vector<int> e(kf1.N);
        for (int i = 0; i < kf1.N; i++) e[i] = vpMatches12[i] ? (int)(vpMatches12[i] - &p2.pool[0]) : -1;
        put(out, e.data(), e.size());
    }
    return ORBmatcher::LastStatus() == 0 ? 0 : 4;
}

// maparchive <in.bin> <out.bin> <dump.txt>: loads a saved map through ORB_SLAM2::MapArchiveB200 (host-only, no GPU needed), writes
// it back, and dumps what the class hands out as text for the Python test to compare with the logical map the file was made from.
static int run_maparchive(int argc, char** argv) {
    if (argc < 5) return 2;
    ORB_SLAM2::MapArchiveB200 ar;
    if (!ar.Load(argv[2])) { fprintf(stderr, "load failed: %s\n", ar.LastError().c_str()); return 3; }
    if (!ar.Save(argv[3])) { fprintf(stderr, "save failed: %s\n", ar.LastError().c_str()); return 4; }
    FILE* f = fopen(argv[4], "w");
    if (!f) return 5;
    fprintf(f, "map %lu %lu %lu %d\n", ar.KeyFramesInMap(), ar.MapPointsInMap(), ar.GetMaxKFid(), (int)ar.LoadValidated());
    for (size_t i = 0; i < ar.KeyFramesInMap(); i++) {
        ORB_SLAM2::MapArchiveB200::KeyFrameData k;
        if (!ar.GetKeyFrame(i, k)) { fclose(f); return 6; }
        unsigned long long dsum = 0;
        for (int r = 0; r < k.mDescriptors.rows; r++)
            for (int c = 0; c < k.mDescriptors.cols; c++) dsum = dsum * 1000003ull + k.mDescriptors.ptr(r)[c];
        long mpsum = 0;
        for (long v : k.mvpMapPointIds) mpsum = mpsum * 31 + v;
        double kx = 0;
        for (const cv::KeyPoint& p : k.mvKeysUn) kx += (double)p.pt.x + 2.0 * (double)p.pt.y + (double)p.angle + p.octave;
        fprintf(f, "kf %lu %lu %d %d %llu %ld %.6f %.9g %.9g %d %zu %zu %zu %zu %zu\n", k.mnId, k.mnFrameId, k.N, k.mDescriptors.rows, dsum, mpsum, kx,
                (double)k.Tcw.at<float>(1, 3), (double)k.mK.at<float>(0, 2), (int)k.hasParent, k.mConnectedKeyFrameWeights.size(),
                k.mvpOrderedConnectedKeyFrames.size(), k.mspChildrens.size(), k.mspLoopEdges.size(), k.gridFeatures.size());
    }
    std::vector<ORB_SLAM2::MapArchiveB200::MapPointData> mps = ar.GetAllMapPoints();
    for (const auto& p : mps) {
        unsigned long long dsum = 0;
        for (int c = 0; c < 32; c++) dsum = dsum * 1000003ull + p.mDescriptor.ptr()[c];
        long osum = 0;
        for (const auto& o : p.mObservations) osum = osum * 31 + o.first * 7 + o.second;
        fprintf(f, "mp %lu %llu %ld %zu %ld %.9g\n", p.mnId, dsum, p.refKFId, p.mObservations.size(), osum, (double)p.mWorldPos.at<float>(2));
    }
    cv::Mat obs;
    std::vector<int> off;
    if (!ar.ObservedDescriptors(obs, off)) { fclose(f); return 7; }
    unsigned long long osum = 0;
    for (int r = 0; r < obs.rows; r++)
        for (int c = 0; c < 32; c++) osum = osum * 1000003ull + obs.ptr(r)[c];
    fprintf(f, "observed %d %llu %d\n", obs.rows, osum, off.empty() ? 0 : off.back());
    fclose(f);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    if (!strcmp(argv[1], "maparchive")) return run_maparchive(argc, argv);
    if (!strcmp(argv[1], "bow")) return run_bow(argc, argv);
    if (!strcmp(argv[1], "extract")) return run_extract(argc, argv);
    if (!strcmp(argv[1], "match")) return run_match(argc, argv);
    if (!strcmp(argv[1], "project")) return run_project(argc, argv);
    if (!strcmp(argv[1], "project2")) return run_project2(argc, argv);
    if (!strcmp(argv[1], "stereo")) return run_stereo(argc, argv);
    if (!strcmp(argv[1], "fuse")) return run_fuse(argc, argv);
    return 2;
}
