// host_api_driver.cc — exercises the C++ drop-in classes (ORB_SLAM2::ORBextractor / ORBmatcher) the way the reference's
// Frame / LocalMapping code does, on inputs written by tests/test_gpu_host_cpp.py; writes results back as raw binaries.
//   driver extract <in.raw> <w> <h> <nfeat> <scale> <nlevels> <ini> <min> <out_prefix> [mask.raw]
//   driver match   <in.bin> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "ORBextractor.h"
#include "ORBmatcher.h"
#include "ORBVocabulary.h"

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

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    if (!strcmp(argv[1], "bow")) return run_bow(argc, argv);
    if (!strcmp(argv[1], "extract")) return run_extract(argc, argv);
    if (!strcmp(argv[1], "match")) return run_match(argc, argv);
    return 2;
}
