// orb_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A plain, serial, OpenCV-free restatement of the reference hot path
//   /root/reference/src/ORBextractor.cc   (ORBextractor ctor, operator(), ComputePyramid,
//                                           ComputeKeyPointsOctTree, DistributeOctTree, IC_Angle,
//                                           computeOrbDescriptor)
//   /root/reference/src/ORBmatcher.cc     (DescriptorDistance, SearchByBoW x2, SearchForTriangulation,
//                                           CheckDistEpipolarLine, ComputeThreeMaxima)
// plus the four OpenCV primitives that path delegates to (cv::FAST, cv::resize INTER_LINEAR,
// cv::GaussianBlur 7x7 sigma 2, cv::fastAtan2), restated from OpenCV 4.13 behaviour.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this
// library, and only as the checker / reported CPU baseline.  The product (orbslam_mapsave_b200/csrc) never
// links, loads or calls it.
//
// PARITY PINNING.  Extractor: oracle/_ref is the reference's own src/ORBextractor.cc compiled unmodified over ref_shim/cvshim.hpp
// (OpenCV API stand-in whose image primitives are the functions of this file); run on a monotonic heap it equals this oracle's
// output bit for bit on all golden configurations (tests/test_reference_ref.py).  Matcher: oracle/_ref/libref_orbmatcher.so is the
// reference's own src/ORBmatcher.cc compiled unmodified over plain-data stand-ins for KeyFrame / Frame / MapPoint; its
// DescriptorDistance, SearchByBoW x2, SearchForTriangulation, ComputeThreeMaxima, SearchByProjection(Frame, MapPoints),
// SearchByProjection(Frame, LastFrame), SearchForInitialization and the relocalisation / loop-closing projections equal this oracle on
// the scenes of the GPU parity tests.  The Fuse x2, SearchBySim3 and ComputeStereoMatches restatements remain "parity unpinned" by
// the reference itself (the reference ships no tests or golden vectors).  Also pinned: every OpenCV primitive below bit-for-bit
// against cv2 4.13.0 golden vectors (tests/golden/*.npz, made by tests/golden/make_golden*.py), and the full extractor against an
// independent Python chain of the real cv2 primitives (same script).
//
// Known, documented deviations from "whatever binary the reference authors ran":
//   * DistributeOctTree sorts (size, node pointer) pairs (ORBextractor.cc:683); pointer order is allocator
//     dependent.  Canonical rule here: the pointer is replaced by the node's creation sequence number.
//   * GaussianBlur follows OpenCV 4.13 (fixed-point 8.8 kernel), the only OpenCV available.
//   * No FMA contraction anywhere (build with -ffp-contract=off).
//
// Build: see oracle/Makefile  (g++ -O2 -ffp-contract=off -shared -fPIC).

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cfloat>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <thread>
#include <utility>
#include <vector>

typedef unsigned char u8;

extern "C" {
struct orc_kp { float x, y, size, angle, response; int octave, class_id; };  // == cv::KeyPoint layout (28 B)
struct orc_xyr { int x, y, r; };
}

static const signed char kPattern[1024] = {
#include "../orbslam_mapsave_b200/csrc/orb_pattern_31.inc"
};

// ----------------------------------------------------------------------------------------------
// Rounding helpers (OpenCV cvRound = round-half-to-even via cvtsd2si; cvFloor/cvCeil as named)
// ----------------------------------------------------------------------------------------------
static inline int cvRoundD(double v) { return (int)std::nearbyint(v); }   // default FE_TONEAREST
static inline int cvRoundF(float v) { return (int)std::nearbyintf(v); }
static inline int cvFloorD(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeilD(double v) { int i = (int)v; return i + (i < v); }

// ----------------------------------------------------------------------------------------------
// cv::FAST(img, kps, threshold, nonmaxSuppression) — FAST-9/16.   (call sites ORBextractor.cc:808,813)
// ----------------------------------------------------------------------------------------------
static const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// Is (x,y) a FAST-9 corner at threshold t: >=9 contiguous ring pixels all < v-t or all > v+t.
static bool fast_is_corner(const u8* p, int stride, int t) {
    const int v = p[0];
    {   // cheap necessary condition first (same idea as OpenCV's own pre-test; does not change the result): any 9 contiguous
        // ring pixels contain at least one of each opposite pair, so both (0,8) and (4,12) must have a darker or a brighter member
        const int lo = v - t, hi = v + t;
        const int a = p[3 * stride], b = p[-3 * stride], c = p[3], d = p[-3];
        const bool dark = (a < lo || b < lo) && (c < lo || d < lo);
        const bool bright = (a > hi || b > hi) && (c > hi || d > hi);
        if (!dark && !bright) return false;
    }
    int ring[25];
    for (int k = 0; k < 16; k++) ring[k] = p[kRingDy[k] * stride + kRingDx[k]];
    for (int k = 16; k < 25; k++) ring[k] = ring[k - 16];
    int cd = 0, cb = 0;
    for (int k = 0; k < 25; k++) {
        if (ring[k] < v - t) { if (++cd > 8) return true; } else cd = 0;
        if (ring[k] > v + t) { if (++cb > 8) return true; } else cb = 0;
    }
    return false;
}

// OpenCV cornerScore<16>(ptr, pixel, threshold): largest threshold for which the pixel stays a corner.
static int fast_corner_score(const u8* p, int stride, int threshold) {
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = v - p[kRingDy[k] * stride + kRingDx[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int a0 = threshold;
    for (int k = 0; k < 16; k += 2) {
        int a = std::min(d[k + 1], d[k + 2]);
        a = std::min(a, d[k + 3]);
        if (a <= a0) continue;
        a = std::min(a, d[k + 4]); a = std::min(a, d[k + 5]); a = std::min(a, d[k + 6]);
        a = std::min(a, d[k + 7]); a = std::min(a, d[k + 8]);
        a0 = std::max(a0, std::min(a, d[k]));
        a0 = std::max(a0, std::min(a, d[k + 9]));
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = std::max(d[k + 1], d[k + 2]);
        b = std::max(b, d[k + 3]); b = std::max(b, d[k + 4]); b = std::max(b, d[k + 5]);
        if (b >= b0) continue;
        b = std::max(b, d[k + 6]); b = std::max(b, d[k + 7]); b = std::max(b, d[k + 8]);
        b0 = std::min(b0, std::max(b, d[k]));
        b0 = std::min(b0, std::max(b, d[k + 9]));
    }
    return -b0 - 1;
}

static void fast9(const u8* img, int w, int h, int stride, int threshold, bool nms, std::vector<orc_xyr>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    std::vector<u8> score((size_t)w * h, 0);
    std::vector<u8> corner((size_t)w * h, 0);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            const u8* p = img + (size_t)y * stride + x;
            if (fast_is_corner(p, stride, threshold)) {
                corner[(size_t)y * w + x] = 1;
                score[(size_t)y * w + x] = (u8)fast_corner_score(p, stride, threshold);
            }
        }
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            if (!corner[(size_t)y * w + x]) continue;
            const int s = score[(size_t)y * w + x];
            if (nms) {
                bool keep = true;
                for (int dy = -1; dy <= 1 && keep; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        if (!dx && !dy) continue;
                        if (!(s > score[(size_t)(y + dy) * w + (x + dx)])) { keep = false; break; }
                    }
                if (!keep) continue;
            }
            out.push_back({x, y, s});
        }
}

// ----------------------------------------------------------------------------------------------
// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for 8UC1   (call site ORBextractor.cc:1123)
// ----------------------------------------------------------------------------------------------
static void resize_linear_u8(const u8* src, int sw, int sh, int sstride, u8* dst, int dw, int dh, int dstride) {
    const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> ialpha(2 * (size_t)dw), ibeta(2 * (size_t)dh);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cvFloorD(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ialpha[2 * dx] = (short)cvRoundF((1.f - fx) * 2048.f);
        ialpha[2 * dx + 1] = (short)cvRoundF(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cvFloorD(fy);
        fy -= sy;
        yofs[dy] = sy;
        ibeta[2 * dy] = (short)cvRoundF((1.f - fy) * 2048.f);
        ibeta[2 * dy + 1] = (short)cvRoundF(fy * 2048.f);
    }
    std::vector<int> r0(dw), r1(dw);
    for (int dy = 0; dy < dh; dy++) {
        const int sy0 = std::min(std::max(yofs[dy], 0), sh - 1);
        const int sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
        const u8* S0 = src + (size_t)sy0 * sstride;
        const u8* S1 = src + (size_t)sy1 * sstride;
        for (int dx = 0; dx < dw; dx++) {
            const int sx = xofs[dx], sx1 = std::min(sx + 1, sw - 1);
            const int a0 = ialpha[2 * dx], a1 = ialpha[2 * dx + 1];
            r0[dx] = S0[sx] * a0 + S0[sx1] * a1;
            r1[dx] = S1[sx] * a0 + S1[sx1] * a1;
        }
        const int b0 = ibeta[2 * dy], b1 = ibeta[2 * dy + 1];
        u8* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++) {
            int v = (((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (u8)std::min(std::max(v, 0), 255);
        }
    }
}

static inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * (n - 1) - p; }
    return p;
}

// cv::copyMakeBorder(src, dst, b,b,b,b, BORDER_REFLECT_101) into a (w+2b)x(h+2b) buffer.
static void copy_make_border101(const u8* src, int w, int h, int sstride, u8* dst, int dstride, int b) {
    for (int y = -b; y < h + b; y++) {
        const u8* S = src + (size_t)reflect101(y, h) * sstride;
        u8* D = dst + (size_t)(y + b) * dstride;
        for (int x = -b; x < w + b; x++) D[x + b] = S[reflect101(x, w)];
    }
}

// ----------------------------------------------------------------------------------------------
// cv::GaussianBlur(img, img, Size(7,7), 2, 2, BORDER_REFLECT_101) for 8UC1, OpenCV 4.13 fixed-point path
// (call site ORBextractor.cc:1089)
// ----------------------------------------------------------------------------------------------
static void gaussian_blur7(const u8* src, int w, int h, int sstride, u8* dst, int dstride) {
    static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
    std::vector<int> H((size_t)w * h);
    std::vector<int> pad(w + 6);
    for (int y = 0; y < h; y++) {
        const u8* S = src + (size_t)y * sstride;
        for (int x = -3; x < w + 3; x++) pad[x + 3] = S[reflect101(x, w)];       // REFLECT_101 once per row, not per tap
        int* Hr = &H[(size_t)y * w];
        for (int x = 0; x < w; x++) {
            const int* p = &pad[x];
            Hr[x] = K[0] * p[0] + K[1] * p[1] + K[2] * p[2] + K[3] * p[3] + K[4] * p[4] + K[5] * p[5] + K[6] * p[6];
        }
    }
    for (int y = 0; y < h; y++) {
        const int* r[7];
        for (int i = -3; i <= 3; i++) r[i + 3] = &H[(size_t)reflect101(y + i, h) * w];
        u8* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            const int s = K[0] * r[0][x] + K[1] * r[1][x] + K[2] * r[2][x] + K[3] * r[3][x] + K[4] * r[4][x] + K[5] * r[5][x] + K[6] * r[6][x];
            D[x] = (u8)((s + 32768) >> 16);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// cv::fastAtan2(y, x) scalar path (degrees)   (call site ORBextractor.cc:102)
// ----------------------------------------------------------------------------------------------
static float fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ----------------------------------------------------------------------------------------------
// ORBextractor restatement
// ----------------------------------------------------------------------------------------------
namespace {

const int PATCH_SIZE = 31, HALF_PATCH_SIZE = 15, EDGE_THRESHOLD = 19;   // ORBextractor.cc:71-73

struct Img {           // a level: ROI (w x h) inside a bordered buffer
    int w = 0, h = 0, stride = 0;
    std::vector<u8> buf;                 // (w+38) x (h+38)
    u8* roi() { return buf.data() + (size_t)EDGE_THRESHOLD * stride + EDGE_THRESHOLD; }
    const u8* roi() const { return buf.data() + (size_t)EDGE_THRESHOLD * stride + EDGE_THRESHOLD; }
};

struct Key { float x, y, response; };    // the fields of cv::KeyPoint the octree touches

struct Node {                            // ExtractorNode, include/ORBextractor.h:37-48
    std::vector<Key> keys;
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    std::list<Node>::iterator lit;
    bool noMore = false;
    long seq = 0;                        // creation sequence: canonical stand-in for the heap pointer
};

// ExtractorNode::DivideNode, ORBextractor.cc:480-536
void divide_node(const Node& p, Node& n1, Node& n2, Node& n3, Node& n4) {
    const int halfX = (int)std::ceil(static_cast<float>(p.URx - p.ULx) / 2);
    const int halfY = (int)std::ceil(static_cast<float>(p.BRy - p.ULy) / 2);
    n1.ULx = p.ULx; n1.ULy = p.ULy;
    n1.URx = p.ULx + halfX; n1.URy = p.ULy;
    n1.BLx = p.ULx; n1.BLy = p.ULy + halfY;
    n1.BRx = p.ULx + halfX; n1.BRy = p.ULy + halfY;
    n2.ULx = n1.URx; n2.ULy = n1.URy;
    n2.URx = p.URx; n2.URy = p.URy;
    n2.BLx = n1.BRx; n2.BLy = n1.BRy;
    n2.BRx = p.URx; n2.BRy = p.ULy + halfY;
    n3.ULx = n1.BLx; n3.ULy = n1.BLy;
    n3.URx = n1.BRx; n3.URy = n1.BRy;
    n3.BLx = p.BLx; n3.BLy = p.BLy;
    n3.BRx = n1.BRx; n3.BRy = p.BLy;
    n4.ULx = n3.URx; n4.ULy = n3.URy;
    n4.URx = n2.BRx; n4.URy = n2.BRy;
    n4.BLx = n3.BRx; n4.BLy = n3.BRy;
    n4.BRx = p.BRx; n4.BRy = p.BRy;
    for (const Key& kp : p.keys) {
        if (kp.x < n1.URx) {
            if (kp.y < n1.BRy) n1.keys.push_back(kp); else n3.keys.push_back(kp);
        } else if (kp.y < n1.BRy) n2.keys.push_back(kp);
        else n4.keys.push_back(kp);
    }
    if (n1.keys.size() == 1) n1.noMore = true;
    if (n2.keys.size() == 1) n2.noMore = true;
    if (n3.keys.size() == 1) n3.noMore = true;
    if (n4.keys.size() == 1) n4.noMore = true;
}

// ORBextractor::DistributeOctTree, ORBextractor.cc:538-762
std::vector<Key> distribute_octtree(const std::vector<Key>& in, int minX, int maxX, int minY, int maxY, int N) {
    long seq = 0;
    const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
    const float hX = static_cast<float>(maxX - minX) / nIni;
    std::list<Node> L;
    std::vector<Node*> ini(nIni);
    for (int i = 0; i < nIni; i++) {
        Node ni;
        ni.ULx = (int)(hX * static_cast<float>(i)); ni.ULy = 0;
        ni.URx = (int)(hX * static_cast<float>(i + 1)); ni.URy = 0;
        ni.BLx = ni.ULx; ni.BLy = maxY - minY;
        ni.BRx = ni.URx; ni.BRy = maxY - minY;
        ni.seq = seq++;
        L.push_back(ni);
        ini[i] = &L.back();
    }
    for (const Key& kp : in) ini[(int)(kp.x / hX)]->keys.push_back(kp);
    for (auto lit = L.begin(); lit != L.end();) {
        if (lit->keys.size() == 1) { lit->noMore = true; ++lit; }
        else if (lit->keys.empty()) lit = L.erase(lit);
        else ++lit;
    }
    bool finish = false;
    typedef std::pair<int, std::pair<long, Node*> > SizeSeqNode;     // (size, (seq, node)): seq replaces the pointer
    std::vector<SizeSeqNode> vSize;
    auto push_child = [&](Node& c, int& nToExpand) {
        if (c.keys.empty()) return;
        c.seq = seq++;
        L.push_front(c);
        if (c.keys.size() > 1) {
            nToExpand++;
            vSize.push_back(std::make_pair((int)c.keys.size(), std::make_pair(L.front().seq, &L.front())));
            L.front().lit = L.begin();
        }
    };
    while (!finish) {
        const int prevSize = (int)L.size();
        auto lit = L.begin();
        int nToExpand = 0;
        vSize.clear();
        while (lit != L.end()) {
            if (lit->noMore) { ++lit; continue; }
            Node n1, n2, n3, n4;
            divide_node(*lit, n1, n2, n3, n4);
            push_child(n1, nToExpand); push_child(n2, nToExpand);
            push_child(n3, nToExpand); push_child(n4, nToExpand);
            lit = L.erase(lit);
        }
        if ((int)L.size() >= N || (int)L.size() == prevSize) {
            finish = true;
        } else if (((int)L.size() + nToExpand * 3) > N) {
            while (!finish) {
                const int prev2 = (int)L.size();
                std::vector<SizeSeqNode> vPrev = vSize;
                vSize.clear();
                std::sort(vPrev.begin(), vPrev.end());
                for (int j = (int)vPrev.size() - 1; j >= 0; j--) {
                    Node n1, n2, n3, n4;
                    Node* pn = vPrev[j].second.second;
                    divide_node(*pn, n1, n2, n3, n4);
                    int dummy = 0;
                    push_child(n1, dummy); push_child(n2, dummy);
                    push_child(n3, dummy); push_child(n4, dummy);
                    L.erase(pn->lit);
                    if ((int)L.size() >= N) break;
                }
                if ((int)L.size() >= N || (int)L.size() == prev2) finish = true;
            }
        }
    }
    std::vector<Key> res;
    res.reserve(L.size());
    for (auto& nd : L) {
        const Key* best = &nd.keys[0];
        float maxR = best->response;
        for (size_t k = 1; k < nd.keys.size(); k++)
            if (nd.keys[k].response > maxR) { best = &nd.keys[k]; maxR = nd.keys[k].response; }
        res.push_back(*best);
    }
    return res;
}

struct Extractor {
    int nfeatures; double scaleFactor; int nlevels, iniThFAST, minThFAST;       // include/ORBextractor.h:102-106
    std::vector<float> sf, isf, s2, is2;
    std::vector<int> quota, umax;
    std::vector<Img> pyr;
    std::vector<std::vector<Key> > cand;       // per level vToDistributeKeys (coords relative to minBorder)
    std::vector<std::vector<orc_kp> > lvlkp;   // per level keypoints after octree + orientation (level coords)
    std::vector<std::vector<u8> > blurred;     // per level blurred clone (w x h, contiguous)

    // ORBextractor::ORBextractor, ORBextractor.cc:409-469
    Extractor(int nf, float sfac, int nl, int ini, int mn)
        : nfeatures(nf), scaleFactor(sfac), nlevels(nl), iniThFAST(ini), minThFAST(mn) {
        sf.resize(nl); s2.resize(nl); isf.resize(nl); is2.resize(nl);
        sf[0] = 1.0f; s2[0] = 1.0f;
        for (int i = 1; i < nl; i++) {
            sf[i] = (float)(sf[i - 1] * scaleFactor);      // float * double -> double -> float
            s2[i] = sf[i] * sf[i];
        }
        for (int i = 0; i < nl; i++) { isf[i] = 1.0f / sf[i]; is2[i] = 1.0f / s2[i]; }
        pyr.resize(nl);
        quota.resize(nl);
        float factor = (float)(1.0f / scaleFactor);
        float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
        int sum = 0;
        for (int l = 0; l < nl - 1; l++) {
            quota[l] = cvRoundF(nDesired);
            sum += quota[l];
            nDesired *= factor;
        }
        quota[nl - 1] = std::max(nfeatures - sum, 0);
        umax.resize(HALF_PATCH_SIZE + 1);
        int v, v0, vmax = cvFloorD(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
        int vmin = cvCeilD(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
        const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
        for (v = 0; v <= vmax; ++v) umax[v] = cvRoundD(std::sqrt(hp2 - v * v));
        for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }

    // ORBextractor::ComputePyramid, ORBextractor.cc:1110-1135
    void compute_pyramid(const u8* image, int cols, int rows, int stride) {
        for (int level = 0; level < nlevels; ++level) {
            const float scale = isf[level];
            const int w = cvRoundF((float)cols * scale), h = cvRoundF((float)rows * scale);
            Img& im = pyr[level];
            im.w = w; im.h = h; im.stride = w + 2 * EDGE_THRESHOLD;
            im.buf.assign((size_t)im.stride * (h + 2 * EDGE_THRESHOLD), 0);
            if (level != 0) {
                const Img& pv = pyr[level - 1];
                std::vector<u8> tmp((size_t)w * h);
                resize_linear_u8(pv.roi(), pv.w, pv.h, pv.stride, tmp.data(), w, h, w);
                copy_make_border101(tmp.data(), w, h, w, im.buf.data(), im.stride, EDGE_THRESHOLD);
            } else {
                copy_make_border101(image, cols, rows, stride, im.buf.data(), im.stride, EDGE_THRESHOLD);
            }
        }
    }

    // IC_Angle, ORBextractor.cc:76-103
    float ic_angle(const Img& im, float px, float py) const {
        int m_01 = 0, m_10 = 0;
        const int step = im.stride;
        const u8* center = im.roi() + (size_t)cvRoundF(py) * step + cvRoundF(px);
        for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
        for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
            int v_sum = 0;
            const int d = umax[v];
            for (int u = -d; u <= d; ++u) {
                const int val_plus = center[u + v * step], val_minus = center[u - v * step];
                v_sum += (val_plus - val_minus);
                m_10 += u * (val_plus + val_minus);
            }
            m_01 += v * v_sum;
        }
        return fast_atan2((float)m_01, (float)m_10);
    }

    // ORBextractor::ComputeKeyPointsOctTree, ORBextractor.cc:764-852
    void compute_keypoints() {
        cand.assign(nlevels, std::vector<Key>());
        lvlkp.assign(nlevels, std::vector<orc_kp>());
        const float W = 30;
        std::vector<orc_xyr> cell;
        for (int level = 0; level < nlevels; ++level) {
            const Img& im = pyr[level];
            const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
            const int maxBorderX = im.w - EDGE_THRESHOLD + 3, maxBorderY = im.h - EDGE_THRESHOLD + 3;
            std::vector<Key>& toDist = cand[level];
            const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
            const int nCols = (int)(width / W), nRows = (int)(height / W);
            const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
            for (int i = 0; i < nRows; i++) {
                const float iniY = (float)(minBorderY + i * hCell);
                float maxY = iniY + hCell + 6;
                if (iniY >= maxBorderY - 3) continue;
                if (maxY > maxBorderY) maxY = (float)maxBorderY;
                for (int j = 0; j < nCols; j++) {
                    const float iniX = (float)(minBorderX + j * wCell);
                    float maxX = iniX + wCell + 6;
                    if (iniX >= maxBorderX - 6) continue;
                    if (maxX > maxBorderX) maxX = (float)maxBorderX;
                    const int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
                    const u8* roi = im.roi() + (size_t)y0 * im.stride + x0;
                    fast9(roi, x1 - x0, y1 - y0, im.stride, iniThFAST, true, cell);
                    if (cell.empty()) fast9(roi, x1 - x0, y1 - y0, im.stride, minThFAST, true, cell);
                    for (const orc_xyr& c : cell)
                        toDist.push_back({(float)c.x + j * wCell, (float)c.y + i * hCell, (float)c.r});
                }
            }
            std::vector<Key> kept = distribute_octtree(toDist, minBorderX, maxBorderX, minBorderY, maxBorderY, quota[level]);
            const int scaledPatchSize = (int)(PATCH_SIZE * sf[level]);
            for (const Key& k : kept) {
                orc_kp kp;
                kp.x = k.x + minBorderX; kp.y = k.y + minBorderY;
                kp.size = (float)scaledPatchSize; kp.angle = -1; kp.response = k.response;
                kp.octave = level; kp.class_id = -1;
                lvlkp[level].push_back(kp);
            }
        }
        for (int level = 0; level < nlevels; ++level)
            for (orc_kp& kp : lvlkp[level]) kp.angle = ic_angle(pyr[level], kp.x, kp.y);
    }

    // computeOrbDescriptor, ORBextractor.cc:107-146
    static void orb_descriptor(const orc_kp& kpt, const u8* img, int step, u8* desc) {
        const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
        const float angle = (float)kpt.angle * factorPI;
        const float a = (float)cosf(angle), b = (float)sinf(angle);
        const u8* center = img + (size_t)cvRoundF(kpt.y) * step + cvRoundF(kpt.x);
        const signed char* pat = kPattern;
        for (int i = 0; i < 32; ++i, pat += 32) {
            int val = 0;
            for (int k = 0; k < 8; k++) {
                const int x0 = pat[4 * k], y0 = pat[4 * k + 1], x1 = pat[4 * k + 2], y1 = pat[4 * k + 3];
                const int t0 = center[cvRoundF(x0 * b + y0 * a) * step + cvRoundF(x0 * a - y0 * b)];
                const int t1 = center[cvRoundF(x1 * b + y1 * a) * step + cvRoundF(x1 * a - y1 * b)];
                val |= (t0 < t1) << k;
            }
            desc[i] = (u8)val;
        }
    }

    // ORBextractor::operator(), ORBextractor.cc:1042-1108.  Returns number of keypoints (or -1 if cap too small).
    int extract(const u8* img, int cols, int rows, int stride, const u8* mask, int mstride,
                orc_kp* kp_out, u8* desc_out, int cap) {
        if (!img || cols <= 0 || rows <= 0) return 0;
        std::vector<u8> image((size_t)cols * rows, 0);           // imageIn.copyTo(image, Mask)
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++)
                if (!mask || mask[(size_t)y * mstride + x]) image[(size_t)y * cols + x] = img[(size_t)y * stride + x];
        compute_pyramid(image.data(), cols, rows, cols);
        compute_keypoints();
        int nk = 0;
        for (int l = 0; l < nlevels; l++) nk += (int)lvlkp[l].size();
        if (nk > cap) return -1;
        blurred.assign(nlevels, std::vector<u8>());
        int offset = 0;
        for (int level = 0; level < nlevels; ++level) {
            std::vector<orc_kp>& kps = lvlkp[level];
            if (kps.empty()) continue;
            const Img& im = pyr[level];
            blurred[level].resize((size_t)im.w * im.h);
            gaussian_blur7(im.roi(), im.w, im.h, im.stride, blurred[level].data(), im.w);
            for (size_t i = 0; i < kps.size(); i++)
                orb_descriptor(kps[i], blurred[level].data(), im.w, desc_out + (size_t)(offset + i) * 32);
            for (size_t i = 0; i < kps.size(); i++) {
                orc_kp kp = kps[i];
                if (level != 0) { const float scale = sf[level]; kp.x *= scale; kp.y *= scale; }
                kp_out[offset + i] = kp;
            }
            offset += (int)kps.size();
        }
        return nk;
    }
};

}  // namespace

// ----------------------------------------------------------------------------------------------
// ORBmatcher restatement (flat-array form of the KeyFrame/Frame fields the functions touch)
// ----------------------------------------------------------------------------------------------
static const int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;      // ORBmatcher.cc:37-39

// ORBmatcher::DescriptorDistance, ORBmatcher.cc:1650-1666 (the bit-hack popcount, verbatim arithmetic)
static int descriptor_distance(const u8* a, const u8* b) {
    int32_t pa[8], pb[8];
    std::memcpy(pa, a, 32); std::memcpy(pb, b, 32);
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        unsigned int v = pa[i] ^ pb[i];
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

// ORBmatcher::ComputeThreeMaxima, ORBmatcher.cc:1604-1645 (on bin sizes)
static void three_maxima(const int* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = histo[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// A DBoW2::FeatureVector flattened: node ids ascending, CSR offsets, feature indices (ascending inside a node).
struct FeatVec { int nnodes; const int* ids; const int* off; const int* feat; };

static inline int rot_bin(float a1, float a2) {
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)std::round(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

extern "C" {

// ---- primitives ----
int orc_fast(const u8* img, int w, int h, int stride, int threshold, int nms, orc_xyr* out, int cap) {
    std::vector<orc_xyr> v;
    fast9(img, w, h, stride, threshold, nms != 0, v);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = v[i];
    return (int)v.size();
}
void orc_resize(const u8* src, int sw, int sh, int sstride, u8* dst, int dw, int dh, int dstride) {
    resize_linear_u8(src, sw, sh, sstride, dst, dw, dh, dstride);
}
void orc_blur(const u8* src, int w, int h, int sstride, u8* dst, int dstride) { gaussian_blur7(src, w, h, sstride, dst, dstride); }
void orc_border101(const u8* src, int w, int h, int sstride, u8* dst, int dstride, int b) { copy_make_border101(src, w, h, sstride, dst, dstride, b); }
float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
int orc_cv_round(double v) { return cvRoundD(v); }
const signed char* orc_pattern() { return kPattern; }

// ---- extractor ----
void* orc_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    return new Extractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void orc_extractor_destroy(void* h) { delete (Extractor*)h; }
void orc_extractor_tables(void* h, float* sf, float* isf, float* s2, float* is2, int* quota, int* umax) {
    Extractor* e = (Extractor*)h;
    for (int i = 0; i < e->nlevels; i++) {
        if (sf) sf[i] = e->sf[i];
        if (isf) isf[i] = e->isf[i];
        if (s2) s2[i] = e->s2[i];
        if (is2) is2[i] = e->is2[i];
        if (quota) quota[i] = e->quota[i];
    }
    if (umax) for (int i = 0; i <= HALF_PATCH_SIZE; i++) umax[i] = e->umax[i];
}
int orc_extractor_extract(void* h, const u8* img, int cols, int rows, int stride, const u8* mask, int mstride,
                          orc_kp* kp_out, u8* desc_out, int cap) {
    return ((Extractor*)h)->extract(img, cols, rows, stride, mask, mstride, kp_out, desc_out, cap);
}
void orc_extractor_level_dims(void* h, int level, int* w, int* hh) {
    Extractor* e = (Extractor*)h; *w = e->pyr[level].w; *hh = e->pyr[level].h;
}
// bordered != 0: copy the whole (w+38)x(h+38) buffer, else the w x h ROI
void orc_extractor_level_copy(void* h, int level, int bordered, u8* dst, int dstride) {
    Extractor* e = (Extractor*)h; const Img& im = e->pyr[level];
    if (bordered) for (int y = 0; y < im.h + 2 * EDGE_THRESHOLD; y++) std::memcpy(dst + (size_t)y * dstride, im.buf.data() + (size_t)y * im.stride, im.stride);
    else for (int y = 0; y < im.h; y++) std::memcpy(dst + (size_t)y * dstride, im.roi() + (size_t)y * im.stride, im.w);
}
int orc_extractor_blurred_copy(void* h, int level, u8* dst, int dstride) {
    Extractor* e = (Extractor*)h; const Img& im = e->pyr[level];
    if (e->blurred[level].empty()) return 0;
    for (int y = 0; y < im.h; y++) std::memcpy(dst + (size_t)y * dstride, e->blurred[level].data() + (size_t)y * im.w, im.w);
    return 1;
}
// FAST candidates of a level in octree input order, coordinates in level pixels (minBorder added back)
int orc_extractor_candidates(void* h, int level, orc_xyr* out, int cap) {
    Extractor* e = (Extractor*)h; const std::vector<Key>& c = e->cand[level];
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) out[i] = {(int)c[i].x + 16, (int)c[i].y + 16, (int)c[i].response};
    return (int)c.size();
}
int orc_extractor_level_keypoints(void* h, int level, orc_kp* out, int cap) {
    Extractor* e = (Extractor*)h; const std::vector<orc_kp>& c = e->lvlkp[level];
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) out[i] = c[i];
    return (int)c.size();
}
// stand-alone octree on an explicit candidate list (coords relative to minBorder, as the reference passes them)
int orc_octree(const float* x, const float* y, const float* resp, int n, int minX, int maxX, int minY, int maxY, int N,
               float* ox, float* oy, float* oresp, int cap) {
    std::vector<Key> in(n);
    for (int i = 0; i < n; i++) in[i] = {x[i], y[i], resp[i]};
    std::vector<Key> r = distribute_octtree(in, minX, maxX, minY, maxY, N);
    for (size_t i = 0; i < r.size() && (int)i < cap; i++) { ox[i] = r[i].x; oy[i] = r[i].y; oresp[i] = r[i].response; }
    return (int)r.size();
}
// descriptor of one keypoint on an already blurred image
void orc_descriptor(const u8* blurred, int step, float x, float y, float angle, u8* desc) {
    orc_kp kp; kp.x = x; kp.y = y; kp.angle = angle;
    Extractor::orb_descriptor(kp, blurred, step, desc);
}

// CPU baseline helper: extract `nframes` frames (contiguous cols x rows each) with `nthreads` threads, one extractor
// instance per thread, frames dealt round-robin.  Returns wall seconds; total keypoints in *nkp_total.
double orc_extract_batch_mt(const u8* frames, int nframes, int cols, int rows, int nfeatures, float scaleFactor, int nlevels,
                            int iniTh, int minTh, int nthreads, long* nkp_total) {
    std::atomic<long> total(0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&, t]() {
            Extractor e(nfeatures, scaleFactor, nlevels, iniTh, minTh);
            const int cap = nfeatures + 4 * nlevels + 64;
            std::vector<orc_kp> kp(cap); std::vector<u8> desc((size_t)cap * 32);
            long s = 0;
            for (int f = t; f < nframes; f += nthreads) {
                int n = e.extract(frames + (size_t)f * cols * rows, cols, rows, cols, nullptr, 0, kp.data(), desc.data(), cap);
                if (n > 0) s += n;
            }
            total += s;
        });
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    if (nkp_total) *nkp_total = total.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- matcher ----
int orc_descriptor_distance(const u8* a, const u8* b) { return descriptor_distance(a, b); }

void orc_three_maxima(const int* histo, int L, int* ind) {
    int i1 = -1, i2 = -1, i3 = -1;
    three_maxima(histo, L, i1, i2, i3);
    ind[0] = i1; ind[1] = i2; ind[2] = i3;
}

// Brute-force top-2 (the inner loop shape of SearchByBoW, ORBmatcher.cc:205-229, over the whole set):
// best_idx = first index attaining the minimum distance, best = that distance, second = second smallest (256 if none).
void orc_hamming_top2(const u8* q, int nq, const u8* db, int ndb, int* best_idx, int* best, int* second) {
    for (int i = 0; i < nq; i++) {
        int b1 = 256, b2 = 256, bi = -1;
        for (int j = 0; j < ndb; j++) {
            const int d = descriptor_distance(q + (size_t)i * 32, db + (size_t)j * 32);
            if (d < b1) { b2 = b1; b1 = d; bi = j; } else if (d < b2) { b2 = d; }
        }
        best_idx[i] = bi; best[i] = b1; second[i] = b2;
    }
}
// multi-threaded variant for the CPU baseline; returns wall seconds
double orc_hamming_top2_mt(const u8* q, int nq, const u8* db, int ndb, int* best_idx, int* best, int* second, int nthreads) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([=]() {
            const int lo = (int)((long)nq * t / nthreads), hi = (int)((long)nq * (t + 1) / nthreads);
            orc_hamming_top2(q + (size_t)lo * 32, hi - lo, db, ndb, best_idx + lo, best + lo, second + lo);
        });
    for (auto& x : th) x.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// SearchByBoW(KeyFrame*, Frame&, ...), ORBmatcher.cc:159-291.
//   valid1[i] = KF feature i has a MapPoint that is not bad.   match21[j] (size n2) = KF feature index whose
//   MapPoint the reference would store in vpMapPointMatches[j], or -1.   Returns nmatches.
int orc_search_bow_kf_f(const u8* desc1, int n1, const u8* valid1, const float* angle1,
                        int nn1, const int* ids1, const int* off1, const int* feat1,
                        const u8* desc2, int n2, const float* angle2,
                        int nn2, const int* ids2, const int* off2, const int* feat2,
                        float nnratio, int checkOri, int* match21) {
    (void)n1;
    for (int j = 0; j < n2; j++) match21[j] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (ids1[a] == ids2[b]) {
            for (int iKF = off1[a]; iKF < off1[a + 1]; iKF++) {
                const int realIdxKF = feat1[iKF];
                if (!valid1[realIdxKF]) continue;
                const u8* dKF = desc1 + (size_t)realIdxKF * 32;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (int iF = off2[b]; iF < off2[b + 1]; iF++) {
                    const int realIdxF = feat2[iF];
                    if (match21[realIdxF] >= 0) continue;
                    const int dist = descriptor_distance(dKF, desc2 + (size_t)realIdxF * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                    else if (dist < bestDist2) { bestDist2 = dist; }
                }
                if (bestDist1 <= TH_LOW) {
                    if (static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
                        match21[bestIdxF] = realIdxKF;
                        if (checkOri) rotHist[rot_bin(angle1[realIdxKF], angle2[bestIdxF])].push_back(bestIdxF);
                        nmatches++;
                    }
                }
            }
            a++; b++;
        } else if (ids1[a] < ids2[b]) {
            a = (int)(std::lower_bound(ids1, ids1 + nn1, ids2[b]) - ids1);
        } else {
            b = (int)(std::lower_bound(ids2, ids2 + nn2, ids1[a]) - ids2);
        }
    }
    if (checkOri) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { match21[j] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// SearchByBoW(KeyFrame*, KeyFrame*, ...), ORBmatcher.cc:525-658.  match12[i] (size n1) = idx2 or -1.
int orc_search_bow_kf_kf(const u8* desc1, int n1, const u8* valid1, const float* angle1,
                         int nn1, const int* ids1, const int* off1, const int* feat1,
                         const u8* desc2, int n2, const u8* valid2, const float* angle2,
                         int nn2, const int* ids2, const int* off2, const int* feat2,
                         float nnratio, int checkOri, int* match12) {
    for (int i = 0; i < n1; i++) match12[i] = -1;
    std::vector<bool> matched2(n2, false);
    std::vector<int> rotHist[HISTO_LENGTH];
    int nmatches = 0, a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (ids1[a] == ids2[b]) {
            for (int i1 = off1[a]; i1 < off1[a + 1]; i1++) {
                const int idx1 = feat1[i1];
                if (!valid1[idx1]) continue;
                const u8* d1 = desc1 + (size_t)idx1 * 32;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int i2 = off2[b]; i2 < off2[b + 1]; i2++) {
                    const int idx2 = feat2[i2];
                    if (matched2[idx2] || !valid2[idx2]) continue;
                    const int dist = descriptor_distance(d1, desc2 + (size_t)idx2 * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) { bestDist2 = dist; }
                }
                if (bestDist1 < TH_LOW) {
                    if (static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
                        match12[idx1] = bestIdx2;
                        matched2[bestIdx2] = true;
                        if (checkOri) rotHist[rot_bin(angle1[idx1], angle2[bestIdx2])].push_back(idx1);
                        nmatches++;
                    }
                }
            }
            a++; b++;
        } else if (ids1[a] < ids2[b]) {
            a = (int)(std::lower_bound(ids1, ids1 + nn1, ids2[b]) - ids1);
        } else {
            b = (int)(std::lower_bound(ids2, ids2 + nn2, ids1[a]) - ids2);
        }
    }
    if (checkOri) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { match12[j] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// CheckDistEpipolarLine, ORBmatcher.cc:140-157   (F12 row-major 3x3 float)
static bool check_dist_epipolar_line(float x1, float y1, float x2, float y2, int octave2, const float* F12, const float* sigma2_2) {
    const float a = x1 * F12[0] + y1 * F12[3] + F12[6];
    const float b = x1 * F12[1] + y1 * F12[4] + F12[7];
    const float c = x1 * F12[2] + y1 * F12[5] + F12[8];
    const float num = a * x2 + b * y2 + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * sigma2_2[octave2];
}

// SearchForTriangulation, ORBmatcher.cc:660-826.
//   hasmp{1,2}[i] = GetMapPoint(i) != NULL; kp = mvKeysUn (x, y, angle, octave); uright = mvuRight;
//   (ex, ey) = epipole of KF1's centre in KF2 (ORBmatcher.cc:667-673, computed by the caller);
//   sf2 / sigma2_2 = pKF2->mvScaleFactors / mvLevelSigma2.  pairs_out = (idx1, idx2) ascending idx1.
int orc_search_triangulation(const u8* desc1, int n1, const u8* hasmp1, const float* uright1,
                             const float* kx1, const float* ky1, const float* ang1,
                             int nn1, const int* ids1, const int* off1, const int* feat1,
                             const u8* desc2, int n2, const u8* hasmp2, const float* uright2,
                             const float* kx2, const float* ky2, const float* ang2, const int* oct2,
                             int nn2, const int* ids2, const int* off2, const int* feat2,
                             const float* F12, float ex, float ey, const float* sf2, const float* sigma2_2,
                             int onlyStereo, int checkOri, int* pairs_out, int* npairs_out) {
    int nmatches = 0;
    std::vector<bool> matched2(n2, false);
    std::vector<int> m12(n1, -1);
    std::vector<int> rotHist[HISTO_LENGTH];
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (ids1[a] == ids2[b]) {
            for (int i1 = off1[a]; i1 < off1[a + 1]; i1++) {
                const int idx1 = feat1[i1];
                if (hasmp1[idx1]) continue;
                const bool bStereo1 = uright1[idx1] >= 0;
                if (onlyStereo) if (!bStereo1) continue;
                const u8* d1 = desc1 + (size_t)idx1 * 32;
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int i2 = off2[b]; i2 < off2[b + 1]; i2++) {
                    const int idx2 = feat2[i2];
                    if (matched2[idx2] || hasmp2[idx2]) continue;
                    const bool bStereo2 = uright2[idx2] >= 0;
                    if (onlyStereo) if (!bStereo2) continue;
                    const int dist = descriptor_distance(d1, desc2 + (size_t)idx2 * 32);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ex - kx2[idx2], distey = ey - ky2[idx2];
                        if (distex * distex + distey * distey < 100 * sf2[oct2[idx2]]) continue;
                    }
                    if (check_dist_epipolar_line(kx1[idx1], ky1[idx1], kx2[idx2], ky2[idx2], oct2[idx2], F12, sigma2_2)) {
                        bestIdx2 = idx2; bestDist = dist;
                    }
                }
                if (bestIdx2 >= 0) {
                    m12[idx1] = bestIdx2;
                    nmatches++;
                    if (checkOri) rotHist[rot_bin(ang1[idx1], ang2[bestIdx2])].push_back(idx1);
                }
            }
            a++; b++;
        } else if (ids1[a] < ids2[b]) {
            a = (int)(std::lower_bound(ids1, ids1 + nn1, ids2[b]) - ids1);
        } else {
            b = (int)(std::lower_bound(ids2, ids2 + nn2, ids1[a]) - ids2);
        }
    }
    if (checkOri) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { m12[j] = -1; nmatches--; }
        }
    }
    int np = 0;
    for (int i = 0; i < n1; i++) {
        if (m12[i] < 0) continue;
        pairs_out[2 * np] = i; pairs_out[2 * np + 1] = m12[i];
        np++;
    }
    *npairs_out = np;
    return nmatches;
}


// ---- DBoW2 vocabulary: TemplatedVocabulary::transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1140-1207, 1231-1272)
// restated on flat node arrays (node 0 = root, parent[i] < i, children in node order).
void orc_voc_transform(int n_nodes, const int* parent, const u8* ndesc, const double* nweight, const u8* is_leaf, int L,
                       const u8* feat, int n, int levelsup, int* word_id, double* weight, int* node_id) {
    std::vector<std::vector<int> > children(n_nodes);
    std::vector<int> word(n_nodes, -1);
    int nw = 0;
    for (int i = 1; i < n_nodes; i++) { children[parent[i]].push_back(i); if (is_leaf[i]) word[i] = nw++; }
    for (int f = 0; f < n; f++) {
        const int nid_level = L - levelsup;
        int nid = 0;                                            // root when nid_level <= 0
        int final_id = 0, current_level = 0;
        do {
            ++current_level;
            const std::vector<int>& nodes = children[final_id];
            if (nodes.empty()) break;                            // (a one-node vocabulary: the reference would crash here)
            final_id = nodes[0];
            double best_d = (double)descriptor_distance(feat + (size_t)f * 32, ndesc + (size_t)final_id * 32);
            for (size_t c = 1; c < nodes.size(); c++) {
                const double d = (double)descriptor_distance(feat + (size_t)f * 32, ndesc + (size_t)nodes[c] * 32);
                if (d < best_d) { best_d = d; final_id = nodes[c]; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (!children[final_id].empty());
        word_id[f] = word[final_id]; weight[f] = nweight[final_id]; node_id[f] = nid;
    }
}
// BowVector as transform(features, v, fv, levelsup) builds it: weighting 0 TF_IDF, 1 TF, 2 IDF, 3 BINARY; scoring 0 L1_NORM, 1 L2_NORM,
// 2 CHI_SQUARE, 3 KL, 4 BHATTACHARYYA, 5 DOT_PRODUCT (mustNormalize: L1 for 0,2,3,4; L2 for 1; none for 5).  Returns the map size.
int orc_voc_bow(int n, const int* word_id, const double* weight, int weighting, int scoring, int* out_word, double* out_value, int cap) {
    std::map<unsigned, double> v;
    const bool must = scoring != 5;
    const bool l2 = scoring == 1;
    for (int i = 0; i < n; i++) {
        if (!(weight[i] > 0)) continue;
        if (weighting == 0 || weighting == 1) {
            std::map<unsigned, double>::iterator it = v.lower_bound((unsigned)word_id[i]);
            if (it != v.end() && !(v.key_comp()((unsigned)word_id[i], it->first))) it->second += weight[i];
            else v.insert(it, std::make_pair((unsigned)word_id[i], weight[i]));
        } else {
            std::map<unsigned, double>::iterator it = v.lower_bound((unsigned)word_id[i]);
            if (it == v.end() || v.key_comp()((unsigned)word_id[i], it->first)) v.insert(it, std::make_pair((unsigned)word_id[i], weight[i]));
        }
    }
    if ((weighting == 0 || weighting == 1) && !v.empty() && !must) {
        const double nd = (double)v.size();
        for (std::map<unsigned, double>::iterator it = v.begin(); it != v.end(); ++it) it->second /= nd;
    }
    if (must) {
        double norm = 0.0;
        if (!l2) for (std::map<unsigned, double>::iterator it = v.begin(); it != v.end(); ++it) norm += std::fabs(it->second);
        else { for (std::map<unsigned, double>::iterator it = v.begin(); it != v.end(); ++it) norm += it->second * it->second; norm = std::sqrt(norm); }
        if (norm > 0.0) for (std::map<unsigned, double>::iterator it = v.begin(); it != v.end(); ++it) it->second /= norm;
    }
    int k = 0;
    for (std::map<unsigned, double>::iterator it = v.begin(); it != v.end() && k < cap; ++it, ++k) { out_word[k] = (int)it->first; out_value[k] = it->second; }
    return (int)v.size();
}


// MapPoint::ComputeDistinctiveDescriptors selection (src/MapPoint.cc:511-541) for one set of N descriptors
int orc_distinctive(const u8* desc, int N) {
    if (N <= 0) return -1;
    std::vector<std::vector<float> > D(N, std::vector<float>(N, 0.f));
    for (int i = 0; i < N; i++) {
        D[i][i] = 0;
        for (int j = i + 1; j < N; j++) {
            const int dij = descriptor_distance(desc + (size_t)i * 32, desc + (size_t)j * 32);
            D[i][j] = (float)dij; D[j][i] = (float)dij;
        }
    }
    int BestMedian = 0x7fffffff, BestIdx = 0;
    for (int i = 0; i < N; i++) {
        std::vector<int> v(D[i].begin(), D[i].end());
        std::sort(v.begin(), v.end());
        const int median = v[(size_t)(0.5 * (N - 1))];
        if (median < BestMedian) { BestMedian = median; BestIdx = i; }
    }
    return BestIdx;
}


// ---- projection / window searches ----------------------------------------------------------------------------
// The Frame/KeyFrame side that is searched, flat (include/Frame.h:98,141-197): same layout as orbm_grid_view of the product ABI.
struct orc_grid_view {
    int n;
    const u8* desc; const float* x; const float* y; const int* octave; const float* angle; const float* uright; const u8* blocked;
    int grid_cols, grid_rows;
    float min_x, min_y, max_x, max_y, inv_w, inv_h;
    const int* cell_offsets; const int* cell_features;
    const float* scale_factors; int n_levels;
};

// Frame::AssignFeaturesToGrid + PosInGrid, src/Frame.cc:341-356, 500-510 (note: round, not floor).  cell index = ix*rows + iy.
// Writes CSR (offsets[cols*rows+1], features[<= n]); returns the number of features placed.
int orc_assign_features_to_grid(int n, const float* x, const float* y, int cols, int rows, float min_x, float min_y, float inv_w,
                                float inv_h, int* offsets, int* features) {
    std::vector<std::vector<int> > grid((size_t)cols * rows);
    for (int i = 0; i < n; i++) {
        const int posX = (int)std::round((x[i] - min_x) * inv_w);
        const int posY = (int)std::round((y[i] - min_y) * inv_h);
        if (posX < 0 || posX >= cols || posY < 0 || posY >= rows) continue;
        grid[(size_t)posX * rows + posY].push_back(i);
    }
    int k = 0;
    offsets[0] = 0;
    for (size_t c = 0; c < grid.size(); c++) {
        for (int i : grid[c]) features[k++] = i;
        offsets[c + 1] = k;
    }
    return k;
}

// Frame::GetFeaturesInArea, src/Frame.cc:445-498 (KeyFrame::GetFeaturesInArea, src/KeyFrame.cc:1311-1350, is the same without
// the level test: call with minLevel = maxLevel = -1).
static void features_in_area(const orc_grid_view& F, float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) {
    out.clear();
    const int nMinCellX = std::max(0, (int)std::floor((x - F.min_x - r) * F.inv_w));
    if (nMinCellX >= F.grid_cols) return;
    const int nMaxCellX = std::min(F.grid_cols - 1, (int)std::ceil((x - F.min_x + r) * F.inv_w));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)std::floor((y - F.min_y - r) * F.inv_h));
    if (nMinCellY >= F.grid_rows) return;
    const int nMaxCellY = std::min(F.grid_rows - 1, (int)std::ceil((y - F.min_y + r) * F.inv_h));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const int c = ix * F.grid_rows + iy;
            for (int j = F.cell_offsets[c]; j < F.cell_offsets[c + 1]; j++) {
                const int idx = F.cell_features[j];
                if (bCheckLevels) {
                    if (F.octave[idx] < minLevel) continue;
                    if (maxLevel >= 0)
                        if (F.octave[idx] > maxLevel) continue;
                }
                const float distx = F.x[idx] - x;
                const float disty = F.y[idx] - y;
                if (std::fabs(distx) < r && std::fabs(disty) < r) out.push_back(idx);
            }
        }
}

// SearchByProjection(Frame&, const vector<MapPoint*>&, th), ORBmatcher.cc:45-129.  owner[F.n]: last map point stored per feature.
int orc_search_projection_map(const orc_grid_view* Fp, int n_points, const u8* in_view, const float* proj_x, const float* proj_y,
                              const float* proj_xr, const int* level, const float* view_cos, const u8* desc, const u8* claims,
                              float th, float nnratio, int* owner) {
    const orc_grid_view& F = *Fp;
    for (int i = 0; i < F.n; i++) owner[i] = -1;
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    std::vector<int> vIndices;
    for (int iMP = 0; iMP < n_points; iMP++) {
        if (!in_view[iMP]) continue;                                   // !mbTrackInView || isBad()
        const int nPredictedLevel = level[iMP];
        float r = view_cos[iMP] > 0.998 ? 2.5f : 4.0f;                 // RadiusByViewingCos, :131-137
        if (bFactor) r *= th;
        features_in_area(F, proj_x[iMP], proj_y[iMP], r * F.scale_factors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, vIndices);
        if (vIndices.empty()) continue;
        const u8* MPdescriptor = desc + (size_t)iMP * 32;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : vIndices) {
            const bool hasObs = owner[idx] >= 0 ? claims[owner[idx]] != 0 : (F.blocked && F.blocked[idx]);
            if (hasObs) continue;                                      // mvpMapPoints[idx] && Observations() > 0
            if (F.uright && F.uright[idx] > 0) {
                const float er = std::fabs(proj_xr[iMP] - F.uright[idx]);
                if (er > r * F.scale_factors[nPredictedLevel]) continue;
            }
            const int dist = descriptor_distance(MPdescriptor, F.desc + (size_t)idx * 32);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = F.octave[idx]; bestIdx = idx; }
            else if (dist < bestDist2) { bestLevel2 = F.octave[idx]; bestDist2 = dist; }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            owner[bestIdx] = iMP;
            nmatches++;
        }
    }
    return nmatches;
}

// SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono), ORBmatcher.cc:1331-1463.
// 3x3 * 3x1 + 3x1 products follow cv::gemm for small float matrices as observed with cv2 4.13: float32, left to right, no FMA
// (tests/golden/make_golden.py pins this).  owner[cur.n]: LastFrame feature index, -1 untouched, -2 set to NULL by the cull.
static inline void rt_apply(const float* T /*3x4*/, const float* p, float* out) {
    for (int r = 0; r < 3; r++) out[r] = ((T[4 * r] * p[0] + T[4 * r + 1] * p[1]) + T[4 * r + 2] * p[2]) + T[4 * r + 3];
}
// exposed for the golden test: out = T[:, :3] * p + T[:, 3] ; outT = -T[:, :3]^T * T[:, 3]
void orc_rt_apply(const float* T, const float* p, float* out) { rt_apply(T, p, out); }
void orc_minus_rt_t(const float* T, float* out) {
    for (int r = 0; r < 3; r++)
        out[r] = (float)(-(((double)T[0 * 4 + r] * T[3] + (double)T[1 * 4 + r] * T[7]) + (double)T[2 * 4 + r] * T[11]));
}
int orc_search_projection_frame(const orc_grid_view* Cp, const float* Tcw, const float* Tlw, float fx, float fy, float cx, float cy,
                                float mbf, float mb, int n_last, const u8* has_point, const float* world, const int* octave,
                                const float* angle, const u8* desc, const u8* claims, float th, int mono, int checkOri, int* owner) {
    const orc_grid_view& CF = *Cp;
    for (int i = 0; i < CF.n; i++) owner[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    // twc = -Rcw.t()*tcw ; tlc = Rlw*twc + tlw  (:1344-1352)
    float twc[3], tlc[3];
    // (the transposed product takes cv::gemm's general path, which accumulates float inputs in double; pinned by prim_gemm3.npz)
    for (int r = 0; r < 3; r++)
        twc[r] = (float)(-(((double)Tcw[0 * 4 + r] * Tcw[3] + (double)Tcw[1 * 4 + r] * Tcw[7]) + (double)Tcw[2 * 4 + r] * Tcw[11]));
    rt_apply(Tlw, twc, tlc);
    const bool bForward = tlc[2] > mb && !mono;
    const bool bBackward = -tlc[2] > mb && !mono;
    std::vector<int> vIndices2;
    for (int i = 0; i < n_last; i++) {
        if (!has_point[i]) continue;                                   // pMP && !mvbOutlier[i]
        float x3Dc[3];
        rt_apply(Tcw, world + 3 * (size_t)i, x3Dc);
        const float xc = x3Dc[0], yc = x3Dc[1];
        const float invzc = 1.0 / x3Dc[2];
        if (invzc < 0) continue;
        float u = fx * xc * invzc + cx;
        float v = fy * yc * invzc + cy;
        if (u < CF.min_x || u > CF.max_x) continue;
        if (v < CF.min_y || v > CF.max_y) continue;
        const int nLastOctave = octave[i];
        float radius = th * CF.scale_factors[nLastOctave];
        if (bForward) features_in_area(CF, u, v, radius, nLastOctave, -1, vIndices2);
        else if (bBackward) features_in_area(CF, u, v, radius, 0, nLastOctave, vIndices2);
        else features_in_area(CF, u, v, radius, nLastOctave - 1, nLastOctave + 1, vIndices2);
        if (vIndices2.empty()) continue;
        const u8* dMP = desc + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            const bool hasObs = owner[i2] >= 0 ? claims[owner[i2]] != 0 : (owner[i2] == -1 && CF.blocked && CF.blocked[i2]);
            if (hasObs) continue;
            if (CF.uright && CF.uright[i2] > 0) {
                const float ur = u - mbf * invzc;
                const float er = std::fabs(ur - CF.uright[i2]);
                if (er > radius) continue;
            }
            const int dist = descriptor_distance(dMP, CF.desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            owner[bestIdx2] = i;
            nmatches++;
            if (checkOri) rotHist[rot_bin(angle[i], CF.angle[bestIdx2])].push_back(bestIdx2);
        }
    }
    if (checkOri) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int j : rotHist[i]) { owner[j] = -2; nmatches--; }
    }
    return nmatches;
}

// SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize), ORBmatcher.cc:408-523.
int orc_search_initialization(const orc_grid_view* F2p, int n1, const u8* desc1, const int* octave1, const float* angle1,
                              float* prev_xy, int windowSize, float nnratio, int checkOri, int* vnMatches12) {
    const orc_grid_view& F2 = *F2p;
    int nmatches = 0;
    for (int i = 0; i < n1; i++) vnMatches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<int> vMatchedDistance(F2.n, INT_MAX), vnMatches21(F2.n, -1);
    std::vector<int> vIndices2;
    for (int i1 = 0; i1 < n1; i1++) {
        const int level1 = octave1[i1];
        if (level1 > 0) continue;
        features_in_area(F2, prev_xy[2 * i1], prev_xy[2 * i1 + 1], (float)windowSize, level1, level1, vIndices2);
        if (vIndices2.empty()) continue;
        const u8* d1 = desc1 + (size_t)i1 * 32;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            const int dist = descriptor_distance(d1, F2.desc + (size_t)i2 * 32);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) { bestDist2 = dist; }
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                vnMatches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (checkOri) rotHist[rot_bin(angle1[i1], F2.angle[bestIdx2])].push_back(i1);
            }
        }
    }
    if (checkOri) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i])
                if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
        }
    }
    for (int i1 = 0; i1 < n1; i1++)
        if (vnMatches12[i1] >= 0) { prev_xy[2 * i1] = F2.x[vnMatches12[i1]]; prev_xy[2 * i1 + 1] = F2.y[vnMatches12[i1]]; }
    return nmatches;
}


// ---- the shared core of the remaining projection searches: explicit windows, best candidate, every accepted match blocks ------
// per query: active, (u, v, r), GetFeaturesInArea level bounds; candidates with F.blocked are skipped (:1548-1549 / :375-376).
// owner[F.n]: query index, -1 untouched, -2 set to NULL by the rotation cull (only with checkOri).
int orc_search_windows(const orc_grid_view* Fp, int nq, const u8* active, const float* u, const float* v, const float* r,
                       const int* min_level, const int* max_level, const u8* desc, const float* angle, int th_dist, int checkOri,
                       int* owner) {
    const orc_grid_view& F = *Fp;
    for (int i = 0; i < F.n; i++) owner[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH], cand;
    for (int q = 0; q < nq; q++) {
        if (!active[q]) continue;
        features_in_area(F, u[q], v[q], r[q], min_level[q], max_level[q], cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : cand) {
            if ((F.blocked && F.blocked[i2]) || owner[i2] >= 0) continue;
            const int dist = descriptor_distance(desc + (size_t)q * 32, F.desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= th_dist) {
            owner[bestIdx2] = q;
            nmatches++;
            if (checkOri) rotHist[rot_bin(angle[q], F.angle[bestIdx2])].push_back(bestIdx2);
        }
    }
    if (checkOri) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int j : rotHist[i]) { owner[j] = -2; nmatches--; }
    }
    return nmatches;
}

// MapPoint::PredictScale, src/MapPoint.cc:633-642: ceil(log(mfMaxDistance / dist) / logScaleFactor), float log (std::log(float)).
// This fork does not clamp the result, and then indexes mvScaleFactors with it (undefined behaviour outside [0, nlevels));
// oracle and product clamp to the valid range, as upstream ORB-SLAM2 does since 2017.
static int predict_scale(float mfMaxDistance, float currentDist, float logScaleFactor, int nlevels) {
    const float ratio = mfMaxDistance / currentDist;
    int nScale = (int)std::ceil(std::log(ratio) / logScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= nlevels) nScale = nlevels - 1;
    return nScale;
}
static inline float norm3(const float* p) {          // cv::norm(3x1 CV_32F): double accumulation, sqrt, then -> float (prim_gemm3.npz)
    return (float)std::sqrt((double)p[0] * p[0] + (double)p[1] * p[1] + (double)p[2] * p[2]);
}
static inline void minus_rt_t(const float* T, float* out) {
    for (int r = 0; r < 3; r++)
        out[r] = (float)(-(((double)T[0 * 4 + r] * T[3] + (double)T[1 * 4 + r] * T[7]) + (double)T[2 * 4 + r] * T[11]));
}

// SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist), ORBmatcher.cc:1465-1602.
// valid[i] = pMP && !isBad() && !sAlreadyFound.count(pMP); CF.blocked = CurrentFrame.mvpMapPoints[i2] != NULL (:1548-1549).
int orc_search_projection_kf(const orc_grid_view* Cp, const float* Tcw, float fx, float fy, float cx, float cy, float logScaleFactor,
                             int n_kf, const u8* valid, const float* world, const float* mf_max, const float* mf_min,
                             const float* angle, const u8* desc, float th, int ORBdist, int checkOri, int* owner) {
    const orc_grid_view& CF = *Cp;
    std::vector<u8> active(n_kf, 0);
    std::vector<float> u(n_kf, 0.f), v(n_kf, 0.f), rad(n_kf, 0.f);
    std::vector<int> minL(n_kf, -1), maxL(n_kf, -1);
    float Ow[3];
    minus_rt_t(Tcw, Ow);
    for (int i = 0; i < n_kf; i++) {
        if (!valid[i]) continue;
        float x3Dc[3];
        const float* x3Dw = world + 3 * (size_t)i;
        rt_apply(Tcw, x3Dw, x3Dc);
        const float xc = x3Dc[0], yc = x3Dc[1];
        const float invzc = 1.0 / x3Dc[2];
        const float uu = fx * xc * invzc + cx;
        const float vv = fy * yc * invzc + cy;
        if (uu < CF.min_x || uu > CF.max_x) continue;
        if (vv < CF.min_y || vv > CF.max_y) continue;
        const float PO[3] = {x3Dw[0] - Ow[0], x3Dw[1] - Ow[1], x3Dw[2] - Ow[2]};
        const float dist3D = norm3(PO);
        const float maxDistance = 1.2f * mf_max[i], minDistance = 0.8f * mf_min[i];
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const int nPredictedLevel = predict_scale(mf_max[i], dist3D, logScaleFactor, CF.n_levels);
        active[i] = 1; u[i] = uu; v[i] = vv;
        rad[i] = th * CF.scale_factors[nPredictedLevel];
        minL[i] = nPredictedLevel - 1; maxL[i] = nPredictedLevel + 1;
    }
    return orc_search_windows(Cp, n_kf, active.data(), u.data(), v.data(), rad.data(), minL.data(), maxL.data(), desc, angle, ORBdist,
                              checkOri, owner);
}

// SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th),
// ORBmatcher.cc:293-406.  Scw: rows 0..2 of the 4x4 similarity.  valid[i] = !isBad() && !spAlreadyFound.count(pMP);
// KF.blocked = vpMatched[idx] != NULL.  Conventions for the cv::Mat expressions that cannot be pinned from Python:
// row.dot(row) and PO.dot(Pn) accumulate in double (cv::Mat::dot); sRcw/scw and col/scw multiply by (float)(1.0/scw).
int orc_search_projection_sim3(const orc_grid_view* Kp, const float* Scw, float fx, float fy, float cx, float cy, float logScaleFactor,
                               int n_pts, const u8* valid, const float* world, const float* mf_max, const float* mf_min,
                               const float* normal, const u8* desc, int th, int* owner) {
    const orc_grid_view& KF = *Kp;
    const float scw = (float)std::sqrt((double)Scw[0] * Scw[0] + (double)Scw[1] * Scw[1] + (double)Scw[2] * Scw[2]);
    const float inv = (float)(1.0 / (double)scw);
    float T[12];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) T[4 * r + c] = Scw[4 * r + c] * inv;
    float Ow[3];
    minus_rt_t(T, Ow);
    std::vector<u8> active(n_pts, 0);
    std::vector<float> u(n_pts, 0.f), v(n_pts, 0.f), rad(n_pts, 0.f);
    std::vector<int> minL(n_pts, -1), maxL(n_pts, -1);
    for (int i = 0; i < n_pts; i++) {
        if (!valid[i]) continue;
        const float* p3Dw = world + 3 * (size_t)i;
        float p3Dc[3];
        rt_apply(T, p3Dw, p3Dc);
        if (p3Dc[2] < 0.0) continue;
        const float invz = 1 / p3Dc[2];
        const float x = p3Dc[0] * invz, y = p3Dc[1] * invz;
        const float uu = fx * x + cx, vv = fy * y + cy;
        if (!(uu >= KF.min_x && uu < KF.max_x && vv >= KF.min_y && vv < KF.max_y)) continue;       // KeyFrame::IsInImage
        const float PO[3] = {p3Dw[0] - Ow[0], p3Dw[1] - Ow[1], p3Dw[2] - Ow[2]};
        const float dist = norm3(PO);
        const float maxDistance = 1.2f * mf_max[i], minDistance = 0.8f * mf_min[i];
        if (dist < minDistance || dist > maxDistance) continue;
        const float* Pn = normal + 3 * (size_t)i;
        const double dot = (double)PO[0] * Pn[0] + (double)PO[1] * Pn[1] + (double)PO[2] * Pn[2];
        if (dot < 0.5 * dist) continue;
        const int nPredictedLevel = predict_scale(mf_max[i], dist, logScaleFactor, KF.n_levels);
        active[i] = 1; u[i] = uu; v[i] = vv;
        rad[i] = th * KF.scale_factors[nPredictedLevel];
        minL[i] = nPredictedLevel - 1; maxL[i] = nPredictedLevel;
    }
    return orc_search_windows(Kp, n_pts, active.data(), u.data(), v.data(), rad.data(), minL.data(), maxL.data(), desc, nullptr, TH_LOW, 0, owner);
}


// ---- Frame::ComputeStereoMatches, src/Frame.cc:584-756 -------------------------------------------------------------
// pyrL / pyrR: the border-less levels of the two extractors' mvImagePyramid (contiguous, stride = w[l]).  kpL = mvKeys, kpR =
// mvKeysRight.  Out: mvuRight, mvDepth (n_left each).  Two undefined behaviours of the reference are given a definition here
// (and identically in the product): rows outside [0, nRows) are not entered in the row table (:604-612 would write out of
// bounds), and an empty vDistIdx skips the median cut (:738-739 would read element 0 of an empty vector).
void orc_stereo_matches(int nlevels, const u8* const* pyrL, const u8* const* pyrR, const int* w, const int* h,
                        const float* scaleFactors, const float* invScaleFactors,
                        const orc_kp* kpL, const u8* descL, int N, const orc_kp* kpR, const u8* descR, int Nr,
                        float mbf, float mb, float* mvuRight, float* mvDepth) {
    (void)nlevels;
    for (int i = 0; i < N; i++) { mvuRight[i] = -1.0f; mvDepth[i] = -1.0f; }
    const int nRows = h[0];
    std::vector<std::vector<size_t> > vRowIndices(nRows, std::vector<size_t>());
    for (int iR = 0; iR < Nr; iR++) {
        const float kpY = kpR[iR].y;
        const float r = 2.0f * scaleFactors[kpR[iR].octave];
        const int maxr = (int)std::ceil(kpY + r);
        const int minr = (int)std::floor(kpY - r);
        for (int yi = minr; yi <= maxr; yi++)
            if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);
    }
    const float minZ = mb;
    const float minD = -3;
    const float maxD = mbf / minZ;
    std::vector<std::pair<int, int> > vDistIdx;
    for (int iL = 0; iL < N; iL++) {
        const orc_kp& kpL_ = kpL[iL];
        const int levelL = kpL_.octave;
        const float vL = kpL_.y, uL = kpL_.x;
        if (!(vL >= 0 && (size_t)vL < (size_t)nRows)) continue;
        const std::vector<size_t>& vCandidates = vRowIndices[(size_t)vL];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD;
        const float maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        size_t bestIdxR = 0;
        const u8* dL = descL + (size_t)iL * 32;
        for (size_t iC = 0; iC < vCandidates.size(); iC++) {
            const size_t iR = vCandidates[iC];
            if (kpR[iR].octave < levelL - 1 || kpR[iR].octave > levelL + 1) continue;
            const float uR = kpR[iR].x;
            if (uR >= minU && uR <= maxU) {
                const int dist = descriptor_distance(dL, descR + iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < TH_HIGH) {
            const float uR0 = kpR[bestIdxR].x;
            const float scaleFactor = invScaleFactors[levelL];
            const float scaleduL = std::round(kpL_.x * scaleFactor);
            const float scaledvL = std::round(kpL_.y * scaleFactor);
            const float scaleduR0 = std::round(uR0 * scaleFactor);
            const int wd = 5;
            const u8* IL = pyrL[levelL];
            const u8* IRimg = pyrR[levelL];
            const int cols = w[levelL];
            const int r0 = (int)(scaledvL - wd), c0 = (int)(scaleduL - wd);
            const float ILc = (float)IL[(size_t)(r0 + wd) * cols + c0 + wd];
            int bestDistS = INT_MAX, bestincR = 0;
            const int L = 5;
            float vDists[2 * 5 + 1];
            const float iniu = scaleduR0 + L - wd;
            const float endu = scaleduR0 + L + wd + 1;
            if (iniu < 0 || endu >= cols) continue;
            // (third definition of reference UB / cv::Mat range assertions: windows leaving the level image are skipped)
            if (r0 < 0 || r0 + 2 * wd + 1 > h[levelL] || c0 < 0 || c0 + 2 * wd + 1 > cols || (int)scaleduR0 - L - wd < 0) continue;
            for (int incR = -L; incR <= +L; incR++) {
                const int cr = (int)(scaleduR0 + incR - wd);
                const float IRc = (float)IRimg[(size_t)(r0 + wd) * cols + cr + wd];
                float dist = 0;                                     // cv::norm(IL, IR, NORM_L1): integer-valued terms, exact in float
                for (int y = 0; y < 2 * wd + 1; y++)
                    for (int x = 0; x < 2 * wd + 1; x++)
                        dist += std::fabs(((float)IL[(size_t)(r0 + y) * cols + c0 + x] - ILc) - ((float)IRimg[(size_t)(r0 + y) * cols + cr + x] - IRc));
                if (dist < bestDistS) { bestDistS = dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1];
            const float dist2 = vDists[L + bestincR];
            const float dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = scaleFactors[levelL] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= 0 && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
                mvDepth[iL] = mbf / disparity;
                mvuRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(bestDistS, iL));
            }
        }
    }
    if (vDistIdx.empty()) return;
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if (vDistIdx[i].first < thDist) break;
        mvuRight[vDistIdx[i].second] = -1;
        mvDepth[vDistIdx[i].second] = -1;
    }
}


// ---- independent window search: the candidate loops of SearchBySim3 (ORBmatcher.cc:1178-1213, 1258-1293) and Fuse (:899-950,
// 1054-1081).  With ur / inv_sigma2 the chi-square reprojection gate of Fuse (:913-944) applies per candidate.
void orc_search_windows_best(const orc_grid_view* Fp, int nq, const u8* active, const float* u, const float* v, const float* r,
                             const int* min_level, const int* max_level, const u8* desc, const float* ur, const float* inv_sigma2,
                             int th_dist, int* best_idx) {
    const orc_grid_view& F = *Fp;
    std::vector<int> cand;
    for (int q = 0; q < nq; q++) {
        best_idx[q] = -1;
        if (!active[q]) continue;
        features_in_area(F, u[q], v[q], r[q], min_level[q], max_level[q], cand);
        int bestDist = 256, bestIdx = -1;
        for (int idx : cand) {
            if (inv_sigma2) {
                const int kpLevel = F.octave[idx];
                const float kpx = F.x[idx], kpy = F.y[idx];
                const float kpr = F.uright ? F.uright[idx] : -1.0f;
                const float ex = u[q] - kpx, ey = v[q] - kpy;
                if (kpr >= 0) {
                    const float er = ur[q] - kpr;
                    const float e2 = ex * ex + ey * ey + er * er;
                    if (e2 * inv_sigma2[kpLevel] > 7.8) continue;
                } else {
                    const float e2 = ex * ex + ey * ey;
                    if (e2 * inv_sigma2[kpLevel] > 5.99) continue;
                }
            }
            const int dist = descriptor_distance(desc + (size_t)q * 32, F.desc + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestDist <= th_dist) best_idx[q] = bestIdx;
    }
}


// ---- Fuse (ORBmatcher.cc:828-972) and Fuse with a similarity (:974-1103): projection + candidate loop for every point that passes
// the entry test; best_idx[i] = bestIdx when bestDist <= TH_LOW, else -1.  The map updates (:952-971, 1083-1099) are host
// bookkeeping on MapPoint / KeyFrame objects and are replayed by the tests from these indices.
//   variant 0: T = [Rcw|tcw], Ow = GetCameraCenter(), invz = 1/z, chi-square gate with bf / inv_sigma2
//   variant 1: T = rows of Scw (decomposed here, :983-987), invz = 1.0/z, no gate
void orc_fuse_search(const orc_grid_view* Kp, int variant, const float* Tin, const float* Ow_in, float fx, float fy, float cx, float cy,
                     float bf, float logScaleFactor, int n, const u8* skip, const float* world, const float* mf_max, const float* mf_min,
                     const float* normal, const u8* desc, float th, const float* inv_sigma2, int* best_idx) {
    const orc_grid_view& KF = *Kp;
    float T[12], Ow[3];
    if (variant == 0) {
        for (int i = 0; i < 12; i++) T[i] = Tin[i];
        for (int i = 0; i < 3; i++) Ow[i] = Ow_in[i];
    } else {
        const float scw = (float)std::sqrt((double)Tin[0] * Tin[0] + (double)Tin[1] * Tin[1] + (double)Tin[2] * Tin[2]);
        const float inv = (float)(1.0 / (double)scw);
        for (int i = 0; i < 12; i++) T[i] = Tin[i] * inv;
        minus_rt_t(T, Ow);
    }
    std::vector<u8> active(n, 0);
    std::vector<float> u(n, 0.f), v(n, 0.f), rad(n, 0.f), ur(n, 0.f);
    std::vector<int> minL(n, -1), maxL(n, -1);
    for (int i = 0; i < n; i++) {
        if (skip[i]) continue;
        const float* p3Dw = world + 3 * (size_t)i;
        float p3Dc[3];
        rt_apply(T, p3Dw, p3Dc);
        if (p3Dc[2] < 0.0f) continue;
        float invz;
        if (variant == 0) invz = 1 / p3Dc[2]; else invz = 1.0 / p3Dc[2];
        const float x = p3Dc[0] * invz, y = p3Dc[1] * invz;
        const float uu = fx * x + cx, vv = fy * y + cy;
        if (!(uu >= KF.min_x && uu < KF.max_x && vv >= KF.min_y && vv < KF.max_y)) continue;
        ur[i] = uu - bf * invz;
        const float maxDistance = 1.2f * mf_max[i], minDistance = 0.8f * mf_min[i];
        const float PO[3] = {p3Dw[0] - Ow[0], p3Dw[1] - Ow[1], p3Dw[2] - Ow[2]};
        const float dist3D = norm3(PO);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const float* Pn = normal + 3 * (size_t)i;
        const double dot = (double)PO[0] * Pn[0] + (double)PO[1] * Pn[1] + (double)PO[2] * Pn[2];
        if (dot < 0.5 * dist3D) continue;
        const int nPredictedLevel = predict_scale(mf_max[i], dist3D, logScaleFactor, KF.n_levels);
        active[i] = 1; u[i] = uu; v[i] = vv;
        rad[i] = th * KF.scale_factors[nPredictedLevel];
        minL[i] = nPredictedLevel - 1; maxL[i] = nPredictedLevel;
    }
    orc_search_windows_best(Kp, n, active.data(), u.data(), v.data(), rad.data(), minL.data(), maxL.data(), desc,
                            variant == 0 ? ur.data() : nullptr, variant == 0 ? inv_sigma2 : nullptr, TH_LOW, best_idx);
}

// One direction of SearchBySim3 (ORBmatcher.cc:1150-1219 / 1222-1299): points seen from camera `from` (Rfw, tfw) moved into camera
// `to` by [sR|t], searched in `to`'s features.  valid[i] = pMP && !vbAlreadyMatched[i] && !isBad().
void orc_sim3_direction(const orc_grid_view* Tp, const float* Rfw, const float* tfw, const float* sR, const float* t, float fx, float fy,
                        float cx, float cy, float logScaleFactor, int n, const u8* valid, const float* world, const float* mf_max,
                        const float* mf_min, const u8* desc, float th, int* vnMatch) {
    const orc_grid_view& KF = *Tp;
    float T1[12], T2[12];
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) { T1[4 * r + c] = Rfw[3 * r + c]; T2[4 * r + c] = sR[3 * r + c]; } T1[4 * r + 3] = tfw[r]; T2[4 * r + 3] = t[r]; }
    std::vector<u8> active(n, 0);
    std::vector<float> u(n, 0.f), v(n, 0.f), rad(n, 0.f);
    std::vector<int> minL(n, -1), maxL(n, -1);
    for (int i = 0; i < n; i++) {
        if (!valid[i]) continue;
        float c1[3], c2[3];
        rt_apply(T1, world + 3 * (size_t)i, c1);
        rt_apply(T2, c1, c2);
        if (c2[2] < 0.0) continue;
        const float invz = 1.0 / c2[2];
        const float x = c2[0] * invz, y = c2[1] * invz;
        const float uu = fx * x + cx, vv = fy * y + cy;
        if (!(uu >= KF.min_x && uu < KF.max_x && vv >= KF.min_y && vv < KF.max_y)) continue;
        const float maxDistance = 1.2f * mf_max[i], minDistance = 0.8f * mf_min[i];
        const float dist3D = norm3(c2);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const int nPredictedLevel = predict_scale(mf_max[i], dist3D, logScaleFactor, KF.n_levels);
        active[i] = 1; u[i] = uu; v[i] = vv;
        rad[i] = th * KF.scale_factors[nPredictedLevel];
        minL[i] = nPredictedLevel - 1; maxL[i] = nPredictedLevel;
    }
    orc_search_windows_best(Tp, n, active.data(), u.data(), v.data(), rad.data(), minL.data(), maxL.data(), desc, nullptr, nullptr, TH_HIGH, vnMatch);
}

// Frame::GetFeaturesInArea on its own (the candidate list every window search starts from), for the comparison with the reference's
// own src/Frame.cc: returns the number of indices, out[0..cap) receives them in the reference's order.
int orc_features_in_area(const orc_grid_view* F, float x, float y, float r, int min_level, int max_level, int* out, int cap) {
    std::vector<int> v;
    features_in_area(*F, x, y, r, min_level, max_level, v);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = v[i];
    return (int)v.size();
}

}  // extern "C"
