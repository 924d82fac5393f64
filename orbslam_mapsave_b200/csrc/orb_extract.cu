// orb_extract.cu — B200 (sm_100a) implementation of ORB_SLAM2::ORBextractor's hot path behind the C ABI of
// include/orb_b200.h.  Replaces src/ORBextractor.cc:409-478 (ctor, tables), :1110-1135 (ComputePyramid),
// :764-852 (ComputeKeyPointsOctTree, cv::FAST per cell with the iniThFAST/minThFAST retry), :538-762
// (DistributeOctTree), :76-103 (IC_Angle), :1088-1089 (GaussianBlur) and :107-146 (computeOrbDescriptor) of the
// reference.  Designed batch-first: every kernel runs over (work item, frame) so that a pass over B frames is ~20
// launches; nothing here is a translation of the reference's serial loops.
//
// HBM layout per frame (all u8):  [level 0 buffer][level 1 buffer]...   each buffer = (h+38) rows x pitch bytes, the
// level's pixel (x,y) at row y+19, column x+32 (32 keeps 4-pixel groups word-aligned; columns [13,32) and
// [32+w, 32+w+19) hold the REFLECT_101 border the reference's copyMakeBorder produces).  The blurred pyramid uses
// the same geometry in a second block.
#include "orb_common.cuh"

#include <cuda.h>            // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <type_traits>
#include <vector>

static thread_local char g_err[512] = "";
void orb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* orb_last_error(void) { return g_err; }
extern "C" int orb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

#define ORBX_MAX_LEVELS 16
#define ORBX_OX 32            // column of level pixel x=0 inside a level buffer
#define ORBX_OY 19            // row of level pixel y=0  (EDGE_THRESHOLD, ORBextractor.cc:73)
#define ORBX_EDGE 19
#define ORBX_MINB 16          // minBorder = EDGE_THRESHOLD-3 (ORBextractor.cc:772)
#define ORBX_FAST_THREADS 128
#define ORBX_FAST_LOCAL_CAP 1024
#define ORBX_OCT_THREADS 1024
#define OCT_U 4

struct LevelPlan {
    int w, h, pitch, brows;
    unsigned off;                         // byte offset of the level buffer inside a frame block
    int nCols, nRows, wCell, hCell, cellBase;
    int maxBX, maxBY;                     // w-16, h-16
    int quota, nIni;
    float hX;
    int candOff, candCap, selOff, selCap;
    float scale, kpSize;
    int xtabOff, ytabOff;
    int fastResize;                       // 1: k_resize (dp2a form) applies, 0: k_resize_generic
};
struct Plan {
    int nlevels, iniTh, minTh;
    int totalCells, candTotal, selTotal, maxNodes, sortN;
    int tilePitch, tileRows, scorePitch, scoreRows;   // FAST shared-memory tile geometry (max over levels)
    int fwBoxW, fwBoxH, fwTileBytes, fwScoreOff, fwPlistOff, fwGlistOff, fwBarOff, fwStride;   // k_fast_tma per-warp shared-memory layout
    int fwScorePitch;                     // score-map pitch of k_fast_tma: 40 or 48 (compile-time variants) or the generic pitch
    int width, height;
    unsigned long long frameBytes;
    int umax[16];
    int blurTileBase[ORBX_MAX_LEVELS + 1];
    LevelPlan lv[ORBX_MAX_LEVELS];
};
struct ResizeTap { int s; short c0, c1; };

// =====================================================================================================
// K1  pyramid.  Level 0 = masked copy + REFLECT_101 border; level l = fixed-point bilinear resize of level l-1
// (cv::resize INTER_LINEAR 8UC1: 11-bit coefficients, the (>>4, *b, >>16, +2, >>2) vertical pass) + border, fused:
// border pixels are evaluated at their reflected coordinate, so each level is produced in one pass.
// One thread = 4 horizontally adjacent buffer bytes -> one 32-bit store.
// =====================================================================================================
__global__ void __launch_bounds__(256) k_level0(const __grid_constant__ Plan P, const u8* __restrict__ images,
                                                const u8* __restrict__ masks, u8* __restrict__ pyr) {
    // ROI of level 0 = input (zeroed where mask == 0, ORBextractor.cc:1048-1053); the border is filled by k_border
    const LevelPlan& L = P.lv[0];
    const int x = (blockIdx.x * 64 + threadIdx.x) * 4, y = blockIdx.y * 4 + threadIdx.y, f = blockIdx.z;
    if (x >= L.w || y >= L.h) return;
    const size_t fo = (size_t)f * P.width * P.height + (size_t)y * P.width;
    const u8* src = images + fo;
    const u8* msk = masks ? masks + fo : nullptr;
    u32 out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int xx = min(x + k, L.w - 1);
        u32 v = __ldg(src + xx);
        if (msk && __ldg(msk + xx) == 0) v = 0;
        out |= v << (8 * k);
    }
    *reinterpret_cast<u32*>(pyr + (size_t)f * P.frameBytes + L.off + (size_t)(y + ORBX_OY) * L.pitch + x + ORBX_OX) = out;
}

// 16-pixel groups with 128-bit loads / stores when the frame width and base pointers are 16-byte aligned (VGA, 720p, 4K ...).  A pure
// copy is bound by bytes in flight: flat indexing over the groups of a frame (no idle lanes at the row end: 640 / 16 = 40 groups do not
// fill 64-lane rows) and FOUR independent loads per thread before the first store.
#define L0_UNROLL 4
__global__ void __launch_bounds__(256) k_level0_v16(const __grid_constant__ Plan P, const u8* __restrict__ images,
                                                    const u8* __restrict__ masks, u8* __restrict__ pyr, u32 gprInv) {
    const LevelPlan& L = P.lv[0];
    const u32 gpr = (u32)L.w >> 4, total = gpr * (u32)L.h;              // groups per row / per frame
    const int f = blockIdx.y;
    const size_t fo = (size_t)f * P.width * P.height;
    u8* dst0 = pyr + (size_t)f * P.frameBytes + L.off + (size_t)ORBX_OY * L.pitch + ORBX_OX;
    const u32 g0 = blockIdx.x * (256 * L0_UNROLL) + threadIdx.x;
    uint4 v[L0_UNROLL], m[L0_UNROLL];
#pragma unroll
    for (int k = 0; k < L0_UNROLL; k++) {
        const u32 g = g0 + k * 256;
        if (g < total) {
            v[k] = __ldg(reinterpret_cast<const uint4*>(images + fo) + g);
            if (masks) m[k] = __ldg(reinterpret_cast<const uint4*>(masks + fo) + g);
        }
    }
#pragma unroll
    for (int k = 0; k < L0_UNROLL; k++) {
        const u32 g = g0 + k * 256;
        if (g >= total) continue;
        if (masks) {
            // per byte: keep the pixel where the mask byte is non-zero (0x80 trick: (m | (m & 0x7f..) + 0x7f..) & 0x80.. marks non-zero bytes)
            auto keep = [](u32 px, u32 mk) {
                const u32 nz = ((mk & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | mk;       // bit 7 of each byte set iff that mask byte != 0
                const u32 full = ((nz & 0x80808080u) >> 7) * 0xFFu;            // 0xFF in the bytes to keep
                return px & full;
            };
            v[k].x = keep(v[k].x, m[k].x); v[k].y = keep(v[k].y, m[k].y); v[k].z = keep(v[k].z, m[k].z); v[k].w = keep(v[k].w, m[k].w);
        }
        const u32 y = __umulhi(g, gprInv), xg = g - y * gpr;          // exact: g * gpr < 2^32 (checked by the caller)
        *reinterpret_cast<uint4*>(dst0 + (size_t)y * L.pitch + 16 * xg) = v[k];
    }
}

// Generic gather form (any scale factor): one thread = 4 adjacent ROI pixels, 4 byte gathers each.
__global__ void __launch_bounds__(256) k_resize_generic(const __grid_constant__ Plan P, int level, u8* __restrict__ pyr,
                                                        const ResizeTap* __restrict__ xtab, const ResizeTap* __restrict__ ytab) {
    const LevelPlan& L = P.lv[level];
    const LevelPlan& S = P.lv[level - 1];
    const int x = (blockIdx.x * 64 + threadIdx.x) * 4, y = blockIdx.y * 4 + threadIdx.y, f = blockIdx.z;
    if (x >= L.w || y >= L.h) return;
    u8* frame = pyr + (size_t)f * P.frameBytes;
    const u8* src = frame + S.off + (size_t)ORBX_OY * S.pitch + ORBX_OX;
    const ResizeTap ty = ytab[L.ytabOff + y];
    const int sy0 = min(max(ty.s, 0), S.h - 1), sy1 = min(max(ty.s + 1, 0), S.h - 1);
    const u8* r0p = src + (size_t)sy0 * S.pitch;
    const u8* r1p = src + (size_t)sy1 * S.pitch;
    u32 out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const ResizeTap tx = xtab[L.xtabOff + min(x + k, L.w - 1)];
        const int sx = tx.s, sx1 = min(sx + 1, S.w - 1);
        const int r0 = r0p[sx] * tx.c0 + r0p[sx1] * tx.c1;
        const int r1 = r1p[sx] * tx.c0 + r1p[sx1] * tx.c1;
        int v = (((ty.c0 * (r0 >> 4)) >> 16) + ((ty.c1 * (r1 >> 4)) >> 16) + 2) >> 2;
        v = min(max(v, 0), 255);
        out |= (u32)v << (8 * k);
    }
    *reinterpret_cast<u32*>(frame + L.off + (size_t)(y + ORBX_OY) * L.pitch + x + ORBX_OX) = out;
}

// Fast form for scale factors <= 2.  A thread produces 4 adjacent pixels of TWO consecutive output rows: the tap tables, the
// PRMT selectors and all address arithmetic are shared by the two rows, and so is the horizontal pass of the source row the
// two outputs have in common (for a 1.2 pyramid the second output row starts where the first one ends 5 times out of 6).
// Horizontal pass of a source row: 3 aligned words -> the 8 bytes starting at the first output's tap (two funnel shifts);
// output k picks its two taps (bytes d, d+1 with d = sx[k] - sx[0] <= 6, checked on the host) with one PRMT and combines them
// with their 11-bit coefficients by one IDP2A.  Vertical pass: the reference's exact (>>4, *b, >>16, +2, >>2) sequence.
__global__ void __launch_bounds__(256) k_resize(const __grid_constant__ Plan P, int level, u8* __restrict__ pyr,
                                                const ResizeTap* __restrict__ xtab, const ResizeTap* __restrict__ ytab, u32 wqInv) {
    const LevelPlan& L = P.lv[level];
    const LevelPlan& S = P.lv[level - 1];
    // flat index over (row pair, group of 4 pixels): no idle lanes at the right edge of a level (widths are not multiples of 128)
    const u32 wq = (u32)(L.w + 3) >> 2, idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= wq * (u32)((L.h + 1) >> 1)) return;
    const int yp = (int)__umulhi(idx, wqInv), x = (int)(idx - (u32)yp * wq) * 4, y = 2 * yp, f = blockIdx.y;
    const bool two = y + 1 < L.h;
    u8* frame = pyr + (size_t)f * P.frameBytes;
    const u8* src = frame + S.off + (size_t)ORBX_OY * S.pitch + ORBX_OX;
    const ResizeTap ty0 = ytab[L.ytabOff + y], ty1 = ytab[L.ytabOff + (two ? y + 1 : y)];
    const int sA0 = min(max(ty0.s, 0), S.h - 1), sA1 = min(max(ty0.s + 1, 0), S.h - 1);
    const int sB0 = min(max(ty1.s, 0), S.h - 1), sB1 = min(max(ty1.s + 1, 0), S.h - 1);
    // x taps of the 4 outputs: xtab rows are padded to a multiple of 4 entries, so two 16-byte loads
    const uint4* tp = reinterpret_cast<const uint4*>(xtab + L.xtabOff + x);
    const uint4 t01 = __ldg(tp), t23 = __ldg(tp + 1);
    const int sx0 = (int)t01.x;
    const u32 cf[4] = {t01.y, t01.w, t23.y, t23.w};
    const u32 sel[4] = {0x10u, (u32)((int)t01.z - sx0) * 0x11u + 0x10u, (u32)((int)t23.x - sx0) * 0x11u + 0x10u,
                        (u32)((int)t23.z - sx0) * 0x11u + 0x10u};
    const int base = sx0 & ~3, sh0 = (sx0 - base) * 8;
    auto hpass = [&](int sy, int (&h)[4]) {                          // h[k] = (p0*c0 + p1*c1) >> 4 for the 4 outputs, source row sy
        // one 32-bit offset and one 64-bit add per source row; the three words are immediates off that address (the compiler
        // otherwise keeps three 64-bit column bases and adds the frame pointer per load: 9 instead of 3 address instructions per row)
        const u32* rp = reinterpret_cast<const u32*>(src + (u32)(sy * S.pitch + base));
        const u32 w0 = rp[0], w1 = rp[1], w2 = rp[2];
        const u32 lo = __funnelshift_r(w0, w1, sh0), hi = __funnelshift_r(w1, w2, sh0);
#pragma unroll
        for (int k = 0; k < 4; k++) h[k] = (int)__dp2a_lo(cf[k], __byte_perm(lo, hi, sel[k]), 0u) >> 4;
    };
    auto vpass = [&](const ResizeTap& t, const int (&h0)[4], const int (&h1)[4]) {
        u32 out = 0;                                                 // <= 255 by construction (coefficients >= 0 summing to 2048 per axis)
#pragma unroll
        for (int k = 0; k < 4; k++) out |= (u32)((((t.c0 * h0[k]) >> 16) + ((t.c1 * h1[k]) >> 16) + 2) >> 2) << (8 * k);
        return out;
    };
    int hA0[4], hA1[4], hB0[4], hB1[4];
    hpass(sA0, hA0);
    hpass(sA1, hA1);
    u8* dst = frame + L.off + (size_t)(y + ORBX_OY) * L.pitch + x + ORBX_OX;
    *reinterpret_cast<u32*>(dst) = vpass(ty0, hA0, hA1);
    if (two) {
        if (sB0 == sA1) {
#pragma unroll
            for (int k = 0; k < 4; k++) hB0[k] = hA1[k];
        } else hpass(sB0, hB0);
        if (sB1 == sA1) {                                            // (only at the clamped bottom rows)
#pragma unroll
            for (int k = 0; k < 4; k++) hB1[k] = hA1[k];
        } else hpass(sB1, hB1);
        *reinterpret_cast<u32*>(dst + L.pitch) = vpass(ty1, hB0, hB1);
    }
}

// REFLECT_101 border of every level (copyMakeBorder, ORBextractor.cc:1125-1132): `bw` columns left/right and `bh` rows
// above/below.  The hot path only needs the 3 pixels the 7x7 blur reads (bw = 4, bh = 3); the full 19-pixel border of the
// API-visible pyramid is produced on demand by orbx_get_pyramid_level(bordered = 1).
__global__ void __launch_bounds__(256) k_border(const __grid_constant__ Plan P, u8* __restrict__ pyr, int bw, int bh, int level0,
                                                int frame0) {
    const int level = level0 + blockIdx.y, f = frame0 + blockIdx.z;
    const LevelPlan& L = P.lv[level];
    const int side = L.h * 2 * bw, rowlen = L.w + 2 * bw, total = side + 2 * bh * rowlen;
    u8* img = pyr + (size_t)f * P.frameBytes + L.off + (size_t)ORBX_OY * L.pitch + ORBX_OX;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        int x, y;
        if (i < side) {
            y = i / (2 * bw);
            const int j = i - y * 2 * bw;
            x = j < bw ? j - bw : L.w + (j - bw);
        } else {
            const int i2 = i - side, r = i2 / rowlen;
            x = i2 - r * rowlen - bw;
            y = r < bh ? r - bh : L.h + (r - bh);
        }
        img[(ptrdiff_t)y * L.pitch + x] = img[(ptrdiff_t)dev_reflect101(y, L.h) * L.pitch + dev_reflect101(x, L.w)];
    }
}

// =====================================================================================================
// K2+K3  FAST-9/16 per grid cell: score map, in-cell 3x3 NMS, iniThFAST -> minThFAST retry, emission.
// One CTA per (cell, frame).  Uses the single-score-map formulation: the FAST score does not depend on
// the threshold, so "cv::FAST(cell, ini) or, if empty, cv::FAST(cell, min)" == NMS over scores computed at minTh,
// then keep score >= ini unless that set is empty.  Candidates carry (cell id, y, x) as their order key, which is
// the reference's vToDistributeKeys order (cell row-major, then FAST's y-then-x order).
// =====================================================================================================
// Packed formulation: two horizontally adjacent pixels per thread, every value a u16x2.  With the shared tile stored as
// u16 pixels in two copies (even-aligned pairs and pairs shifted by one pixel) every (ring pixel of x, ring pixel of x+1)
// pair is ONE aligned 32-bit shared load.  d' = 256 + v - ring is one 32-bit subtract (no borrow between halves), and
//   A = max over the 16 arcs of min(d over 9 contiguous)      B = max over arcs of min(-d) = -(min over arcs of max(d))
// are 2 x (16 + 16 + 8) three-input packed min/max (VIMNMX3.U16x2 on sm_100a).  score = max(A, B) - 1 and the pixel is a
// FAST-9 corner at threshold t iff score >= t — no segment test, no branches, identical to cv::FAST + cornerScore.
__device__ __forceinline__ u32 min3x2(u32 a, u32 b, u32 c) { return __vimin3_u16x2(a, b, c); }
__device__ __forceinline__ u32 max3x2(u32 a, u32 b, u32 c) { return __vimax3_u16x2(a, b, c); }

__device__ __forceinline__ void fast_score_pair(const u32 (&r)[16], u32 v, int& s0, int& s1) {
    const u32 vb = v + 0x01000100u;
    u32 d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = vb - r[k];
    u32 lo3[16], hi3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        lo3[k] = min3x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
        hi3[k] = max3x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    }
    u32 lo9[16], hi9[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        lo9[k] = min3x2(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]);
        hi9[k] = max3x2(hi3[k], hi3[(k + 3) & 15], hi3[(k + 6) & 15]);
    }
    u32 A = max3x2(lo9[0], lo9[1], lo9[2]), Bm = min3x2(hi9[0], hi9[1], hi9[2]);
#pragma unroll
    for (int k = 3; k < 15; k += 2) { A = max3x2(A, lo9[k], lo9[k + 1]); Bm = min3x2(Bm, hi9[k], hi9[k + 1]); }
    A = __vmaxu2(A, lo9[15]);
    Bm = __vminu2(Bm, hi9[15]);
    s0 = max((int)(A & 0xFFFF) - 256, 256 - (int)(Bm & 0xFFFF)) - 1;
    s1 = max((int)(A >> 16) - 256, 256 - (int)(Bm >> 16)) - 1;
}

// TPP / SPP > 0: compile-time tile / score pitches (ring loads become one base register + immediates);
// 0: runtime pitches from the plan (cells wider than 40 px, i.e. pyramid levels narrower than ~100 px).
//
// Structure of one CTA (= one valid grid cell of one frame; the cell list is precomputed on the host):
//   load    cell ROI -> shared tile of u16 pixels (aligned 32-bit global loads, PRMT expansion)
//   phase A every pixel pair: the necessary condition "for each of 3 opposite ring pairs (k, k+8) at least one member is
//           darker than v-t / brighter than v+t" on packed u16x2 values (20-30 % of the pixels of the synthetic frames pass); passing pixels are
//           compacted into a list (warp ballot + one shared atomic per warp iteration)
//   phase B full 16-arc score (fast_score_pair) for the listed pixels only, two arbitrary pixels per thread
//   NMS     3x3 strict maximum over the listed corners; iniThFAST -> minThFAST retry; emission
#define ORBX_FAST_TPP 28
#define ORBX_FAST_SPP 48
template <int TPP, int SPP>
__global__ void __launch_bounds__(ORBX_FAST_THREADS) k_fast(const __grid_constant__ Plan P, const uint4* __restrict__ cells,
                                                             const u8* __restrict__ pyr, uint2* __restrict__ cand,
                                                             int* __restrict__ candCount, int* __restrict__ status) {
    extern __shared__ __align__(16) u8 smem[];
    __shared__ u32 s_list[ORBX_FAST_LOCAL_CAP];
    __shared__ int s_nlist, s_nloc, s_nini, s_base, s_emit;

    const uint4 ce = __ldg(cells + blockIdx.x);
    const int iniX = (int)(ce.x & 0xFFFF), iniY = (int)(ce.x >> 16);
    const int tw = (int)(ce.y & 0xFF), th = (int)((ce.y >> 8) & 0xFF), l = (int)((ce.y >> 16) & 0xFF), c = (int)ce.z;
    const int dw = tw - 6, dh = th - 6, f = blockIdx.y;
    const LevelPlan& L = P.lv[l];

    const int TPp = TPP > 0 ? TPP : P.tilePitch, SP = SPP > 0 ? SPP : P.scorePitch;   // tile pitch in pixel PAIRS (u32), score pitch in bytes
    u32* tile = reinterpret_cast<u32*>(smem);                         // pair j of a row = pixels (2j, 2j+1) as u16 each
    u8* score = reinterpret_cast<u8*>(tile + (size_t)P.tileRows * TPp);
    u16* plist = reinterpret_cast<u16*>(score + (size_t)P.scoreRows * SP);
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) { s_nlist = 0; s_nloc = 0; s_nini = 0; s_emit = 0; }

    // ---- load: 16 lanes walk the words of a row, 8 rows per pass
    const int bx0 = iniX + ORBX_OX, gx0 = bx0 & ~3, lead = bx0 - gx0;
    {
        const int words = (lead + tw + 3) >> 2;
        const int pitchW = L.pitch >> 2;
        const u32* src = reinterpret_cast<const u32*>(pyr + (size_t)f * P.frameBytes + L.off + (size_t)(iniY + ORBX_OY) * L.pitch + gx0);
        const int wx = tid & 15;
        for (int r = tid >> 4; r < th; r += ORBX_FAST_THREADS / 16)
            for (int wi = wx; wi < words; wi += 16) {
                const u32 w = __ldg(src + r * pitchW + wi);
                *reinterpret_cast<uint2*>(tile + r * TPp + 2 * wi) = make_uint2(__byte_perm(w, 0, 0x4140), __byte_perm(w, 0, 0x4342));
            }
        // score map: pixel (px, py) of the domain lives at byte (py + 1) * SP + px + 2; everything else stays 0
        for (int i = tid; i < (((dh + 2) * SP) >> 2); i += ORBX_FAST_THREADS) reinterpret_cast<u32*>(score)[i] = 0;
    }
    __syncthreads();

    const int t = P.minTh;
    const int X0 = lead + 3;                                          // tile x of domain pixel 0
    // ---- phase A
    {
        const int off = X0 & 1, npr = (off + dw + 1) >> 1;
        const u32 inv = 0xFFFFFFFFu / (u32)npr + 1;
        const u32* cE = tile + 3 * TPp + (X0 >> 1);
        const int hiT = 256 + t, loT = 256 - t;
        const int ntask = npr * dh;
        for (int task0 = 0; task0 < ntask; task0 += ORBX_FAST_THREADS) {
            const int task = task0 + tid;
            bool pass0 = false, pass1 = false;
            int row = 0, px0 = 0;
            if (task < ntask) {
                row = npr == 1 ? task : (int)__umulhi((u32)task, inv);
                const int p = task - row * npr;
                px0 = 2 * p - off;
                const u32* pe = cE + row * TPp + p;
                const u32 vb = pe[0] + 0x01000100u;
                const u32 d0 = vb - pe[3 * TPp], d8 = vb - pe[-3 * TPp];
                const u32 d2 = vb - pe[2 * TPp + 1], d10 = vb - pe[-2 * TPp - 1];
                const u32 d6 = vb - pe[-2 * TPp + 1], d14 = vb - pe[2 * TPp - 1];
                const u32 md = min3x2(__vmaxu2(d0, d8), __vmaxu2(d2, d10), __vmaxu2(d6, d14));    // dark arc possible iff > 256+t
                const u32 mb = max3x2(__vminu2(d0, d8), __vminu2(d2, d10), __vminu2(d6, d14));    // bright arc possible iff < 256-t
                pass0 = ((int)(md & 0xFFFF) > hiT || (int)(mb & 0xFFFF) < loT) && px0 >= 0;
                pass1 = ((int)(md >> 16) > hiT || (int)(mb >> 16) < loT) && px0 + 1 < dw;
            }
            const u32 b0 = __ballot_sync(0xffffffffu, pass0), b1 = __ballot_sync(0xffffffffu, pass1);
            const int n0 = __popc(b0), n = n0 + __popc(b1);
            if (n) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_nlist, n);
                base = __shfl_sync(0xffffffffu, base, 0);
                const u32 lt = (1u << lane) - 1;
                if (pass0) plist[base + __popc(b0 & lt)] = (u16)((row << 8) | px0);
                if (pass1) plist[base + n0 + __popc(b1 & lt)] = (u16)((row << 8) | (px0 + 1));
            }
        }
    }
    __syncthreads();

    // ---- phase B: exact score of the listed pixels, two per thread (halves of the u16x2 lanes)
    const int nlist = s_nlist;
    {
        const u16* t16 = reinterpret_cast<const u16*>(tile) + 3 * (2 * TPp) + X0;
        const int RP = 2 * TPp;                                       // row pitch in u16 pixels
#define RING(dx, dy) ((u32)q0[(dy) * RP + (dx)] | ((u32)q1[(dy) * RP + (dx)] << 16))
        for (int i = tid; 2 * i < nlist; i += ORBX_FAST_THREADS) {
            const int e0 = plist[2 * i], e1 = plist[min(2 * i + 1, nlist - 1)];
            const u16* q0 = t16 + (e0 >> 8) * RP + (e0 & 0xFF);
            const u16* q1 = t16 + (e1 >> 8) * RP + (e1 & 0xFF);
            u32 r[16];
            r[0] = RING(0, 3); r[1] = RING(1, 3); r[2] = RING(2, 2); r[3] = RING(3, 1);
            r[4] = RING(3, 0); r[5] = RING(3, -1); r[6] = RING(2, -2); r[7] = RING(1, -3);
            r[8] = RING(0, -3); r[9] = RING(-1, -3); r[10] = RING(-2, -2); r[11] = RING(-3, -1);
            r[12] = RING(-3, 0); r[13] = RING(-3, 1); r[14] = RING(-2, 2); r[15] = RING(-1, 3);
            const u32 v = RING(0, 0);
            int s0, s1;
            fast_score_pair(r, v, s0, s1);
            if (s0 >= t) score[((e0 >> 8) + 1) * SP + (e0 & 0xFF) + 2] = (u8)s0;
            if (s1 >= t) score[((e1 >> 8) + 1) * SP + (e1 & 0xFF) + 2] = (u8)s1;
        }
#undef RING
    }
    __syncthreads();

    // ---- 3x3 non-maximum suppression over the listed pixels (neighbours outside the cell's domain count as 0)
    for (int e = tid; e < nlist; e += ORBX_FAST_THREADS) {
        const int py = plist[e] >> 8, px = plist[e] & 0xFF;
        const u8* sp = score + (py + 1) * SP + px + 2;
        const int s = sp[0];
        if (s == 0) continue;
        const bool keep = s > sp[-1] && s > sp[1] && s > sp[-SP - 1] && s > sp[-SP] && s > sp[-SP + 1] &&
                          s > sp[SP - 1] && s > sp[SP] && s > sp[SP + 1];
        if (!keep) continue;
        const int slot = atomicAdd(&s_nloc, 1);
        if (slot < ORBX_FAST_LOCAL_CAP) s_list[slot] = ((u32)s << 16) | ((u32)py << 8) | (u32)px;
        if (s >= P.iniTh) atomicAdd(&s_nini, 1);
    }
    __syncthreads();
    const int nloc = min(s_nloc, ORBX_FAST_LOCAL_CAP);
    if (nloc == 0) return;
    const int T = s_nini > 0 ? P.iniTh : P.minTh;                     // the iniThFAST -> minThFAST retry (:811-815)
    const int nEmit = s_nini > 0 ? s_nini : nloc;
    if (tid == 0) {
        const int base = atomicAdd(&candCount[f * P.nlevels + l], nEmit);
        if (base + nEmit > L.candCap) atomicOr(status, ORB_DEV_CAND_OVERFLOW);
        s_base = base;
    }
    __syncthreads();
    uint2* out = cand + (size_t)f * P.candTotal + L.candOff;
    for (int e = tid; e < nloc; e += ORBX_FAST_THREADS) {
        const u32 k = s_list[e];
        const int s = k >> 16;
        if (s < T) continue;
        const int slot = s_base + atomicAdd(&s_emit, 1);
        if (slot >= L.candCap) continue;
        const int x = iniX + 3 + (int)(k & 0xFF), y = iniY + 3 + (int)((k >> 8) & 0xFF);
        out[slot] = make_uint2((u32)x | ((u32)y << 16), ((u32)s << 24) | (u32)c);
    }
}

// ---------------------------------------------------------------------------------------------------
// k_fast_tma: the same per-cell algorithm with ONE WARP per cell and a persistent grid.  Each warp streams its cells through
// two shared-memory tiles filled by TMA (cp.async.bulk.tensor.3d on a per-level tensor map of the pyramid, completion on an
// mbarrier), so the load of the next cell overlaps the work on the current one and costs no per-thread instructions.
// Nothing inside a cell needs a CTA barrier or a shared atomic: compaction is ballot + popc on a warp-uniform counter.
// ---------------------------------------------------------------------------------------------------
#define ORBX_FW_MAXWARPS 32          // warps per CTA are chosen per launch (blockDim.x / 32): the warps of a CTA never meet at a barrier,
                                     // so ONE CTA per SM with as many warps as its shared memory holds gives the highest occupancy
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u32 bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    u32 done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(u32 dst, const CUtensorMap* map, int x, int y, int z, u32 bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}

// explicit shared-space accesses from a 32-bit shared address (generic pointers make the compiler rebuild the shared::cluster window
// base inside the loop); OFF is an immediate
template <int OFF> __device__ __forceinline__ u32 lds32i(u32 addr) {
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF> __device__ __forceinline__ u32 lds8i(u32 addr) {
    u32 v;
    asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ void sts8(u32 addr, u32 v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts32(u32 addr, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts16(u32 addr, u32 v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory"); }

// Phase A (the quick test) works on FOUR horizontally adjacent pixels per lane, one byte each: the tile is read as aligned 32-bit
// words, |v - ring| of the four pixels is ONE VABSDIFF4, and "differs by more than t" is the carry into bit 7 of
// ((d & 0x7f) + (127 - t)) | d  (exact for t <= 127; larger thresholds are tested at 127, which is still a necessary condition).
// A FAST-9 corner needs, for each of the 4 opposite ring pairs (0,8) (2,10) (4,12) (6,14), at least one member that differs from the
// centre by more than t (any 9 contiguous ring pixels contain one member of every opposite pair).  On the synthetic frames this
// sign-agnostic 4-pair test passes 13.4 % of the pixels at t = 20 (the signed 3-pair u16x2 test it replaces passed 20.1 % at twice
// the instructions per pixel; FAST-9 corners are 5.6 %): tools/fast_filter_rates.py.
template <int BOXW, int SPP>
__global__ void __launch_bounds__(32 * ORBX_FW_MAXWARPS) k_fast_tma(const __grid_constant__ Plan P, const CUtensorMap* __restrict__ maps,
                                                                  const uint4* __restrict__ cells, int nCells, int nf,
                                                                  uint2* __restrict__ cand, int* __restrict__ candCount,
                                                                  int* __restrict__ status, int* __restrict__ workCounter,
                                                                  unsigned long long cellsInv40) {
    extern __shared__ __align__(128) u8 smem_fw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int BW = BOXW > 0 ? BOXW : P.fwBoxW, SP = SPP > 0 ? SPP : P.fwScorePitch;
    u8* wbase = smem_fw + (size_t)warp * P.fwStride;
    u8* score = wbase + P.fwScoreOff;
    u16* plist = reinterpret_cast<u16*>(wbase + P.fwPlistOff);
    const u32 bar0 = smem_u32(wbase + P.fwBarOff), bar1 = bar0 + 8;
    const u32 boxBytes = (u32)(BW * P.fwBoxH);
    const int nItems = nCells * nf;
    const int wpc = blockDim.x >> 5, Wt = gridDim.x * wpc;
    int item = blockIdx.x * wpc + warp;
    if (item >= nItems) return;

    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // item -> (frame, cell): one 64-bit multiply by ceil(2^40 / nCells) (exact while item * nCells < 2^40) instead of an integer division
    auto split = [&](int it, int& f, int& ci) {
        f = (int)(((unsigned long long)(u32)it * cellsInv40) >> 40);
        ci = it - f * nCells;
    };
    auto issue = [&](int f, int ci) {          // lane 0 only
        const uint4 ce = __ldg(cells + ci);
        mbar_expect_tx(bar0, boxBytes);
        // TMA needs the box start 16-byte aligned in the innermost dimension: load from the aligned column, keep the lead
        tma_load_3d(smem_u32(wbase), maps + ((ce.y >> 16) & 0xFF), ((int)(ce.x & 0xFFFF) + ORBX_OX) & ~15, (int)(ce.x >> 16) + ORBX_OY, f, bar0);
    };
    int f, cidx;
    split(item, f, cidx);
    if (lane == 0) issue(f, cidx);
    u32 ph0 = 0;
    const u32 lt = (1u << lane) - 1;
    // the score map (pixel (px, py) at byte (py + 1) * SP + px + 2) is zeroed once; every cell clears the entries it wrote
    for (int i = lane; i < ((P.scoreRows * SP) >> 2); i += 32) reinterpret_cast<u32*>(score)[i] = 0;
    const u32 scoreS = smem_u32(score);
    // Cells differ in cost (texture, the minThFAST retry), so after its first, statically assigned cell a warp takes its work
    // from a global counter; the id of the next cell is fetched one cell ahead so that its TMA load can be issued early.
    int nextItem = 0;
    if (lane == 0) nextItem = Wt + atomicAdd(workCounter, 1);
    nextItem = __shfl_sync(0xffffffffu, nextItem, 0);

    for (; item < nItems;) {
        const uint4 ce = __ldg(cells + cidx);
        const int iniX = (int)(ce.x & 0xFFFF), iniY = (int)(ce.x >> 16);
        const int tw = (int)(ce.y & 0xFF), th = (int)((ce.y >> 8) & 0xFF), l = (int)((ce.y >> 16) & 0xFF), c = (int)ce.z;
        const int dw = tw - 6, dh = th - 6;
        mbar_wait(bar0, ph0);
        ph0 ^= 1;
        __syncwarp();
        const u8* tile = wbase;

        // Like the reference (ORBextractor.cc:807-815) the cell is first searched at iniThFAST and only if that leaves no keypoint
        // at minThFAST.  NMS among corners >= ini is unaffected by weaker neighbours (they can never block a stronger pixel), so the
        // ini pass is exact on its own, and it lets the cheap phase-A test reject ~87 % of the pixels instead of ~77 %.
        const int X0 = ((iniX + ORBX_OX) & 15) + 3;                  // domain pixel 0 sits at tile (X0, 3)
        // 4-pixel groups aligned to tile words: group g of a row covers tile x [g0 + 4g, g0 + 4g + 4), i.e. domain px [4g - offq, ...)
        const int g0 = X0 & ~3, offq = X0 - g0, ng = (offq + dw + 3) >> 2, ntask = ng * dh;
        const u32 tileS = smem_u32(wbase), plS = smem_u32(plist), glS = smem_u32(wbase + P.fwGlistOff);
        const u32 inv = ce.w;                                         // 2^32 / ng + 1, from the host
        int nKeep = 0, nc = 0, T = P.iniTh;
        for (int pass = 0; pass < 2; pass++) {
            const int t = pass ? P.minTh : P.iniTh;
            T = t;
            // ---- phase A
            int nl = 0, gh = 0, gn = 0;                               // pixel-list length; head / fill of the 64-entry group ring
            // expand up to 32 group entries of the ring (from its head) into the pixel list (order is irrelevant: every entry carries
            // its coordinates).  The ring keeps the group list at 256 bytes instead of one entry per group of the cell, which is what
            // lets 30 instead of 27 warps share an SM's shared memory.
            auto expand = [&](int cnt) {
                const u32 ge = lane < cnt ? lds32i<0>(glS + 4 * (u32)((gh + lane) & 63)) : 0u;
                const u32 fb = ge & 0x80808080u;
                int pos = __popc(fb), tot;
                // exclusive prefix sum of the per-lane pixel counts: SHFL.UP delivers the in-range predicate with the value
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int v;
                    asm volatile("{\n\t.reg .pred p;\n\tshfl.sync.up.b32 %0|p, %1, %2, 0, 0xffffffff;\n\t@p add.s32 %1, %1, %0;\n\t}"
                                 : "=r"(v), "+r"(pos) : "r"(o));
                }
                tot = __shfl_sync(0xffffffffu, pos, 31);
                pos = nl + pos - __popc(fb);
                const u32 px0 = (ge & 0x3F7Fu) - 4u;                   // row << 8 | lo  (as the u16 the list holds)
                if (fb & 0x80u) { sts16(plS + 2 * (u32)pos, px0); pos++; }
                if (fb & 0x8000u) { sts16(plS + 2 * (u32)pos, px0 + 1); pos++; }
                if (fb & 0x800000u) { sts16(plS + 2 * (u32)pos, px0 + 2); pos++; }
                if ((int)fb < 0) sts16(plS + 2 * (u32)pos, px0 + 3);
                nl += tot;
            };
            {
                // no `& 0x7f` before the add: a byte with d >= 129 + t carries into its left neighbour, whose test then reads
                // d >= t instead of d > t — weaker, hence still a necessary condition (its own bit 7 comes from `| d`)
                const u32 Kt = (u32)(127 - min(t, 127)) * 0x01010101u;
                const u32 cE = tileS + 3 * BW + g0;
                for (int task0 = 0; task0 < ntask; task0 += 32) {
                    const int task = task0 + lane;
                    u32 fl = 0;                                        // bit 7 of byte i: pixel i of the group passes
                    int ent = 0;
                    if (task < ntask) {
                        const int row = ng == 1 ? task : (int)__umulhi((u32)task, inv);   // (inv overflows for ng == 1)
                        const int g = task - row * ng;
                        const int lo = 4 * g - offq;                  // domain px of the group's first pixel (-3 .. dw-1)
                        ent = (row << 8) + lo;
                        const u32 pw = cE + (u32)(row * BW + 4 * g);
                        u32 cl, cc, cr, ul, uc, ur, dl, dc, dr, r0, r8;
                        if (BOXW > 0) {                               // compile-time pitch: one address register, immediate offsets
                            cl = lds32i<-4>(pw); cc = lds32i<0>(pw); cr = lds32i<4>(pw);
                            ul = lds32i<2 * BOXW - 4>(pw); uc = lds32i<2 * BOXW>(pw); ur = lds32i<2 * BOXW + 4>(pw);
                            dl = lds32i<-2 * BOXW - 4>(pw); dc = lds32i<-2 * BOXW>(pw); dr = lds32i<-2 * BOXW + 4>(pw);
                            r0 = lds32i<3 * BOXW>(pw); r8 = lds32i<-3 * BOXW>(pw);
                        } else {
                            const u32 pu = pw + 2 * BW, pd = pw - 2 * BW;
                            cl = lds32i<-4>(pw); cc = lds32i<0>(pw); cr = lds32i<4>(pw);
                            ul = lds32i<-4>(pu); uc = lds32i<0>(pu); ur = lds32i<4>(pu);
                            dl = lds32i<-4>(pd); dc = lds32i<0>(pd); dr = lds32i<4>(pd);
                            r0 = lds32i<0>(pw + 3 * BW); r8 = lds32i<0>(pw - 3 * BW);
                        }
                        const u32 r4 = __byte_perm(cc, cr, 0x6543), r12 = __byte_perm(cl, cc, 0x4321);
                        const u32 r2 = __byte_perm(uc, ur, 0x5432), r14 = __byte_perm(ul, uc, 0x5432);
                        const u32 r6 = __byte_perm(dc, dr, 0x5432), r10 = __byte_perm(dl, dc, 0x5432);
#define FAR(r, dv) const u32 dv = __vabsdiffu4(cc, r), dv##x = dv + Kt
                        FAR(r0, a0); FAR(r8, a8); FAR(r2, a2); FAR(r10, a10); FAR(r4, a4); FAR(r12, a12); FAR(r6, a6); FAR(r14, a14);
#undef FAR
                        // pair (k, k+8) has a far member: bit 7 of (x_k | d_k | x_k8 | d_k8); all four pairs: AND
                        const u32 p0 = a0x | a0 | a8x, p1 = a2x | a2 | a10x, p2 = a4x | a4 | a12x, p3 = a6x | a6 | a14x;
                        u32 m = 0x80808080u;                          // pixels of the group that lie inside the domain
                        if (g == 0) m = 0x80808080u << (8 * offq);
                        if (g == ng - 1) m &= 0x80808080u >> (8 * (4 * ng - offq - dw));
                        fl = (p0 | a8) & (p1 | a10) & (p2 | a12) & (p3 | a14) & m;
                    }
                    // compaction at GROUP level (one ballot per iteration): entry = row << 8 | (lo + 4) with the pass flags left in
                    // bits 7 / 15 / 23 / 31 (lo + 4 <= 67 and row <= 63 keep those bits free); expanded to pixels below
                    const u32 bg = __ballot_sync(0xffffffffu, fl != 0u);
                    if (bg == 0u) continue;                            // warp-uniform: flat neighbourhoods leave nothing to compact
                    if (fl) sts32(glS + 4 * (u32)((gh + gn + __popc(bg & lt)) & 63), (u32)(ent + 4) | fl);
                    gn += __popc(bg);
                    if (gn >= 32) {                                    // warp-uniform
                        __syncwarp();
                        expand(32);
                        gh = (gh + 32) & 63; gn -= 32;
                        __syncwarp();
                    }
                }
                __syncwarp();
                if (gn > 0) expand(gn);
            }
            __syncwarp();

            // ---- phase B: exact score of the listed pixels, ONE pixel per lane with d = v - ring in the low half and -d in the
            // high half of every register (x = K + ring * 0xFFFF, K = (256 + v) | (256 - v) << 16), so that
            //   A = max over arcs of min9(d)   and   B = max over arcs of min9(-d)
            // come out of one min3/max3 sequence (40 VIMNMX3.U16x2 per pixel).  Corners are compacted in place at the list front.
            nc = 0;
            {
                const u8* t8 = tile + 3 * BW + X0;
                for (int e0 = 0; e0 < nl; e0 += 32) {
                    const int e = e0 + lane;
                    int ent = 0, sc = 0;
                    if (e < nl) {
                        ent = plist[e];
                        const u8* q = t8 + (ent >> 8) * BW + (ent & 0xFF);
                        const u32 v = q[0];
                        const u32 K = (256u + v) | ((256u - v) << 16);
#define RING(dx, dy) (K + (u32)q[(dy) * BW + (dx)] * 0xFFFFu)
                        u32 d[16];
                        d[0] = RING(0, 3); d[1] = RING(1, 3); d[2] = RING(2, 2); d[3] = RING(3, 1);
                        d[4] = RING(3, 0); d[5] = RING(3, -1); d[6] = RING(2, -2); d[7] = RING(1, -3);
                        d[8] = RING(0, -3); d[9] = RING(-1, -3); d[10] = RING(-2, -2); d[11] = RING(-3, -1);
                        d[12] = RING(-3, 0); d[13] = RING(-3, 1); d[14] = RING(-2, 2); d[15] = RING(-1, 3);
#undef RING
                        u32 m3[16], m9[16];
#pragma unroll
                        for (int k = 0; k < 16; k++) m3[k] = min3x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
#pragma unroll
                        for (int k = 0; k < 16; k++) m9[k] = min3x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
                        u32 A = max3x2(m9[0], m9[1], m9[2]);
#pragma unroll
                        for (int k = 3; k < 15; k += 2) A = max3x2(A, m9[k], m9[k + 1]);
                        A = __vmaxu2(A, m9[15]);
                        sc = (int)max(A & 0xFFFF, A >> 16) - 257;    // max(A, B) - 1
                        if (sc < t) sc = 0;
                    }
                    __syncwarp();                                      // everyone has read its entry before the in-place compaction
                    const u32 bm = __ballot_sync(0xffffffffu, sc > 0);
                    if (sc > 0) {
                        score[((ent >> 8) + 1) * SP + (ent & 0xFF) + 2] = (u8)sc;
                        plist[nc + __popc(bm & lt)] = (u16)ent;
                    }
                    nc += __popc(bm);
                }
            }
            __syncwarp();

            // ---- 3x3 NMS over the corners, branch-free (8 neighbours -> one maximum -> one compare); the survivors go to the group
            // list's buffer (free by now) so that plist[0, nc) keeps the corners for the clean-up of the score map
            nKeep = 0;
            for (int e0 = 0; e0 < nc; e0 += 32) {
                const int e = e0 + lane;
                bool keep = false;
                int ent = 0;
                if (e < nc) {
                    ent = plist[e];
                    const u32 sp = scoreS + (u32)(((ent >> 8) + 1) * SP + (ent & 0xFF) + 2);
                    u32 s, n0, n1, n2, n3, n4, n5, n6, n7;
                    if (SPP > 0) {
                        s = lds8i<0>(sp); n0 = lds8i<-1>(sp); n1 = lds8i<1>(sp);
                        n2 = lds8i<-SPP - 1>(sp); n3 = lds8i<-SPP>(sp); n4 = lds8i<-SPP + 1>(sp);
                        n5 = lds8i<SPP - 1>(sp); n6 = lds8i<SPP>(sp); n7 = lds8i<SPP + 1>(sp);
                    } else {
                        const u32 su = sp - SP, sd = sp + SP;
                        s = lds8i<0>(sp); n0 = lds8i<-1>(sp); n1 = lds8i<1>(sp);
                        n2 = lds8i<-1>(su); n3 = lds8i<0>(su); n4 = lds8i<1>(su);
                        n5 = lds8i<-1>(sd); n6 = lds8i<0>(sd); n7 = lds8i<1>(sd);
                    }
                    keep = s > __vimax3_u32(__vimax3_u32(n0, n1, n2), __vimax3_u32(n3, n4, n5), max(n6, n7));
                }
                const u32 bk = __ballot_sync(0xffffffffu, keep);
                if (keep) sts16(glS + 2 * (u32)(nKeep + __popc(bk & lt)), (u32)ent);
                nKeep += __popc(bk);
            }
            if (nKeep > 0 || P.minTh >= P.iniTh) break;               // :811 `if(vKeysCell.empty())` -> retry at minThFAST
            __syncwarp();
        }
        __syncwarp();
        // the tile is no longer needed: fetch the next cell's tile now, it lands while emission / score clearing run
        const int fCur = f;                                           // (l, c, iniX ... of the cell in hand are locals)
        item = nextItem;
        if (item < nItems) {
            split(item, f, cidx);
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // order our generic reads before the async-proxy write
                issue(f, cidx);
                nextItem = Wt + atomicAdd(workCounter, 1);
            }
        }
        nextItem = __shfl_sync(0xffffffffu, nextItem, 0);
        if (nKeep) {                                                  // every survivor of the pass that produced them is emitted (score >= T)
            const LevelPlan& L = P.lv[l];
            int base = 0;
            if (lane == 0) {
                base = atomicAdd(&candCount[fCur * P.nlevels + l], nKeep);
                if (base + nKeep > L.candCap) atomicOr(status, ORB_DEV_CAND_OVERFLOW);
            }
            base = __shfl_sync(0xffffffffu, base, 0);
            uint2* out = cand + (size_t)fCur * P.candTotal + L.candOff;
            const u16* kept = reinterpret_cast<const u16*>(wbase + P.fwGlistOff);
            for (int e = lane; e < nKeep; e += 32) {
                const int ent = kept[e], py = ent >> 8, px = ent & 0xFF;
                const int s = score[(py + 1) * SP + px + 2];
                if (base + e < L.candCap && s >= T)
                    out[base + e] = make_uint2((u32)(iniX + 3 + px) | ((u32)(iniY + 3 + py) << 16), ((u32)s << 24) | (u32)c);
            }
        }
        __syncwarp();
        // leave the score map clean: the corners of the last pass are a superset of every entry this cell wrote (a minThFAST retry
        // re-scores the iniThFAST corners)
        for (int e = lane; e < nc; e += 32) {
            const int ent = plist[e];
            sts8(scoreS + (u32)(((ent >> 8) + 1) * SP + (ent & 0xFF) + 2), 0u);
        }
        __syncwarp();
    }
}

// =====================================================================================================
// K4  DistributeOctTree, one CTA per (level, frame).  The reference's std::list walk is restated as data-parallel
// passes: node membership is a pure function of the root box and (x,y), list order after a pass is
// [children in reverse creation order][untouched nodes in their old order], and the sequential "expand the largest
// node until there are N" tail is a sort + prefix sum + cut (see DESIGN.md for the proof sketch).
// Tie-break among equal-size nodes = creation sequence (the reference's is the heap pointer; canonical rule).
// =====================================================================================================
struct OctSmem {
    int4* boxA; int4* boxB;
    int* cntA; int* cntB;
    int* cc;        // 4 per node: child counts, then child list positions
    int* aux;       // per node scratch (nch / creation prefix)
    int* aux2;      // per node scratch (stay rank / sorted nch)
    int* div;       // per node: divides in this step
    u64* sortb;     // sortN entries
};

__device__ int block_excl_scan(int* a, int n, int* wsum) {
    // in-place exclusive prefix sum of a[0..n) by the whole CTA; returns the total.  wsum: 34 ints of smem.
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int K = (n + T - 1) / T;
    const int beg = min(tid * K, n), end = min(beg + K, n);
    int s = 0;
    for (int i = beg; i < end; i++) s += a[i];
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        const int ws = lane < (T >> 5) ? wsum[lane] : 0;
        int wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
        wsum[lane] = wi - ws;
        if (lane == 31) wsum[32] = wi;
    }
    __syncthreads();
    int base = wsum[w] + inc - s;
    for (int i = beg; i < end; i++) { const int v = a[i]; a[i] = base; base += v; }
    const int total = wsum[32];
    __syncthreads();
    return total;
}

__device__ void block_bitonic_sort(u64* a, int n /*pow2*/) {
    for (int k = 2; k <= n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const u64 x = a[i], y = a[p];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[p] = x; }
                }
            }
            __syncthreads();
        }
}

__global__ void __launch_bounds__(ORBX_OCT_THREADS) k_octree(const __grid_constant__ Plan P, const uint2* __restrict__ cand,
                                                              const int* __restrict__ candCount, u32* __restrict__ nodeOf,
                                                              uint2* __restrict__ sel, int* __restrict__ selCount,
                                                              int* __restrict__ status) {
    extern __shared__ __align__(16) u8 smem[];
    __shared__ int s_wsum[34];
    __shared__ int s_nToExpand, s_cut, s_rootMap[64];
    const int l = blockIdx.x, f = blockIdx.y, tid = threadIdx.x, T = blockDim.x;
    const LevelPlan& L = P.lv[l];
    const int M = P.maxNodes, N = L.quota;
    const int n = min(candCount[f * P.nlevels + l], L.candCap);
    const uint2* keys = cand + (size_t)f * P.candTotal + L.candOff;
    u32* nof = nodeOf + (size_t)f * P.candTotal + L.candOff;
    uint2* out = sel + (size_t)f * P.selTotal + L.selOff;

    OctSmem S;
    {
        u8* p = smem;
        S.boxA = (int4*)p; p += sizeof(int4) * M;
        S.boxB = (int4*)p; p += sizeof(int4) * M;
        S.sortb = (u64*)p; p += sizeof(u64) * P.sortN;
        S.cntA = (int*)p; p += 4 * M;
        S.cntB = (int*)p; p += 4 * M;
        S.cc = (int*)p; p += 16 * M;
        S.aux = (int*)p; p += 4 * M;
        S.aux2 = (int*)p; p += 4 * M;
        S.div = (int*)p; p += 4 * M;
    }
    if (n == 0) { if (tid == 0) selCount[f * P.nlevels + l] = 0; return; }

    // ---- roots (ORBextractor.cc:542-584)
    const int nIni = L.nIni;
    const int bw = L.maxBX - ORBX_MINB, bh = L.maxBY - ORBX_MINB;
    for (int i = tid; i < nIni; i += T) S.cntB[i] = 0;
    __syncthreads();
    // Key loops: keys and node labels live in global memory (L2), so every loop issues the loads of OCT_U keys per thread before it
    // touches any of them — one memory round trip per OCT_U keys instead of one (two, where the key was loaded after its label) per key.
    for (int k0 = tid; k0 < n; k0 += OCT_U * T) {
        u32 kx[OCT_U];
#pragma unroll
        for (int u = 0; u < OCT_U; u++) { const int k = k0 + u * T; kx[u] = k < n ? keys[k].x : 0u; }
#pragma unroll
        for (int u = 0; u < OCT_U; u++) {
            const int k = k0 + u * T;
            if (k >= n) break;
            const int xr = (int)(kx[u] & 0xFFFF) - ORBX_MINB;
            int r = (int)__fdiv_rn((float)xr, L.hX);
            r = min(max(r, 0), nIni - 1);
            nof[k] = (u32)r;
            atomicAdd(&S.cntB[r], 1);
        }
    }
    __syncthreads();
    if (tid == 0) {
        int m = 0;
        for (int i = 0; i < nIni; i++) {
            if (S.cntB[i] > 0) {
                S.boxA[m] = make_int4((int)(L.hX * (float)i), (int)(L.hX * (float)(i + 1)), 0, bh);
                S.cntA[m] = S.cntB[i];
                s_rootMap[i] = m++;
            } else s_rootMap[i] = -1;
        }
        s_cut = m;
    }
    __syncthreads();
    int size = s_cut;
    for (int k0 = tid; k0 < n; k0 += OCT_U * T) {
        u32 v[OCT_U];
#pragma unroll
        for (int u = 0; u < OCT_U; u++) { const int k = k0 + u * T; v[u] = k < n ? nof[k] : 0u; }
#pragma unroll
        for (int u = 0; u < OCT_U; u++) { const int k = k0 + u * T; if (k < n) nof[k] = (u32)s_rootMap[v[u]]; }
    }
    __syncthreads();
    (void)bw;

    int4* box = S.boxA; int4* nbox = S.boxB;
    int* cnt = S.cntA; int* ncnt = S.cntB;
    bool sweep = false;
    while (true) {
        const int prevSize = size;
        // 1. quadrant of every key that lives in a multi-key node; child counts
        for (int i = tid; i < 4 * size; i += T) S.cc[i] = 0;
        if (tid == 0) { s_nToExpand = 0; s_cut = 0x7fffffff; }
        __syncthreads();
        for (int k0 = tid; k0 < n; k0 += OCT_U * T) {
            u32 v[OCT_U], kx[OCT_U];
#pragma unroll
            for (int u = 0; u < OCT_U; u++) {
                const int k = k0 + u * T;
                v[u] = k < n ? nof[k] : 0u;
                kx[u] = k < n ? keys[k].x : 0u;
            }
#pragma unroll
            for (int u = 0; u < OCT_U; u++) {
                const int k = k0 + u * T;
                if (k >= n) break;
                const int nd = (int)(v[u] & 0xFFFF);
                if (cnt[nd] > 1) {
                    const int x = (int)(kx[u] & 0xFFFF) - ORBX_MINB, y = (int)(kx[u] >> 16) - ORBX_MINB;
                    const int4 b = box[nd];
                    const int mx = b.x + ((b.y - b.x + 1) >> 1), my = b.z + ((b.w - b.z + 1) >> 1);
                    const int q = (x < mx ? 0 : 1) + (y < my ? 0 : 2);
                    atomicAdd(&S.cc[nd * 4 + q], 1);
                    nof[k] = (u32)nd | ((u32)q << 16);
                }
            }
        }
        __syncthreads();
        // 2. number of non-empty children per multi node
        for (int i = tid; i < size; i += T) {
            int nch = 0;
            if (cnt[i] > 1) nch = (S.cc[4 * i] > 0) + (S.cc[4 * i + 1] > 0) + (S.cc[4 * i + 2] > 0) + (S.cc[4 * i + 3] > 0);
            S.aux[i] = nch;
        }
        __syncthreads();
        int C, newSize;
        if (!sweep) {
            // full pass (ORBextractor.cc:597-664): every multi node divides, processed in list order
            for (int i = tid; i < size; i += T) {
                S.div[i] = cnt[i] > 1;
                if (cnt[i] > 1) {
                    const int e = (S.cc[4 * i] > 1) + (S.cc[4 * i + 1] > 1) + (S.cc[4 * i + 2] > 1) + (S.cc[4 * i + 3] > 1);
                    if (e) atomicAdd(&s_nToExpand, e);
                }
            }
            __syncthreads();
            C = block_excl_scan(S.aux, size, s_wsum);          // aux[i] = creation index of node i's first child
        } else {
            // final phase sweep (ORBextractor.cc:675-736): multi nodes sorted by (size, creation seq) ascending and walked
            // from the back == (size desc, list position asc); stop as soon as the list holds N nodes.
            for (int i = tid; i < P.sortN; i += T) S.sortb[i] = ~0ull;
            for (int i = tid; i < size; i += T) { S.aux2[i] = cnt[i] > 1; S.div[i] = 0; }
            __syncthreads();
            const int ncand = block_excl_scan(S.aux2, size, s_wsum);
            for (int i = tid; i < size; i += T)
                if (cnt[i] > 1) S.sortb[S.aux2[i]] = ((u64)(0xFFFFFFFFu - (u32)cnt[i]) << 32) | (u32)i;
            __syncthreads();
            if (ncand == 0) break;                                             // nothing left to expand (size == prevSize)
            int sn = 2;
            while (sn < ncand) sn <<= 1;
            block_bitonic_sort(S.sortb, sn);
            for (int t2 = tid; t2 < ncand; t2 += T) S.aux2[t2] = S.aux[(int)(S.sortb[t2] & 0xFFFFFFFFu)];
            __syncthreads();
            // inclusive list size after expanding sorted node t: size + sum_{u<=t}(nch_u - 1)
            for (int t2 = tid; t2 < ncand; t2 += T) S.cc[t2] = S.aux2[t2];     // keep nch_t (cc is free now)
            __syncthreads();
            block_excl_scan(S.aux2, ncand, s_wsum);                            // aux2[t] = q_t
            for (int t2 = tid; t2 < ncand; t2 += T) {
                const int incl = size + S.aux2[t2] + S.cc[t2] - (t2 + 1);
                if (incl >= N) atomicMin(&s_cut, t2);
            }
            __syncthreads();
            const int cut = min(s_cut, ncand - 1);
            C = S.aux2[cut] + S.cc[cut];
            __syncthreads();
            for (int t2 = tid; t2 <= cut; t2 += T) {
                const int i = (int)(S.sortb[t2] & 0xFFFFFFFFu);
                S.div[i] = 1;
                S.aux[i] = S.aux2[t2];                                         // creation index of its first child
            }
            __syncthreads();
            // recompute the child counts the scan scratch overwrote
            for (int i = tid; i < 4 * size; i += T) S.cc[i] = 0;
            __syncthreads();
            for (int k0 = tid; k0 < n; k0 += OCT_U * T) {
                u32 v[OCT_U];
#pragma unroll
                for (int u = 0; u < OCT_U; u++) { const int k = k0 + u * T; v[u] = k < n ? nof[k] : 0u; }
#pragma unroll
                for (int u = 0; u < OCT_U; u++) {
                    if (k0 + u * T >= n) break;
                    const int nd = (int)(v[u] & 0xFFFF);
                    if (cnt[nd] > 1) atomicAdd(&S.cc[nd * 4 + (int)((v[u] >> 16) & 3)], 1);
                }
            }
            __syncthreads();
        }
        // 3. list positions: children of dividing nodes go to the front in reverse creation order, the rest keep
        //    their order behind them
        for (int i = tid; i < size; i += T) S.aux2[i] = S.div[i] ? 0 : 1;
        __syncthreads();
        const int nStay = block_excl_scan(S.aux2, size, s_wsum);
        newSize = C + nStay;
        if (newSize > M) {                                                     // cannot happen (bound N+3); be safe
            if (tid == 0) atomicOr(status, ORB_DEV_NODE_OVERFLOW);
            newSize = size;
            break;
        }
        for (int i = tid; i < size; i += T) {
            if (S.div[i]) {
                const int4 b = box[i];
                const int mx = b.x + ((b.y - b.x + 1) >> 1), my = b.z + ((b.w - b.z + 1) >> 1);
                int q = S.aux[i];
                const int c0 = S.cc[4 * i], c1 = S.cc[4 * i + 1], c2 = S.cc[4 * i + 2], c3 = S.cc[4 * i + 3];
                int pos;
                if (c0 > 0) { pos = C - 1 - q++; nbox[pos] = make_int4(b.x, mx, b.z, my); ncnt[pos] = c0; S.cc[4 * i] = pos; }
                if (c1 > 0) { pos = C - 1 - q++; nbox[pos] = make_int4(mx, b.y, b.z, my); ncnt[pos] = c1; S.cc[4 * i + 1] = pos; }
                if (c2 > 0) { pos = C - 1 - q++; nbox[pos] = make_int4(b.x, mx, my, b.w); ncnt[pos] = c2; S.cc[4 * i + 2] = pos; }
                if (c3 > 0) { pos = C - 1 - q++; nbox[pos] = make_int4(mx, b.y, my, b.w); ncnt[pos] = c3; S.cc[4 * i + 3] = pos; }
            } else {
                const int pos = C + S.aux2[i];
                nbox[pos] = box[i];
                ncnt[pos] = cnt[i];
                S.aux2[i] = pos;
            }
        }
        __syncthreads();
        for (int k0 = tid; k0 < n; k0 += OCT_U * T) {
            u32 v[OCT_U];
#pragma unroll
            for (int u = 0; u < OCT_U; u++) { const int k = k0 + u * T; v[u] = k < n ? nof[k] : 0u; }
#pragma unroll
            for (int u = 0; u < OCT_U; u++) {
                const int k = k0 + u * T;
                if (k >= n) break;
                const int nd = (int)(v[u] & 0xFFFF);
                nof[k] = (u32)(S.div[nd] ? S.cc[nd * 4 + (int)((v[u] >> 16) & 3)] : S.aux2[nd]);
            }
        }
        const int nToExpand = s_nToExpand;
        __syncthreads();
        size = newSize;
        { int4* tb = box; box = nbox; nbox = tb; int* tc = cnt; cnt = ncnt; ncnt = tc; }
        if (size >= N || size == prevSize) break;                              // ORBextractor.cc:668, 733
        if (!sweep && size + 3 * nToExpand > N) sweep = true;                  // :672
    }

    // ---- one keypoint per node: max response, first in input order wins ties (ORBextractor.cc:741-759)
    u64* best = (u64*)S.cc;
    for (int i = tid; i < size; i += T) best[i] = 0ull;
    __syncthreads();
    for (int k0 = tid; k0 < n; k0 += OCT_U * T) {
        u32 nd[OCT_U];
        uint2 ky[OCT_U];
#pragma unroll
        for (int u = 0; u < OCT_U; u++) {
            const int k = k0 + u * T;
            nd[u] = k < n ? nof[k] : 0u;
            ky[u] = k < n ? keys[k] : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < OCT_U; u++) {
            if (k0 + u * T >= n) break;
            const uint2 key = ky[u];
            const u64 order = ((u64)(key.y & 0xFFFFFFu) << 32) | ((u64)(key.x >> 16) << 16) | (u64)(key.x & 0xFFFF);
            const u64 v = ((u64)(key.y >> 24) << 56) | (0x00FFFFFFFFFFFFFFull - order);
            atomicMax(&best[nd[u] & 0xFFFF], v);
        }
    }
    __syncthreads();
    for (int i = tid; i < size; i += T) {
        const u64 v = best[i];
        const u64 order = 0x00FFFFFFFFFFFFFFull - (v & 0x00FFFFFFFFFFFFFFull);
        const u32 x = (u32)(order & 0xFFFF), y = (u32)((order >> 16) & 0xFFFF);
        if (i < L.selCap) out[i] = make_uint2(x | (y << 16), (u32)(v >> 56));
    }
    if (tid == 0) selCount[f * P.nlevels + l] = min(size, L.selCap);
}

// =====================================================================================================
// K5  GaussianBlur 7x7 sigma 2 (OpenCV 4.13 fixed-point: taps [18,34,48,56,48,34,18]/256 each way,
// out = (sum + 32768) >> 16).  Reads the level WITHOUT a materialised border (BORDER_REFLECT_101 is applied to the taps of edge tiles),
// writes the blurred level in the same geometry.  Tile 128x128 outputs per CTA; separable through shared memory.
// =====================================================================================================
#define BL_TW 128
#define BL_TH 128
// All levels in one launch: blockIdx.x walks the per-level tile lists (Plan::blurTileBase).
// Horizontal pass: the 7 taps of an output are two IDP4A over byte windows cut from three aligned words with funnel shifts (no
// byte unpacking); a thread does two vertically adjacent rows and stores their exact results (<= 65280) packed as u16 pairs.
// Vertical pass: with rows paired that way the 7 vertical taps of an output are four IDP2A (tap pairs (18,34)(48,56)(48,34)(18,0)
// for even rows, (0,18)(34,48)(56,48)(34,18) for odd rows) on four LDS.128 shared by the two outputs of a thread.
__global__ void __launch_bounds__(256) k_blur(const __grid_constant__ Plan P, const u8* __restrict__ pyr, u8* __restrict__ blur) {
    __shared__ uint4 s_p[(BL_TH + 6) / 2][BL_TW / 4 + 1];             // [row pair][4-column group]: h(row 2j) | h(row 2j+1) << 16 (+1: rows start on different banks)
    int level = 0;
    while (level + 1 < P.nlevels && (int)blockIdx.x >= P.blurTileBase[level + 1]) level++;
    const LevelPlan& L = P.lv[level];
    const int t = blockIdx.x - P.blurTileBase[level];
    const int tilesX = (L.w + BL_TW - 1) / BL_TW;
    const int ty = t / tilesX, tx = t - ty * tilesX;
    const int x0 = tx * BL_TW, y0 = ty * BL_TH, f = blockIdx.y;
    const int tid = threadIdx.x;
    const int pitchW = L.pitch >> 2;
    const u32* src = reinterpret_cast<const u32*>(pyr + (size_t)f * P.frameBytes + L.off) + ORBX_OY * pitchW + ((x0 + ORBX_OX) >> 2);
    const u32 K1 = 18u | (34u << 8) | (48u << 16) | (56u << 24);      // taps for x-3, x-2, x-1, x
    const u32 K2 = 48u | (34u << 8) | (18u << 16);                    // taps for x+1, x+2, x+3
    const int rows = min(BL_TH, L.h - y0);                            // output rows of this tile
    const int pairsNeeded = (rows + 6 + 1) >> 1;
    const int wcols = min(BL_TW / 4, (L.w - x0 + 3) >> 2);            // 4-column groups that hold image pixels
    // BORDER_REFLECT_101 without a materialised border: rows are read at their reflected index; in the first / last 4-pixel group
    // of a level row the words that would hold border pixels are rebuilt from their neighbours with one PRMT (selectors by
    // k = valid pixels in the last group: bytes of (a, b) = 0..7; kept pixel i -> 4 + i, reflected pixel p -> 2w - 2 - p)
    const int kLast = ((L.w - 1) & 3) + 1, gLast = ((L.w - 1) >> 2) - (x0 >> 2);     // (gLast relative to this tile's first group)
    const u32 selB = kLast == 1 ? 0x1234u : kLast == 2 ? 0x3454u : kLast == 3 ? 0x5654u : 0x7654u;
    const u32 selC = kLast == 1 ? 0x0000u : kLast == 2 ? 0x0012u : kLast == 3 ? 0x1234u : 0x3456u;
    // work items are (row pair, column group) over the groups that hold image pixels only: edge tiles of a level (widths are not
    // multiples of 128) keep every lane busy.  j = i / wcols by multiplication (exact for i < 2048).
    const u32 winv = (65536u + (u32)wcols - 1u) / (u32)wcols;
    // interior tiles (CTA-uniform) take the plain path; only tiles touching a level edge pay for the reflection
    const bool edgeTile = y0 < 3 || y0 + 2 * pairsNeeded - 3 > L.h || x0 == 0 || gLast <= wcols;
    auto hpass = [&](auto edge) {
        constexpr bool EDGE = decltype(edge)::value;
        for (int i = tid; i < pairsNeeded * wcols; i += 256) {
            const int j = (int)(((u32)i * winv) >> 16), wx = i - j * wcols;
            // REFLECT_101 of a row index in [-3, h + 3]: min(|p|, 2 (h - 1) - |p|)
            const int pa = y0 - 3 + 2 * j, ya = EDGE ? min(abs(pa), 2 * (L.h - 1) - abs(pa)) : pa;
            const int yb = EDGE ? min(abs(pa + 1), 2 * (L.h - 1) - abs(pa + 1)) : ya + 1;
            const u32* rp = src + (u32)(ya * pitchW + wx);           // one 32-bit offset, then immediates (as in k_resize)
            const u32* rq = EDGE ? src + (u32)(yb * pitchW + wx) : rp + pitchW;
            u32 a0 = __ldg(rp - 1), b0 = __ldg(rp), c0 = __ldg(rp + 1);
            u32 a1 = __ldg(rq - 1), b1 = __ldg(rq), c1 = __ldg(rq + 1);
            if (EDGE) {
                if (x0 == 0 && wx == 0) { a0 = __byte_perm(b0, c0, 0x1234); a1 = __byte_perm(b1, c1, 0x1234); }  // px -4..-1 = px 4..1
                if (wx == gLast) {
                    c0 = __byte_perm(a0, b0, selC); b0 = __byte_perm(a0, b0, selB);
                    c1 = __byte_perm(a1, b1, selC); b1 = __byte_perm(a1, b1, selB);
                } else if (wx == gLast - 1) { c0 = __byte_perm(b0, c0, selB); c1 = __byte_perm(b1, c1, selB); }
            }
            uint4 h;
            h.x = __dp4a(__funnelshift_r(a0, b0, 8), K1, __dp4a(__funnelshift_r(b0, c0, 8), K2, 0u)) |
                  (__dp4a(__funnelshift_r(a1, b1, 8), K1, __dp4a(__funnelshift_r(b1, c1, 8), K2, 0u)) << 16);
            h.y = __dp4a(__funnelshift_r(a0, b0, 16), K1, __dp4a(__funnelshift_r(b0, c0, 16), K2, 0u)) |
                  (__dp4a(__funnelshift_r(a1, b1, 16), K1, __dp4a(__funnelshift_r(b1, c1, 16), K2, 0u)) << 16);
            h.z = __dp4a(__funnelshift_r(a0, b0, 24), K1, __dp4a(__funnelshift_r(b0, c0, 24), K2, 0u)) |
                  (__dp4a(__funnelshift_r(a1, b1, 24), K1, __dp4a(__funnelshift_r(b1, c1, 24), K2, 0u)) << 16);
            h.w = __dp4a(b0, K1, __dp4a(c0, K2, 0u)) | (__dp4a(b1, K1, __dp4a(c1, K2, 0u)) << 16);
            s_p[j][wx] = h;
        }
    };
    if (edgeTile) hpass(std::true_type{}); else hpass(std::false_type{});
    __syncthreads();
    u8* dst = blur + (size_t)f * P.frameBytes + L.off + (size_t)(y0 + ORBX_OY) * L.pitch + x0 + ORBX_OX;
    // tap pairs as the two low bytes of the IDP2A weight operand
    const u32 E0 = 18u | (34u << 8), E1 = 48u | (56u << 8), E2 = 48u | (34u << 8), E3 = 18u;          // even output row
    const u32 O0 = 18u << 8, O1 = 34u | (48u << 8), O2 = 56u | (48u << 8), O3 = 34u | (18u << 8);     // odd output row
    for (int i = tid; i < ((rows + 1) >> 1) * wcols; i += 256) {
        const int j = (int)(((u32)i * winv) >> 16), wx = i - j * wcols;   // output rows 2j and 2j+1 use row pairs j .. j+3
        const uint4 p0 = s_p[j][wx], p1 = s_p[j + 1][wx], p2 = s_p[j + 2][wx], p3 = s_p[j + 3][wx];
#define VSUM(c, W0, W1, W2, W3) __dp2a_lo(p0.c, W0, __dp2a_lo(p1.c, W1, __dp2a_lo(p2.c, W2, __dp2a_lo(p3.c, W3, 32768u))))
        const u32 e = (VSUM(x, E0, E1, E2, E3) >> 16) | ((VSUM(y, E0, E1, E2, E3) >> 16) << 8) |
                      ((VSUM(z, E0, E1, E2, E3) >> 16) << 16) | ((VSUM(w, E0, E1, E2, E3) >> 16) << 24);
        u8* const drow = dst + (u32)((2 * j) * L.pitch + 4 * wx);
        *reinterpret_cast<u32*>(drow) = e;
        if (2 * j + 1 < rows) {
            const u32 o = (VSUM(x, O0, O1, O2, O3) >> 16) | ((VSUM(y, O0, O1, O2, O3) >> 16) << 8) |
                          ((VSUM(z, O0, O1, O2, O3) >> 16) << 16) | ((VSUM(w, O0, O1, O2, O3) >> 16) << 24);
            *reinterpret_cast<u32*>(drow + L.pitch) = o;
        }
#undef VSUM
    }
}

// =====================================================================================================
// K6  IC_Angle + rBRIEF, one warp per selected keypoint.  Orientation: int32 moments over the radius-15 disc of the
// UNBLURRED level, cv::fastAtan2 polynomial with explicit round-to-nearest mul/add (no FMA contraction).
// Descriptor: lane i computes byte i (8 point pairs) of the 256-bit string from the BLURRED level.
// Writes the final cv::KeyPoint (coordinates scaled to level 0) and descriptor at their level-major position.
// =====================================================================================================
__constant__ signed char c_pattern[1024] = {
#include "orb_pattern_31.inc"
};

__device__ __forceinline__ float dev_fast_atan2(float y, float x) {
    const float s = (float)(180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, (float)DBL_EPSILON));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, (float)DBL_EPSILON));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// sin and cos of x in [0, 2 pi] in double: quadrant by k = rint(x * 2/pi), Cody-Waite reduction with a two-part pi/2, the fdlibm
// kernel polynomials on [-pi/4, pi/4] (< 1 ulp in double).  Only the float rounding of the results is used, so this agrees with a
// correctly rounded cosf / sinf except within ~1e-16 (relative) of a float rounding boundary.  ~30 instructions against ~80 of the
// library's sincos (large-argument and special-case paths that cannot occur here).
__device__ __forceinline__ void dev_sincos_0_2pi(double x, double& s, double& c) {
    const int q = __double2int_rn(x * 0.63661977236758138243);
    const double k = (double)q;
    double r = fma(-k, 1.57079632679489655800e+00, x);
    r = fma(-k, 6.12323399573676603587e-17, r);
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double sn = fma(r * z, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double cn = fma(z * z, pc, fma(z, -0.5, 1.0));
    const double ss = (q & 1) ? cn : sn, cc = (q & 1) ? sn : cn;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}

#define DESC_MAXWARPS 32      // warps per CTA are chosen per launch (blockDim.x / 32): one CTA per SM with as many warps as fit, like k_fast_tma
#define DESC_PW 11            // aligned words per staged blurred-patch row: covers kx-18 .. kx+18 for any alignment
#define DESC_AW 9             // aligned words per staged disc row: covers kx-15 .. kx+15 for any alignment
#define DESC_PATCH_WORDS (37 * DESC_PW)
#define DESC_DISC_WORDS (31 * DESC_AW)
#define DESC_BUF_WORDS (DESC_PATCH_WORDS + DESC_DISC_WORDS)
#define DESC_SMEM_BYTES(W) (32 * 36 * 4 + 2 * (W) * DESC_BUF_WORDS * 4)
// TMA staging (k_describe<true>): the two neighbourhoods of a keypoint are two boxes of the per-level tensor maps, dropped into
// shared memory by one elected lane (no per-lane address arithmetic).  A box has to start at a 16-byte aligned pixel column (an
// unaligned start raises "illegal instruction" on sm_100a), so boxes are 64 / 48 bytes wide and start at the aligned column at or
// before the neighbourhood's first pixel; buffers are 128-byte aligned.
#define DESC_TMA_PW 64                    // 15 + 37 <= 64
#define DESC_TMA_AW 48                    // 15 + 31 <= 48
#define DESC_TMA_PATCH_BYTES 2432         // 37 x 64 = 2368, rounded up to 128
#define DESC_TMA_DISC_BYTES 1536          // 31 x 48 = 1488
#define DESC_TMA_BUF_BYTES (DESC_TMA_PATCH_BYTES + DESC_TMA_DISC_BYTES)
#define DESC_TMA_TX_BYTES (37 * DESC_TMA_PW + 31 * DESC_TMA_AW)
#define DESC_TMA_SMEM_BYTES(W) (32 * 36 * 4 + 2 * (W) * DESC_TMA_BUF_BYTES + 2 * (W) * 8)
// Persistent warps over the work items (frame, slot): slot r of a frame is position r of the per-frame selection buffer
// (level l owns [selOff[l], selOff[l] + selCap[l])), so the selection entry and the per-level counts of an item are two
// INDEPENDENT loads whose addresses follow from the item number alone.  Software pipeline per warp:
//   item i+2: those two loads are issued;   item i+1: its two neighbourhoods are copied to shared memory with cp.async (no
//   registers, no waiting);   item i: orientation + descriptor from the other shared-memory buffer.
struct DescItem { uint2 k; int cntv, f, r; bool valid; };
__device__ __forceinline__ void cp_async4(u32 dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
template <bool TMA>
__global__ void __launch_bounds__(32 * DESC_MAXWARPS) k_describe(const __grid_constant__ Plan P, const u8* __restrict__ pyr,
                                                              const u8* __restrict__ blur, const uint2* __restrict__ sel,
                                                              const int* __restrict__ selCount, orbx_keypoint* __restrict__ kpOut,
                                                              u8* __restrict__ descOut, int* __restrict__ nOut, int cap,
                                                              int* __restrict__ status, int nf, const CUtensorMap* __restrict__ maps,
                                                              int* __restrict__ workCounter, int DESC_CHUNK) {
    extern __shared__ __align__(128) u8 smem_desc[];
    float* s_pat = reinterpret_cast<float*>(smem_desc);            // pattern as float (no I2F in the tap loop); row stride 36: conflict-free LDS.128
    u32* s_buf = reinterpret_cast<u32*>(smem_desc + 32 * 36 * 4);  // [2][DESC_WARPS][buffer]: blurred 37x37 + unblurred 31x31 neighbourhoods
    constexpr int BUF_WORDS = TMA ? DESC_TMA_BUF_BYTES / 4 : DESC_BUF_WORDS;
    constexpr int PATCH_WORDS = TMA ? DESC_TMA_PATCH_BYTES / 4 : DESC_PATCH_WORDS;
    constexpr int PWB = TMA ? DESC_TMA_PW : DESC_PW * 4, AWB = TMA ? DESC_TMA_AW : DESC_AW * 4;   // row pitches in bytes
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, DESC_WARPS = blockDim.x >> 5;
    // TMA: one mbarrier per (warp, buffer), completed by the two box loads of an item
    const u32 bars = smem_u32(smem_desc + 32 * 36 * 4 + 2 * DESC_WARPS * DESC_TMA_BUF_BYTES) + (u32)warp * 16;
    u32 phases = 0;                                               // bit b: parity of buffer b's barrier
    if (TMA) {
        if (lane == 0) {
            mbar_init(bars, 1);
            mbar_init(bars + 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_pat[(i >> 5) * 36 + (i & 31)] = (float)c_pattern[i];
    __syncthreads();
    const int nItems = nf * P.selTotal, GW = gridDim.x * DESC_WARPS;

    // the item stream of this warp: chunks of DESC_CHUNK (4) consecutive items — consecutive selection slots of one frame.  The first
    // chunk is the warp's own number, further chunks come from a global counter, so the warps in flight always work on one contiguous
    // window of items (neighbouring keypoints of a few frames: their neighbourhoods share L2 lines) however many warps an SM holds and
    // however the schedulers' loads differ.  Measured per 4096 frames: fixed stride 3.12 ms at 28 warps per SM but 3.6-3.7 ms at 25-27
    // (uneven warps per scheduler) ; claims of 1 item 4.35 ms (one atomic per item), 2: 3.39, 4: 3.12, 8: 3.11, 16: 3.14, 32: 3.36
    // (a warp walking 32 slots alone loses the sharing between concurrent warps).  (frame, slot) advance incrementally inside a chunk.
    int ldIt = 0, ldEnd = 0, ldF = 0, ldR = 0;
    bool ldDone = false;
    auto claim = [&](int chunk) {
        const long long base = (long long)chunk * DESC_CHUNK;
        if (base >= nItems) { ldDone = true; return; }
        ldIt = (int)base; ldEnd = min(ldIt + DESC_CHUNK, nItems);
        ldF = ldIt / P.selTotal; ldR = ldIt - ldF * P.selTotal;
    };
    claim(blockIdx.x * DESC_WARPS + warp);
    auto load_next = [&]() {                                       // stage A: two independent loads
        if (!ldDone && ldIt >= ldEnd) {
            int c = 0;
            if (lane == 0) c = GW + atomicAdd(workCounter, 1);
            claim(__shfl_sync(0xffffffffu, c, 0));
        }
        DescItem d;
        d.k = make_uint2(0u, 0u); d.cntv = 0; d.f = ldF; d.r = ldR; d.valid = !ldDone;
        if (d.valid) {
            d.k = __ldg(sel + ldIt);
            d.cntv = lane < P.nlevels ? __ldg(selCount + ldF * P.nlevels + lane) : 0;
            ldIt++; ldR++;
            if (ldR >= P.selTotal) { ldR = 0; ldF++; }
        }
        return d;
    };
    // resolves (level, index in level, output position) of an item; returns false for empty slots
    auto resolve = [&](const DescItem& d, int& f, int& l, int& pos, int& total) {
        if (!d.valid) return false;
        f = d.f;
        const int r = d.r;
        // level = number of levels 1.. whose first slot is <= r (lanes compare in parallel); counts: two warp reductions
        l = __popc(__ballot_sync(0xffffffffu, lane >= 1 && lane < P.nlevels && r >= P.lv[lane & (ORBX_MAX_LEVELS - 1)].selOff));
        const int idx = r - P.lv[l].selOff;
        total = __reduce_add_sync(0xffffffffu, d.cntv);
        const int before = __reduce_add_sync(0xffffffffu, lane < l ? d.cntv : 0), cnt = __shfl_sync(0xffffffffu, d.cntv, l);
        pos = before + idx;
        if (r == 0 && lane == 0) {                                 // slot 0 of every frame reports the frame's keypoint count
            nOut[f] = total;
            if (total > cap) atomicOr(status, ORB_DEV_OUT_OVERFLOW);
        }
        return idx < cnt && pos < cap;
    };
    auto stage = [&](int buf, int f, int l, const uint2 k) {       // stage B: both neighbourhoods to shared memory, asynchronously
        const LevelPlan& L = P.lv[l];
        const int kx = (int)(k.x & 0xFFFF), ky = (int)(k.x >> 16);
        if (TMA) {
            if (lane == 0) {
                const u32 bar = bars + 8 * buf;
                const u32 dst = smem_u32(s_buf + (size_t)(buf * DESC_WARPS + warp) * BUF_WORDS);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // our generic reads of this buffer precede the async write
                mbar_expect_tx(bar, DESC_TMA_TX_BYTES);
                tma_load_3d(dst, maps + P.nlevels + l, (kx - 18 + ORBX_OX) & ~15, ky - 18 + ORBX_OY, f, bar);
                tma_load_3d(dst + DESC_TMA_PATCH_BYTES, maps + 2 * P.nlevels + l, (kx - 15 + ORBX_OX) & ~15, ky - 15 + ORBX_OY, f, bar);
            }
            return;
        }
        const int pitchW = L.pitch >> 2;
        const size_t lofs = (size_t)f * P.frameBytes + L.off;
        const int pxs = kx - 18 + ORBX_OX, axs = kx - 15 + ORBX_OX;
        const u32* bsrc = reinterpret_cast<const u32*>(blur + lofs) + (ky - 18 + ORBX_OY) * pitchW + (pxs >> 2);
        const u32* asrc = reinterpret_cast<const u32*>(pyr + lofs) + (ky - 15 + ORBX_OY) * pitchW + (axs >> 2);
        const u32 dst = smem_u32(s_buf + (size_t)(buf * DESC_WARPS + warp) * BUF_WORDS);
#pragma unroll
        for (int j = 0; j < 13; j++) {
            const int i = lane + 32 * j, r = i / DESC_PW, w = i - r * DESC_PW;
            if (i < DESC_PATCH_WORDS) cp_async4(dst + 4 * i, bsrc + r * pitchW + w);
        }
#pragma unroll
        for (int j = 0; j < 9; j++) {
            const int i = lane + 32 * j, r = i / DESC_AW, w = i - r * DESC_AW;
            if (i < DESC_DISC_WORDS) cp_async4(dst + 4 * (DESC_PATCH_WORDS + i), asrc + r * pitchW + w);
        }
    };

    // prologue
    DescItem d0 = load_next(), d1 = load_next();
    int f0 = 0, l0 = 0, pos0 = 0, tot0 = 0;
    bool ok0 = resolve(d0, f0, l0, pos0, tot0);
    if (ok0) stage(0, f0, l0, d0.k);
    if (!TMA) asm volatile("cp.async.commit_group;" ::: "memory");
    int buf = 0;
    while (d0.valid) {                                             // (chunks are claimed in increasing order: after the first item past the end all are)
        // stage A for item+2 (consumed two iterations from now), stage B for item+1
        const DescItem d2 = load_next();
        int f1 = 0, l1 = 0, pos1 = 0, tot1 = 0;
        const bool ok1 = resolve(d1, f1, l1, pos1, tot1);
        if (ok1) stage(buf ^ 1, f1, l1, d1.k);
        if (TMA) {
            if (ok0) { mbar_wait(bars + 8 * buf, (phases >> buf) & 1u); phases ^= 1u << buf; }
        } else {
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");   // the copies of the CURRENT item have landed
        }
        __syncwarp();
        if (ok0) {
            const int l = l0, f = f0, pos = pos0;
            const uint2 k = d0.k;
            const LevelPlan& L = P.lv[l];
            const int kx = (int)(k.x & 0xFFFF), ky = (int)(k.x >> 16);
            const int shift = (kx - 18 + ORBX_OX) & (TMA ? 15 : 3), ashift = (kx - 15 + ORBX_OX) & (TMA ? 15 : 3);
            const u32* patch = s_buf + (size_t)(buf * DESC_WARPS + warp) * BUF_WORDS;
            const u32* disc = patch + PATCH_WORDS;

            // ---- IC_Angle (ORBextractor.cc:76-103): lane = column u (-15..15); the disc is symmetric (:461-468 makes umax its own
            // transpose), so column u spans rows |v| <= umax[|u|]
            int m10 = 0, m01 = 0;
            if (lane < 31) {
                const int u = lane - 15, vlim = P.umax[u < 0 ? -u : u];
                // explicit shared-space loads from a 32-bit address: with generic pointers the compiler rebuilt a shared::cluster window
                // address (S2R SR_CgaCtaId + LEA) for every one of the 31 predicated loads of the TMA variant
                const u32 cen = smem_u32(disc) + 15 * AWB + 15 + ashift + u;
                int colsum = 0;
#pragma unroll
                for (int v = -15; v <= 15; v++) {
                    int I = 0;
                    if ((v < 0 ? -v : v) <= vlim) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(I) : "r"(cen + v * AWB));
                    colsum += I;
                    m01 += v * I;
                }
                m10 = u * colsum;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
            const float angle = dev_fast_atan2((float)m01, (float)m10);

            // ---- computeOrbDescriptor (ORBextractor.cc:107-146); cos for lane 0, sin for lane 1 (double, then rounded to float)
            const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
            const float ang = __fmul_rn(angle, factorPI);
            float cs = 0.f;
            if (lane < 2) {                                         // lane 0: cos, lane 1: sin (double, then rounded to float)
                double sd, cd;
                dev_sincos_0_2pi((double)ang, sd, cd);
                cs = (float)(lane == 0 ? cd : sd);
            }
            const float a = __shfl_sync(0xffffffffu, cs, 0), b = __shfl_sync(0xffffffffu, cs, 1);
            const float4* pat = reinterpret_cast<const float4*>(s_pat + lane * 36);
            u32 val = 0;
            // round-half-even without the XU pipe: x + 1.5*2^23 leaves rint(x) in the low mantissa bits (|x| < 2^22), == cvRound;
            // the bias of both coordinates (0x4B400000 each) is folded into the patch address once: byte (r, q) of the patch is at
            // centerB + asint(r + MAGIC) * PWB + asint(q + MAGIC)   (32-bit shared address arithmetic, wraps)
            const float MAGIC = 12582912.f;
            const u32 centerB = smem_u32(patch) + (u32)(18 * PWB + 18 + shift) - 0x4B400000u * (u32)(PWB + 1);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 pp = pat[j];                               // x0, y0, x1, y1
                const u32 r0 = (u32)__float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(pp.x, b), __fmul_rn(pp.y, a)), MAGIC));
                const u32 q0 = (u32)__float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(pp.x, a), __fmul_rn(pp.y, b)), MAGIC));
                const u32 r1 = (u32)__float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(pp.z, b), __fmul_rn(pp.w, a)), MAGIC));
                const u32 q1 = (u32)__float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(pp.z, a), __fmul_rn(pp.w, b)), MAGIC));
                const u32 t0 = lds8i<0>(centerB + r0 * (u32)PWB + q0), t1 = lds8i<0>(centerB + r1 * (u32)PWB + q1);
                val |= (u32)(t0 < t1) << j;
            }
            descOut[((size_t)f * cap + pos) * 32 + lane] = (u8)val;
            if (lane == 0) {
                orbx_keypoint kp;
                kp.x = (float)kx; kp.y = (float)ky;
                if (l != 0) { kp.x = __fmul_rn(kp.x, L.scale); kp.y = __fmul_rn(kp.y, L.scale); }   // :1098-1104
                kp.size = L.kpSize; kp.angle = angle; kp.response = (float)k.y; kp.octave = l; kp.class_id = -1;
                kpOut[(size_t)f * cap + pos] = kp;
            }
        }
        __syncwarp();                                              // everyone is done with buffer `buf` before it is refilled
        d0 = d1; d1 = d2;
        ok0 = ok1; f0 = f1; l0 = l1; pos0 = pos1;
        buf ^= 1;
    }
    if (!TMA) asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// =====================================================================================================
// Host side: planning (the ORBextractor constructor's tables + geometry), workspace, launch sequencing
// =====================================================================================================
static inline int h_cvRoundF(float v) { return (int)nearbyintf(v); }
static inline int h_cvRoundD(double v) { return (int)nearbyint(v); }
static inline int h_cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int h_cvCeil(double v) { int i = (int)v; return i + (i < v); }

// staging slots of the pipelined host-buffer path (orbx_extract_batch): with H2D time ~ compute time two slots run the copy and
// compute streams in lock-step (each waits for the other's previous chunk), so any jitter stalls both; four let the copies run ahead
#define ORBX_MAX_SLOTS 6
#define ORBX_MAX_HELPERS 3

// Host worker pool of the pageable-caller path of orbx_extract_batch: the reference's callers hand over ordinary cv::Mat memory, which the
// DMA engines cannot read; one thread copying it into pinned staging moves 6-9 GB/s (17-27 k VGA frames/s end to end), twelve threads with
// non-temporal stores 49 GB/s (134 k frames/s; page-locked caller buffers reach 170 k).  run(n, fn) calls fn(0..n-1) on the pool's threads and the caller's; it returns when all items are done.
#if defined(__x86_64__)
#include <emmintrin.h>
#endif
// memcpy into pinned staging with non-temporal stores: the destination is only ever read by the DMA engine, so pulling its lines into the
// cache first (the read-for-ownership of an ordinary store) is a third of the memory traffic of the copy for nothing
static void copy_stream(u8* dst, const u8* src, size_t n) {
#if defined(__x86_64__)
    const size_t head = std::min(n, (size_t)((16 - ((uintptr_t)dst & 15)) & 15));
    if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i)), b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32)), d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
    }
    if (i < n) memcpy(dst + i, src + i, n - i);
    _mm_sfence();
#else
    memcpy(dst, src, n);
#endif
}

class HostPool {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv, cvDone;
    const std::function<void(int)>* fn = nullptr;
    std::atomic<int> next{0};
    int total = 0, active = 0;
    unsigned long long gen = 0;
    bool stop = false;
    void worker() {
        unsigned long long seen = 0;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return stop || gen != seen; });
            if (stop) return;
            seen = gen;
            const std::function<void(int)>* f = fn;
            const int n = total;
            lk.unlock();
            for (int i; (i = next.fetch_add(1)) < n;) (*f)(i);
            lk.lock();
            if (--active == 0) cvDone.notify_one();
        }
    }
public:
    explicit HostPool(int threads) {
        for (int i = 1; i < threads; i++) th.emplace_back([this] { worker(); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (std::thread& t : th) t.join();
    }
    int threads() const { return (int)th.size() + 1; }
    void run(int n, const std::function<void(int)>& f) {
        if (n <= 0) return;
        { std::lock_guard<std::mutex> lk(mu); fn = &f; total = n; next.store(0); active = (int)th.size(); gen++; }
        cv.notify_all();
        for (int i; (i = next.fetch_add(1)) < n;) f(i);
        std::unique_lock<std::mutex> lk(mu);
        cvDone.wait(lk, [&] { return active == 0; });
    }
};

struct orbx_extractor {
    int nfeatures, nlevels, iniTh, minTh, device, maxBatch, candPerCell;
    double scaleFactor;
    std::vector<float> sf, isf, s2, is2;
    std::vector<int> quota;
    Plan plan;
    cudaStream_t stream = nullptr;
    cudaStream_t sBlur = nullptr;          // the blur of a pass runs here, next to FAST / octree, when those leave most SMs idle
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    // device workspace
    u8 *d_pyr = nullptr, *d_blur = nullptr;
    // host-buffer path: two staging slots so that H2D(chunk i+1), compute(chunk i) and D2H(chunk i-1) overlap
    u8 *d_in[ORBX_MAX_SLOTS] = {}, *d_mask[ORBX_MAX_SLOTS] = {}, *d_desc[ORBX_MAX_SLOTS] = {};
    orbx_keypoint* d_kp[ORBX_MAX_SLOTS] = {};
    int* d_n[ORBX_MAX_SLOTS] = {};
    int* h_n = nullptr;                    // pinned staging for the per-frame counts
    int* h_status = nullptr;               // pinned mirror of the device status word
    u8* h_stage = nullptr;                 // pinned staging for the latency path when the caller's buffers are pageable
    size_t h_stage_cap = 0;
    // pipelined path, pageable caller buffers: pinned staging per slot (frames in; keypoint + descriptor rows out) filled / drained
    // by a pool of host threads
    u8 *h_in[ORBX_MAX_SLOTS] = {}, *h_out[ORBX_MAX_SLOTS] = {};
    HostPool* pool = nullptr;
    size_t h_n_cap = 0;
    cudaStream_t sH2D = nullptr, sD2H = nullptr;
    cudaEvent_t evH2D[ORBX_MAX_SLOTS] = {}, evComp[ORBX_MAX_SLOTS] = {}, evD2H[ORBX_MAX_SLOTS] = {};
    uint2 *d_cand = nullptr, *d_sel = nullptr;
    u32* d_nodeOf = nullptr;
    int *d_candCount = nullptr, *d_selCount = nullptr, *d_status = nullptr, *d_workCounter = nullptr;
    ResizeTap *d_xtab = nullptr, *d_ytab = nullptr;
    CUtensorMap* d_maps = nullptr;         // one TMA descriptor per pyramid level (k_fast_tma)
    bool useTma = false, descTma = false;
    size_t fwSmem = 0;
    int fwGrid = 0, descGrid = 0, fwWarps = 0, descWarps = 8, smCount = 148;
    uint4* d_cells = nullptr;              // valid FAST cells: {iniX | iniY<<16, tw | th<<8 | level<<16, cell id, 0}
    int nCells = 0;
    int capInternal = 0;
    size_t fastSmem = 0, octSmem = 0;
    bool fastConst = false;
    long long launches = 0;
    int lastFrames = 0;
    u8* d_stereo = nullptr;                // scratch of orbx_stereo_matches
    size_t stereoCap = 0;
    std::mutex mu;
    // orbx_extract_batch over several passes: helper handles (own stream and workspace each) take chunks in turn, so that the
    // launch gaps and kernel tails of one pass (13 dependent kernels, ~0.1 ms per pass) are filled by another pass's work
    orbx_extractor* helpers[ORBX_MAX_HELPERS] = {};
    int nHelpers = 0;
};

static void build_resize_taps(int ssize, int dsize, bool clampX, std::vector<ResizeTap>& out) {
    // cv::resize INTER_LINEAR coefficient tables (OpenCV imgproc resize.cpp): scale = 1/((double)dsize/ssize)
    const double inv_scale = (double)dsize / ssize, scale = 1. / inv_scale;
    for (int d = 0; d < dsize; d++) {
        float fx = (float)((d + 0.5) * scale - 0.5);
        int s = h_cvFloor(fx);
        fx -= s;
        if (clampX) {
            if (s < 0) { fx = 0; s = 0; }
            if (s >= ssize - 1) { fx = 0; s = ssize - 1; }
        }
        ResizeTap t;
        t.s = s;
        t.c0 = (short)h_cvRoundF((1.f - fx) * 2048.f);
        t.c1 = (short)h_cvRoundF(fx * 2048.f);
        out.push_back(t);
    }
}

static int make_plan(orbx_extractor* ex, int width, int height) {
    Plan& P = ex->plan;
    memset(&P, 0, sizeof(P));
    const int nl = ex->nlevels;
    P.nlevels = nl; P.iniTh = ex->iniTh; P.minTh = ex->minTh; P.width = width; P.height = height;
    // umax (ORBextractor.cc:451-468)
    {
        int v, v0, vmax = h_cvFloor(15 * sqrt(2.f) / 2 + 1), vmin = h_cvCeil(15 * sqrt(2.f) / 2);
        const double hp2 = 15 * 15;
        for (v = 0; v <= vmax; ++v) P.umax[v] = h_cvRoundD(sqrt(hp2 - v * v));
        for (v = 15, v0 = 0; v >= vmin; --v) {
            while (P.umax[v0] == P.umax[v0 + 1]) ++v0;
            P.umax[v] = v0;
            ++v0;
        }
    }
    std::vector<ResizeTap> xt, yt;
    std::vector<uint4> cells;
    size_t off = 0;
    int cellBase = 0, candOff = 0, selOff = 0, maxNodes = 8, maxCW = 1, maxCH = 1;
    for (int l = 0; l < nl; l++) {
        LevelPlan& L = P.lv[l];
        L.w = h_cvRoundF((float)width * ex->isf[l]);                       // ORBextractor.cc:1114-1115
        L.h = h_cvRoundF((float)height * ex->isf[l]);
        ORB_REQUIRE(L.w >= 62 && L.h >= 62 && L.w <= 65000 && L.h <= 65000, ORB_ERR_GEOMETRY,
                    "pyramid level %d is %dx%d; the reference needs every level >= 62 px (ORBextractor.cc:783-786)", l, L.w, L.h);
        L.pitch = (int)orb_align_up((size_t)ORBX_OX + L.w + ORBX_EDGE, 64);
        L.brows = L.h + 2 * ORBX_EDGE;
        L.off = (unsigned)off;
        off += orb_align_up((size_t)L.pitch * L.brows, 256);
        L.maxBX = L.w - ORBX_MINB; L.maxBY = L.h - ORBX_MINB;
        const float fw = (float)(L.maxBX - ORBX_MINB), fh = (float)(L.maxBY - ORBX_MINB);   // :779-786
        L.nCols = (int)(fw / 30.f); L.nRows = (int)(fh / 30.f);
        L.wCell = (int)ceilf(fw / L.nCols); L.hCell = (int)ceilf(fh / L.nRows);
        ORB_REQUIRE(L.wCell <= 120 && L.hCell <= 120, ORB_ERR_GEOMETRY, "cell too large");
        L.cellBase = cellBase;
        cellBase += L.nCols * L.nRows;
        maxCW = std::max(maxCW, L.wCell); maxCH = std::max(maxCH, L.hCell);
        L.quota = ex->quota[l];
        L.nIni = (int)roundf((float)(L.maxBX - ORBX_MINB) / (float)(L.maxBY - ORBX_MINB));  // :542
        ORB_REQUIRE(L.nIni >= 1 && L.nIni <= 64, ORB_ERR_GEOMETRY,
                    "level %d aspect ratio gives nIni=%d root nodes (reference divides by zero at 0)", l, L.nIni);
        L.hX = (float)(L.maxBX - ORBX_MINB) / L.nIni;                                       // :544
        L.candOff = candOff;
        {   // worst case: 3x3 NMS survivors are never 8-adjacent, so a cell domain dw x dh holds <= ceil(dw/2)*ceil(dh/2).
            // HBM is plentiful (180 GB): size for the bound so the candidate list can never overflow.
            long bound = 0;
            for (int i = 0; i < L.nRows; i++) {
                const int iniY = ORBX_MINB + i * L.hCell;
                if (iniY >= L.maxBY - 3) continue;
                const int dh = std::min(iniY + L.hCell + 6, L.maxBY) - iniY - 6;
                for (int j = 0; j < L.nCols; j++) {
                    const int iniX = ORBX_MINB + j * L.wCell;
                    if (iniX >= L.maxBX - 6) continue;
                    const int dw = std::min(iniX + L.wCell + 6, L.maxBX) - iniX - 6;
                    if (dw > 0 && dh > 0) bound += (long)((dw + 1) / 2) * ((dh + 1) / 2);
                }
            }
            for (int i = 0; i < L.nRows; i++) {          // the valid cells, in the reference's visiting order (:788-805)
                const int iniY = ORBX_MINB + i * L.hCell;
                if (iniY >= L.maxBY - 3) continue;
                const int th = std::min(iniY + L.hCell + 6, L.maxBY) - iniY;
                for (int j = 0; j < L.nCols; j++) {
                    const int iniX = ORBX_MINB + j * L.wCell;
                    if (iniX >= L.maxBX - 6) continue;
                    const int tw = std::min(iniX + L.wCell + 6, L.maxBX) - iniX;
                    if (tw - 6 <= 0 || th - 6 <= 0) continue;        // cv::FAST on a ROI < 7 px finds nothing
                    // .w: 2^32 / ng + 1 with ng = 4-pixel groups per row of the cell's domain as k_fast_tma lays them out (task -> row, group)
                    const int offq = (((iniX + ORBX_OX) & 15) + 3) & 3, ng = (offq + (tw - 6) + 3) >> 2;
                    cells.push_back(make_uint4((unsigned)iniX | ((unsigned)iniY << 16), (unsigned)tw | ((unsigned)th << 8) | ((unsigned)l << 16),
                                               (unsigned)(i * L.nCols + j), ng > 1 ? 0xFFFFFFFFu / (unsigned)ng + 1u : 0u));
                }
            }
            if (ex->candPerCell > 0) bound = std::min<long>(bound, (long)L.nCols * L.nRows * ex->candPerCell);
            L.candCap = (int)std::max<long>(bound, 1);
        }
        candOff += L.candCap;
        const int nodeBound = std::max(L.quota + 3, 4 * L.nIni) + 1;
        L.selOff = selOff; L.selCap = nodeBound;
        selOff += nodeBound;
        maxNodes = std::max(maxNodes, nodeBound);
        L.scale = ex->sf[l];
        L.kpSize = (float)(int)(31 * ex->sf[l]);                                            // :837
        L.xtabOff = (int)xt.size(); L.ytabOff = (int)yt.size();
        if (l > 0) {
            build_resize_taps(P.lv[l - 1].w, L.w, true, xt);
            build_resize_taps(P.lv[l - 1].h, L.h, false, yt);
            // fast form needs: the taps of each group of 4 outputs lie within the 8 source bytes that start at the first one's tap,
            // and every coefficient is in [0, 2048] (no saturation possible)
            L.fastResize = 1;
            for (int x = 0; x < L.w; x += 4) {
                const int s0 = xt[L.xtabOff + x].s;
                for (int k = 0; k < 4 && x + k < L.w; k++) {
                    const ResizeTap& t = xt[L.xtabOff + x + k];
                    if (t.s < s0 || t.s + 1 - s0 > 7 || s0 < 0 || t.c0 < 0 || t.c1 < 0 || t.c0 + t.c1 != 2048) L.fastResize = 0;
                }
            }
            for (int y = 0; y < L.h; y++) {
                const ResizeTap& t = yt[L.ytabOff + y];
                if (t.c0 < 0 || t.c1 < 0 || t.c0 + t.c1 != 2048) L.fastResize = 0;
            }
            while (xt.size() % 4) xt.push_back(xt.back());           // pad so a thread can always load 4 taps
        }
    }
    ORB_REQUIRE(maxNodes <= 60000, ORB_ERR_ARG, "nfeatures too large for the octree kernel");
    P.blurTileBase[0] = 0;
    for (int l = 0; l < nl; l++)
        P.blurTileBase[l + 1] = P.blurTileBase[l] + orb_div_up(P.lv[l].w, BL_TW) * orb_div_up(P.lv[l].h, BL_TH);
    P.totalCells = cellBase; P.candTotal = candOff; P.selTotal = selOff; P.maxNodes = maxNodes;
    P.sortN = 2; while (P.sortN < maxNodes) P.sortN <<= 1;
    P.frameBytes = off;
    // FAST tile: pixel pairs (u32) per row = 2 * words, words = ceil((3 + cellW + 6 + 1) / 4); two copies; + score map + corner list
    P.tilePitch = 2 * (int)((3 + maxCW + 6 + 3) / 4) + 2; P.tileRows = maxCH + 6;
    P.scorePitch = (int)orb_align_up(maxCW + 8, 4); P.scoreRows = maxCH + 2;
    ex->fastConst = P.tilePitch <= ORBX_FAST_TPP && P.scorePitch <= ORBX_FAST_SPP;
    if (ex->fastConst) { P.tilePitch = ORBX_FAST_TPP; P.scorePitch = ORBX_FAST_SPP; }
    ex->fastSmem = (size_t)P.tilePitch * P.tileRows * 4 + (size_t)P.scorePitch * P.scoreRows + (size_t)(maxCW + 2) * maxCH * 2 + 16;
    ex->nCells = (int)cells.size();
    ex->octSmem = (size_t)maxNodes * (16 * 2 + 4 * 2 + 16 + 4 * 3) + (size_t)P.sortN * 8 + 64;
    ORB_REQUIRE(ex->octSmem <= 220 * 1024, ORB_ERR_ARG, "nfeatures too large for the octree kernel's shared memory");
    ex->capInternal = selOff;

    const int B = ex->maxBatch;
    ORB_CUDA_TRY(cudaMalloc(&ex->d_pyr, (size_t)B * P.frameBytes));
    ORB_CUDA_TRY(cudaMalloc(&ex->d_blur, (size_t)B * P.frameBytes));
    ORB_CUDA_TRY(cudaMalloc(&ex->d_cand, (size_t)B * P.candTotal * sizeof(uint2)));
    ORB_CUDA_TRY(cudaMalloc(&ex->d_nodeOf, (size_t)B * P.candTotal * sizeof(u32)));
    ORB_CUDA_TRY(cudaMalloc(&ex->d_sel, (size_t)B * P.selTotal * sizeof(uint2)));
    ORB_CUDA_TRY(cudaMalloc(&ex->d_candCount, (size_t)B * nl * sizeof(int)));
    ORB_CUDA_TRY(cudaMalloc(&ex->d_workCounter, 2 * sizeof(int)));
    ORB_CUDA_TRY(cudaMalloc(&ex->d_selCount, (size_t)B * nl * sizeof(int)));
    ORB_CUDA_TRY(cudaMalloc(&ex->d_status, sizeof(int)));
    ORB_CUDA_TRY(cudaMemset(ex->d_status, 0, sizeof(int)));
    ORB_CUDA_TRY(cudaMemset(ex->d_blur, 0, (size_t)B * P.frameBytes));
    ORB_CUDA_TRY(cudaMalloc(&ex->d_xtab, std::max<size_t>(xt.size(), 1) * sizeof(ResizeTap)));
    ORB_CUDA_TRY(cudaMalloc(&ex->d_ytab, std::max<size_t>(yt.size(), 1) * sizeof(ResizeTap)));
    ORB_CUDA_TRY(cudaMalloc(&ex->d_cells, std::max<size_t>(cells.size(), 1) * sizeof(uint4)));
    if (!cells.empty()) ORB_CUDA_TRY(cudaMemcpy(ex->d_cells, cells.data(), cells.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    if (!xt.empty()) ORB_CUDA_TRY(cudaMemcpy(ex->d_xtab, xt.data(), xt.size() * sizeof(ResizeTap), cudaMemcpyHostToDevice));
    if (!yt.empty()) ORB_CUDA_TRY(cudaMemcpy(ex->d_ytab, yt.data(), yt.size() * sizeof(ResizeTap), cudaMemcpyHostToDevice));
    // ---- k_fast_tma: per-warp shared-memory layout and one tensor map per level over the pyramid workspace
    {
        const int boxW = (int)orb_align_up(maxCW + 6 + 15, 16), boxH = maxCH + 6;   // + up to 15 lead bytes (16-byte aligned box start)
        P.fwBoxW = boxW; P.fwBoxH = boxH;
        P.fwTileBytes = (int)orb_align_up((size_t)boxW * boxH, 128);
        P.fwScoreOff = P.fwTileBytes;                                 // one tile buffer: more warps per SM beat double buffering here
        // score map: pixel (px, py) of the domain (px < cell width, py < cell height) at byte (py + 1) * pitch + px + 2, one zero column
        // on either side for the 3x3 NMS: pitch >= cell width + 3.  Pitches 40 and 48 exist as compile-time variants.
        P.fwScorePitch = (boxW == 64 && ex->fastConst) ? (maxCW + 3 <= 40 ? 40 : 48) : P.scorePitch;
        P.fwPlistOff = P.fwScoreOff + (int)orb_align_up((size_t)P.fwScorePitch * P.scoreRows, 16);
        P.fwGlistOff = P.fwPlistOff + (int)orb_align_up((size_t)maxCW * maxCH * 2, 16);            // one u16 per domain pixel: every pixel may pass
        // group ring (64 x u32) during phase A; afterwards the same bytes hold the 3x3-NMS survivors (u16; at most one per 2x2 block)
        P.fwBarOff = P.fwGlistOff + (int)orb_align_up(std::max<size_t>(256, (size_t)((maxCW + 1) / 2 + 1) * ((maxCH + 1) / 2 + 1) * 2), 16);
        P.fwStride = (int)orb_align_up((size_t)P.fwBarOff + 16, 128);
        {   // warps per SM = what its shared memory holds (one CTA per SM in batch passes), at most 32
            cudaDeviceProp prop;
            ORB_CUDA_TRY(cudaGetDeviceProperties(&prop, ex->device));
            static const int envW = [] { const char* e = getenv("ORBX_FW_WARPS"); return e ? atoi(e) : 0; }();
            ex->fwWarps = std::min(ORBX_FW_MAXWARPS, (int)(prop.sharedMemPerBlockOptin / (size_t)P.fwStride));
            if (envW > 0) ex->fwWarps = std::min(ex->fwWarps, envW);
            ex->smCount = prop.multiProcessorCount;
        }
        ex->fwSmem = (size_t)P.fwStride * std::max(ex->fwWarps, 1);
        ex->useTma = false;
        const char* env = getenv("ORBX_FAST_TMA");
        const bool want = !(env && atoi(env) == 0);
        if (want && boxW <= 256 && boxH <= 256 && maxCW <= 62 && maxCH <= 62 && ex->fwWarps >= 1) {
            typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
                qres == cudaDriverEntryPointSuccess) {
                // maps [0, nl): FAST cell tiles of the pyramid; [nl, 2nl): 64 x 37 boxes of the blurred pyramid and [2nl, 3nl): 48 x 31
                // boxes of the pyramid, the two neighbourhoods k_describe<true> stages per keypoint
                std::vector<CUtensorMap> maps(3 * nl);
                bool ok = true;
                for (int l = 0; l < nl && ok; l++) {
                    const LevelPlan& L = P.lv[l];
                    cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)L.brows, (cuuint64_t)ex->maxBatch};
                    cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)P.frameBytes};
                    cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1};
                    cuuint32_t boxP[3] = {DESC_TMA_PW, 37, 1}, boxA[3] = {DESC_TMA_AW, 31, 1};
                    cuuint32_t estr[3] = {1, 1, 1};
                    ok = ((EncodeFn)fn)(&maps[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ex->d_pyr + L.off, dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS &&
                         ((EncodeFn)fn)(&maps[nl + l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ex->d_blur + L.off, dims, strides, boxP, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS &&
                         ((EncodeFn)fn)(&maps[2 * nl + l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ex->d_pyr + L.off, dims, strides, boxA, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
                }
                if (ok) {
                    ORB_CUDA_TRY(cudaMalloc(&ex->d_maps, 3 * nl * sizeof(CUtensorMap)));
                    ORB_CUDA_TRY(cudaMemcpy(ex->d_maps, maps.data(), 3 * nl * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
                    ex->useTma = true;
                }
            } else cudaGetLastError();
        }
        if (ex->useTma) {
            cudaDeviceProp prop;
            ORB_CUDA_TRY(cudaGetDeviceProperties(&prop, ex->device));
            ex->fwGrid = prop.multiProcessorCount;
            static std::mutex amu2;
            static size_t maxFwDev[64] = {0};                         // the opt-in is per function AND per device
            size_t& maxFw = maxFwDev[ex->device & 63];
            std::lock_guard<std::mutex> lk(amu2);
            if (ex->fwSmem > maxFw) {
                ORB_CUDA_TRY(cudaFuncSetAttribute(k_fast_tma<64, 40>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ex->fwSmem));
                ORB_CUDA_TRY(cudaFuncSetAttribute(k_fast_tma<64, ORBX_FAST_SPP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ex->fwSmem));
                ORB_CUDA_TRY(cudaFuncSetAttribute(k_fast_tma<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ex->fwSmem));
                maxFw = ex->fwSmem;
            }
        }
    }
    {   // k_describe: persistent grid, DESC_SMEM_BYTES of dynamic shared memory per CTA
        cudaDeviceProp prop;
        ORB_CUDA_TRY(cudaGetDeviceProperties(&prop, ex->device));
        static const bool wantDescTma = [] { const char* e = getenv("ORBX_DESC_TMA"); return !(e && atoi(e) == 0); }();
        ex->descTma = ex->useTma && wantDescTma;
        // warps per SM = what its shared memory holds next to ONE copy of the pattern table (one CTA per SM in batch passes)
        static const int envDW = [] { const char* e = getenv("ORBX_DESC_WARPS"); return e ? atoi(e) : 0; }();
        const size_t perWarp = ex->descTma ? (size_t)(2 * DESC_TMA_BUF_BYTES + 16) : (size_t)(2 * DESC_BUF_WORDS * 4);
        ex->descWarps = std::max(1, std::min(DESC_MAXWARPS, (int)((prop.sharedMemPerBlockOptin - 32 * 36 * 4) / perWarp)));
        if (envDW > 0) ex->descWarps = std::min(ex->descWarps, envDW);
        ex->descGrid = prop.multiProcessorCount;
        static std::mutex amu3;
        static bool descOptIn[64] = {false};
        std::lock_guard<std::mutex> lk(amu3);
        if (!descOptIn[ex->device & 63]) {
            ORB_CUDA_TRY(cudaFuncSetAttribute(k_describe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
            ORB_CUDA_TRY(cudaFuncSetAttribute(k_describe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
            descOptIn[ex->device & 63] = true;
        }
    }
    {   // dynamic shared-memory opt-in is per function, shared by all handles: only ever raise it
        static std::mutex amu;
        static size_t maxFastDev[64] = {0}, maxOctDev[64] = {0};     // per function AND per device
        size_t &maxFast = maxFastDev[ex->device & 63], &maxOct = maxOctDev[ex->device & 63];
        std::lock_guard<std::mutex> lk(amu);
        if (ex->fastSmem > maxFast) {
            ORB_CUDA_TRY(cudaFuncSetAttribute(k_fast<ORBX_FAST_TPP, ORBX_FAST_SPP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ex->fastSmem));
            ORB_CUDA_TRY(cudaFuncSetAttribute(k_fast<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ex->fastSmem));
            maxFast = ex->fastSmem;
        }
        if (ex->octSmem > maxOct) {
            ORB_CUDA_TRY(cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ex->octSmem));
            maxOct = ex->octSmem;
        }
    }
    return ORB_OK;
}

// ORBextractor::ORBextractor, src/ORBextractor.cc:414-445: mvScaleFactor[i] = mvScaleFactor[i-1] * scaleFactor (double product -> float),
// sigma^2, inverses; per-level quotas as a geometric series, cvRound per level, remainder to the last level
static void compute_tables(int nfeatures, float scale_factor, int nlevels, float* sf, float* isf, float* s2, float* is2, int* quota) {
    const double scaleFactor = scale_factor;
    std::vector<float> f(nlevels), q(nlevels);
    f[0] = 1.0f; q[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) { f[i] = (float)(f[i - 1] * scaleFactor); q[i] = f[i] * f[i]; }
    for (int i = 0; i < nlevels; i++) {
        if (sf) sf[i] = f[i];
        if (s2) s2[i] = q[i];
        if (isf) isf[i] = 1.0f / f[i];
        if (is2) is2[i] = 1.0f / q[i];
    }
    if (!quota) return;
    float factor = (float)(1.0f / scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) { quota[l] = h_cvRoundF(nDesired); sum += quota[l]; nDesired *= factor; }
    quota[nlevels - 1] = std::max(nfeatures - sum, 0);
}

extern "C" int orbx_compute_tables(int nfeatures, float scale_factor, int nlevels, float* scale, float* inv_scale, float* sigma2,
                                   float* inv_sigma2, int* quota) {
    ORB_REQUIRE(nfeatures > 0 && nlevels >= 1 && nlevels <= ORBX_MAX_LEVELS && scale_factor > 1.f, ORB_ERR_ARG, "bad extractor parameters");
    compute_tables(nfeatures, scale_factor, nlevels, scale, inv_scale, sigma2, inv_sigma2, quota);
    return ORB_OK;
}

extern "C" int orbx_create(orbx_extractor** out, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th,
                           int width, int height, int max_batch, int device) {
    ORB_REQUIRE(out, ORB_ERR_ARG, "out is NULL");
    *out = nullptr;
    ORB_REQUIRE(nfeatures > 0 && nlevels >= 1 && nlevels <= ORBX_MAX_LEVELS && scale_factor > 1.f && ini_th > 0 && min_th > 0 &&
                ini_th < 255 && min_th <= ini_th && width > 0 && height > 0 && max_batch >= 1,
                ORB_ERR_ARG, "bad extractor parameters");
    ORB_REQUIRE(orb_device_count() > device && device >= 0, ORB_ERR_CUDA, "CUDA device %d not available (no CPU fallback)", device);
    ORB_CUDA_TRY(cudaSetDevice(device));
    orbx_extractor* ex = new orbx_extractor();
    ex->nfeatures = nfeatures; ex->nlevels = nlevels; ex->iniTh = ini_th; ex->minTh = min_th; ex->device = device;
    ex->maxBatch = max_batch; ex->scaleFactor = scale_factor; ex->candPerCell = 0;   // 0 = size for the worst case
    if (const char* e = getenv("ORBX_CAND_PER_CELL")) ex->candPerCell = std::max(8, atoi(e));
    // scale tables and per-level quotas (ORBextractor.cc:414-445)
    ex->sf.resize(nlevels); ex->s2.resize(nlevels); ex->isf.resize(nlevels); ex->is2.resize(nlevels); ex->quota.resize(nlevels);
    compute_tables(nfeatures, scale_factor, nlevels, ex->sf.data(), ex->isf.data(), ex->s2.data(), ex->is2.data(), ex->quota.data());
    int rc = make_plan(ex, width, height);
    if (rc == ORB_OK && cudaStreamCreateWithFlags(&ex->stream, cudaStreamNonBlocking) != cudaSuccess) {
        orb_set_error("cudaStreamCreate failed"); rc = ORB_ERR_CUDA;
    }
    if (rc != ORB_OK) { orbx_destroy(ex); return rc; }
    *out = ex;
    return ORB_OK;
}

extern "C" void orbx_destroy(orbx_extractor* ex) {
    if (!ex) return;
    for (int i = 0; i < ORBX_MAX_HELPERS; i++) if (ex->helpers[i]) { orbx_destroy(ex->helpers[i]); ex->helpers[i] = nullptr; }
    cudaSetDevice(ex->device);
    cudaFree(ex->d_pyr); cudaFree(ex->d_blur); cudaFree(ex->d_stereo);
    for (int i = 0; i < ORBX_MAX_SLOTS; i++) {
        cudaFree(ex->d_in[i]); cudaFree(ex->d_mask[i]); cudaFree(ex->d_desc[i]); cudaFree(ex->d_kp[i]); cudaFree(ex->d_n[i]);
        if (ex->evH2D[i]) cudaEventDestroy(ex->evH2D[i]);
        if (ex->evComp[i]) cudaEventDestroy(ex->evComp[i]);
        if (ex->evD2H[i]) cudaEventDestroy(ex->evD2H[i]);
    }
    if (ex->h_n) cudaFreeHost(ex->h_n);
    if (ex->h_status) cudaFreeHost(ex->h_status);
    if (ex->h_stage) cudaFreeHost(ex->h_stage);
    for (int i = 0; i < ORBX_MAX_SLOTS; i++) { if (ex->h_in[i]) cudaFreeHost(ex->h_in[i]); if (ex->h_out[i]) cudaFreeHost(ex->h_out[i]); }
    delete ex->pool;
    if (ex->sH2D) cudaStreamDestroy(ex->sH2D);
    if (ex->sD2H) cudaStreamDestroy(ex->sD2H);
    cudaFree(ex->d_cand); cudaFree(ex->d_sel); cudaFree(ex->d_nodeOf); cudaFree(ex->d_candCount); cudaFree(ex->d_workCounter); cudaFree(ex->d_selCount);
    cudaFree(ex->d_status); cudaFree(ex->d_xtab); cudaFree(ex->d_ytab); cudaFree(ex->d_cells); cudaFree(ex->d_maps);
    if (ex->stream) cudaStreamDestroy(ex->stream);
    if (ex->sBlur) cudaStreamDestroy(ex->sBlur);
    if (ex->evFork) cudaEventDestroy(ex->evFork);
    if (ex->evJoin) cudaEventDestroy(ex->evJoin);
    delete ex;
}

extern "C" int orbx_tables(const orbx_extractor* ex, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* quota) {
    ORB_REQUIRE(ex, ORB_ERR_ARG, "null handle");
    for (int i = 0; i < ex->nlevels; i++) {
        if (scale) scale[i] = ex->sf[i];
        if (inv_scale) inv_scale[i] = ex->isf[i];
        if (sigma2) sigma2[i] = ex->s2[i];
        if (inv_sigma2) inv_sigma2[i] = ex->is2[i];
        if (quota) quota[i] = ex->quota[i];
    }
    return ORB_OK;
}
extern "C" int orbx_level_size(const orbx_extractor* ex, int level, int* w, int* h) {
    ORB_REQUIRE(ex && level >= 0 && level < ex->nlevels, ORB_ERR_ARG, "bad level");
    if (w) *w = ex->plan.lv[level].w;
    if (h) *h = ex->plan.lv[level].h;
    return ORB_OK;
}
extern "C" int orbx_max_keypoints(const orbx_extractor* ex) { return ex ? ex->capInternal : 0; }
extern "C" long long orbx_launch_count(const orbx_extractor* ex) {
    if (!ex) return 0;
    long long n = ex->launches;
    for (int i = 0; i < ORBX_MAX_HELPERS; i++) if (ex->helpers[i]) n += ex->helpers[i]->launches;
    return n;
}

// One device pass over nf <= maxBatch frames.
static int run_pass(orbx_extractor* ex, const u8* d_images, const u8* d_masks, int nf, orbx_keypoint* d_kp, u8* d_desc,
                    int cap, int* d_n, int stages, cudaStream_t st) {
    const Plan& P = ex->plan;
    const int nl = P.nlevels;
    if (stages & ORBX_STAGE_PYRAMID) {
        {
            const LevelPlan& L = P.lv[0];
            const bool v16 = (L.w % 16 == 0) && ((uintptr_t)d_images % 16 == 0) && (!d_masks || (uintptr_t)d_masks % 16 == 0) &&
                             (unsigned long long)(L.w >> 4) * (L.w >> 4) * L.h < (1ull << 32);
            if (v16) {
                const u32 gpr = (u32)L.w >> 4;
                dim3 g(orb_div_up((int)gpr * L.h, 256 * L0_UNROLL), nf);
                k_level0_v16<<<g, 256, 0, st>>>(P, d_images, d_masks, ex->d_pyr, 0xFFFFFFFFu / gpr + 1);
            } else {
                dim3 g(orb_div_up(L.w, 256), orb_div_up(L.h, 4), nf), b(64, 4);
                k_level0<<<g, b, 0, st>>>(P, d_images, d_masks, ex->d_pyr);
            }
            ex->launches++;
        }
        for (int l = 1; l < nl; l++) {
            const LevelPlan& L = P.lv[l];
            dim3 g(orb_div_up(L.w, 256), orb_div_up(L.h, 4), nf), b(64, 4);
            // y = idx / wq as umulhi(idx, 2^32 / wq + 1) is exact while idx * wq < 2^32 (true up to ~6000 x 6000 levels)
            if (L.fastResize && (unsigned long long)((L.w + 3) >> 2) * ((L.w + 3) >> 2) * L.h < (1ull << 32)) {
                const u32 wq = (u32)(L.w + 3) >> 2;
                dim3 gf(orb_div_up((int)wq * ((L.h + 1) >> 1), 256), nf);
                k_resize<<<gf, 256, 0, st>>>(P, l, ex->d_pyr, ex->d_xtab, ex->d_ytab, 0xFFFFFFFFu / wq + 1);
            }
            else k_resize_generic<<<g, b, 0, st>>>(P, l, ex->d_pyr, ex->d_xtab, ex->d_ytab);
            ex->launches++;
        }
        // (no border pass: nothing on the hot path reads border pixels — the blur reflects its own taps, which costs it 0.09 us per
        // frame; a separate pass over the border words costs 0.18 us per frame because every word it touches is a DRAM sector of its
        // own.  The 19-pixel border of the API-visible pyramid is produced on demand by orbx_get_pyramid(bordered).)
    }
    // The blur depends on the pyramid only.  In a complete pass it is forked onto a second stream next to the kernels that cannot fill
    // the machine: FAST when the pass has fewer cells than one CTA per SM takes (single frames), else the octree when it has fewer than
    // two CTAs per SM (one CTA per (level, frame): 4K passes, single frames) — k_octree at 4K runs at 30 % SM throughput for 11 % of the
    // pass.  In VGA batch passes every kernel fills the machine and the streams would only take turns: no fork.
    int forkAt = 0;                                                    // 0: none, 1: before FAST, 2: before the octree
    static const bool wantFork = [] { const char* e = getenv("ORBX_BLUR_FORK"); return !(e && e[0] == '0'); }();
    if (wantFork && stages == ORBX_STAGE_ALL) {
        if (ex->useTma && ex->nCells * nf < ex->smCount * 4) forkAt = 1;
        else if (nl * nf < 2 * ex->smCount) forkAt = 2;
        if (forkAt && !ex->sBlur) {
            ORB_CUDA_TRY(cudaStreamCreateWithFlags(&ex->sBlur, cudaStreamNonBlocking));
            ORB_CUDA_TRY(cudaEventCreateWithFlags(&ex->evFork, cudaEventDisableTiming));
            ORB_CUDA_TRY(cudaEventCreateWithFlags(&ex->evJoin, cudaEventDisableTiming));
        }
    }
    auto launch_blur = [&](cudaStream_t sb) {
        dim3 g(P.blurTileBase[nl], nf);
        k_blur<<<g, 256, 0, sb>>>(P, ex->d_pyr, ex->d_blur);
        ex->launches++;
    };
    auto fork_blur = [&]() -> int {
        ORB_CUDA_TRY(cudaEventRecord(ex->evFork, st));
        ORB_CUDA_TRY(cudaStreamWaitEvent(ex->sBlur, ex->evFork, 0));
        launch_blur(ex->sBlur);
        ORB_CUDA_TRY(cudaEventRecord(ex->evJoin, ex->sBlur));
        return ORB_OK;
    };
    if (forkAt == 1) { int rcf = fork_blur(); if (rcf != ORB_OK) return rcf; }
    if (stages & ORBX_STAGE_FAST) {
        ORB_CUDA_TRY(cudaMemsetAsync(ex->d_candCount, 0, (size_t)nf * nl * sizeof(int), st));
        ORB_CUDA_TRY(cudaMemsetAsync(ex->d_workCounter, 0, sizeof(int), st));
        dim3 g(ex->nCells, nf);
        if (ex->nCells > 0 && ex->useTma) {
            const int items = ex->nCells * nf;
            // batch passes: one CTA of fwWarps warps per SM; small passes (single frames): fewer warps per CTA so that every SM gets work
            const int wpc = std::min(ex->fwWarps, std::max(4, orb_div_up(items, ex->smCount)));
            const int grid = std::min(ex->smCount * std::max(1, ex->fwWarps / wpc), orb_div_up(items, wpc));
            const size_t smem = (size_t)P.fwStride * wpc;
            const unsigned long long inv40 = ((1ull << 40) + (unsigned long long)ex->nCells - 1) / (unsigned long long)ex->nCells;
            if (P.fwBoxW == 64 && ex->fastConst && P.fwScorePitch == 40)
                k_fast_tma<64, 40><<<grid, 32 * wpc, smem, st>>>(P, ex->d_maps, ex->d_cells, ex->nCells, nf, ex->d_cand,
                                                                 ex->d_candCount, ex->d_status, ex->d_workCounter, inv40);
            else if (P.fwBoxW == 64 && ex->fastConst)
                k_fast_tma<64, ORBX_FAST_SPP><<<grid, 32 * wpc, smem, st>>>(P, ex->d_maps, ex->d_cells, ex->nCells, nf, ex->d_cand,
                                                                            ex->d_candCount, ex->d_status, ex->d_workCounter, inv40);
            else
                k_fast_tma<0, 0><<<grid, 32 * wpc, smem, st>>>(P, ex->d_maps, ex->d_cells, ex->nCells, nf, ex->d_cand,
                                                               ex->d_candCount, ex->d_status, ex->d_workCounter, inv40);
        } else if (ex->nCells > 0) {
            if (ex->fastConst)
                k_fast<ORBX_FAST_TPP, ORBX_FAST_SPP><<<g, ORBX_FAST_THREADS, ex->fastSmem, st>>>(P, ex->d_cells, ex->d_pyr, ex->d_cand, ex->d_candCount, ex->d_status);
            else
                k_fast<0, 0><<<g, ORBX_FAST_THREADS, ex->fastSmem, st>>>(P, ex->d_cells, ex->d_pyr, ex->d_cand, ex->d_candCount, ex->d_status);
        }
        ex->launches++;
    }
    if (forkAt == 2) { int rcf = fork_blur(); if (rcf != ORB_OK) return rcf; }
    if (stages & ORBX_STAGE_OCTREE) {
        dim3 g(nl, nf);
        // a (level, frame) CTA at VGA size holds ~1400 candidates at most: in a batch pass (plenty of CTAs) 128 threads keep the many
        // barrier-separated passes short (0.37 vs 0.63 us/frame); large images and single-frame latency calls want the full 512
        // (1024 threads for large images: a 4K level 0 holds ~100 k candidates and the key loops are pure memory latency)
        const int octT = (P.width * P.height <= 500000 && nf >= 16) ? 128 : (P.width * P.height <= 500000 ? 512 : ORBX_OCT_THREADS);
        k_octree<<<g, octT, ex->octSmem, st>>>(P, ex->d_cand, ex->d_candCount, ex->d_nodeOf, ex->d_sel,
                                                           ex->d_selCount, ex->d_status);
        ex->launches++;
    }
    if (forkAt) ORB_CUDA_TRY(cudaStreamWaitEvent(st, ex->evJoin, 0));
    else if (stages & ORBX_STAGE_BLUR) launch_blur(st);
    if (stages & ORBX_STAGE_DESCRIBE) {
        const int items = nf * P.selTotal;
        // batch passes: one CTA of descWarps warps per SM; small passes (single frames): fewer warps per CTA so that every SM gets work
        const int wpc = std::min(ex->descWarps, std::max(4, orb_div_up(items, ex->smCount)));
        const int grid = std::min(ex->smCount * std::max(1, ex->descWarps / wpc), orb_div_up(items, wpc));
        ORB_CUDA_TRY(cudaMemsetAsync(ex->d_workCounter + 1, 0, sizeof(int), st));
        static const int descChunk = [] { const char* e = getenv("ORBX_DESC_CHUNK"); return std::max(1, e ? atoi(e) : 4); }();
        if (ex->descTma)
            k_describe<true><<<grid, 32 * wpc, DESC_TMA_SMEM_BYTES(wpc), st>>>(P, ex->d_pyr, ex->d_blur, ex->d_sel, ex->d_selCount, d_kp, d_desc,
                                                                              d_n, cap, ex->d_status, nf, ex->d_maps, ex->d_workCounter + 1, descChunk);
        else
            k_describe<false><<<grid, 32 * wpc, DESC_SMEM_BYTES(wpc), st>>>(P, ex->d_pyr, ex->d_blur, ex->d_sel, ex->d_selCount, d_kp, d_desc,
                                                                           d_n, cap, ex->d_status, nf, nullptr, ex->d_workCounter + 1, descChunk);
        ex->launches++;
    }
    ORB_CUDA_TRY(cudaGetLastError());
    ex->lastFrames = nf;
    return ORB_OK;
}

extern "C" int orbx_run_stages_device(orbx_extractor* ex, const uint8_t* d_images, int n_frames, const uint8_t* d_masks,
                                      orbx_keypoint* d_kp_out, uint8_t* d_desc_out, int cap, int* d_n_out, int stage_mask,
                                      void* stream) {
    ORB_REQUIRE(ex && d_images && d_kp_out && d_desc_out && d_n_out && n_frames >= 0 && cap > 0, ORB_ERR_ARG, "bad arguments");
    std::lock_guard<std::mutex> lk(ex->mu);
    ORB_CUDA_TRY(cudaSetDevice(ex->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ex->stream;
    const size_t fpx = (size_t)ex->plan.width * ex->plan.height;
    // (Alternating the passes between two handles on forked streams was measured and is slower here: 167 k vs 170 k frames/s —
    // two concurrent passes double the working set; profiles/README.md.)
    for (int f0 = 0; f0 < n_frames; f0 += ex->maxBatch) {
        const int nf = std::min(ex->maxBatch, n_frames - f0);
        int rc = run_pass(ex, d_images + f0 * fpx, d_masks ? d_masks + f0 * fpx : nullptr, nf, d_kp_out + (size_t)f0 * cap,
                          d_desc_out + (size_t)f0 * cap * 32, cap, d_n_out + f0, stage_mask, st);
        if (rc != ORB_OK) return rc;
    }
    return ORB_OK;
}

extern "C" int orbx_extract_batch_device(orbx_extractor* ex, const uint8_t* d_images, int n_frames, const uint8_t* d_masks,
                                         orbx_keypoint* d_kp_out, uint8_t* d_desc_out, int cap, int* d_n_out, void* stream) {
    return orbx_run_stages_device(ex, d_images, n_frames, d_masks, d_kp_out, d_desc_out, cap, d_n_out, ORBX_STAGE_ALL, stream);
}

extern "C" int orbx_check_status(orbx_extractor* ex) {
    ORB_REQUIRE(ex, ORB_ERR_ARG, "null handle");
    int s = 0;
    ORB_CUDA_TRY(cudaSetDevice(ex->device));
    ORB_CUDA_TRY(cudaMemcpy(&s, ex->d_status, sizeof(int), cudaMemcpyDeviceToHost));
    if (s) ORB_CUDA_TRY(cudaMemset(ex->d_status, 0, sizeof(int)));
    for (int i = 0; i < ORBX_MAX_HELPERS; i++)
        if (ex->helpers[i]) {
            int s2 = 0;
            ORB_CUDA_TRY(cudaMemcpy(&s2, ex->helpers[i]->d_status, sizeof(int), cudaMemcpyDeviceToHost));
            if (s2) ORB_CUDA_TRY(cudaMemset(ex->helpers[i]->d_status, 0, sizeof(int)));
            s |= s2;
        }
    ORB_REQUIRE(!(s & (ORB_DEV_CAND_OVERFLOW | ORB_DEV_NODE_OVERFLOW)), ORB_ERR_OVERFLOW,
                "FAST candidate buffer overflow (ORBX_CAND_PER_CELL=%d caps it; unset it to size for the worst case)", ex->candPerCell);
    ORB_REQUIRE(!(s & ORB_DEV_OUT_OVERFLOW), ORB_ERR_CAPACITY, "keypoint output capacity too small (need orbx_max_keypoints())");
    return ORB_OK;
}

// `image_ptrs` (or NULL): one pointer per frame instead of base + frame_stride — frames held as separate allocations
// (std::vector<cv::Mat>); `images` then only stands for "frame 0" in the pinned / pageable test.
static int extract_batch_impl(orbx_extractor* ex, const uint8_t* images, const uint8_t* const* image_ptrs, int n_frames, int width, int height,
                              int stride, size_t frame_stride, const uint8_t* masks, int mask_stride, size_t mask_frame_stride,
                              orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* n_out) {
    ORB_REQUIRE(ex && images && kp_out && desc_out && n_out && n_frames >= 0 && cap > 0, ORB_ERR_ARG, "bad arguments");
    auto frame_at = [&](int f) -> const u8* { return image_ptrs ? image_ptrs[f] : images + (size_t)f * frame_stride; };
    ORB_REQUIRE(width == ex->plan.width && height == ex->plan.height, ORB_ERR_ARG,
                "image is %dx%d but the handle was planned for %dx%d", width, height, ex->plan.width, ex->plan.height);
    ORB_REQUIRE(stride >= width && (!masks || mask_stride >= width), ORB_ERR_ARG, "stride < width");
    if (n_frames == 0) return ORB_OK;
    const int B = ex->maxBatch, icap = ex->capInternal;
    const size_t fpx = (size_t)width * height;
    std::lock_guard<std::mutex> lk(ex->mu);
    ORB_CUDA_TRY(cudaSetDevice(ex->device));
    static const int maxSlots = [] { const char* e = getenv("ORBX_E2E_SLOTS"); const int v = e ? atoi(e) : ORBX_MAX_SLOTS;
                                     return std::max(2, std::min(ORBX_MAX_SLOTS, v)); }();      // must exceed the number of compute streams
    const int nslots = n_frames > B ? maxSlots : 1;
    for (int i = 0; i < nslots; i++) {
        if (!ex->d_in[i]) {
            ORB_CUDA_TRY(cudaMalloc(&ex->d_in[i], (size_t)B * fpx));
            ORB_CUDA_TRY(cudaMalloc(&ex->d_kp[i], (size_t)B * icap * sizeof(orbx_keypoint)));
            ORB_CUDA_TRY(cudaMalloc(&ex->d_desc[i], (size_t)B * icap * 32));
            ORB_CUDA_TRY(cudaMalloc(&ex->d_n[i], (size_t)B * sizeof(int)));
            ORB_CUDA_TRY(cudaEventCreateWithFlags(&ex->evH2D[i], cudaEventDisableTiming));
            ORB_CUDA_TRY(cudaEventCreateWithFlags(&ex->evComp[i], cudaEventDisableTiming));
            ORB_CUDA_TRY(cudaEventCreateWithFlags(&ex->evD2H[i], cudaEventDisableTiming));
        }
        if (masks && !ex->d_mask[i]) ORB_CUDA_TRY(cudaMalloc(&ex->d_mask[i], (size_t)B * fpx));
    }
    if (!ex->sH2D) {
        ORB_CUDA_TRY(cudaStreamCreateWithFlags(&ex->sH2D, cudaStreamNonBlocking));
        ORB_CUDA_TRY(cudaStreamCreateWithFlags(&ex->sD2H, cudaStreamNonBlocking));
    }
    if (ex->h_n_cap < (size_t)n_frames) {
        if (ex->h_n) ORB_CUDA_TRY(cudaFreeHost(ex->h_n));
        ex->h_n = nullptr; ex->h_n_cap = 0;
        ORB_CUDA_TRY(cudaMallocHost(&ex->h_n, (size_t)n_frames * sizeof(int)));
        ex->h_n_cap = n_frames;
    }
    const int ccap = std::min(cap, icap);             // keypoints copied back per frame (the rest of a row cannot be used)
    if (nslots > 1) {
        static const int nStreams = [] { const char* e = getenv("ORBX_E2E_STREAMS"); const int v = e ? atoi(e) : 2;
                                         return std::max(1, std::min(1 + ORBX_MAX_HELPERS, v)); }();
        for (int i = 0; i < nStreams - 1; i++)
            if (!ex->helpers[i]) {
                int rc = orbx_create(&ex->helpers[i], ex->nfeatures, (float)ex->scaleFactor, ex->nlevels, ex->iniTh, ex->minTh, width, height,
                                     B, ex->device);
                if (rc != ORB_OK) return rc;
            }
        ex->nHelpers = nStreams - 1;
    }
    if (!ex->h_status) ORB_CUDA_TRY(cudaMallocHost(&ex->h_status, sizeof(int)));
    if (n_frames <= B) {
        // ---- latency path (one device pass, e.g. the per-frame call of Frame::ExtractORB): a single stream, a single
        // synchronisation; pageable caller buffers (cv::Mat, std::vector) go through one pinned staging block so that every copy
        // is a true asynchronous DMA
        cudaStream_t st = ex->stream;
        auto pinned = [](const void* p) {
            cudaPointerAttributes a;
            if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
            return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
        };
        const bool tight = !image_ptrs && (size_t)stride == (size_t)width && frame_stride == fpx;
        const bool stageIn = !(tight && pinned(images)) && (size_t)n_frames * fpx <= (4u << 20);   // large frames: direct copy is cheaper
        const bool stageOut = !(pinned(kp_out) && pinned(desc_out)) && (size_t)n_frames * icap <= 65536;
        const size_t inBytes = (size_t)n_frames * fpx, kpBytes = (size_t)n_frames * icap * sizeof(orbx_keypoint), dBytes = (size_t)n_frames * icap * 32;
        const size_t need = (stageIn ? orb_align_up(inBytes, 256) : 0) + (stageOut ? orb_align_up(kpBytes, 256) + dBytes : 0);
        if (need > ex->h_stage_cap) {
            if (ex->h_stage) ORB_CUDA_TRY(cudaFreeHost(ex->h_stage));
            ex->h_stage = nullptr; ex->h_stage_cap = 0;
            ORB_CUDA_TRY(cudaMallocHost(&ex->h_stage, need));
            ex->h_stage_cap = need;
        }
        u8* hIn = ex->h_stage;
        u8* hKp = ex->h_stage + (stageIn ? orb_align_up(inBytes, 256) : 0);
        u8* hDesc = hKp + orb_align_up(kpBytes, 256);
        if (stageIn) {
            for (int f = 0; f < n_frames; f++)
                for (int y = 0; y < height; y++)
                    memcpy(hIn + f * fpx + (size_t)y * width, frame_at(f) + (size_t)y * stride, width);
            ORB_CUDA_TRY(cudaMemcpyAsync(ex->d_in[0], hIn, inBytes, cudaMemcpyHostToDevice, st));
        } else if (tight) {
            ORB_CUDA_TRY(cudaMemcpyAsync(ex->d_in[0], images, inBytes, cudaMemcpyHostToDevice, st));
        } else {
            for (int f = 0; f < n_frames; f++)
                ORB_CUDA_TRY(cudaMemcpy2DAsync(ex->d_in[0] + f * fpx, width, frame_at(f), stride, width, height,
                                               cudaMemcpyHostToDevice, st));
        }
        if (masks)
            for (int f = 0; f < n_frames; f++)
                ORB_CUDA_TRY(cudaMemcpy2DAsync(ex->d_mask[0] + f * fpx, width, masks + (size_t)f * mask_frame_stride, mask_stride, width,
                                               height, cudaMemcpyHostToDevice, st));
        int rc = run_pass(ex, ex->d_in[0], masks ? ex->d_mask[0] : nullptr, n_frames, ex->d_kp[0], ex->d_desc[0], icap, ex->d_n[0],
                          ORBX_STAGE_ALL, st);
        if (rc != ORB_OK) return rc;
        ORB_CUDA_TRY(cudaMemcpyAsync(ex->h_n, ex->d_n[0], n_frames * sizeof(int), cudaMemcpyDeviceToHost, st));
        ORB_CUDA_TRY(cudaMemcpyAsync(ex->h_status, ex->d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
        if (stageOut) {
            ORB_CUDA_TRY(cudaMemcpyAsync(hKp, ex->d_kp[0], kpBytes, cudaMemcpyDeviceToHost, st));
            ORB_CUDA_TRY(cudaMemcpyAsync(hDesc, ex->d_desc[0], dBytes, cudaMemcpyDeviceToHost, st));
        } else {
            ORB_CUDA_TRY(cudaMemcpy2DAsync(kp_out, (size_t)cap * sizeof(orbx_keypoint), ex->d_kp[0], (size_t)icap * sizeof(orbx_keypoint),
                                           (size_t)ccap * sizeof(orbx_keypoint), n_frames, cudaMemcpyDeviceToHost, st));
            ORB_CUDA_TRY(cudaMemcpy2DAsync(desc_out, (size_t)cap * 32, ex->d_desc[0], (size_t)icap * 32, (size_t)ccap * 32, n_frames,
                                           cudaMemcpyDeviceToHost, st));
        }
        ORB_CUDA_TRY(cudaStreamSynchronize(st));
        const int sflags = *ex->h_status;
        if (sflags) ORB_CUDA_TRY(cudaMemset(ex->d_status, 0, sizeof(int)));
        ORB_REQUIRE(!(sflags & (ORB_DEV_CAND_OVERFLOW | ORB_DEV_NODE_OVERFLOW)), ORB_ERR_OVERFLOW,
                    "FAST candidate buffer overflow (ORBX_CAND_PER_CELL=%d caps it; unset it to size for the worst case)", ex->candPerCell);
        for (int f = 0; f < n_frames; f++) {
            const int n = ex->h_n[f];
            n_out[f] = n;
            ORB_REQUIRE(n <= cap, ORB_ERR_CAPACITY, "frame %d has %d keypoints but cap is %d", f, n, cap);
            if (stageOut && n > 0) {
                memcpy(kp_out + (size_t)f * cap, hKp + (size_t)f * icap * sizeof(orbx_keypoint), (size_t)n * sizeof(orbx_keypoint));
                memcpy(desc_out + (size_t)f * cap * 32, hDesc + (size_t)f * icap * 32, (size_t)n * 32);
            }
        }
        return ORB_OK;
    }
    // Chunk schedule.  Measured with ORBX_E2E_TRACE (tools/e2e_trace.py): the compute stream is the busy one (a pass costs about
    // 0.18 ms + 5.7 us per 640x480 frame, the copy 5.6 us per frame), so the step is  fill + sum of passes + drain  with
    // fill = H2D of the first chunk and drain = D2H of the last one.  Few large passes minimise the sum; a small first and last
    // chunk (64 frames) minimise fill and drain.  ORBX_E2E_RAMP=0 restores equal chunks (A/B measurements).
    std::vector<int> sizes;
    {
        static const bool ramp = [] { const char* e = getenv("ORBX_E2E_RAMP"); return !(e && e[0] == '0'); }();
        const int edge = (ramp && B > 64 && n_frames >= 2 * B) ? 64 : 0;
        if (edge) sizes.push_back(edge);
        for (int mid = n_frames - 2 * edge; mid > 0; mid -= B) sizes.push_back(std::min(B, mid));
        if (edge) sizes.push_back(edge);
    }
    // ORBX_E2E_TRACE=1: per-chunk device timeline of the three streams on stderr (a profiling aid, not used by the product path)
    static const bool trace = [] { const char* e = getenv("ORBX_E2E_TRACE"); return e && e[0] == '1'; }();
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t s) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        tev.push_back(e);
    };
    // Pageable caller memory (what the reference's callers hold: cv::Mat, std::vector): cudaMemcpyAsync would fall back to the driver's
    // single-threaded staging and block the issuing thread.  Instead the frames of a chunk are copied into the slot's pinned staging by the
    // host pool (rows packed on the way, so strided images cost nothing extra) and the results leave through pinned staging too, only the
    // n valid keypoints of every frame being copied on to the caller.  ORBX_HOST_THREADS sets the pool size.
    auto pageable = [](const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
        return a.type == cudaMemoryTypeUnregistered;
    };
    static const int poolThreads = [] {
        const char* e = getenv("ORBX_HOST_THREADS");
        // default: three quarters of the hardware threads (measured on a 16-thread host: 26 k / 87 k / 117 k / 134 k / 131 k frames/s with
        // 1 / 4 / 8 / 12 / 16 threads), shared between the processes of a one-process-per-GPU launch (torchrun's LOCAL_WORLD_SIZE)
        const int hw = (int)std::thread::hardware_concurrency();
        const char* lw = getenv("LOCAL_WORLD_SIZE");
        const int procs = std::max(1, lw ? atoi(lw) : 1);
        const int v = e ? atoi(e) : std::min(16, std::max(1, hw * 3 / 4 / procs));
        return std::max(1, std::min(64, v));
    }();
    const bool stageIn = poolThreads > 0 && pageable(images);
    const bool stageOut = poolThreads > 0 && (pageable(kp_out) || pageable(desc_out));
    const size_t kpRow = (size_t)ccap * sizeof(orbx_keypoint), dRow = (size_t)ccap * 32, outKp = orb_align_up((size_t)B * kpRow, 256);
    if (stageIn || stageOut) {
        if (!ex->pool) ex->pool = new HostPool(poolThreads);
        for (int i = 0; i < nslots; i++) {
            if (stageIn && !ex->h_in[i]) ORB_CUDA_TRY(cudaMallocHost(&ex->h_in[i], (size_t)B * fpx));
            if (stageOut && !ex->h_out[i]) ORB_CUDA_TRY(cudaMallocHost(&ex->h_out[i], (size_t)B * icap * (sizeof(orbx_keypoint) + 32) + 256));
        }
    }
    std::vector<int> slotF0(nslots, -1), slotNf(nslots, 0);               // the chunk whose results sit in a slot's output staging
    auto drain = [&](int sl) -> int {                                      // pinned staging -> caller arrays, valid entries only
        if (slotF0[sl] < 0) return ORB_OK;
        ORB_CUDA_TRY(cudaEventSynchronize(ex->evD2H[sl]));
        const int b0 = slotF0[sl];
        const u8 *hk = ex->h_out[sl], *hd = ex->h_out[sl] + outKp;
        ex->pool->run(slotNf[sl], [&](int f) {
            const int n = std::min(std::max(ex->h_n[b0 + f], 0), ccap);
            memcpy(kp_out + (size_t)(b0 + f) * cap, hk + (size_t)f * kpRow, (size_t)n * sizeof(orbx_keypoint));
            memcpy(desc_out + (size_t)(b0 + f) * cap * 32, hd + (size_t)f * dRow, (size_t)n * 32);
        });
        slotF0[sl] = -1;
        return ORB_OK;
    };
    int chunk = 0, f0 = 0;
    for (size_t ci = 0; ci < sizes.size(); f0 += sizes[ci], ci++, chunk++) {
        const int nf = sizes[ci], sl = chunk % nslots;
        // the handle (stream + workspace) computing this chunk; the last chunk is always ex's own, so that the views of the last pass
        // (orbx_get_pyramid*, orbx_stereo_matches) keep their meaning
        const int hsel = (int)((sizes.size() - 1 - ci) % (size_t)(ex->nHelpers + 1));
        orbx_extractor* cx = hsel ? ex->helpers[hsel - 1] : ex;
        // ---- H2D (slot's input buffer is free once the compute of chunk-nslots has finished)
        if (chunk >= nslots) ORB_CUDA_TRY(cudaStreamWaitEvent(ex->sH2D, ex->evComp[sl], 0));
        mark(ex->sH2D);
        if (stageIn) {
            if (chunk >= nslots) ORB_CUDA_TRY(cudaEventSynchronize(ex->evH2D[sl]));        // the slot's previous copy has left the staging
            u8* hin = ex->h_in[sl];
            const int rowsPer = std::max(1, (int)((256u << 10) / (size_t)width));            // ~256 KB work items
            const int perFrame = orb_div_up(height, rowsPer);
            ex->pool->run(nf * perFrame, [&](int it) {
                const int f = it / perFrame, y0 = (it - f * perFrame) * rowsPer, y1 = std::min(height, y0 + rowsPer);
                const u8* src = frame_at(f0 + f);
                u8* dst = hin + (size_t)f * fpx;
                static const bool nt = [] { const char* e = getenv("ORBX_HOST_NT"); return !(e && e[0] == '0'); }();
                if (!nt) {
                    if (stride == width) memcpy(dst + (size_t)y0 * width, src + (size_t)y0 * width, (size_t)(y1 - y0) * width);
                    else for (int y = y0; y < y1; y++) memcpy(dst + (size_t)y * width, src + (size_t)y * stride, width);
                } else if (stride == width) copy_stream(dst + (size_t)y0 * width, src + (size_t)y0 * width, (size_t)(y1 - y0) * width);
                else for (int y = y0; y < y1; y++) copy_stream(dst + (size_t)y * width, src + (size_t)y * stride, width);
            });
            ORB_CUDA_TRY(cudaMemcpyAsync(ex->d_in[sl], hin, (size_t)nf * fpx, cudaMemcpyHostToDevice, ex->sH2D));
        } else if (!image_ptrs && (size_t)stride == (size_t)width && frame_stride == fpx) {
            ORB_CUDA_TRY(cudaMemcpyAsync(ex->d_in[sl], images + (size_t)f0 * frame_stride, (size_t)nf * fpx, cudaMemcpyHostToDevice, ex->sH2D));
        } else {
            for (int f = 0; f < nf; f++)
                ORB_CUDA_TRY(cudaMemcpy2DAsync(ex->d_in[sl] + f * fpx, width, frame_at(f0 + f), stride, width,
                                               height, cudaMemcpyHostToDevice, ex->sH2D));
        }
        if (masks)
            for (int f = 0; f < nf; f++)
                ORB_CUDA_TRY(cudaMemcpy2DAsync(ex->d_mask[sl] + f * fpx, width, masks + (size_t)(f0 + f) * mask_frame_stride, mask_stride,
                                               width, height, cudaMemcpyHostToDevice, ex->sH2D));
        mark(ex->sH2D);
        ORB_CUDA_TRY(cudaEventRecord(ex->evH2D[sl], ex->sH2D));
        // ---- compute (slot's output buffers are free once the D2H of chunk-nslots has finished)
        ORB_CUDA_TRY(cudaStreamWaitEvent(cx->stream, ex->evH2D[sl], 0));
        if (chunk >= nslots) ORB_CUDA_TRY(cudaStreamWaitEvent(cx->stream, ex->evD2H[sl], 0));
        mark(cx->stream);
        int rc = run_pass(cx, ex->d_in[sl], masks ? ex->d_mask[sl] : nullptr, nf, ex->d_kp[sl], ex->d_desc[sl], icap, ex->d_n[sl],
                          ORBX_STAGE_ALL, cx->stream);
        if (rc != ORB_OK) return rc;
        mark(cx->stream);
        ORB_CUDA_TRY(cudaEventRecord(ex->evComp[sl], cx->stream));
        // ---- D2H: counts, then the padded keypoint / descriptor rows in one 2-D copy each
        ORB_CUDA_TRY(cudaStreamWaitEvent(ex->sD2H, ex->evComp[sl], 0));
        mark(ex->sD2H);
        ORB_CUDA_TRY(cudaMemcpyAsync(ex->h_n + f0, ex->d_n[sl], nf * sizeof(int), cudaMemcpyDeviceToHost, ex->sD2H));
        if (stageOut) {
            int rcd = drain(sl);                                      // the slot's previous results go to the caller first
            if (rcd != ORB_OK) return rcd;
            ORB_CUDA_TRY(cudaMemcpy2DAsync(ex->h_out[sl], kpRow, ex->d_kp[sl], (size_t)icap * sizeof(orbx_keypoint), kpRow, nf,
                                           cudaMemcpyDeviceToHost, ex->sD2H));
            ORB_CUDA_TRY(cudaMemcpy2DAsync(ex->h_out[sl] + outKp, dRow, ex->d_desc[sl], (size_t)icap * 32, dRow, nf, cudaMemcpyDeviceToHost,
                                           ex->sD2H));
            slotF0[sl] = f0; slotNf[sl] = nf;
        } else {
            ORB_CUDA_TRY(cudaMemcpy2DAsync(kp_out + (size_t)f0 * cap, (size_t)cap * sizeof(orbx_keypoint), ex->d_kp[sl],
                                           (size_t)icap * sizeof(orbx_keypoint), (size_t)ccap * sizeof(orbx_keypoint), nf,
                                           cudaMemcpyDeviceToHost, ex->sD2H));
            ORB_CUDA_TRY(cudaMemcpy2DAsync(desc_out + (size_t)f0 * cap * 32, (size_t)cap * 32, ex->d_desc[sl], (size_t)icap * 32,
                                           (size_t)ccap * 32, nf, cudaMemcpyDeviceToHost, ex->sD2H));
        }
        mark(ex->sD2H);
        ORB_CUDA_TRY(cudaEventRecord(ex->evD2H[sl], ex->sD2H));
    }
    if (stageOut)
        for (int k = 0; k < nslots; k++) {                            // remaining slots, oldest chunk first
            int rcd = drain((chunk + k) % nslots);
            if (rcd != ORB_OK) return rcd;
        }
    ORB_CUDA_TRY(cudaStreamSynchronize(ex->sD2H));
    ORB_CUDA_TRY(cudaStreamSynchronize(ex->stream));
    for (int i = 0; i < ex->nHelpers; i++) ORB_CUDA_TRY(cudaStreamSynchronize(ex->helpers[i]->stream));
    if (trace) {
        fprintf(stderr, "chunk frames   h2d[start end]   compute[start end]   d2h[start end]  (ms since the first copy started)\n");
        for (size_t c = 0; c < sizes.size(); c++) {
            float t[6];
            for (int k = 0; k < 6; k++) cudaEventElapsedTime(&t[k], tev[0], tev[6 * c + k]);
            fprintf(stderr, "%3zu %5d   %7.3f %7.3f   %7.3f %7.3f   %7.3f %7.3f\n", c, sizes[c], t[0], t[1], t[2], t[3], t[4], t[5]);
        }
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    int s = 0;
    ORB_CUDA_TRY(cudaMemcpy(&s, ex->d_status, sizeof(int), cudaMemcpyDeviceToHost));
    if (s) ORB_CUDA_TRY(cudaMemset(ex->d_status, 0, sizeof(int)));
    for (int i = 0; i < ex->nHelpers; i++) {
        int s2 = 0;
        ORB_CUDA_TRY(cudaMemcpy(&s2, ex->helpers[i]->d_status, sizeof(int), cudaMemcpyDeviceToHost));
        if (s2) ORB_CUDA_TRY(cudaMemset(ex->helpers[i]->d_status, 0, sizeof(int)));
        s |= s2;
    }
    ORB_REQUIRE(!(s & (ORB_DEV_CAND_OVERFLOW | ORB_DEV_NODE_OVERFLOW)), ORB_ERR_OVERFLOW,
                "FAST candidate buffer overflow (ORBX_CAND_PER_CELL=%d caps it; unset it to size for the worst case)", ex->candPerCell);
    for (int f = 0; f < n_frames; f++) {
        n_out[f] = ex->h_n[f];
        ORB_REQUIRE(ex->h_n[f] <= cap, ORB_ERR_CAPACITY, "frame %d has %d keypoints but cap is %d", f, ex->h_n[f], cap);
    }
    return ORB_OK;
}

extern "C" int orbx_extract_batch(orbx_extractor* ex, const uint8_t* images, int n_frames, int width, int height, int stride,
                                  size_t frame_stride, const uint8_t* masks, int mask_stride, size_t mask_frame_stride,
                                  orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* n_out) {
    return extract_batch_impl(ex, images, nullptr, n_frames, width, height, stride, frame_stride, masks, mask_stride, mask_frame_stride, kp_out,
                              desc_out, cap, n_out);
}
extern "C" int orbx_extract_batch_ptrs(orbx_extractor* ex, const uint8_t* const* images, int n_frames, int width, int height, int stride,
                                       orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* n_out) {
    ORB_REQUIRE(ex && kp_out && desc_out && n_out && n_frames >= 0 && cap > 0 && (images || n_frames == 0), ORB_ERR_ARG, "bad arguments");
    if (n_frames == 0) return ORB_OK;
    for (int f = 0; f < n_frames; f++) ORB_REQUIRE(images[f], ORB_ERR_ARG, "image %d is NULL", f);
    return extract_batch_impl(ex, images[0], images, n_frames, width, height, stride, (size_t)stride * height, nullptr, 0, 0, kp_out, desc_out,
                              cap, n_out);
}

// Page-lock / unlock caller-owned host memory (cudaHostRegister) so that the batch call can DMA straight from / into it: from
// ordinary pageable arrays every copy goes through the driver's staging buffers (measured: 27 k instead of 169 k frames/s end to end).
extern "C" int orbx_host_register(void* p, size_t bytes) {
    ORB_REQUIRE(p && bytes > 0, ORB_ERR_ARG, "bad arguments");
    ORB_REQUIRE(orb_device_count() > 0, ORB_ERR_CUDA, "no CUDA device (no CPU fallback)");
    ORB_CUDA_TRY(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return ORB_OK;
}
extern "C" int orbx_host_unregister(void* p) {
    ORB_REQUIRE(p, ORB_ERR_ARG, "bad arguments");
    ORB_CUDA_TRY(cudaHostUnregister(p));
    return ORB_OK;
}

extern "C" int orbx_extract(orbx_extractor* ex, const uint8_t* image, int width, int height, int stride, const uint8_t* mask,
                            int mask_stride, orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* n_out) {
    return orbx_extract_batch(ex, image, 1, width, height, stride, (size_t)stride * height, mask, mask_stride,
                              (size_t)mask_stride * height, kp_out, desc_out, cap, n_out);
}

static int copy_level(orbx_extractor* ex, const u8* base, int frame, int level, int bordered, uint8_t* dst, int dst_stride) {
    ORB_REQUIRE(ex && dst && level >= 0 && level < ex->nlevels && frame >= 0 && frame < ex->lastFrames, ORB_ERR_ARG, "bad frame/level");
    const LevelPlan& L = ex->plan.lv[level];
    ORB_CUDA_TRY(cudaSetDevice(ex->device));
    const u8* src = base + (size_t)frame * ex->plan.frameBytes + L.off;
    if (bordered) {
        ORB_REQUIRE(dst_stride >= L.w + 38, ORB_ERR_ARG, "dst_stride too small");
        if (base == ex->d_pyr) {      // materialise the full 19-px REFLECT_101 border of this level (hot path keeps only 4x3)
            std::lock_guard<std::mutex> lk(ex->mu);
            dim3 g(8, 1, 1);
            k_border<<<g, 256, 0, ex->stream>>>(ex->plan, ex->d_pyr, ORBX_EDGE, ORBX_EDGE, level, frame);
            ex->launches++;
            ORB_CUDA_TRY(cudaGetLastError());
            ORB_CUDA_TRY(cudaStreamSynchronize(ex->stream));
        }
        ORB_CUDA_TRY(cudaMemcpy2D(dst, dst_stride, src + ORBX_OX - ORBX_EDGE, L.pitch, L.w + 38, L.h + 38, cudaMemcpyDeviceToHost));
    } else {
        ORB_REQUIRE(dst_stride >= L.w, ORB_ERR_ARG, "dst_stride too small");
        ORB_CUDA_TRY(cudaMemcpy2D(dst, dst_stride, src + (size_t)ORBX_OY * L.pitch + ORBX_OX, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    }
    return ORB_OK;
}
extern "C" int orbx_get_pyramid_level(orbx_extractor* ex, int frame, int level, int bordered, uint8_t* dst, int dst_stride) {
    return copy_level(ex, ex ? ex->d_pyr : nullptr, frame, level, bordered, dst, dst_stride);
}
extern "C" int orbx_get_pyramid(orbx_extractor* ex, int frame, int bordered, uint8_t* const* dst, const int* dst_stride) {
    ORB_REQUIRE(ex && dst && dst_stride && frame >= 0 && frame < ex->lastFrames, ORB_ERR_ARG, "bad frame");
    std::lock_guard<std::mutex> lk(ex->mu);
    ORB_CUDA_TRY(cudaSetDevice(ex->device));
    const int nl = ex->nlevels, b = bordered ? 38 : 0;
    if (bordered) {                  // full 19-px REFLECT_101 border of every level in one launch
        dim3 g(8, nl, 1);
        k_border<<<g, 256, 0, ex->stream>>>(ex->plan, ex->d_pyr, ORBX_EDGE, ORBX_EDGE, 0, frame);
        ex->launches++;
        ORB_CUDA_TRY(cudaGetLastError());
    }
    // device -> pinned staging (asynchronous DMA), one synchronisation, then plain row copies into the caller's (pageable) images
    size_t total = 0;
    for (int l = 0; l < nl; l++) total += orb_align_up((size_t)(ex->plan.lv[l].w + b) * (ex->plan.lv[l].h + b), 256);
    if (total > ex->h_stage_cap) {
        if (ex->h_stage) ORB_CUDA_TRY(cudaFreeHost(ex->h_stage));
        ex->h_stage = nullptr; ex->h_stage_cap = 0;
        ORB_CUDA_TRY(cudaMallocHost(&ex->h_stage, total));
        ex->h_stage_cap = total;
    }
    size_t o = 0;
    for (int l = 0; l < nl; l++) {
        const LevelPlan& L = ex->plan.lv[l];
        ORB_REQUIRE(dst[l] && dst_stride[l] >= L.w + b, ORB_ERR_ARG, "dst_stride too small");
        const u8* src = ex->d_pyr + (size_t)frame * ex->plan.frameBytes + L.off + (bordered ? (size_t)(ORBX_OX - ORBX_EDGE) : (size_t)ORBX_OY * L.pitch + ORBX_OX);
        ORB_CUDA_TRY(cudaMemcpy2DAsync(ex->h_stage + o, L.w + b, src, L.pitch, L.w + b, L.h + b, cudaMemcpyDeviceToHost, ex->stream));
        o += orb_align_up((size_t)(L.w + b) * (L.h + b), 256);
    }
    ORB_CUDA_TRY(cudaStreamSynchronize(ex->stream));
    o = 0;
    for (int l = 0; l < nl; l++) {
        const LevelPlan& L = ex->plan.lv[l];
        for (int y = 0; y < L.h + b; y++) memcpy(dst[l] + (size_t)y * dst_stride[l], ex->h_stage + o + (size_t)y * (L.w + b), L.w + b);
        o += orb_align_up((size_t)(L.w + b) * (L.h + b), 256);
    }
    return ORB_OK;
}
extern "C" int orbx_get_blurred_level(orbx_extractor* ex, int frame, int level, uint8_t* dst, int dst_stride) {
    return copy_level(ex, ex ? ex->d_blur : nullptr, frame, level, 0, dst, dst_stride);
}

extern "C" int orbx_get_candidates(orbx_extractor* ex, int frame, int level, orbx_candidate* out, int cap, int* n_out) {
    ORB_REQUIRE(ex && n_out && level >= 0 && level < ex->nlevels && frame >= 0 && frame < ex->lastFrames, ORB_ERR_ARG, "bad frame/level");
    const LevelPlan& L = ex->plan.lv[level];
    ORB_CUDA_TRY(cudaSetDevice(ex->device));
    int n = 0;
    ORB_CUDA_TRY(cudaMemcpy(&n, ex->d_candCount + frame * ex->nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    *n_out = n;
    n = std::min(n, L.candCap);
    std::vector<uint2> k(n);
    if (n) ORB_CUDA_TRY(cudaMemcpy(k.data(), ex->d_cand + (size_t)frame * ex->plan.candTotal + L.candOff, n * sizeof(uint2), cudaMemcpyDeviceToHost));
    // present in the reference's order: (cell id, y, x)
    std::sort(k.begin(), k.end(), [](const uint2& a, const uint2& b) {
        const unsigned long long ka = ((unsigned long long)(a.y & 0xFFFFFF) << 32) | ((unsigned long long)(a.x >> 16) << 16) | (a.x & 0xFFFF);
        const unsigned long long kb = ((unsigned long long)(b.y & 0xFFFFFF) << 32) | ((unsigned long long)(b.x >> 16) << 16) | (b.x & 0xFFFF);
        return ka < kb;
    });
    for (int i = 0; i < n && i < cap; i++) { out[i].x = k[i].x & 0xFFFF; out[i].y = k[i].x >> 16; out[i].response = k[i].y >> 24; }
    return ORB_OK;
}

// =====================================================================================================
// Frame::ComputeStereoMatches (src/Frame.cc:584-756) on the device-resident pyramids of two extractor handles —
// the only consumer of the public mvImagePyramid, so a stereo front-end no longer downloads the pyramids at all.
// One warp per left keypoint, no cross-keypoint dependency until the final median cut:
//   1. candidates = right keypoints whose row band [floor(y-r), ceil(y+r)], r = 2*scale[octave], holds row int(vL) (:599-612;
//      the reference's row table lists them in ascending index, so the packed (distance, index) minimum is its "first wins"),
//      octave within +-1, uR in [uL - mbf/mb, uL + 3]; best descriptor distance < TH_HIGH (:644-668);
//   2. 11x11 SAD of the centre-subtracted patches at the keypoint's pyramid level for the 11 shifts -5..+5 (:680-715): integer
//      arithmetic (the reference's float L1 norm of integer-valued differences is exact);
//   3. parabola fit, disparity test, depth (:717-747), float32 with explicit _rn operations;
//   4. k_stereo_median: median of the accepted SADs (rank size/2) by a 17-step counting search, cut at 1.5f*1.4f*median (:750-765).
struct StereoParams {
    const u8* pyrL; const u8* pyrR;
    const orbx_keypoint* kpL; const orbx_keypoint* kpR; const u8* descL; const u8* descR;
    int nL, nR, nRows;
    float mbf, maxD;
    float sf[ORBX_MAX_LEVELS], isf[ORBX_MAX_LEVELS];
};
#define ST_WARPS 8
__global__ void __launch_bounds__(32 * ST_WARPS) k_stereo_prepare(StereoParams S, int4* __restrict__ rinfo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.nR) return;
    const orbx_keypoint k = S.kpR[i];
    const float r = __fmul_rn(2.0f, S.sf[k.octave]);
    int4 v;
    v.x = max((int)floorf(__fsub_rn(k.y, r)), 0);                    // rows outside [0, nRows) are not entered in the table
    v.y = min((int)ceilf(__fadd_rn(k.y, r)), S.nRows - 1);
    v.z = k.octave;
    v.w = __float_as_int(k.x);
    rinfo[i] = v;
}

__global__ void __launch_bounds__(32 * ST_WARPS) k_stereo(const __grid_constant__ Plan P, StereoParams S, const int4* __restrict__ rinfo,
                                                          float* __restrict__ uRight, float* __restrict__ depth, int* __restrict__ sad) {
    const int lane = threadIdx.x & 31, iL = blockIdx.x * ST_WARPS + (threadIdx.x >> 5);
    if (iL >= S.nL) return;
    const orbx_keypoint kL = S.kpL[iL];
    const int levelL = kL.octave;
    const float uL = kL.x, vL = kL.y;
    float outU = -1.0f, outD = -1.0f;
    int outS = -1;
    const float minU = __fsub_rn(uL, S.maxD), maxU = __fsub_rn(uL, -3.0f);
    const bool rowOk = vL >= 0 && (int)vL < S.nRows;
    if (rowOk && !(maxU < 0)) {
        const int row = (int)vL;
        u32 dl[8];
        {
            const uint4 lo = __ldg(reinterpret_cast<const uint4*>(S.descL + (size_t)iL * 32)), hi = __ldg(reinterpret_cast<const uint4*>(S.descL + (size_t)iL * 32) + 1);
            dl[0] = lo.x; dl[1] = lo.y; dl[2] = lo.z; dl[3] = lo.w; dl[4] = hi.x; dl[5] = hi.y; dl[6] = hi.z; dl[7] = hi.w;
        }
        u32 best = ((u32)100 << 22);                                  // TH_HIGH, strict '<' (:646, 661)
        for (int j = lane; j < S.nR; j += 32) {
            const int4 ri = __ldg(&rinfo[j]);
            if (row < ri.x || row > ri.y) continue;
            if (ri.z < levelL - 1 || ri.z > levelL + 1) continue;
            const float uR = __int_as_float(ri.w);
            if (!(uR >= minU && uR <= maxU)) continue;
            const uint4 lo = __ldg(reinterpret_cast<const uint4*>(S.descR + (size_t)j * 32)), hi = __ldg(reinterpret_cast<const uint4*>(S.descR + (size_t)j * 32) + 1);
            const int d = __popc(dl[0] ^ lo.x) + __popc(dl[1] ^ lo.y) + __popc(dl[2] ^ lo.z) + __popc(dl[3] ^ lo.w) +
                          __popc(dl[4] ^ hi.x) + __popc(dl[5] ^ hi.y) + __popc(dl[6] ^ hi.z) + __popc(dl[7] ^ hi.w);
            best = min(best, ((u32)d << 22) | (u32)j);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        if ((best >> 22) < 100u) {
            const int bestIdxR = (int)(best & 0x3FFFFFu);
            const float uR0 = S.kpR[bestIdxR].x;
            const float scaleFactor = S.isf[levelL];
            const float scaleduL = roundf(__fmul_rn(kL.x, scaleFactor));
            const float scaledvL = roundf(__fmul_rn(kL.y, scaleFactor));
            const float scaleduR0 = roundf(__fmul_rn(uR0, scaleFactor));
            const LevelPlan& Lp = P.lv[levelL];
            const float iniu = __fadd_rn(scaleduR0, 0.0f), endu = __fadd_rn(scaleduR0, 11.0f);      // +L-w, +L+w+1
            // the reference only tests the right window (:697-699); windows that leave the level image (impossible for the
            // extractor's own keypoints, which stay 19 level-pixels away from the border) are skipped instead of being read
            const int r0 = (int)scaledvL - 5, c0 = (int)scaleduL - 5, cr0 = (int)scaleduR0 - 10;
            const bool inside = r0 >= 0 && r0 + 11 <= Lp.h && c0 >= 0 && c0 + 11 <= Lp.w;
            if (!(iniu < 0 || endu >= (float)Lp.w) && inside && cr0 >= 0) {
                const u8* bl = S.pyrL + Lp.off + (size_t)ORBX_OY * Lp.pitch + ORBX_OX;
                const u8* br = S.pyrR + Lp.off + (size_t)ORBX_OY * Lp.pitch + ORBX_OX;
                const int ILc = bl[(size_t)(r0 + 5) * Lp.pitch + c0 + 5];
                int acc[11];
#pragma unroll
                for (int k = 0; k < 11; k++) acc[k] = 0;
                for (int p = lane; p < 121; p += 32) {
                    const int y = p / 11, x = p - y * 11;
                    const int a = (int)bl[(size_t)(r0 + y) * Lp.pitch + c0 + x] - ILc;
                    const u8* rr = br + (size_t)(r0 + y) * Lp.pitch + cr0 + x;
                    const u8* rc = br + (size_t)(r0 + 5) * Lp.pitch + cr0 + 5;
#pragma unroll
                    for (int k = 0; k < 11; k++) acc[k] += abs(a - ((int)rr[k] - (int)rc[k]));
                }
#pragma unroll
                for (int k = 0; k < 11; k++)
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
                int bestS = 0x7fffffff, bestinc = 0;
#pragma unroll
                for (int k = 0; k < 11; k++)
                    if (acc[k] < bestS) { bestS = acc[k]; bestinc = k - 5; }
                if (bestinc != -5 && bestinc != 5) {
                    float dist1 = 0.f, dist2 = 0.f, dist3 = 0.f;
#pragma unroll
                    for (int k = 1; k < 10; k++)
                        if (k - 5 == bestinc) { dist1 = (float)acc[k - 1]; dist2 = (float)acc[k]; dist3 = (float)acc[k + 1]; }
                    const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2)));
                    const float deltaR = __fdiv_rn(__fsub_rn(dist1, dist3), den);
                    if (!(deltaR < -1 || deltaR > 1)) {
                        float bestuR = __fmul_rn(S.sf[levelL], __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
                        float disparity = __fsub_rn(uL, bestuR);
                        if (disparity >= 0 && disparity < S.maxD) {
                            if (disparity <= 0) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }
                            outD = __fdiv_rn(S.mbf, disparity);
                            outU = bestuR;
                            outS = bestS;
                        }
                    }
                }
            }
        }
    }
    if (lane == 0) { uRight[iL] = outU; depth[iL] = outD; sad[iL] = outS; }
}

__global__ void __launch_bounds__(1024) k_stereo_median(int n, const int* __restrict__ sad, float* __restrict__ uRight, float* __restrict__ depth) {
    __shared__ int s_cnt, s_valid;
    const int tid = threadIdx.x;
    if (tid == 0) s_valid = 0;
    __syncthreads();
    int c = 0;
    for (int i = tid; i < n; i += blockDim.x) c += sad[i] >= 0;
    if (c) atomicAdd(&s_valid, c);
    __syncthreads();
    const int nv = s_valid;
    if (nv == 0) return;
    const int rank = nv / 2;                                         // vDistIdx[size/2] of the ascending sort
    int lo = 0, hi = 1 << 17;                                        // smallest v with #{sad <= v} >= rank + 1  (SAD <= 121*510)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        __syncthreads();
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        c = 0;
        for (int i = tid; i < n; i += blockDim.x) { const int v = sad[i]; c += (v >= 0 && v <= mid); }
        if (c) atomicAdd(&s_cnt, c);
        __syncthreads();
        if (s_cnt >= rank + 1) hi = mid; else lo = mid + 1;
    }
    const float thDist = __fmul_rn(__fmul_rn(1.5f, 1.4f), (float)lo);
    for (int i = tid; i < n; i += blockDim.x)
        if (sad[i] >= 0 && !((float)sad[i] < thDist)) { uRight[i] = -1.0f; depth[i] = -1.0f; }
}

extern "C" int orbx_stereo_matches(orbx_extractor* left, orbx_extractor* right, int frame_left, int frame_right,
                                   const orbx_keypoint* kp_left, const uint8_t* desc_left, int n_left,
                                   const orbx_keypoint* kp_right, const uint8_t* desc_right, int n_right,
                                   float mbf, float mb, float* u_right, float* depth) {
    ORB_REQUIRE(left && right && n_left >= 0 && n_right >= 0, ORB_ERR_ARG, "bad arguments");
    ORB_REQUIRE(n_left == 0 || (kp_left && desc_left && u_right && depth), ORB_ERR_ARG, "null left array");
    ORB_REQUIRE(n_right == 0 || (kp_right && desc_right), ORB_ERR_ARG, "null right array");
    ORB_REQUIRE(left->device == right->device && left->nlevels == right->nlevels && left->plan.width == right->plan.width &&
                left->plan.height == right->plan.height && left->scaleFactor == right->scaleFactor, ORB_ERR_ARG,
                "the two extractors must share device, image size, level count and scale factor");
    ORB_REQUIRE(frame_left >= 0 && frame_left < left->lastFrames && frame_right >= 0 && frame_right < right->lastFrames, ORB_ERR_ARG,
                "no pyramid for that frame: run the extraction first");
    ORB_REQUIRE(n_right <= 0x3FFFFF, ORB_ERR_ARG, "too many right keypoints");
    for (int i = 0; i < n_left; i++) ORB_REQUIRE(kp_left[i].octave >= 0 && kp_left[i].octave < left->nlevels, ORB_ERR_ARG, "left octave out of range");
    for (int i = 0; i < n_right; i++) ORB_REQUIRE(kp_right[i].octave >= 0 && kp_right[i].octave < right->nlevels, ORB_ERR_ARG, "right octave out of range");
    if (n_left == 0) return ORB_OK;
    ORB_CUDA_TRY(cudaSetDevice(left->device));
    std::lock_guard<std::mutex> lk(left->mu);
    cudaStream_t st = left->stream;
    // scratch: one allocation per call sized by the keypoint counts (a few hundred KB)
    const size_t bL = (size_t)n_left * sizeof(orbx_keypoint), bR = (size_t)std::max(n_right, 1) * sizeof(orbx_keypoint);
    const size_t offKL = 0, offKR = orb_align_up(offKL + bL, 256), offDL = orb_align_up(offKR + bR, 256),
                 offDR = orb_align_up(offDL + (size_t)n_left * 32, 256), offRI = orb_align_up(offDR + (size_t)std::max(n_right, 1) * 32, 256),
                 offU = orb_align_up(offRI + (size_t)std::max(n_right, 1) * 16, 256), offD = orb_align_up(offU + (size_t)n_left * 4, 256),
                 offS = orb_align_up(offD + (size_t)n_left * 4, 256), total = orb_align_up(offS + (size_t)n_left * 4, 256);
    if (total > left->stereoCap) {                                   // grow-only scratch kept in the handle
        if (left->d_stereo) ORB_CUDA_TRY(cudaFree(left->d_stereo));
        left->d_stereo = nullptr;
        left->stereoCap = orb_align_up(total + (total >> 1), 1 << 16);
        ORB_CUDA_TRY(cudaMalloc(&left->d_stereo, left->stereoCap));
    }
    u8* d = left->d_stereo;
    ORB_CUDA_TRY(cudaMemcpyAsync(d + offKL, kp_left, bL, cudaMemcpyHostToDevice, st));
    ORB_CUDA_TRY(cudaMemcpyAsync(d + offDL, desc_left, (size_t)n_left * 32, cudaMemcpyHostToDevice, st));
    if (n_right) {
        ORB_CUDA_TRY(cudaMemcpyAsync(d + offKR, kp_right, (size_t)n_right * sizeof(orbx_keypoint), cudaMemcpyHostToDevice, st));
        ORB_CUDA_TRY(cudaMemcpyAsync(d + offDR, desc_right, (size_t)n_right * 32, cudaMemcpyHostToDevice, st));
    }
    // the right handle's last pass must be complete before its pyramid is read on the left handle's stream
    ORB_CUDA_TRY(cudaStreamSynchronize(right->stream));
    StereoParams S;
    S.pyrL = left->d_pyr + (size_t)frame_left * left->plan.frameBytes;
    S.pyrR = right->d_pyr + (size_t)frame_right * right->plan.frameBytes;
    S.kpL = reinterpret_cast<const orbx_keypoint*>(d + offKL); S.kpR = reinterpret_cast<const orbx_keypoint*>(d + offKR);
    S.descL = d + offDL; S.descR = d + offDR;
    S.nL = n_left; S.nR = n_right; S.nRows = left->plan.lv[0].h;
    S.mbf = mbf; S.maxD = mbf / mb;
    for (int l = 0; l < left->nlevels; l++) { S.sf[l] = left->sf[l]; S.isf[l] = left->isf[l]; }
    int4* rinfo = reinterpret_cast<int4*>(d + offRI);
    float* dU = reinterpret_cast<float*>(d + offU);
    float* dD = reinterpret_cast<float*>(d + offD);
    int* dS = reinterpret_cast<int*>(d + offS);
    if (n_right) k_stereo_prepare<<<orb_div_up(n_right, 32 * ST_WARPS), 32 * ST_WARPS, 0, st>>>(S, rinfo);
    k_stereo<<<orb_div_up(n_left, ST_WARPS), 32 * ST_WARPS, 0, st>>>(left->plan, S, rinfo, dU, dD, dS);
    k_stereo_median<<<1, 1024, 0, st>>>(n_left, dS, dU, dD);
    left->launches += 3;
    ORB_CUDA_TRY(cudaGetLastError());
    ORB_CUDA_TRY(cudaMemcpyAsync(u_right, dU, (size_t)n_left * 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA_TRY(cudaMemcpyAsync(depth, dD, (size_t)n_left * 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA_TRY(cudaStreamSynchronize(st));
    return ORB_OK;
}
