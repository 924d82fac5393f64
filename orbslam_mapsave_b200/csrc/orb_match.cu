// orb_match.cu — B200 (sm_100a) implementation of ORB_SLAM2::ORBmatcher's Hamming searches behind the C ABI of
// include/orb_b200.h.  Replaces src/ORBmatcher.cc:1650-1666 (DescriptorDistance), :159-291 and :525-658 (SearchByBoW),
// :660-826 (SearchForTriangulation), :140-157 (CheckDistEpipolarLine), :1604-1645 (ComputeThreeMaxima).
//
// The distance kernel is LOP3(XOR) + POPC on descriptors held as 8 x u32 in registers (queries) and shared memory
// (database tile, broadcast reads); top-2 is kept branch-free as two packed (distance<<23 | index) keys.
// Integer pipe work only: no tensor cores.
#include "orb_common.cuh"
#include "orb_match_common.cuh"

#include <algorithm>
#include <mutex>
#include <vector>

// ---------------------------------------------------------------------------------------------------
// DescriptorDistance for n pairs
// ---------------------------------------------------------------------------------------------------
__global__ void k_pair_distance(const u8* __restrict__ a, const u8* __restrict__ b, int n, int* __restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 x[8], y[8];
    load_desc(a + (size_t)i * 32, x);
    load_desc(b + (size_t)i * 32, y);
    dist[i] = ham256(x, y);
}

// ---------------------------------------------------------------------------------------------------
// Brute-force top-2, batched over independent (query set, database set) problems.
// CTA = 128 threads, QPT queries per thread in registers; database streamed through a 256-descriptor smem tile.
// ---------------------------------------------------------------------------------------------------
#define T2_THREADS 128
#define T2_DBT 256

// Packed key of one pair: (Hamming distance << 23) + database index.  The 8 XOR words are first compressed with four
// carry-save adders (LOP3 0x96 = sum, LOP3 0xE8 = majority: 3 words of weight w become one of weight w and one of weight 2w),
// so a pair needs 4 POPC instead of 8 at the price of 8 more LOP3.  POPC issues at 16 lanes/clk/SM and LOP3 at 64, which makes
// this the balance point of the two pipes (16 LOP3 + 2.5 min/max on the ALU pipe vs 4 POPC on the XU pipe; measured on B200:
// 556 G pairs/s with 8 POPC, 832 G with 5, 868 G with 4).  The weighted sum and the shift into the key are multiply-adds with
// the multipliers in constant memory: they issue on the FMA pipe (with immediates ptxas turns them into ALU-pipe LEAs).
// Exact: sum_i popc(x_i) = popc(s2) + popc(s3) + 2 popc(t) + 4 popc(f).
// (explicit lop3/xor so that the compiler keeps the 8 + 2-per-adder form instead of re-associating the XORs into more LOP3s)
__device__ __forceinline__ u32 xor3(u32 a, u32 b, u32 c) { u32 r; asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ u32 maj3(u32 a, u32 b, u32 c) { u32 r; asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__constant__ u32 c_keymul[3] = {1u << KEY_SHIFT, 2u << KEY_SHIFT, 4u << KEY_SHIFT};
__device__ __forceinline__ u32 ham_key(const u32 (&q)[8], const u32 (&w)[8], u32 idx) {
    u32 x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) asm("xor.b32 %0, %1, %2;" : "=r"(x[i]) : "r"(q[i]), "r"(w[i]));
    const u32 s1 = xor3(x[0], x[1], x[2]), c1 = maj3(x[0], x[1], x[2]);
    const u32 s2 = xor3(x[3], x[4], x[5]), c2 = maj3(x[3], x[4], x[5]);
    const u32 s3 = xor3(x[6], x[7], s1), c3 = maj3(x[6], x[7], s1);
    const u32 t = xor3(c1, c2, c3), f = maj3(c1, c2, c3);
    u32 key = (u32)(__popc(s2) + __popc(s3)) * c_keymul[0] + idx;
    key = (u32)__popc(t) * c_keymul[1] + key;
    key = (u32)__popc(f) * c_keymul[2] + key;
    return key;
}

template <int QPT>
__device__ __forceinline__ void top2_scan_tile(const uint4* __restrict__ s_db, int cnt, int jbase, const u32 (&q)[QPT][8],
                                               u32 (&best)[QPT], u32 (&sec)[QPT]) {
    int j = 0;
    for (; j + 2 <= cnt; j += 2) {                   // two database rows per step: their keys enter the top-2 as a sorted pair (5 min/max per 2 pairs)
        const uint4 a0 = s_db[2 * j], a1 = s_db[2 * j + 1], b0 = s_db[2 * j + 2], b1 = s_db[2 * j + 3];
        const u32 wa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const u32 wb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < QPT; i++) {
            const u32 ka = ham_key(q[i], wa, (u32)(jbase + j));
            const u32 kb = ham_key(q[i], wb, (u32)(jbase + j + 1));
            const u32 lo = min(ka, kb), hi = max(ka, kb);
            sec[i] = min(min(sec[i], hi), max(best[i], lo));
            best[i] = min(best[i], lo);
        }
    }
    if (j < cnt) {
        const uint4 a0 = s_db[2 * j], a1 = s_db[2 * j + 1];
        const u32 wa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < QPT; i++) {
            const u32 key = ham_key(q[i], wa, (u32)(jbase + j));
            sec[i] = min(sec[i], max(best[i], key));
            best[i] = min(best[i], key);
        }
    }
}

template <int QPT>
__global__ void __launch_bounds__(T2_THREADS) k_top2(const u8* __restrict__ q, const int* __restrict__ q_off,
                                                     const int* __restrict__ q_cnt, const u8* __restrict__ db,
                                                     const int* __restrict__ db_off, const int* __restrict__ db_cnt,
                                                     int* __restrict__ best_idx, int* __restrict__ best_dist,
                                                     int* __restrict__ second_dist) {
    __shared__ uint4 s_db[T2_DBT * 2];
    const int p = blockIdx.y, tid = threadIdx.x;
    const int nq = q_cnt[p], qo = q_off[p], nd = db_cnt[p];
    const int q0 = blockIdx.x * (T2_THREADS * QPT);
    if (q0 >= nq) return;
    const u8* dbp = db + (size_t)db_off[p] * 32;
    u32 qr[QPT][8], best[QPT], sec[QPT];
#pragma unroll
    for (int i = 0; i < QPT; i++) {
        const int qi = min(q0 + i * T2_THREADS + tid, nq - 1);
        load_desc(q + (size_t)(qo + qi) * 32, qr[i]);
        best[i] = KEY_NONE; sec[i] = KEY_NONE;
    }
    for (int j0 = 0; j0 < nd; j0 += T2_DBT) {
        const int cnt = min(T2_DBT, nd - j0);
        __syncthreads();
        for (int i = tid; i < cnt * 2; i += T2_THREADS) s_db[i] = __ldg(reinterpret_cast<const uint4*>(dbp + (size_t)j0 * 32) + i);
        __syncthreads();
        top2_scan_tile<QPT>(s_db, cnt, j0, qr, best, sec);
    }
#pragma unroll
    for (int i = 0; i < QPT; i++) {
        const int qi = q0 + i * T2_THREADS + tid;
        if (qi < nq) {
            best_idx[qo + qi] = best[i] == KEY_NONE ? -1 : (int)(best[i] & KEY_IDX_MASK);
            best_dist[qo + qi] = key_dist(best[i]);
            second_dist[qo + qi] = key_dist(sec[i]);
        }
    }
}

// Latency form for ONE small problem (e.g. 2000 x 2000 between consecutive keyframes): the database is cut into `nslice` slices
// so that the grid fills the GPU; each CTA keeps the packed top-2 keys of its slice, a second tiny kernel merges them.
__global__ void __launch_bounds__(T2_THREADS) k_top2_slice(const u8* __restrict__ q, int nq, const u8* __restrict__ db, int nd,
                                                           int slice_len, u32* __restrict__ part /*[nslice][2][nq]*/) {
    __shared__ uint4 s_db[T2_DBT * 2];
    const int tid = threadIdx.x, qi0 = blockIdx.x * T2_THREADS + tid, sl = blockIdx.y;
    const int j_begin = sl * slice_len, j_end = min(nd, j_begin + slice_len);
    u32 qr[1][8], best[1] = {KEY_NONE}, sec[1] = {KEY_NONE};
    load_desc(q + (size_t)min(qi0, nq - 1) * 32, qr[0]);
    for (int j0 = j_begin; j0 < j_end; j0 += T2_DBT) {
        const int cnt = min(T2_DBT, j_end - j0);
        __syncthreads();
        for (int i = tid; i < cnt * 2; i += T2_THREADS) s_db[i] = __ldg(reinterpret_cast<const uint4*>(db + (size_t)j0 * 32) + i);
        __syncthreads();
        top2_scan_tile<1>(s_db, cnt, j0, qr, best, sec);
    }
    if (qi0 < nq) {
        part[((size_t)sl * 2 + 0) * nq + qi0] = best[0];
        part[((size_t)sl * 2 + 1) * nq + qi0] = sec[0];
    }
}
__global__ void k_top2_merge(const u32* __restrict__ part, int nq, int nslice, int* __restrict__ best_idx, int* __restrict__ best_dist,
                             int* __restrict__ second_dist) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    u32 best = KEY_NONE, sec = KEY_NONE;
    for (int sl = 0; sl < nslice; sl++) {          // keys are unique (index in the low bits), so the packed top-2 merge is exact
        const u32 b = part[((size_t)sl * 2 + 0) * nq + qi], s2 = part[((size_t)sl * 2 + 1) * nq + qi];
        sec = min(min(sec, s2), max(best, b));
        best = min(best, b);
    }
    best_idx[qi] = best == KEY_NONE ? -1 : (int)(best & KEY_IDX_MASK);
    best_dist[qi] = key_dist(best);
    second_dist[qi] = key_dist(sec);
}

// ---------------------------------------------------------------------------------------------------
// All-pairs keyframe matching (config 4): CTA = (query tile of keyframe q, database keyframe k).
// count[q][k] += #queries passing  best <= th_low && best < ratio * second ; optional global nearest keyframe.
// ---------------------------------------------------------------------------------------------------
template <int QPT>
__global__ void __launch_bounds__(T2_THREADS) k_allpairs(const u8* __restrict__ desc, int n_kf, int per_kf, int q_begin,
                                                         int th_low, float ratio, u32* __restrict__ count_words,
                                                         u32* __restrict__ best_packed) {
    __shared__ uint4 s_db[T2_DBT * 2];
    __shared__ int s_cnt;
    const int tid = threadIdx.x, k = blockIdx.y, qk = q_begin + blockIdx.z;
    if (k == qk) return;
    const int q0 = blockIdx.x * (T2_THREADS * QPT);
    if (tid == 0) s_cnt = 0;
    const u8* qp = desc + (size_t)qk * per_kf * 32;
    const u8* dbp = desc + (size_t)k * per_kf * 32;
    u32 qr[QPT][8], best[QPT], sec[QPT];
#pragma unroll
    for (int i = 0; i < QPT; i++) {
        const int qi = min(q0 + i * T2_THREADS + tid, per_kf - 1);
        load_desc(qp + (size_t)qi * 32, qr[i]);
        best[i] = KEY_NONE; sec[i] = KEY_NONE;
    }
    for (int j0 = 0; j0 < per_kf; j0 += T2_DBT) {
        const int cnt = min(T2_DBT, per_kf - j0);
        __syncthreads();
        for (int i = tid; i < cnt * 2; i += T2_THREADS) s_db[i] = __ldg(reinterpret_cast<const uint4*>(dbp + (size_t)j0 * 32) + i);
        __syncthreads();
        top2_scan_tile<QPT>(s_db, cnt, j0, qr, best, sec);
    }
    int ok = 0;
#pragma unroll
    for (int i = 0; i < QPT; i++) {
        const int qi = q0 + i * T2_THREADS + tid;
        if (qi < per_kf) {
            const int b = key_dist(best[i]), s = key_dist(sec[i]);
            ok += (b <= th_low && (float)b < __fmul_rn(ratio, (float)s)) ? 1 : 0;
            if (best_packed) atomicMin(&best_packed[(size_t)blockIdx.z * per_kf + qi], ((u32)b << KEY_SHIFT) | (u32)k);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ok += __shfl_xor_sync(0xffffffffu, ok, o);
    if ((tid & 31) == 0 && ok) atomicAdd(&s_cnt, ok);
    __syncthreads();
    if (tid == 0 && s_cnt) {
        const size_t e = (size_t)blockIdx.z * n_kf + k;                 // uint16 element index
        atomicAdd(&count_words[e >> 1], (u32)s_cnt << ((e & 1) * 16));
    }
}
__global__ void k_unpack_best(const u32* __restrict__ packed, int n, int* __restrict__ kf, int* __restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 v = packed[i];
    kf[i] = v == KEY_NONE ? -1 : (int)(v & KEY_IDX_MASK);
    dist[i] = key_dist(v);
}

// ---------------------------------------------------------------------------------------------------
// Node-constrained searches.  A FeatureVector arrives as CSR (ids ascending, offsets, features).
// ---------------------------------------------------------------------------------------------------
struct DevView {
    int n;
    const u8* desc; const u8* flag; const float* angle; const float* x; const float* y; const int* octave; const float* uright;
    int nn; const int* ids; const int* off; const int* feat;
};

__device__ __forceinline__ int find_node(const int* ids, int n, int id) {     // index of id in ascending ids, or -1
    int lo = 0, hi = n;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (ids[m] < id) lo = m + 1; else hi = m; }
    return (lo < n && ids[lo] == id) ? lo : -1;
}
__global__ void k_three_maxima(const int* histo, int L, int* ind) {
    int a, b, c;
    three_maxima_dev(histo, L, a, b, c);
    ind[0] = a; ind[1] = b; ind[2] = c;
}

// SearchByBoW, both overloads.  Two launches per call: (1) one WARP per shared vocabulary node, spread over as many CTAs as
// there are nodes / 4 (the greedy claim `vpMapPointMatches[realIdxF]` / `vbMatched2[idx2]` only couples features of the same
// node, and a feature belongs to exactly one node, so nodes are independent); inside a node the side-1 features are walked in
// order and the side-2 scan is spread over the lanes with a (distance, position) top-2 warp reduction; (2) one CTA applies the
// rotation-histogram cull (ComputeThreeMaxima) and counts.
//   KFKF = false: SearchByBoW(KF, Frame)  -> out[j in side 2] = idx1, accept best <= TH_LOW        (:231)
//   KFKF = true : SearchByBoW(KF1, KF2)   -> out[i in side 1] = idx2, accept best <  TH_LOW, side-2 flag filter (:601)
// Scratch (zeroed by the host before the launch): taken2[B.n], hist[HISTO_LENGTH], nmatch[1]; out[] preset to -1.
#define SB_WARPS 4
template <bool KFKF>
__device__ __forceinline__ void search_bow_nodes_body(const DevView& A, const DevView& B, float nnratio, int checkOri,
                                                      int* __restrict__ out, int* __restrict__ taken2,
                                                      int* __restrict__ binOf, int* __restrict__ hist,
                                                      int* __restrict__ nmatch) {
    const int lane = threadIdx.x & 31, a = blockIdx.x * SB_WARPS + (threadIdx.x >> 5);
    if (a >= A.nn) return;
    const int b = find_node(B.ids, B.nn, A.ids[a]);
    if (b < 0) return;
    const int o2 = B.off[b], c2 = B.off[b + 1] - o2;
    int nm = 0;
    for (int i1 = A.off[a]; i1 < A.off[a + 1]; i1++) {
        const int idx1 = A.feat[i1];
        if (!A.flag[idx1]) continue;
        u32 d1[8];
        load_desc(A.desc + (size_t)idx1 * 32, d1);
        u32 best = KEY_NONE, sec = KEY_NONE;
        for (int j = lane; j < c2; j += 32) {
            const int idx2 = B.feat[o2 + j];
            if (taken2[idx2]) continue;
            if (KFKF && !B.flag[idx2]) continue;
            u32 d2[8];
            load_desc(B.desc + (size_t)idx2 * 32, d2);
            const u32 key = ((u32)ham256(d1, d2) << KEY_SHIFT) | (u32)j;
            sec = min(sec, max(best, key));
            best = min(best, key);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const u32 ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, sec, o);
            sec = min(min(sec, os), max(best, ob));
            best = min(best, ob);
        }
        const int bd1 = key_dist(best), bd2 = key_dist(sec);
        const bool th = KFKF ? (bd1 < ORBM_TH_LOW) : (bd1 <= ORBM_TH_LOW);
        if (th && (float)bd1 < __fmul_rn(nnratio, (float)bd2)) {
            const int idx2 = B.feat[o2 + (int)(best & KEY_IDX_MASK)];
            if (lane == 0) {
                taken2[idx2] = 1;
                const int slot = KFKF ? idx1 : idx2;
                out[slot] = KFKF ? idx2 : idx1;
                if (checkOri) {
                    const int bin = rot_bin(A.angle[idx1], B.angle[idx2]);
                    binOf[slot] = bin;
                    atomicAdd(&hist[bin], 1);
                }
                nm++;
            }
        }
        __syncwarp();
    }
    if (lane == 0 && nm) atomicAdd(nmatch, nm);
}
template <bool KFKF>
__global__ void __launch_bounds__(32 * SB_WARPS) k_search_bow_nodes(DevView A, DevView B, float nnratio, int checkOri,
                                                                    int* __restrict__ out, int* __restrict__ taken2,
                                                                    int* __restrict__ binOf, int* __restrict__ hist,
                                                                    int* __restrict__ nmatch) {
    search_bow_nodes_body<KFKF>(A, B, nnratio, checkOri, out, taken2, binOf, hist, nmatch);
}
// Batched form: blockIdx.y = pair of the batch (one keyframe / frame against N candidates, e.g. the candidate loops of
// Tracking::Relocalization, src/Tracking.cc:1621-1643, and LoopClosing::ComputeSim3, src/LoopClosing.cc:240-266); the pairs are
// independent searches, so the grid is simply (nodes, pairs) and the whole batch is two launches.
struct BowJob { DevView A, B; int *out, *taken2, *binOf, *hist, *nmatch; int nOut; };
template <bool KFKF>
__global__ void __launch_bounds__(32 * SB_WARPS) k_search_bow_nodes_batch(const BowJob* __restrict__ jobs, float nnratio, int checkOri) {
    const BowJob& J = jobs[blockIdx.y];
    search_bow_nodes_body<KFKF>(J.A, J.B, nnratio, checkOri, J.out, J.taken2, J.binOf, J.hist, J.nmatch);
}

__device__ __forceinline__ void search_bow_finish_body(int nOut, int checkOri, int* __restrict__ out, const int* __restrict__ binOf,
                                                       const int* __restrict__ hist, int* __restrict__ nmatch) {
    __shared__ int s_ind[3], s_drop;
    const int tid = threadIdx.x;
    if (!checkOri) return;
    if (tid == 0) { int a, b, c; three_maxima_dev(hist, ORBM_HISTO_LENGTH, a, b, c); s_ind[0] = a; s_ind[1] = b; s_ind[2] = c; s_drop = 0; }
    __syncthreads();
    int drop = 0;
    for (int i = tid; i < nOut; i += blockDim.x)
        if (out[i] >= 0) {
            const int bin = binOf[i];
            if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) { out[i] = -1; drop++; }
        }
    if (drop) atomicAdd(&s_drop, drop);
    __syncthreads();
    if (tid == 0) *nmatch -= s_drop;
}
__global__ void __launch_bounds__(256) k_search_bow_finish(int nOut, int checkOri, int* __restrict__ out, const int* __restrict__ binOf,
                                                           const int* __restrict__ hist, int* __restrict__ nmatch) {
    search_bow_finish_body(nOut, checkOri, out, binOf, hist, nmatch);
}
__global__ void __launch_bounds__(256) k_search_bow_finish_batch(const BowJob* __restrict__ jobs, int checkOri) {
    const BowJob& J = jobs[blockIdx.x];
    search_bow_finish_body(J.nOut, checkOri, J.out, J.binOf, J.hist, J.nmatch);
}

// SearchForTriangulation.  No cross-query coupling (vbMatched2 is never set in the reference, :680,728), so one thread
// per side-1 CSR slot walks its node's side-2 features serially, which keeps the reference's "last candidate wins a
// tie" (`dist>bestDist` rejects, :741) and "bestDist only moves when the epipolar test passes" (:754-758).  Launch 1 spreads
// the slots over the GPU; launch 2 (one CTA) applies the orientation cull and the ordered compaction (:818-823).
// Scratch zeroed by the host: hist[HISTO_LENGTH], cnt[2]; m12[] preset to -1.
struct TriParams { float F[9]; float ex, ey; int onlyStereo, checkOri; };

__device__ __forceinline__ void search_tri_slots_body(const DevView& A, const DevView& B, const TriParams& tp, const float* __restrict__ sf2,
                                                      const float* __restrict__ sigma2, int* __restrict__ m12,
                                                      int* __restrict__ hist, int* __restrict__ cnt) {
    const int total1 = A.nn > 0 ? A.off[A.nn] : 0;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total1) return;
    int lo = 0, hi = A.nn;                                       // node a with off[a] <= p < off[a+1]
    while (hi - lo > 1) { const int m = (lo + hi) >> 1; if (A.off[m] <= p) lo = m; else hi = m; }
    const int b = find_node(B.ids, B.nn, A.ids[lo]);
    if (b < 0) return;
    const int idx1 = A.feat[p];
    if (A.flag[idx1]) return;                                    // already has a MapPoint (:706-708)
    const bool st1 = A.uright[idx1] >= 0;
    if (tp.onlyStereo && !st1) return;
    const float x1 = A.x[idx1], y1 = A.y[idx1];
    u32 d1[8];
    load_desc(A.desc + (size_t)idx1 * 32, d1);
    // epipolar line l = x1' F12 (:143-145)
    const float la = __fadd_rn(__fadd_rn(__fmul_rn(x1, tp.F[0]), __fmul_rn(y1, tp.F[3])), tp.F[6]);
    const float lb = __fadd_rn(__fadd_rn(__fmul_rn(x1, tp.F[1]), __fmul_rn(y1, tp.F[4])), tp.F[7]);
    const float lc = __fadd_rn(__fadd_rn(__fmul_rn(x1, tp.F[2]), __fmul_rn(y1, tp.F[5])), tp.F[8]);
    const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
    int bestDist = ORBM_TH_LOW, bestIdx2 = -1;
    for (int i2 = B.off[b]; i2 < B.off[b + 1]; i2++) {
        const int idx2 = B.feat[i2];
        if (B.flag[idx2]) continue;
        const bool st2 = B.uright[idx2] >= 0;
        if (tp.onlyStereo && !st2) continue;
        u32 d2[8];
        load_desc(B.desc + (size_t)idx2 * 32, d2);
        const int dist = ham256(d1, d2);
        if (dist > ORBM_TH_LOW || dist > bestDist) continue;
        const float x2 = B.x[idx2], y2 = B.y[idx2];
        const int oc = B.octave[idx2];
        if (!st1 && !st2) {
            const float dx = __fsub_rn(tp.ex, x2), dy = __fsub_rn(tp.ey, y2);
            if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, sf2[oc])) continue;
        }
        if (den == 0) continue;
        const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, x2), __fmul_rn(lb, y2)), lc);
        const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
        if ((double)dsqr < 3.84 * (double)sigma2[oc]) { bestIdx2 = idx2; bestDist = dist; }
    }
    if (bestIdx2 >= 0) {
        m12[idx1] = bestIdx2;
        atomicAdd(&cnt[1], 1);
        if (tp.checkOri) atomicAdd(&hist[rot_bin(A.angle[idx1], B.angle[bestIdx2])], 1);
    }
}
__global__ void __launch_bounds__(128) k_search_tri_slots(DevView A, DevView B, TriParams tp, const float* __restrict__ sf2,
                                                          const float* __restrict__ sigma2, int* __restrict__ m12,
                                                          int* __restrict__ hist, int* __restrict__ cnt) {
    search_tri_slots_body(A, B, tp, sf2, sigma2, m12, hist, cnt);
}
// Batched form (blockIdx.y = neighbour): LocalMapping::CreateNewMapPoints matches the new keyframe against up to 20 neighbours
// (src/LocalMapping.cc:215-268), each with its own fundamental matrix and epipole.
struct TriJob { DevView A, B; TriParams tp; const float *sf2, *sigma2; int *m12, *hist, *pairs, *cnt; };
__global__ void __launch_bounds__(128) k_search_tri_slots_batch(const TriJob* __restrict__ jobs) {
    const TriJob& J = jobs[blockIdx.y];
    search_tri_slots_body(J.A, J.B, J.tp, J.sf2, J.sigma2, J.m12, J.hist, J.cnt);
}

__device__ __forceinline__ void search_tri_finish_body(const DevView& A, const DevView& B, int checkOri, int* __restrict__ m12,
                                                       const int* __restrict__ hist, int* __restrict__ pairs,
                                                       int* __restrict__ cnt) {
    __shared__ int s_ind[3], s_drop, s_wsum[34];
    const int tid = threadIdx.x;
    if (checkOri) {
        if (tid == 0) { int a, b, c; three_maxima_dev(hist, ORBM_HISTO_LENGTH, a, b, c); s_ind[0] = a; s_ind[1] = b; s_ind[2] = c; s_drop = 0; }
        __syncthreads();
        int drop = 0;
        for (int i = tid; i < A.n; i += blockDim.x)
            if (m12[i] >= 0) {
                const int bin = rot_bin(A.angle[i], B.angle[m12[i]]);
                if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) { m12[i] = -1; drop++; }
            }
        if (drop) atomicAdd(&s_drop, drop);
        __syncthreads();
        if (tid == 0) cnt[1] -= s_drop;
    }
    // ordered compaction of (i, m12[i]) (:818-823): chunked block scan
    int base = 0;
    const int lane = tid & 31, w = tid >> 5;
    for (int c0 = 0; c0 < A.n; c0 += blockDim.x) {
        const int i = c0 + tid;
        const int has = (i < A.n && m12[i] >= 0) ? 1 : 0;
        int inc = has;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
        if (lane == 31) s_wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            const int ws = lane < (int)(blockDim.x >> 5) ? s_wsum[lane] : 0;
            int wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
            s_wsum[lane] = wi - ws;
            if (lane == 31) s_wsum[32] = wi;
        }
        __syncthreads();
        if (has) { const int pos = base + s_wsum[w] + inc - 1; pairs[2 * pos] = i; pairs[2 * pos + 1] = m12[i]; }
        base += s_wsum[32];
        __syncthreads();
    }
    if (tid == 0) cnt[0] = base;
}
__global__ void __launch_bounds__(256) k_search_tri_finish(DevView A, DevView B, int checkOri, int* __restrict__ m12,
                                                           const int* __restrict__ hist, int* __restrict__ pairs,
                                                           int* __restrict__ cnt) {
    search_tri_finish_body(A, B, checkOri, m12, hist, pairs, cnt);
}
__global__ void __launch_bounds__(256) k_search_tri_finish_batch(const TriJob* __restrict__ jobs) {
    const TriJob& J = jobs[blockIdx.x];
    search_tri_finish_body(J.A, J.B, J.tp.checkOri, J.m12, J.hist, J.pairs, J.cnt);
}

// MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:483-548), one warp per map point.  Row i of the N x N distance matrix is
// produced by the lanes (descriptor i broadcast), kept in shared memory, and its median = the smallest value v with
// #{d <= v} >= int(0.5*(N-1)) + 1, found by a 9-step binary search over v in [0, 256] with ballot counts.
#define DD_WARPS 4
#define DD_MAXN 1024
__global__ void __launch_bounds__(32 * DD_WARPS) k_distinctive(const u8* __restrict__ desc, const int* __restrict__ offsets, int n_sets,
                                                                int* __restrict__ best_idx) {
    __shared__ u16 s_row[DD_WARPS][DD_MAXN];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, s = blockIdx.x * DD_WARPS + w;
    if (s >= n_sets) return;
    const int o = offsets[s], N = offsets[s + 1] - o;
    if (N <= 0) { if (lane == 0) best_idx[s] = -1; return; }
    const int m = (int)(0.5 * (N - 1));
    int bestMedian = 0x7fffffff, bestIdx = 0;
    for (int i = 0; i < N; i++) {
        u32 di[8];
        load_desc(desc + (size_t)(o + i) * 32, di);
        for (int j = lane; j < N; j += 32) {
            u32 dj[8];
            load_desc(desc + (size_t)(o + j) * 32, dj);
            s_row[w][j] = (u16)ham256(di, dj);
        }
        __syncwarp();
        int lo = 0, hi = 256;                                            // smallest v with count(d <= v) >= m + 1
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            int c = 0;
            for (int j0 = 0; j0 < N; j0 += 32) {
                const int j = j0 + lane;
                c += __popc(__ballot_sync(0xffffffffu, j < N && (int)s_row[w][j] <= mid));
            }
            if (c >= m + 1) hi = mid; else lo = mid + 1;
        }
        if (lo < bestMedian) { bestMedian = lo; bestIdx = i; }
        __syncwarp();
    }
    if (lane == 0) best_idx[s] = bestIdx;
}

// POPC issue-rate microbenchmark: 8 independent xor+popc chains per thread, register resident.
__global__ void k_popc_peak(u32* out, int iters) {
    u32 a[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x; acc[i] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) { acc[i] += __popc(a[i] ^ acc[i]); }
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// =====================================================================================================
// Host side
// =====================================================================================================
extern "C" int orbm_descriptor_distance(const uint8_t* a, const uint8_t* b, int n, int* dist, int device) {
    ORB_REQUIRE(a && b && dist && n >= 0, ORB_ERR_ARG, "bad arguments");
    int rc = check_device(device);
    if (rc) return rc;
    if (n == 0) return ORB_OK;
    Arena& A = g_arena;
    rc = A.ensure(device, 2 * pad((size_t)n * 32) + pad((size_t)n * 4));
    if (rc) return rc;
    const u8 *da, *db;
    if ((rc = upload(A, a, (size_t)n * 32, &da))) return rc;
    if ((rc = upload(A, b, (size_t)n * 32, &db))) return rc;
    if ((rc = A.flush())) return rc;
    int* dd = A.take<int>(n);
    k_pair_distance<<<orb_div_up(n, 256), 256, 0, A.stream>>>(da, db, n, dd);
    ORB_CUDA_TRY(cudaGetLastError());
    if ((rc = A.fetch(dist, dd, (size_t)n))) return rc;
    return A.finish();
}

static int launch_top2(const u8* d_q, const int* d_q_off, const int* d_q_cnt, const u8* d_db, const int* d_db_off,
                       const int* d_db_cnt, int npairs, int max_q, int* d_bi, int* d_bd, int* d_sd, cudaStream_t st) {
    if (npairs == 0 || max_q == 0) return ORB_OK;
    ORB_REQUIRE(npairs <= 65535, ORB_ERR_ARG, "npairs > 65535");
    // enough CTAs to fill 148 SMs decides how many queries each thread keeps in registers
    const long ctas4 = (long)orb_div_up(max_q, T2_THREADS * 4) * npairs;
    if (ctas4 >= 148 * 4) {
        dim3 g(orb_div_up(max_q, T2_THREADS * 4), npairs);
        k_top2<4><<<g, T2_THREADS, 0, st>>>(d_q, d_q_off, d_q_cnt, d_db, d_db_off, d_db_cnt, d_bi, d_bd, d_sd);
    } else if ((long)orb_div_up(max_q, T2_THREADS * 2) * npairs >= 148 * 2) {
        dim3 g(orb_div_up(max_q, T2_THREADS * 2), npairs);
        k_top2<2><<<g, T2_THREADS, 0, st>>>(d_q, d_q_off, d_q_cnt, d_db, d_db_off, d_db_cnt, d_bi, d_bd, d_sd);
    } else {
        dim3 g(orb_div_up(max_q, T2_THREADS), npairs);
        k_top2<1><<<g, T2_THREADS, 0, st>>>(d_q, d_q_off, d_q_cnt, d_db, d_db_off, d_db_cnt, d_bi, d_bd, d_sd);
    }
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

extern "C" int orbm_hamming_top2_batch_device(const uint8_t* d_q, const int* d_q_off, const int* d_q_cnt, const uint8_t* d_db,
                                              const int* d_db_off, const int* d_db_cnt, int npairs, int max_q, int* d_best_idx,
                                              int* d_best_dist, int* d_second_dist, void* stream) {
    ORB_REQUIRE(d_q && d_q_off && d_q_cnt && d_db && d_db_off && d_db_cnt && d_best_idx && d_best_dist && d_second_dist &&
                npairs >= 0 && max_q >= 0, ORB_ERR_ARG, "bad arguments");
    return launch_top2(d_q, d_q_off, d_q_cnt, d_db, d_db_off, d_db_cnt, npairs, max_q, d_best_idx, d_best_dist, d_second_dist,
                       (cudaStream_t)stream);
}

extern "C" int orbm_hamming_top2(const uint8_t* q, int nq, const uint8_t* db, int ndb, int* best_idx, int* best_dist,
                                 int* second_dist, int device) {
    ORB_REQUIRE(q && (db || ndb == 0) && best_idx && best_dist && second_dist && nq >= 0 && ndb >= 0, ORB_ERR_ARG, "bad arguments");
    ORB_REQUIRE(ndb <= (int)KEY_IDX_MASK, ORB_ERR_ARG, "database larger than 8388607 descriptors");
    int rc = check_device(device);
    if (rc) return rc;
    if (nq == 0) return ORB_OK;
    Arena& A = g_arena;
    rc = A.ensure(device, pad((size_t)nq * 32) + pad((size_t)ndb * 32) + 3 * pad((size_t)nq * 4) + pad((size_t)32 * nq * 4) + 4 * 256);
    if (rc) return rc;
    const u8 *dq, *ddb;
    if ((rc = upload(A, q, (size_t)nq * 32, &dq))) return rc;
    if ((rc = upload(A, db, (size_t)ndb * 32, &ddb))) return rc;
    const int meta[4] = {0, nq, 0, ndb};
    const int* dmeta;
    if ((rc = upload(A, meta, 4, &dmeta))) return rc;
    u32* u32part = nullptr;
    if ((rc = A.flush())) return rc;
    int *bi = A.take<int>(nq), *bd = A.take<int>(nq), *sd = A.take<int>(nq);
    if ((long)nq * ndb <= (64L << 20)) u32part = A.take<u32>((size_t)16 * 2 * nq);      // latency form only for small problems
    const int qtiles = orb_div_up(nq, T2_THREADS);
    int nslice = std::min(16, std::max(1, 296 / qtiles));
    int slice_len = (int)orb_align_up((size_t)orb_div_up(std::max(ndb, 1), nslice), 64);
    nslice = orb_div_up(std::max(ndb, 1), slice_len);
    if (nslice > 1 && u32part) {
        dim3 g(qtiles, nslice);
        k_top2_slice<<<g, T2_THREADS, 0, A.stream>>>(dq, nq, ddb, ndb, slice_len, u32part);
        k_top2_merge<<<orb_div_up(nq, 256), 256, 0, A.stream>>>(u32part, nq, nslice, bi, bd, sd);
        ORB_CUDA_TRY(cudaGetLastError());
    } else {
        rc = launch_top2(dq, dmeta, dmeta + 1, ddb, dmeta + 2, dmeta + 3, 1, nq, bi, bd, sd, A.stream);
        if (rc) return rc;
    }
    if ((rc = A.fetch(best_idx, bi, (size_t)nq))) return rc;
    if ((rc = A.fetch(best_dist, bd, (size_t)nq))) return rc;
    if ((rc = A.fetch(second_dist, sd, (size_t)nq))) return rc;
    return A.finish();
}

extern "C" int orbm_allpairs_device(const uint8_t* d_desc, int n_kf, int per_kf, int q_begin, int q_end, int th_low, float ratio,
                                    uint16_t* d_count, int* d_best_kf, int* d_best_dist, void* stream) {
    ORB_REQUIRE(d_desc && d_count && n_kf > 0 && per_kf > 0 && q_begin >= 0 && q_end >= q_begin && q_end <= n_kf, ORB_ERR_ARG, "bad arguments");
    ORB_REQUIRE(n_kf <= 65535 && (q_end - q_begin) <= 65535 && n_kf <= (int)KEY_IDX_MASK, ORB_ERR_ARG, "too many keyframes per call");
    ORB_REQUIRE((d_best_kf == nullptr) == (d_best_dist == nullptr), ORB_ERR_ARG, "d_best_kf and d_best_dist go together");
    ORB_REQUIRE(((uintptr_t)d_count & 3) == 0, ORB_ERR_ARG, "d_count must be 4-byte aligned");
    // two 16-bit counts share one 32-bit word that the kernel updates with atomicAdd: a count can reach per_kf
    ORB_REQUIRE(per_kf <= 65535, ORB_ERR_ARG, "per_kf = %d: the match counts are 16 bit", per_kf);
    cudaStream_t st = (cudaStream_t)stream;
    const int nq = q_end - q_begin;
    if (nq == 0) return ORB_OK;
    const size_t cnt_bytes = orb_align_up((size_t)nq * n_kf * 2, 4);
    ORB_CUDA_TRY(cudaMemsetAsync(d_count, 0, cnt_bytes, st));
    u32* packed = reinterpret_cast<u32*>(d_best_dist);           // packed (dist<<23|kf) during the kernel, unpacked after
    if (packed) ORB_CUDA_TRY(cudaMemsetAsync(packed, 0xFF, (size_t)nq * per_kf * 4, st));
    if (per_kf >= T2_THREADS * 4) {
        dim3 g(orb_div_up(per_kf, T2_THREADS * 4), n_kf, nq);
        k_allpairs<4><<<g, T2_THREADS, 0, st>>>(d_desc, n_kf, per_kf, q_begin, th_low, ratio, reinterpret_cast<u32*>(d_count), packed);
    } else {
        dim3 g(orb_div_up(per_kf, T2_THREADS), n_kf, nq);
        k_allpairs<1><<<g, T2_THREADS, 0, st>>>(d_desc, n_kf, per_kf, q_begin, th_low, ratio, reinterpret_cast<u32*>(d_count), packed);
    }
    ORB_CUDA_TRY(cudaGetLastError());
    if (packed) {
        const int n = nq * per_kf;
        k_unpack_best<<<orb_div_up(n, 256), 256, 0, st>>>(packed, n, d_best_kf, d_best_dist);
        ORB_CUDA_TRY(cudaGetLastError());
    }
    return ORB_OK;
}

// All-pairs keyframe matching on several GPUs of one node driven from ONE process (SURVEY §8e as a C entry: the C++ host of a
// map-merging / loop-closing service has no torch.distributed).  One host thread per device.  The descriptor database is uploaded once,
// to the first device, and replicated from there device-to-device (cudaMemcpyPeerAsync: NVLink where peer access exists); query keyframes
// are sharded contiguously over the devices; the gather of the per-shard match tables into the caller's table is the only exchange.
// The same device may be listed more than once (its shards then run on separate streams), which is how the single-GPU tests cover the
// sharding logic.
#include <thread>
extern "C" int orbm_allpairs_multi(const uint8_t* desc, int n_kf, int per_kf, int th_low, float ratio, const int* devices, int n_devices,
                                   uint16_t* count_out, int* best_kf_out, int* best_dist_out) {
    ORB_REQUIRE(desc && count_out && devices && n_devices > 0 && n_devices <= 64 && n_kf > 0 && per_kf > 0, ORB_ERR_ARG, "bad arguments");
    ORB_REQUIRE((best_kf_out == nullptr) == (best_dist_out == nullptr), ORB_ERR_ARG, "best_kf_out and best_dist_out go together");
    ORB_REQUIRE(n_kf <= 65535 && per_kf <= 65535, ORB_ERR_ARG, "n_kf / per_kf above 65535");
    const int ndev = orb_device_count();
    ORB_REQUIRE(ndev > 0, ORB_ERR_CUDA, "no CUDA device (no CPU fallback)");
    for (int i = 0; i < n_devices; i++) ORB_REQUIRE(devices[i] >= 0 && devices[i] < ndev, ORB_ERR_CUDA, "CUDA device %d not available", devices[i]);
    const size_t dbBytes = (size_t)n_kf * per_kf * 32;
    struct Shard {
        int dev = 0, q0 = 0, q1 = 0, rc = ORB_OK;
        u8* d_desc = nullptr; uint16_t* d_count = nullptr; int *d_bk = nullptr, *d_bd = nullptr;
        cudaStream_t st = nullptr; cudaEvent_t ready = nullptr;
        bool ownsDb = false;
        char err[256] = "";
    };
    std::vector<Shard> sh(n_devices);
    const int base = n_kf / n_devices, rem = n_kf % n_devices;
    for (int i = 0; i < n_devices; i++) {
        sh[i].dev = devices[i];
        sh[i].q0 = i * base + std::min(i, rem);
        sh[i].q1 = sh[i].q0 + base + (i < rem ? 1 : 0);
    }
    // phase 1 (this thread): allocations, the one upload, and the device-to-device replication, all asynchronous
    auto fail = [&](Shard& s, cudaError_t e, const char* what) {
        s.rc = ORB_ERR_CUDA;
        snprintf(s.err, sizeof(s.err), "%s failed on device %d: %s", what, s.dev, cudaGetErrorString(e));
        cudaGetLastError();
    };
#define MG_TRY(s, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(s, e_, #call); goto done; } } while (0)
    {
        for (int i = 0; i < n_devices; i++) {
            Shard& s = sh[i];
            MG_TRY(s, cudaSetDevice(s.dev));
            MG_TRY(s, cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
            MG_TRY(s, cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
            // devices listed twice share the first listing's copy of the database
            int first = i;
            for (int j = 0; j < i; j++) if (sh[j].dev == s.dev) { first = j; break; }
            if (first == i) { MG_TRY(s, cudaMalloc(&s.d_desc, dbBytes)); s.ownsDb = true; }
            else s.d_desc = sh[first].d_desc;
            const int nq = s.q1 - s.q0;
            MG_TRY(s, cudaMalloc(&s.d_count, orb_align_up((size_t)std::max(nq, 1) * n_kf * 2, 4)));
            if (best_kf_out) {
                MG_TRY(s, cudaMalloc(&s.d_bk, (size_t)std::max(nq, 1) * per_kf * 4));
                MG_TRY(s, cudaMalloc(&s.d_bd, (size_t)std::max(nq, 1) * per_kf * 4));
            }
        }
        MG_TRY(sh[0], cudaSetDevice(sh[0].dev));
        MG_TRY(sh[0], cudaMemcpyAsync(sh[0].d_desc, desc, dbBytes, cudaMemcpyHostToDevice, sh[0].st));
        MG_TRY(sh[0], cudaEventRecord(sh[0].ready, sh[0].st));
        for (int i = 1; i < n_devices; i++) {
            Shard& s = sh[i];
            MG_TRY(s, cudaSetDevice(s.dev));
            int src = 0;
            for (int j = 0; j < i; j++) if (sh[j].dev == s.dev) { src = j; break; }
            MG_TRY(s, cudaStreamWaitEvent(s.st, sh[src].ready, 0));        // the copy this shard reads (its own device's, or device 0's) has landed
            if (s.ownsDb) {
                MG_TRY(s, cudaMemcpyPeerAsync(s.d_desc, s.dev, sh[0].d_desc, sh[0].dev, dbBytes, s.st));
            }
            MG_TRY(s, cudaEventRecord(s.ready, s.st));
        }
    }
    // phase 2: one host thread per shard — kernel, download of its rows straight into the caller's table
    {
        std::vector<std::thread> th;
        for (int i = 0; i < n_devices; i++)
            th.emplace_back([&, i] {
                Shard& s = sh[i];
                const int nq = s.q1 - s.q0;
                auto chk = [&](cudaError_t e, const char* what) { if (e != cudaSuccess && s.rc == ORB_OK) fail(s, e, what); return e == cudaSuccess; };
                if (!chk(cudaSetDevice(s.dev), "cudaSetDevice")) return;
                if (nq > 0) {
                    const int rc = orbm_allpairs_device(s.d_desc, n_kf, per_kf, s.q0, s.q1, th_low, ratio, s.d_count, s.d_bk, s.d_bd, s.st);
                    if (rc != ORB_OK) { s.rc = rc; snprintf(s.err, sizeof(s.err), "%s", orb_last_error()); return; }
                    if (!chk(cudaMemcpyAsync(count_out + (size_t)s.q0 * n_kf, s.d_count, (size_t)nq * n_kf * 2, cudaMemcpyDeviceToHost, s.st), "D2H counts")) return;
                    if (best_kf_out) {
                        if (!chk(cudaMemcpyAsync(best_kf_out + (size_t)s.q0 * per_kf, s.d_bk, (size_t)nq * per_kf * 4, cudaMemcpyDeviceToHost, s.st), "D2H best_kf")) return;
                        if (!chk(cudaMemcpyAsync(best_dist_out + (size_t)s.q0 * per_kf, s.d_bd, (size_t)nq * per_kf * 4, cudaMemcpyDeviceToHost, s.st), "D2H best_dist")) return;
                    }
                }
                chk(cudaStreamSynchronize(s.st), "cudaStreamSynchronize");
            });
        for (std::thread& t : th) t.join();
    }
done:
#undef MG_TRY
    int rc = ORB_OK;
    for (int i = 0; i < n_devices; i++) {
        Shard& s = sh[i];
        if (s.rc != ORB_OK && rc == ORB_OK) { rc = s.rc; orb_set_error("%s", s.err); }
        if (cudaSetDevice(s.dev) != cudaSuccess) { cudaGetLastError(); continue; }
        if (s.st) cudaStreamSynchronize(s.st);
    }
    for (int i = 0; i < n_devices; i++) {
        Shard& s = sh[i];
        if (cudaSetDevice(s.dev) != cudaSuccess) { cudaGetLastError(); continue; }
        if (s.ownsDb) cudaFree(s.d_desc);
        cudaFree(s.d_count); cudaFree(s.d_bk); cudaFree(s.d_bd);
        if (s.ready) cudaEventDestroy(s.ready);
        if (s.st) cudaStreamDestroy(s.st);
        cudaGetLastError();
    }
    return rc;
}

static size_t view_bytes(const orbm_view* v, bool tri) {
    size_t b = pad((size_t)v->n * 32) + pad(v->n) + pad((size_t)v->n * 4);
    if (tri) b += 4 * pad((size_t)v->n * 4);
    b += pad((size_t)v->fv.n_nodes * 4) + pad((size_t)(v->fv.n_nodes + 1) * 4);
    const int nf = v->fv.n_nodes > 0 ? v->fv.offsets[v->fv.n_nodes] : 0;
    b += pad((size_t)nf * 4);
    return b;
}
static int check_view(const orbm_view* v, bool needFlag, bool tri) {
    ORB_REQUIRE(v && v->n >= 0 && v->fv.n_nodes >= 0, ORB_ERR_ARG, "bad view");
    ORB_REQUIRE(v->n == 0 || v->desc, ORB_ERR_ARG, "view.desc is NULL");
    ORB_REQUIRE(!needFlag || v->n == 0 || v->flag, ORB_ERR_ARG, "view.flag is NULL");
    ORB_REQUIRE(v->n == 0 || v->angle, ORB_ERR_ARG, "view.angle is NULL");
    ORB_REQUIRE(!tri || v->n == 0 || (v->x && v->y && v->octave && v->uright), ORB_ERR_ARG, "triangulation needs x, y, octave, uright");
    ORB_REQUIRE(v->fv.n_nodes == 0 || (v->fv.node_ids && v->fv.offsets && v->fv.features), ORB_ERR_ARG, "bad feature vector");
    const int nf = v->fv.n_nodes > 0 ? v->fv.offsets[v->fv.n_nodes] : 0;
    for (int i = 0; i < v->fv.n_nodes; i++) {
        ORB_REQUIRE(v->fv.offsets[i] <= v->fv.offsets[i + 1], ORB_ERR_ARG, "feature vector offsets not monotone");
        ORB_REQUIRE(i == 0 || v->fv.node_ids[i - 1] < v->fv.node_ids[i], ORB_ERR_ARG, "feature vector node ids not ascending");
    }
    for (int i = 0; i < nf; i++) ORB_REQUIRE(v->fv.features[i] >= 0 && v->fv.features[i] < v->n, ORB_ERR_ARG, "feature index out of range");
    return ORB_OK;
}
static int upload_view(Arena& A, const orbm_view* v, bool tri, DevView* d) {
    int rc;
    d->n = v->n; d->nn = v->fv.n_nodes;
    d->flag = nullptr; d->x = d->y = d->uright = nullptr; d->octave = nullptr;
    if ((rc = upload(A, v->desc, (size_t)v->n * 32, &d->desc))) return rc;
    if (v->flag) { if ((rc = upload(A, v->flag, (size_t)v->n, &d->flag))) return rc; }
    if ((rc = upload(A, v->angle, (size_t)v->n, &d->angle))) return rc;
    if (tri) {
        if ((rc = upload(A, v->x, (size_t)v->n, &d->x))) return rc;
        if ((rc = upload(A, v->y, (size_t)v->n, &d->y))) return rc;
        if ((rc = upload(A, v->octave, (size_t)v->n, &d->octave))) return rc;
        if ((rc = upload(A, v->uright, (size_t)v->n, &d->uright))) return rc;
    }
    const int nf = v->fv.n_nodes > 0 ? v->fv.offsets[v->fv.n_nodes] : 0;
    if ((rc = upload(A, v->fv.node_ids, (size_t)v->fv.n_nodes, &d->ids))) return rc;
    if (v->fv.n_nodes > 0) { if ((rc = upload(A, v->fv.offsets, (size_t)v->fv.n_nodes + 1, &d->off))) return rc; }
    else { static const int zero = 0; if ((rc = upload(A, &zero, 1, &d->off))) return rc; }
    if ((rc = upload(A, v->fv.features, (size_t)nf, &d->feat))) return rc;
    return ORB_OK;
}

template <bool KFKF>
static int search_bow(const orbm_view* v1, const orbm_view* v2, float nnratio, int checkOri, int* match, int* n_matches, int device) {
    ORB_REQUIRE(match && n_matches, ORB_ERR_ARG, "null output");
    int rc = check_device(device);
    if (rc) return rc;
    if ((rc = check_view(v1, true, false))) return rc;
    if ((rc = check_view(v2, KFKF, false))) return rc;
    const int nOut = KFKF ? v1->n : v2->n;
    Arena& A = g_arena;
    rc = A.ensure(device, view_bytes(v1, false) + view_bytes(v2, false) + 2 * pad((size_t)nOut * 4) + pad((size_t)v2->n * 4) + 1024);
    if (rc) return rc;
    DevView d1, d2;
    if ((rc = upload_view(A, v1, false, &d1))) return rc;
    if ((rc = upload_view(A, v2, false, &d2))) return rc;
    if ((rc = A.flush())) return rc;
    int* d_out = A.take<int>(std::max(nOut, 1));
    int* d_bin = A.take<int>(std::max(nOut, 1));
    int* d_zero = A.take<int>((size_t)std::max(v2->n, 1) + ORBM_HISTO_LENGTH + 1);     // taken2 | hist | nmatch, one memset
    int* d_taken = d_zero;
    int* d_hist = d_zero + std::max(v2->n, 1);
    int* d_nm = d_hist + ORBM_HISTO_LENGTH;
    ORB_CUDA_TRY(cudaMemsetAsync(d_out, 0xFF, (size_t)std::max(nOut, 1) * 4, A.stream));
    ORB_CUDA_TRY(cudaMemsetAsync(d_zero, 0, ((size_t)std::max(v2->n, 1) + ORBM_HISTO_LENGTH + 1) * 4, A.stream));
    if (v1->fv.n_nodes > 0)
        k_search_bow_nodes<KFKF><<<orb_div_up(v1->fv.n_nodes, SB_WARPS), 32 * SB_WARPS, 0, A.stream>>>(d1, d2, nnratio, checkOri, d_out,
                                                                                                     d_taken, d_bin, d_hist, d_nm);
    k_search_bow_finish<<<1, 256, 0, A.stream>>>(nOut, checkOri, d_out, d_bin, d_hist, d_nm);
    ORB_CUDA_TRY(cudaGetLastError());
    if ((rc = A.fetch(match, d_out, (size_t)nOut))) return rc;
    if ((rc = A.fetch(n_matches, d_nm, 1))) return rc;
    return A.finish();
}

extern "C" int orbm_search_by_bow_kf_frame(const orbm_view* kf, const orbm_view* frame, float nnratio, int check_orientation,
                                           int* match21, int* n_matches, int device) {
    return search_bow<false>(kf, frame, nnratio, check_orientation, match21, n_matches, device);
}
extern "C" int orbm_search_by_bow_kf_kf(const orbm_view* kf1, const orbm_view* kf2, float nnratio, int check_orientation,
                                        int* match12, int* n_matches, int device) {
    return search_bow<true>(kf1, kf2, nnratio, check_orientation, match12, n_matches, device);
}

extern "C" int orbm_search_for_triangulation(const orbm_view* kf1, const orbm_view* kf2, const float* F12, float ex, float ey,
                                             const float* scale_factors2, const float* level_sigma2_2, int n_levels2, int only_stereo,
                                             int check_orientation, int* pairs_out, int* n_pairs, int* n_matches, int device) {
    ORB_REQUIRE(F12 && scale_factors2 && level_sigma2_2 && n_levels2 > 0 && pairs_out && n_pairs && n_matches, ORB_ERR_ARG, "bad arguments");
    int rc = check_device(device);
    if (rc) return rc;
    if ((rc = check_view(kf1, true, true))) return rc;
    if ((rc = check_view(kf2, true, true))) return rc;
    for (int i = 0; i < kf2->n; i++) ORB_REQUIRE(kf2->octave[i] >= 0 && kf2->octave[i] < n_levels2, ORB_ERR_ARG, "octave out of range");
    Arena& A = g_arena;
    rc = A.ensure(device, view_bytes(kf1, true) + view_bytes(kf2, true) + 3 * pad((size_t)kf1->n * 4) + 2 * pad((size_t)n_levels2 * 4) + 1024);
    if (rc) return rc;
    DevView d1, d2;
    if ((rc = upload_view(A, kf1, true, &d1))) return rc;
    if ((rc = upload_view(A, kf2, true, &d2))) return rc;
    const float *d_sf, *d_s2;
    if ((rc = upload(A, scale_factors2, (size_t)n_levels2, &d_sf))) return rc;
    if ((rc = upload(A, level_sigma2_2, (size_t)n_levels2, &d_s2))) return rc;
    if ((rc = A.flush())) return rc;
    int* d_m12 = A.take<int>(std::max(kf1->n, 1));
    int* d_pairs = A.take<int>(std::max(2 * kf1->n, 1));
    int* d_cnt = A.take<int>(2 + ORBM_HISTO_LENGTH);          // cnt[2] | hist, one memset
    int* d_hist = d_cnt + 2;
    TriParams tp;
    for (int i = 0; i < 9; i++) tp.F[i] = F12[i];
    tp.ex = ex; tp.ey = ey; tp.onlyStereo = only_stereo; tp.checkOri = check_orientation;
    ORB_CUDA_TRY(cudaMemsetAsync(d_m12, 0xFF, (size_t)std::max(kf1->n, 1) * 4, A.stream));
    ORB_CUDA_TRY(cudaMemsetAsync(d_cnt, 0, (2 + ORBM_HISTO_LENGTH) * 4, A.stream));
    {
        const int total1 = kf1->fv.n_nodes > 0 ? kf1->fv.offsets[kf1->fv.n_nodes] : 0;
        if (total1 > 0) k_search_tri_slots<<<orb_div_up(total1, 128), 128, 0, A.stream>>>(d1, d2, tp, d_sf, d_s2, d_m12, d_hist, d_cnt);
    }
    k_search_tri_finish<<<1, 256, 0, A.stream>>>(d1, d2, check_orientation, d_m12, d_hist, d_pairs, d_cnt);
    ORB_CUDA_TRY(cudaGetLastError());
    int cnt[2] = {0, 0};
    if (A.staged) {            // small call: bring counts and the whole pair buffer back with one synchronisation
        std::vector<int> tmp(std::max(2 * kf1->n, 1));
        if ((rc = A.fetch(cnt, d_cnt, 2))) return rc;
        if ((rc = A.fetch(tmp.data(), d_pairs, (size_t)2 * kf1->n))) return rc;
        if ((rc = A.finish())) return rc;
        if (cnt[0]) memcpy(pairs_out, tmp.data(), (size_t)cnt[0] * 8);
    } else {
        if ((rc = A.fetch(cnt, d_cnt, 2))) return rc;
        if ((rc = A.finish())) return rc;
        if (cnt[0]) {
            if ((rc = A.fetch(pairs_out, d_pairs, (size_t)cnt[0] * 2))) return rc;
            if ((rc = A.finish())) return rc;
        }
    }
    *n_pairs = cnt[0]; *n_matches = cnt[1];
    return ORB_OK;
}

// ---- batched node-constrained searches: one view against N others, two launches for the whole batch -----------------------------
template <bool KFKF>
static int search_bow_batch(const orbm_view* anchor, const orbm_view* others, int n_others, float nnratio, int checkOri, int* match_out,
                            int* n_matches, int device) {
    // KFKF = false: SearchByBoW(others[i] (KeyFrame), anchor (Frame)); KFKF = true: SearchByBoW(anchor (KF1), others[i] (KF2)).
    // Either way the output rows are anchor->n long.
    ORB_REQUIRE(match_out && n_matches && n_others >= 0 && (n_others == 0 || others), ORB_ERR_ARG, "bad arguments");
    int rc = check_device(device);
    if (rc) return rc;
    if ((rc = check_view(anchor, KFKF, false))) return rc;
    for (int i = 0; i < n_others; i++) if ((rc = check_view(&others[i], true, false))) return rc;
    if (n_others == 0) return ORB_OK;
    const int nOut = anchor->n, nOutP = std::max(nOut, 1);
    size_t bytes = view_bytes(anchor, false) + pad((size_t)n_others * sizeof(BowJob)) + pad((size_t)n_others * 4) + 1024, maxTaken = 0;
    int maxNodes = 0;
    for (int i = 0; i < n_others; i++) {
        const int n2 = KFKF ? others[i].n : anchor->n;                     // side 2 of pair i
        bytes += view_bytes(&others[i], false) + 2 * pad((size_t)nOutP * 4) + pad(((size_t)std::max(n2, 1) + ORBM_HISTO_LENGTH + 1) * 4);
        maxTaken += (size_t)std::max(n2, 1) + ORBM_HISTO_LENGTH + 1;
        maxNodes = std::max(maxNodes, KFKF ? anchor->fv.n_nodes : others[i].fv.n_nodes);
    }
    Arena& A = g_arena;
    if ((rc = A.ensure(device, bytes))) return rc;
    DevView dA;
    if ((rc = upload_view(A, anchor, false, &dA))) return rc;
    std::vector<BowJob> jobs(n_others);
    for (int i = 0; i < n_others; i++) {
        DevView dO;
        if ((rc = upload_view(A, &others[i], false, &dO))) return rc;
        jobs[i].A = KFKF ? dA : dO;
        jobs[i].B = KFKF ? dO : dA;
        jobs[i].nOut = nOut;
    }
    const BowJob* d_jobs = nullptr;
    BowJob* d_jobs_w = A.take<BowJob>(n_others);                            // filled below, after the output pointers are known
    (void)d_jobs;
    // outputs: out rows back to back (one fetch), then bins, then the zeroed scratch of every pair back to back (one memset)
    int* d_out = A.take<int>((size_t)nOutP * n_others);
    int* d_bin = A.take<int>((size_t)nOutP * n_others);
    int* d_zero = A.take<int>(maxTaken);
    int* d_nm = A.take<int>(n_others);
    size_t zo = 0;
    for (int i = 0; i < n_others; i++) {
        const int n2 = std::max(KFKF ? others[i].n : anchor->n, 1);
        jobs[i].out = d_out + (size_t)i * nOutP;
        jobs[i].binOf = d_bin + (size_t)i * nOutP;
        jobs[i].taken2 = d_zero + zo;
        jobs[i].hist = d_zero + zo + n2;
        jobs[i].nmatch = d_nm + i;
        zo += (size_t)n2 + ORBM_HISTO_LENGTH + 1;
    }
    if (A.staged) memcpy(A.hbase + (reinterpret_cast<u8*>(d_jobs_w) - A.base), jobs.data(), jobs.size() * sizeof(BowJob));
    if ((rc = A.flush())) return rc;                                        // (flush sends everything taken so far, the job table included)
    if (!A.staged) ORB_CUDA_TRY(cudaMemcpyAsync(d_jobs_w, jobs.data(), jobs.size() * sizeof(BowJob), cudaMemcpyHostToDevice, A.stream));
    ORB_CUDA_TRY(cudaMemsetAsync(d_out, 0xFF, (size_t)nOutP * n_others * 4, A.stream));
    ORB_CUDA_TRY(cudaMemsetAsync(d_zero, 0, maxTaken * 4, A.stream));
    ORB_CUDA_TRY(cudaMemsetAsync(d_nm, 0, (size_t)n_others * 4, A.stream));
    if (maxNodes > 0) {
        dim3 g(orb_div_up(maxNodes, SB_WARPS), n_others);
        k_search_bow_nodes_batch<KFKF><<<g, 32 * SB_WARPS, 0, A.stream>>>(d_jobs_w, nnratio, checkOri);
    }
    k_search_bow_finish_batch<<<n_others, 256, 0, A.stream>>>(d_jobs_w, checkOri);
    ORB_CUDA_TRY(cudaGetLastError());
    if (nOut > 0) {
        if (nOutP == nOut) { if ((rc = A.fetch(match_out, d_out, (size_t)nOut * n_others))) return rc; }
    }
    if ((rc = A.fetch(n_matches, d_nm, (size_t)n_others))) return rc;
    return A.finish();
}

extern "C" int orbm_search_by_bow_batch(const orbm_view* anchor, const orbm_view* others, int n_others, int mode, float nnratio,
                                        int check_orientation, int* match_out, int* n_matches, int device) {
    ORB_REQUIRE(mode == 0 || mode == 1, ORB_ERR_ARG, "mode must be 0 (KeyFrame others[i], Frame anchor) or 1 (KeyFrame anchor, KeyFrame others[i])");
    return mode ? search_bow_batch<true>(anchor, others, n_others, nnratio, check_orientation, match_out, n_matches, device)
                : search_bow_batch<false>(anchor, others, n_others, nnratio, check_orientation, match_out, n_matches, device);
}

extern "C" int orbm_search_for_triangulation_batch(const orbm_view* anchor, const orbm_view* others, int n_others, const float* F12,
                                                   const float* epipoles, const float* scale_factors2, const float* level_sigma2_2,
                                                   int n_levels2, int only_stereo, int check_orientation, int* pairs_out, int* n_pairs,
                                                   int* n_matches, int device) {
    ORB_REQUIRE(n_others >= 0 && (n_others == 0 || (others && F12 && epipoles && scale_factors2 && level_sigma2_2)) && n_levels2 > 0 &&
                pairs_out && n_pairs && n_matches, ORB_ERR_ARG, "bad arguments");
    int rc = check_device(device);
    if (rc) return rc;
    if ((rc = check_view(anchor, true, true))) return rc;
    for (int i = 0; i < n_others; i++) {
        if ((rc = check_view(&others[i], true, true))) return rc;
        for (int j = 0; j < others[i].n; j++)
            ORB_REQUIRE(others[i].octave[j] >= 0 && others[i].octave[j] < n_levels2, ORB_ERR_ARG, "octave out of range");
    }
    if (n_others == 0) return ORB_OK;
    const int n1 = anchor->n, n1P = std::max(n1, 1);
    size_t bytes = view_bytes(anchor, true) + pad((size_t)n_others * sizeof(TriJob)) + 2 * pad((size_t)n_others * n_levels2 * 4) + 1024;
    for (int i = 0; i < n_others; i++) bytes += view_bytes(&others[i], true) + 3 * pad((size_t)n1P * 4) + pad((2 + ORBM_HISTO_LENGTH) * 4);
    Arena& A = g_arena;
    if ((rc = A.ensure(device, bytes))) return rc;
    DevView dA;
    if ((rc = upload_view(A, anchor, true, &dA))) return rc;
    const float *d_sf, *d_s2;
    if ((rc = upload(A, scale_factors2, (size_t)n_others * n_levels2, &d_sf))) return rc;
    if ((rc = upload(A, level_sigma2_2, (size_t)n_others * n_levels2, &d_s2))) return rc;
    std::vector<TriJob> jobs(n_others);
    const int total1 = anchor->fv.n_nodes > 0 ? anchor->fv.offsets[anchor->fv.n_nodes] : 0;
    for (int i = 0; i < n_others; i++) {
        if ((rc = upload_view(A, &others[i], true, &jobs[i].B))) return rc;
        jobs[i].A = dA;
        for (int k = 0; k < 9; k++) jobs[i].tp.F[k] = F12[9 * (size_t)i + k];
        jobs[i].tp.ex = epipoles[2 * (size_t)i]; jobs[i].tp.ey = epipoles[2 * (size_t)i + 1];
        jobs[i].tp.onlyStereo = only_stereo; jobs[i].tp.checkOri = check_orientation;
        jobs[i].sf2 = d_sf + (size_t)i * n_levels2;
        jobs[i].sigma2 = d_s2 + (size_t)i * n_levels2;
    }
    TriJob* d_jobs = A.take<TriJob>(n_others);
    int* d_m12 = A.take<int>((size_t)n1P * n_others);
    int* d_pairs = A.take<int>((size_t)2 * n1P * n_others);
    int* d_cnt = A.take<int>((size_t)(2 + ORBM_HISTO_LENGTH) * n_others);   // per pair: cnt[2] | hist
    for (int i = 0; i < n_others; i++) {
        jobs[i].m12 = d_m12 + (size_t)i * n1P;
        jobs[i].pairs = d_pairs + (size_t)2 * i * n1P;
        jobs[i].cnt = d_cnt + (size_t)i * (2 + ORBM_HISTO_LENGTH);
        jobs[i].hist = jobs[i].cnt + 2;
    }
    if (A.staged) memcpy(A.hbase + (reinterpret_cast<u8*>(d_jobs) - A.base), jobs.data(), jobs.size() * sizeof(TriJob));
    if ((rc = A.flush())) return rc;
    if (!A.staged) ORB_CUDA_TRY(cudaMemcpyAsync(d_jobs, jobs.data(), jobs.size() * sizeof(TriJob), cudaMemcpyHostToDevice, A.stream));
    ORB_CUDA_TRY(cudaMemsetAsync(d_m12, 0xFF, (size_t)n1P * n_others * 4, A.stream));
    ORB_CUDA_TRY(cudaMemsetAsync(d_cnt, 0, (size_t)(2 + ORBM_HISTO_LENGTH) * n_others * 4, A.stream));
    if (total1 > 0) {
        dim3 g(orb_div_up(total1, 128), n_others);
        k_search_tri_slots_batch<<<g, 128, 0, A.stream>>>(d_jobs);
    }
    k_search_tri_finish_batch<<<n_others, 256, 0, A.stream>>>(d_jobs);
    ORB_CUDA_TRY(cudaGetLastError());
    std::vector<int> cnt((size_t)(2 + ORBM_HISTO_LENGTH) * n_others), tmp((size_t)2 * n1P * n_others);
    if ((rc = A.fetch(cnt.data(), d_cnt, cnt.size()))) return rc;
    if (n1 > 0 && (rc = A.fetch(tmp.data(), d_pairs, tmp.size()))) return rc;
    if ((rc = A.finish())) return rc;
    for (int i = 0; i < n_others; i++) {
        n_pairs[i] = cnt[(size_t)i * (2 + ORBM_HISTO_LENGTH)];
        n_matches[i] = cnt[(size_t)i * (2 + ORBM_HISTO_LENGTH) + 1];
        if (n_pairs[i]) memcpy(pairs_out + (size_t)2 * i * n1, tmp.data() + (size_t)2 * i * n1P, (size_t)n_pairs[i] * 8);
    }
    return ORB_OK;
}

extern "C" int orbm_distinctive_descriptors(const uint8_t* desc, const int* offsets, int n_sets, int* best_idx, int device) {
    ORB_REQUIRE(offsets && best_idx && n_sets >= 0, ORB_ERR_ARG, "bad arguments");
    int rc = check_device(device);
    if (rc) return rc;
    if (n_sets == 0) return ORB_OK;
    const int total = offsets[n_sets];
    ORB_REQUIRE(total >= 0 && (desc || total == 0), ORB_ERR_ARG, "bad arguments");
    for (int s2 = 0; s2 < n_sets; s2++)
        ORB_REQUIRE(offsets[s2 + 1] >= offsets[s2] && offsets[s2 + 1] - offsets[s2] <= DD_MAXN, ORB_ERR_ARG,
                    "set %d has %d descriptors (limit %d)", s2, offsets[s2 + 1] - offsets[s2], DD_MAXN);
    Arena& A = g_arena;
    if ((rc = A.ensure(device, pad((size_t)total * 32) + pad((size_t)(n_sets + 1) * 4) + pad((size_t)n_sets * 4)))) return rc;
    const u8* dd;
    const int* doff;
    if ((rc = upload(A, desc, (size_t)total * 32, &dd))) return rc;
    if ((rc = upload(A, offsets, (size_t)n_sets + 1, &doff))) return rc;
    if ((rc = A.flush())) return rc;
    int* dbest = A.take<int>(n_sets);
    k_distinctive<<<orb_div_up(n_sets, DD_WARPS), 32 * DD_WARPS, 0, A.stream>>>(dd, doff, n_sets, dbest);
    ORB_CUDA_TRY(cudaGetLastError());
    if ((rc = A.fetch(best_idx, dbest, (size_t)n_sets))) return rc;
    return A.finish();
}

extern "C" int orbm_three_maxima(const int* histo, int n_bins, int* ind, int device) {
    ORB_REQUIRE(histo && ind && n_bins > 0, ORB_ERR_ARG, "bad arguments");
    int rc = check_device(device);
    if (rc) return rc;
    Arena& A = g_arena;
    if ((rc = A.ensure(device, pad((size_t)n_bins * 4) + 256))) return rc;
    const int* dh;
    if ((rc = upload(A, histo, (size_t)n_bins, &dh))) return rc;
    if ((rc = A.flush())) return rc;
    int* di = A.take<int>(3);
    k_three_maxima<<<1, 1, 0, A.stream>>>(dh, n_bins, di);
    ORB_CUDA_TRY(cudaGetLastError());
    if ((rc = A.fetch(ind, di, 3))) return rc;
    return A.finish();
}

extern "C" int orbm_popc_peak(int device, double* popc_per_second, double* sm_clock_hz_used) {
    ORB_REQUIRE(popc_per_second, ORB_ERR_ARG, "null output");
    int rc = check_device(device);
    if (rc) return rc;
    ORB_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    ORB_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    u32* d_out;
    ORB_CUDA_TRY(cudaMalloc(&d_out, (size_t)blocks * threads * 4));
    cudaEvent_t e0, e1;
    ORB_CUDA_TRY(cudaEventCreate(&e0));
    ORB_CUDA_TRY(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        ORB_CUDA_TRY(cudaEventRecord(e0));
        k_popc_peak<<<blocks, threads>>>(d_out, iters);
        ORB_CUDA_TRY(cudaEventRecord(e1));
        ORB_CUDA_TRY(cudaEventSynchronize(e1));
        float ms;
        ORB_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best_ms = std::min(best_ms, ms);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    *popc_per_second = (double)blocks * threads * iters * 8.0 / (best_ms * 1e-3);
    if (sm_clock_hz_used) { int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device); *sm_clock_hz_used = khz * 1e3; }
    return ORB_OK;
}
