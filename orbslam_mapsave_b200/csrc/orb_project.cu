// orb_project.cu — B200 (sm_100a) implementation of the ORBmatcher searches that look inside a window of a Frame's /
// KeyFrame's feature grid: SearchByProjection(Frame, MapPoints) src/ORBmatcher.cc:45-129,
// SearchByProjection(CurrentFrame, LastFrame) :1331-1463, SearchForInitialization :408-523, on top of
// Frame::GetFeaturesInArea src/Frame.cc:445-498.
//
// The reference walks its queries serially because an accepted match changes what later queries may take
// (F.mvpMapPoints[idx] / vMatchedDistance[idx]).  Here every call is three launches:
//   1. k_win_*        one thread per query: the function-specific part (projection, radius, level range, stereo gate) -> Win
//   2. k_win_candidates  one warp per query: GetFeaturesInArea in the reference's order (cells ix-major, iy-minor, push order
//                     inside a cell; lanes = cells, order-preserving write through a warp scan) + the Hamming distances
//   3. k_win_resolve  one CTA: replays the serial claims as data-parallel ROUNDS.  In a round every unresolved query picks its
//                     best (and second) candidate among those still available; a query is FINAL when no earlier unresolved
//                     query could still take one of those two (minq[c] = lowest unresolved query that holds c with a distance
//                     it could be accepted with).  The lowest unresolved query is always final, so the loop terminates, and
//                     claims on one feature are finalised in query order, which is all the serial loop guarantees.
//                     Typical calls need 2-6 rounds.
// Integer / float32 work only; float expressions use explicit _rn operations (no FMA contraction), like the reference built
// without -ffp-contract.
#include "orb_match_common.cuh"

#include <climits>

#define WL_SHIFT 22                         // candidate entry = distance << 22 | feature index
#define WL_IDX_MASK 0x3FFFFFu
#define WIN_ACTIVE 1
#define WIN_CLAIMS 2
#define WIN_GATE 4
#define WIN_CHI2 8                          // Fuse: per-candidate reprojection-error gate (src/ORBmatcher.cc:913-944)

enum { MODE_BEST = 0, MODE_TOP2_LEVEL = 1, MODE_INIT = 2 };

struct DevGrid {
    int n;
    const u8* desc; const float* x; const float* y; const int* octave; const float* angle; const float* uright; const u8* blocked;
    int cols, rows;
    float min_x, min_y, max_x, max_y, inv_w, inv_h;
    const int* cell_off; const int* cell_feat;
    const float* sf; int n_levels;
    const float* inv_sigma2;                // mvInvLevelSigma2 (Fuse only)
};
struct __align__(16) Win { float u, v, r, ur, tol; int minL, maxL, flags; };

// ---- 1. windows --------------------------------------------------------------------------------------------------------------
// SearchByProjection(Frame, MapPoints): ORBmatcher.cc:50-66 (radius by viewing cosine, window r*scale, levels [l-1, l])
__global__ void k_win_map(DevGrid G, int nq, const u8* __restrict__ in_view, const float* __restrict__ px, const float* __restrict__ py,
                          const float* __restrict__ pxr, const int* __restrict__ level, const float* __restrict__ vcos,
                          const u8* __restrict__ claims, float th, Win* __restrict__ win) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    Win w = {0.f, 0.f, 0.f, 0.f, 0.f, -1, -1, 0};
    if (in_view[q]) {
        const int l = level[q];
        float r = ((double)vcos[q] > 0.998) ? 2.5f : 4.0f;
        if ((double)th != 1.0) r = __fmul_rn(r, th);
        w.u = px[q]; w.v = py[q];
        w.r = __fmul_rn(r, G.sf[l]);
        w.ur = pxr[q]; w.tol = w.r;
        w.minL = l - 1; w.maxL = l;
        w.flags = WIN_ACTIVE | (claims[q] ? WIN_CLAIMS : 0) | (G.uright ? WIN_GATE : 0);
    }
    win[q] = w;
}

// SearchByProjection(CurrentFrame, LastFrame): ORBmatcher.cc:1357-1396
struct FrameProj { float T[12]; float fx, fy, cx, cy, mbf, th; int forward, backward; };
__device__ __forceinline__ void win_frame_one(const DevGrid& G, const FrameProj& P, int q, const u8* __restrict__ has_point,
                                              const float* __restrict__ world, const int* __restrict__ octave,
                                              const u8* __restrict__ claims, Win* __restrict__ win) {
    Win w = {0.f, 0.f, 0.f, 0.f, 0.f, -1, -1, 0};
    if (has_point[q]) {
        const float X = world[3 * q], Y = world[3 * q + 1], Z = world[3 * q + 2];
        float c[3];
#pragma unroll
        for (int r = 0; r < 3; r++)
            c[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P.T[4 * r], X), __fmul_rn(P.T[4 * r + 1], Y)), __fmul_rn(P.T[4 * r + 2], Z)), P.T[4 * r + 3]);
        const float invzc = (float)__ddiv_rn(1.0, (double)c[2]);
        bool ok = !(invzc < 0);
        const float u = __fadd_rn(__fmul_rn(__fmul_rn(P.fx, c[0]), invzc), P.cx);
        const float v = __fadd_rn(__fmul_rn(__fmul_rn(P.fy, c[1]), invzc), P.cy);
        if (u < G.min_x || u > G.max_x) ok = false;
        if (v < G.min_y || v > G.max_y) ok = false;
        if (ok) {
            const int o = octave[q];
            w.u = u; w.v = v;
            w.r = __fmul_rn(P.th, G.sf[o]);
            w.ur = __fsub_rn(u, __fmul_rn(P.mbf, invzc)); w.tol = w.r;
            if (P.forward) { w.minL = o; w.maxL = -1; }
            else if (P.backward) { w.minL = 0; w.maxL = o; }
            else { w.minL = o - 1; w.maxL = o + 1; }
            w.flags = WIN_ACTIVE | (claims[q] ? WIN_CLAIMS : 0) | (G.uright ? WIN_GATE : 0);
        }
    }
    win[q] = w;
}
__global__ void k_win_frame(DevGrid G, FrameProj P, int nq, const u8* __restrict__ has_point, const float* __restrict__ world,
                            const int* __restrict__ octave, const u8* __restrict__ claims, Win* __restrict__ win) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) win_frame_one(G, P, q, has_point, world, octave, claims, win);
}

// SearchForInitialization: ORBmatcher.cc:422-430 (level-0 features only, window = windowSize around vbPrevMatched)
__global__ void k_win_init(int nq, const int* __restrict__ octave1, const float* __restrict__ prev_xy, float window, Win* __restrict__ win) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    Win w = {0.f, 0.f, 0.f, 0.f, 0.f, -1, -1, 0};
    const int l = octave1[q];
    if (!(l > 0)) {
        w.u = prev_xy[2 * q]; w.v = prev_xy[2 * q + 1]; w.r = window;
        w.minL = l; w.maxL = l;
        w.flags = WIN_ACTIVE | WIN_CLAIMS;
    }
    win[q] = w;
}

// Explicit windows (the shared core of the relocalisation / loop-closing projection searches): every accepted match claims.
__global__ void k_win_explicit(int nq, const u8* __restrict__ active, const float* __restrict__ u, const float* __restrict__ v,
                               const float* __restrict__ r, const int* __restrict__ minL, const int* __restrict__ maxL,
                               const float* __restrict__ ur, int flags, Win* __restrict__ win) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    Win w = {0.f, 0.f, 0.f, 0.f, 0.f, -1, -1, 0};
    if (active[q]) {
        w.u = u[q]; w.v = v[q]; w.r = r[q]; w.minL = minL[q]; w.maxL = maxL[q]; w.flags = WIN_ACTIVE | flags;
        if (ur) w.ur = ur[q];
    }
    win[q] = w;
}

// ---- 2. candidates: Frame::GetFeaturesInArea (src/Frame.cc:445-498) + the static `continue`s of the candidate loops ------------
__device__ __forceinline__ bool cand_pass(const DevGrid& G, const Win& w, bool checkLevels, int idx) {
    if (G.blocked && G.blocked[idx]) return false;
    if (checkLevels) {
        const int o = G.octave[idx];
        if (o < w.minL) return false;
        if (w.maxL >= 0 && o > w.maxL) return false;
    }
    const float dx = __fsub_rn(G.x[idx], w.u), dy = __fsub_rn(G.y[idx], w.v);
    if (!(fabsf(dx) < w.r && fabsf(dy) < w.r)) return false;
    if (w.flags & WIN_GATE) {
        const float ur = G.uright[idx];
        if (ur > 0 && fabsf(__fsub_rn(w.ur, ur)) > w.tol) return false;
    }
    if (w.flags & WIN_CHI2) {                                           // chi-square gate on the reprojection error, 2 or 3 dof
        const float ex = __fsub_rn(w.u, G.x[idx]), ey = __fsub_rn(w.v, G.y[idx]);
        const float kpr = G.uright ? G.uright[idx] : -1.0f;
        const float is2 = G.inv_sigma2[G.octave[idx]];
        if (kpr >= 0) {
            const float er = __fsub_rn(w.ur, kpr);
            const float e2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(er, er));
            if ((double)__fmul_rn(e2, is2) > 7.8) return false;
        } else {
            const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
            if ((double)__fmul_rn(e2, is2) > 5.99) return false;
        }
    }
    return true;
}

#define WC_WARPS 4
__device__ __forceinline__ void win_candidates_one(const DevGrid& G, const Win* __restrict__ win, const u8* __restrict__ qdesc, int q,
                                                   int lane, int stride, int drop_above, u32* __restrict__ list, int* __restrict__ cnt) {
    const Win w = win[q];
    int total = 0;
    if (w.flags & WIN_ACTIVE) {
        const int minCX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(w.u, G.min_x), w.r), G.inv_w)));
        const int maxCX = min(G.cols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(w.u, G.min_x), w.r), G.inv_w)));
        const int minCY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(w.v, G.min_y), w.r), G.inv_h)));
        const int maxCY = min(G.rows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(w.v, G.min_y), w.r), G.inv_h)));
        if (minCX < G.cols && maxCX >= 0 && minCY < G.rows && maxCY >= 0 && maxCX >= minCX && maxCY >= minCY) {
            const bool checkLevels = (w.minL > 0) || (w.maxL >= 0);
            u32 qd[8];
            load_desc(qdesc + (size_t)q * 32, qd);
            const int ncy = maxCY - minCY + 1, T = (maxCX - minCX + 1) * ncy;
            u32* out = list + (size_t)q * stride;
            for (int t0 = 0; t0 < T; t0 += 32) {
                const int t = t0 + lane;
                int b = 0, e = 0;
                if (t < T) {
                    const int c = (minCX + t / ncy) * G.rows + (minCY + t % ncy);
                    b = G.cell_off[c]; e = G.cell_off[c + 1];
                }
                // distances of this lane's cell; entries above drop_above can never be accepted (best-only modes)
                int n = 0;
                for (int j = b; j < e; j++) {
                    const int idx = G.cell_feat[j];
                    if (!cand_pass(G, w, checkLevels, idx)) continue;
                    u32 d2[8];
                    load_desc(G.desc + (size_t)idx * 32, d2);
                    if (ham256(qd, d2) <= drop_above) n++;
                }
                int inc = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
                int p = total + inc - n;
                for (int j = b; j < e && n > 0; j++) {
                    const int idx = G.cell_feat[j];
                    if (!cand_pass(G, w, checkLevels, idx)) continue;
                    u32 d2[8];
                    load_desc(G.desc + (size_t)idx * 32, d2);
                    const int d = ham256(qd, d2);
                    if (d <= drop_above) out[p++] = ((u32)d << WL_SHIFT) | (u32)idx;
                }
                total += __shfl_sync(0xffffffffu, inc, 31);
            }
        }
    }
    if (lane == 0) cnt[q] = total;
}
__global__ void __launch_bounds__(32 * WC_WARPS) k_win_candidates(DevGrid G, const Win* __restrict__ win, const u8* __restrict__ qdesc,
                                                                  int nq, int stride, int drop_above, u32* __restrict__ list,
                                                                  int* __restrict__ cnt) {
    const int lane = threadIdx.x & 31, q = blockIdx.x * WC_WARPS + (threadIdx.x >> 5);
    if (q < nq) win_candidates_one(G, win, qdesc, q, lane, stride, drop_above, list, cnt);
}

// ---- 3. resolution ------------------------------------------------------------------------------------------------------------
// State per searched feature c:
//   claimq[c] = the first query that claimed c (INT_MAX = nobody).  MODE_BEST / MODE_TOP2_LEVEL: c is available to query q iff
//     claimq[c] > q, i.e. the test is relative to q's place in the serial order, so a claim finalised early by a LATER query
//     never hides c from an earlier one (queries whose point has no observations take a feature without claiming it,
//     ORBmatcher.cc:87-89).
//   MODE_INIT: thr[c] = vMatchedDistance[c] (:448, 471); an entry (d, c) is available iff d < thr[c].  Every claim carries a
//     distance <= TH_LOW, so for an entry with d > TH_LOW that is the same as "nobody before q has claimed c" = claimq[c] > q
//     (relative again), and for d <= TH_LOW the absolute thr[c] is exact because a later claim on c waits (minq) for every
//     earlier unresolved query that holds c with such a distance.
// minq[c] = lowest unresolved claiming query that holds c, available, with a distance it could be accepted with (d <= th_dist).
// g lanes per query (g = 1..32 by mean list length), lanes over the candidate list.
struct ResolveParams { int nq, nt, stride, mode, th_dist, checkOri; float nnratio; };
#define WR_THREADS 1024
#define WR_WARPS (WR_THREADS / 32)

__device__ __forceinline__ bool entry_avail(bool init, int th_dist, int d, int idx, int q, const int* thr, const int* claimq) {
    return (init && d <= th_dist) ? d < thr[idx] : claimq[idx] > q;
}

__device__ __forceinline__ void win_resolve_cta(const ResolveParams& P, const u32* __restrict__ list, const int* __restrict__ cnt,
                                                const Win* __restrict__ win, const int* __restrict__ t_octave,
                                                const float* __restrict__ q_angle, const float* __restrict__ t_angle,
                                                int* __restrict__ thr, int* __restrict__ minq, int* __restrict__ claimq,
                                                int* __restrict__ state, int* __restrict__ dec, int* __restrict__ qbin,
                                                int* __restrict__ owner, int* __restrict__ match, int* __restrict__ out_cnt) {
    __shared__ int s_unres, s_hist[ORBM_HISTO_LENGTH], s_nm, s_ind[3], s_rounds, s_entries, s_active;
    const int tid = threadIdx.x;
    const bool init = P.mode == MODE_INIT, needSecond = P.mode != MODE_BEST;
    if (tid < ORBM_HISTO_LENGTH) s_hist[tid] = 0;
    if (tid == 0) { s_nm = 0; s_rounds = 0; s_entries = 0; s_active = 0; }
    __syncthreads();
    for (int c = tid; c < P.nt; c += WR_THREADS) { thr[c] = 257; claimq[c] = INT_MAX; owner[c] = -1; }
    {
        int e = 0, a = 0;
        for (int q = tid; q < P.nq; q += WR_THREADS) {
            match[q] = -1; dec[q] = -1; state[q] = cnt[q] == 0 ? 2 : 0;
            e += cnt[q]; a += cnt[q] != 0;
        }
        if (a) { atomicAdd(&s_entries, e); atomicAdd(&s_active, a); }
    }
    __syncthreads();
    // lanes per query: a power of two near (mean list length / 4), so short lists keep many queries in flight
    int g = 1;
    while (g < 32 && (long)g * 4 * max(s_active, 1) < (long)s_entries) g <<= 1;
    const int sub = tid & (g - 1), qpw = 32 / g, q_first = (tid >> 5) * qpw + ((tid & 31) / g), q_step = WR_WARPS * qpw;
    for (;;) {
        if (tid == 0) s_unres = 0;
        for (int c = tid; c < P.nt; c += WR_THREADS) minq[c] = INT_MAX;
        __syncthreads();
        for (int q = q_first; q < P.nq; q += q_step) {
            if (state[q] != 0 || !(win[q].flags & WIN_CLAIMS)) continue;
            const u32* l = list + (size_t)q * P.stride;
            const int n = cnt[q];
            for (int k = sub; k < n; k += g) {
                const u32 e = l[k];
                const int d = (int)(e >> WL_SHIFT), idx = (int)(e & WL_IDX_MASK);
                if (d <= P.th_dist && entry_avail(init, P.th_dist, d, idx, q, thr, claimq)) atomicMin(&minq[idx], q);
            }
        }
        __syncthreads();
        for (int q0 = (tid >> 5) * qpw; q0 < P.nq; q0 += q_step) {      // warp-uniform trip count: shuffles below use the full mask
            const int q = q0 + ((tid & 31) / g);
            const bool live = q < P.nq && state[q] == 0;
            const u32* l = list + (size_t)(live ? q : 0) * P.stride;
            const int n = live ? cnt[q] : 0;
            u32 best = KEY_NONE, sec = KEY_NONE;                        // (distance << 22 | position): first in list order wins ties
            for (int k = sub; k < n; k += g) {
                const u32 e = l[k];
                const int d = (int)(e >> WL_SHIFT), idx = (int)(e & WL_IDX_MASK);
                if (!entry_avail(init, P.th_dist, d, idx, q, thr, claimq)) continue;
                const u32 key = ((u32)d << WL_SHIFT) | (u32)k;
                sec = min(sec, max(best, key));
                best = min(best, key);
            }
            for (int o = g >> 1; o > 0; o >>= 1) {
                const u32 ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, sec, o);
                sec = min(min(sec, os), max(best, ob));
                best = min(best, ob);
            }
            if (!live || sub != 0) continue;
            const int bidx = best == KEY_NONE ? -1 : (int)(l[best & WL_IDX_MASK] & WL_IDX_MASK);
            const int sidx = sec == KEY_NONE ? -1 : (int)(l[sec & WL_IDX_MASK] & WL_IDX_MASK);
            const bool fin = (bidx < 0 || minq[bidx] >= q) && (!needSecond || sidx < 0 || minq[sidx] >= q);
            if (!fin) { s_unres = 1; continue; }
            const int bd = best == KEY_NONE ? 256 : (int)(best >> WL_SHIFT);
            bool accept = bidx >= 0 && bd <= P.th_dist;
            if (accept && P.mode == MODE_TOP2_LEVEL) {                  // ORBmatcher.cc:117-121
                const int sd = sec == KEY_NONE ? 256 : (int)(sec >> WL_SHIFT);
                const int bl = t_octave[bidx], sl = sidx >= 0 ? t_octave[sidx] : -1;
                if (bl == sl && (float)bd > __fmul_rn(P.nnratio, (float)sd)) accept = false;
            } else if (accept && init) {                                // ORBmatcher.cc:461-463
                const float sd = sec == KEY_NONE ? (float)INT_MAX : (float)(int)(sec >> WL_SHIFT);
                if (!((float)bd < __fmul_rn(sd, P.nnratio))) accept = false;
            }
            dec[q] = accept ? ((bd << WL_SHIFT) | bidx) : -1;
            state[q] = 1;
        }
        __syncthreads();
        for (int q = tid; q < P.nq; q += WR_THREADS) {
            if (state[q] != 1) continue;
            state[q] = 2;
            if (dec[q] < 0) continue;
            const int bidx = dec[q] & WL_IDX_MASK, bd = dec[q] >> WL_SHIFT;
            if (init) {                                                 // ORBmatcher.cc:465-473 (at most one claim per feature per round)
                const int prev = owner[bidx];
                if (prev >= 0) { match[prev] = -1; atomicSub(&s_nm, 1); }
                owner[bidx] = q;
                thr[bidx] = bd;
                atomicMin(&claimq[bidx], q);
            } else {
                atomicMax(&owner[bidx], q);
                if (win[q].flags & WIN_CLAIMS) atomicMin(&claimq[bidx], q);
            }
            match[q] = bidx;
            atomicAdd(&s_nm, 1);
            if (P.checkOri) {
                const int bin = rot_bin(q_angle[q], t_angle[bidx]);
                qbin[q] = bin;
                atomicAdd(&s_hist[bin], 1);
            }
        }
        __syncthreads();
        if (tid == 0) s_rounds++;
        if (!s_unres) break;
        __syncthreads();
    }
    if (P.checkOri) {                                                   // rotation consistency (:1446-1458, :493-512)
        if (tid == 0) { int a, b, c; three_maxima_dev(s_hist, ORBM_HISTO_LENGTH, a, b, c); s_ind[0] = a; s_ind[1] = b; s_ind[2] = c; }
        __syncthreads();
        for (int q = tid; q < P.nq; q += WR_THREADS) {
            if (init ? match[q] < 0 : dec[q] < 0) continue;
            const int bin = qbin[q];
            if (bin == s_ind[0] || bin == s_ind[1] || bin == s_ind[2]) continue;
            if (init) match[q] = -1;
            else owner[dec[q] & WL_IDX_MASK] = -2;
            atomicSub(&s_nm, 1);
        }
        __syncthreads();
    }
    if (tid == 0) { out_cnt[0] = s_nm; out_cnt[1] = s_rounds; }
}
__global__ void __launch_bounds__(WR_THREADS) k_win_resolve(ResolveParams P, const u32* __restrict__ list, const int* __restrict__ cnt,
                                                            const Win* __restrict__ win, const int* __restrict__ t_octave,
                                                            const float* __restrict__ q_angle, const float* __restrict__ t_angle,
                                                            int* __restrict__ thr, int* __restrict__ minq, int* __restrict__ claimq,
                                                            int* __restrict__ state, int* __restrict__ dec, int* __restrict__ qbin,
                                                            int* __restrict__ owner, int* __restrict__ match, int* __restrict__ out_cnt) {
    win_resolve_cta(P, list, cnt, win, t_octave, q_angle, t_angle, thr, minq, claimq, state, dec, qbin, owner, match, out_cnt);
}

// ---- many-frame form: one job = one (CurrentFrame, LastFrame) pair; blockIdx.y (blockIdx.x in the resolve) selects the job ----
struct WinWork {
    Win* win; int* owner; int* match; int* out_cnt;                    // (owner, match, out_cnt) stay inside the mirrored part
    u32* list; int* cnt; int* thr; int* minq; int* claimq; int* state; int* dec; int* qbin;
};
struct FrameJobDev {
    DevGrid G; FrameProj P; int nq, stride;
    const u8* has_point; const float* world; const int* octave; const u8* claims; const u8* desc; const float* angle;
    WinWork w;
};
__global__ void k_win_frame_batch(const FrameJobDev* __restrict__ jobs) {
    const FrameJobDev& J = jobs[blockIdx.y];
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < J.nq) win_frame_one(J.G, J.P, q, J.has_point, J.world, J.octave, J.claims, J.w.win);
}
__global__ void __launch_bounds__(32 * WC_WARPS) k_win_candidates_batch(const FrameJobDev* __restrict__ jobs, int drop_above) {
    const FrameJobDev& J = jobs[blockIdx.y];
    const int lane = threadIdx.x & 31, q = blockIdx.x * WC_WARPS + (threadIdx.x >> 5);
    if (q < J.nq) win_candidates_one(J.G, J.w.win, J.desc, q, lane, J.stride, drop_above, J.w.list, J.w.cnt);
}
__global__ void __launch_bounds__(WR_THREADS) k_win_resolve_batch(const FrameJobDev* __restrict__ jobs, int mode, int th_dist, int checkOri,
                                                                  float nnratio) {
    const FrameJobDev& J = jobs[blockIdx.x];
    ResolveParams P;
    P.nq = J.nq; P.nt = J.G.n; P.stride = J.stride; P.mode = mode; P.th_dist = th_dist; P.checkOri = checkOri; P.nnratio = nnratio;
    win_resolve_cta(P, J.w.list, J.w.cnt, J.w.win, J.G.octave, J.angle, J.G.angle, J.w.thr, J.w.minq, J.w.claimq, J.w.state, J.w.dec, J.w.qbin,
                    J.w.owner, J.w.match, J.w.out_cnt);
}

// =====================================================================================================
// Host side
// =====================================================================================================
static int check_grid(const orbm_grid_view* g, bool needAngle) {
    ORB_REQUIRE(g && g->n >= 0, ORB_ERR_ARG, "bad grid view");
    ORB_REQUIRE(g->n <= (int)WL_IDX_MASK, ORB_ERR_ARG, "more than 4194303 features");
    ORB_REQUIRE(g->n == 0 || (g->desc && g->x && g->y && g->octave), ORB_ERR_ARG, "grid view: desc/x/y/octave are required");
    ORB_REQUIRE(!needAngle || g->n == 0 || g->angle, ORB_ERR_ARG, "grid view: angle is required for the orientation check");
    ORB_REQUIRE(g->grid_cols > 0 && g->grid_rows > 0 && g->cell_offsets && (g->cell_features || g->n == 0), ORB_ERR_ARG, "grid view: bad grid");
    const int nc = g->grid_cols * g->grid_rows;
    ORB_REQUIRE(g->cell_offsets[0] == 0, ORB_ERR_ARG, "grid view: cell_offsets[0] != 0");
    for (int c = 0; c < nc; c++) ORB_REQUIRE(g->cell_offsets[c] <= g->cell_offsets[c + 1], ORB_ERR_ARG, "grid view: offsets not monotone");
    for (int i = 0; i < g->cell_offsets[nc]; i++)
        ORB_REQUIRE(g->cell_features[i] >= 0 && g->cell_features[i] < g->n, ORB_ERR_ARG, "grid view: feature index out of range");
    ORB_REQUIRE(g->n_levels > 0 && g->scale_factors, ORB_ERR_ARG, "grid view: scale_factors are required");
    for (int i = 0; i < g->n; i++) ORB_REQUIRE(g->octave[i] >= 0 && g->octave[i] < g->n_levels, ORB_ERR_ARG, "grid view: octave out of range");
    return ORB_OK;
}
static size_t grid_bytes(const orbm_grid_view* g) {
    const size_t n = (size_t)g->n, nc = (size_t)g->grid_cols * g->grid_rows;
    return pad(n * 32) + 5 * pad(n * 4) + pad(n) + pad((nc + 1) * 4) + pad((size_t)g->cell_offsets[nc] * 4) + pad((size_t)g->n_levels * 4);
}
static int upload_grid(Arena& A, const orbm_grid_view* g, DevGrid* d) {
    int rc;
    const size_t n = (size_t)g->n, nc = (size_t)g->grid_cols * g->grid_rows;
    d->n = g->n; d->cols = g->grid_cols; d->rows = g->grid_rows;
    d->min_x = g->min_x; d->min_y = g->min_y; d->max_x = g->max_x; d->max_y = g->max_y; d->inv_w = g->inv_w; d->inv_h = g->inv_h;
    d->n_levels = g->n_levels;
    d->angle = nullptr; d->uright = nullptr; d->blocked = nullptr; d->inv_sigma2 = nullptr;
    if ((rc = upload(A, g->desc, n * 32, &d->desc))) return rc;
    if ((rc = upload(A, g->x, n, &d->x))) return rc;
    if ((rc = upload(A, g->y, n, &d->y))) return rc;
    if ((rc = upload(A, g->octave, n, &d->octave))) return rc;
    if (g->angle) { if ((rc = upload(A, g->angle, n, &d->angle))) return rc; }
    if (g->uright) { if ((rc = upload(A, g->uright, n, &d->uright))) return rc; }
    if (g->blocked) { if ((rc = upload(A, g->blocked, n, &d->blocked))) return rc; }
    if ((rc = upload(A, g->cell_offsets, nc + 1, &d->cell_off))) return rc;
    if ((rc = upload(A, g->cell_features, (size_t)g->cell_offsets[nc], &d->cell_feat))) return rc;
    if ((rc = upload(A, g->scale_factors, (size_t)g->n_levels, &d->sf))) return rc;
    return ORB_OK;
}

// outputs + scratch of one search, taken from the arena after the inputs have been flushed
static size_t work_small_bytes(int nq, int nt) { return pad((size_t)nt * 4) + pad((size_t)nq * 4) + pad(8); }
static size_t work_scratch_bytes(int nq, int nt) {
    return pad((size_t)nq * sizeof(Win)) + pad((size_t)nq * (size_t)std::max(nt, 1) * 4) + 4 * pad((size_t)nq * 4) + 3 * pad((size_t)nt * 4);
}
static void take_work_small(Arena& A, int nq, int nt, WinWork* w) {
    w->owner = A.take<int>(std::max(nt, 1));
    w->match = A.take<int>(std::max(nq, 1));
    w->out_cnt = A.take<int>(2);
}
static void take_work_scratch(Arena& A, int nq, int nt, WinWork* w) {
    w->win = A.take<Win>(std::max(nq, 1));
    w->list = A.take<u32>((size_t)std::max(nq, 1) * std::max(nt, 1));
    w->cnt = A.take<int>(std::max(nq, 1));
    w->state = A.take<int>(std::max(nq, 1));
    w->dec = A.take<int>(std::max(nq, 1));
    w->qbin = A.take<int>(std::max(nq, 1));
    w->thr = A.take<int>(std::max(nt, 1));
    w->minq = A.take<int>(std::max(nt, 1));
    w->claimq = A.take<int>(std::max(nt, 1));
}
static void take_work(Arena& A, int nq, int nt, WinWork* w) {
    take_work_small(A, nq, nt, w);
    take_work_scratch(A, nq, nt, w);
}
static int run_search(Arena& A, const DevGrid& G, const WinWork& w, const u8* d_qdesc, const float* d_qangle, int nq, int mode,
                      int th_dist, float nnratio, int checkOri) {
    const int stride = std::max(G.n, 1);
    const int drop_above = mode == MODE_BEST ? th_dist : 256;
    if (nq > 0) k_win_candidates<<<orb_div_up(nq, WC_WARPS), 32 * WC_WARPS, 0, A.stream>>>(G, w.win, d_qdesc, nq, stride, drop_above, w.list, w.cnt);
    ResolveParams P;
    P.nq = nq; P.nt = G.n; P.stride = stride; P.mode = mode; P.th_dist = th_dist; P.checkOri = checkOri; P.nnratio = nnratio;
    k_win_resolve<<<1, WR_THREADS, 0, A.stream>>>(P, w.list, w.cnt, w.win, G.octave, d_qangle, G.angle, w.thr, w.minq, w.claimq, w.state, w.dec,
                                                  w.qbin, w.owner, w.match, w.out_cnt);
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

extern "C" int orbm_search_by_projection_map(const orbm_grid_view* frame, int n_points, const uint8_t* in_view, const float* proj_x,
                                             const float* proj_y, const float* proj_xr, const int* level, const float* view_cos,
                                             const uint8_t* desc, const uint8_t* claims, float th, float nnratio, int* owner,
                                             int* n_matches, int device) {
    ORB_REQUIRE(owner && n_matches && n_points >= 0, ORB_ERR_ARG, "bad arguments");
    ORB_REQUIRE(n_points == 0 || (in_view && proj_x && proj_y && proj_xr && level && view_cos && desc && claims), ORB_ERR_ARG, "null map point array");
    int rc = check_device(device);
    if (rc) return rc;
    if ((rc = check_grid(frame, false))) return rc;
    for (int i = 0; i < n_points; i++)
        ORB_REQUIRE(!in_view[i] || (level[i] >= 0 && level[i] < frame->n_levels), ORB_ERR_ARG, "predicted level out of range");
    const int nq = n_points, nt = frame->n;
    Arena& A = g_arena;
    if ((rc = A.ensure(device, grid_bytes(frame) + 2 * pad(nq) + 4 * pad((size_t)nq * 4) + pad((size_t)nq * 4) + pad((size_t)nq * 32) +
                                   work_small_bytes(nq, nt), work_scratch_bytes(nq, nt)))) return rc;
    DevGrid G;
    if ((rc = upload_grid(A, frame, &G))) return rc;
    const u8 *d_inview, *d_claims, *d_desc;
    const float *d_px, *d_py, *d_pxr, *d_vc;
    const int* d_level;
    if ((rc = upload(A, in_view, (size_t)nq, &d_inview))) return rc;
    if ((rc = upload(A, claims, (size_t)nq, &d_claims))) return rc;
    if ((rc = upload(A, proj_x, (size_t)nq, &d_px))) return rc;
    if ((rc = upload(A, proj_y, (size_t)nq, &d_py))) return rc;
    if ((rc = upload(A, proj_xr, (size_t)nq, &d_pxr))) return rc;
    if ((rc = upload(A, view_cos, (size_t)nq, &d_vc))) return rc;
    if ((rc = upload(A, level, (size_t)nq, &d_level))) return rc;
    if ((rc = upload(A, desc, (size_t)nq * 32, &d_desc))) return rc;
    if ((rc = A.flush())) return rc;
    WinWork w;
    take_work(A, nq, nt, &w);
    if (nq > 0) k_win_map<<<orb_div_up(nq, 256), 256, 0, A.stream>>>(G, nq, d_inview, d_px, d_py, d_pxr, d_level, d_vc, d_claims, th, w.win);
    if ((rc = run_search(A, G, w, d_desc, nullptr, nq, MODE_TOP2_LEVEL, ORBM_TH_HIGH, nnratio, 0))) return rc;
    int cnt[2] = {0, 0};
    if ((rc = A.fetch(owner, w.owner, (size_t)nt))) return rc;
    if ((rc = A.fetch(cnt, w.out_cnt, 2))) return rc;
    if ((rc = A.finish())) return rc;
    *n_matches = cnt[0];
    return ORB_OK;
}

extern "C" int orbm_search_by_projection_frame(const orbm_grid_view* cur, const float* Tcw_cur, const float* Tcw_last, float fx, float fy,
                                               float cx, float cy, float mbf, float mb, int n_last, const uint8_t* has_point,
                                               const float* world, const int* octave, const float* angle, const uint8_t* desc,
                                               const uint8_t* claims, float th, int mono, int check_orientation, int* owner,
                                               int* n_matches, int device) {
    ORB_REQUIRE(owner && n_matches && n_last >= 0 && Tcw_cur && Tcw_last, ORB_ERR_ARG, "bad arguments");
    ORB_REQUIRE(n_last == 0 || (has_point && world && octave && desc && claims && (angle || !check_orientation)), ORB_ERR_ARG, "null LastFrame array");
    int rc = check_device(device);
    if (rc) return rc;
    if ((rc = check_grid(cur, check_orientation != 0))) return rc;
    for (int i = 0; i < n_last; i++)
        ORB_REQUIRE(!has_point[i] || (octave[i] >= 0 && octave[i] < cur->n_levels), ORB_ERR_ARG, "LastFrame octave out of range");
    // twc = -Rcw^T tcw ; tlc = Rlw twc + tlw (src/ORBmatcher.cc:1344-1355); A*x+b: float32, left to right (cv::gemm on 3x3 floats)
    FrameProj P;
    for (int i = 0; i < 12; i++) P.T[i] = Tcw_cur[i];
    float twc[3];
    for (int r = 0; r < 3; r++)           // transposed product: cv::gemm's general path accumulates in double
        twc[r] = (float)(-(((double)Tcw_cur[r] * Tcw_cur[3] + (double)Tcw_cur[4 + r] * Tcw_cur[7]) + (double)Tcw_cur[8 + r] * Tcw_cur[11]));
    const float tlcz = ((Tcw_last[8] * twc[0] + Tcw_last[9] * twc[1]) + Tcw_last[10] * twc[2]) + Tcw_last[11];
    P.forward = (tlcz > mb && !mono) ? 1 : 0;
    P.backward = (-tlcz > mb && !mono) ? 1 : 0;
    P.fx = fx; P.fy = fy; P.cx = cx; P.cy = cy; P.mbf = mbf; P.th = th;
    const int nq = n_last, nt = cur->n;
    Arena& A = g_arena;
    if ((rc = A.ensure(device, grid_bytes(cur) + 2 * pad(nq) + pad((size_t)nq * 12) + 2 * pad((size_t)nq * 4) + pad((size_t)nq * 32) +
                                   work_small_bytes(nq, nt), work_scratch_bytes(nq, nt)))) return rc;
    DevGrid G;
    if ((rc = upload_grid(A, cur, &G))) return rc;
    const u8 *d_has, *d_claims, *d_desc;
    const float *d_world, *d_angle = nullptr;
    const int* d_oct;
    if ((rc = upload(A, has_point, (size_t)nq, &d_has))) return rc;
    if ((rc = upload(A, claims, (size_t)nq, &d_claims))) return rc;
    if ((rc = upload(A, world, (size_t)nq * 3, &d_world))) return rc;
    if ((rc = upload(A, octave, (size_t)nq, &d_oct))) return rc;
    if (check_orientation) { if ((rc = upload(A, angle, (size_t)nq, &d_angle))) return rc; }
    if ((rc = upload(A, desc, (size_t)nq * 32, &d_desc))) return rc;
    if ((rc = A.flush())) return rc;
    WinWork w;
    take_work(A, nq, nt, &w);
    if (nq > 0) k_win_frame<<<orb_div_up(nq, 256), 256, 0, A.stream>>>(G, P, nq, d_has, d_world, d_oct, d_claims, w.win);
    if ((rc = run_search(A, G, w, d_desc, d_angle, nq, MODE_BEST, ORBM_TH_HIGH, 0.f, check_orientation))) return rc;
    int cnt[2] = {0, 0};
    if ((rc = A.fetch(owner, w.owner, (size_t)nt))) return rc;
    if ((rc = A.fetch(cnt, w.out_cnt, 2))) return rc;
    if ((rc = A.finish())) return rc;
    *n_matches = cnt[0];
    return ORB_OK;
}

// Many-frame form of the per-frame tracking search: n_jobs independent (CurrentFrame, LastFrame) pairs — several cameras or sessions
// tracked by one process — in ONE call: all inputs in one upload, three launches (windows, candidates: grid.y = job; resolve: one CTA
// per job, so the serial claim replay of every frame runs on its own SM), one download.  Same device code as the single call.
static void frame_proj(const orbm_frame_search_job& j, FrameProj* P) {
    // twc = -Rcw^T tcw ; tlc = Rlw twc + tlw (src/ORBmatcher.cc:1344-1355); A*x+b: float32, left to right (cv::gemm on 3x3 floats)
    const float *Tc = j.Tcw_cur, *Tl = j.Tcw_last;
    for (int i = 0; i < 12; i++) P->T[i] = Tc[i];
    float twc[3];
    for (int r = 0; r < 3; r++)           // transposed product: cv::gemm's general path accumulates in double
        twc[r] = (float)(-(((double)Tc[r] * Tc[3] + (double)Tc[4 + r] * Tc[7]) + (double)Tc[8 + r] * Tc[11]));
    const float tlcz = ((Tl[8] * twc[0] + Tl[9] * twc[1]) + Tl[10] * twc[2]) + Tl[11];
    P->forward = (tlcz > j.mb && !j.mono) ? 1 : 0;
    P->backward = (-tlcz > j.mb && !j.mono) ? 1 : 0;
    P->fx = j.fx; P->fy = j.fy; P->cx = j.cx; P->cy = j.cy; P->mbf = j.mbf; P->th = j.th;
}

extern "C" int orbm_search_by_projection_frame_batch(orbm_frame_search_job* jobs, int n_jobs, int check_orientation, int device) {
    ORB_REQUIRE(n_jobs >= 0 && (jobs || n_jobs == 0), ORB_ERR_ARG, "bad arguments");
    int rc = check_device(device);
    if (rc) return rc;
    if (n_jobs == 0) return ORB_OK;
    size_t bytes = pad((size_t)n_jobs * sizeof(FrameJobDev)), scratch = 0;
    int maxNq = 0;
    for (int k = 0; k < n_jobs; k++) {
        const orbm_frame_search_job& j = jobs[k];
        ORB_REQUIRE(j.cur && j.owner && j.n_last >= 0 && j.Tcw_cur && j.Tcw_last, ORB_ERR_ARG, "job %d: bad arguments", k);
        ORB_REQUIRE(j.n_last == 0 || (j.has_point && j.world && j.octave && j.desc && j.claims && (j.angle || !check_orientation)), ORB_ERR_ARG,
                    "job %d: null LastFrame array", k);
        if ((rc = check_grid(j.cur, check_orientation != 0))) return rc;
        for (int i = 0; i < j.n_last; i++)
            ORB_REQUIRE(!j.has_point[i] || (j.octave[i] >= 0 && j.octave[i] < j.cur->n_levels), ORB_ERR_ARG, "job %d: LastFrame octave out of range", k);
        const size_t nq = (size_t)j.n_last;
        bytes += grid_bytes(j.cur) + 2 * pad(nq) + pad(nq * 12) + 2 * pad(nq * 4) + pad(nq * 32) + work_small_bytes(j.n_last, j.cur->n);
        scratch += work_scratch_bytes(j.n_last, j.cur->n);
        maxNq = std::max(maxNq, j.n_last);
    }
    Arena& A = g_arena;
    if ((rc = A.ensure(device, bytes, scratch, 64u << 20))) return rc;
    std::vector<FrameJobDev> J(n_jobs);
    for (int k = 0; k < n_jobs; k++) {
        const orbm_frame_search_job& j = jobs[k];
        FrameJobDev& D = J[k];
        const size_t nq = (size_t)j.n_last;
        frame_proj(j, &D.P);
        D.nq = j.n_last; D.stride = std::max(j.cur->n, 1);
        D.angle = nullptr;
        if ((rc = upload_grid(A, j.cur, &D.G))) return rc;
        if ((rc = upload(A, j.has_point, nq, &D.has_point))) return rc;
        if ((rc = upload(A, j.claims, nq, &D.claims))) return rc;
        if ((rc = upload(A, j.world, nq * 3, &D.world))) return rc;
        if ((rc = upload(A, j.octave, nq, &D.octave))) return rc;
        if (check_orientation) { if ((rc = upload(A, j.angle, nq, &D.angle))) return rc; }
        if ((rc = upload(A, j.desc, nq * 32, &D.desc))) return rc;
    }
    // the job table follows the inputs; the outputs (mirrored) and the scratch come after it, so their addresses are known now
    FrameJobDev* d_jobs = A.take<FrameJobDev>(n_jobs);
    const size_t inEnd = A.used;
    for (int k = 0; k < n_jobs; k++) take_work_small(A, jobs[k].n_last, jobs[k].cur->n, &J[k].w);
    for (int k = 0; k < n_jobs; k++) take_work_scratch(A, jobs[k].n_last, jobs[k].cur->n, &J[k].w);
    if (A.staged) memcpy(A.hbase + (reinterpret_cast<u8*>(d_jobs) - A.base), J.data(), (size_t)n_jobs * sizeof(FrameJobDev));
    else ORB_CUDA_TRY(cudaMemcpyAsync(d_jobs, J.data(), (size_t)n_jobs * sizeof(FrameJobDev), cudaMemcpyHostToDevice, A.stream));
    if ((rc = A.flush(inEnd))) return rc;
    if (maxNq > 0) {
        k_win_frame_batch<<<dim3(orb_div_up(maxNq, 256), n_jobs), 256, 0, A.stream>>>(d_jobs);
        k_win_candidates_batch<<<dim3(orb_div_up(maxNq, WC_WARPS), n_jobs), 32 * WC_WARPS, 0, A.stream>>>(d_jobs, ORBM_TH_HIGH);
    }
    k_win_resolve_batch<<<n_jobs, WR_THREADS, 0, A.stream>>>(d_jobs, MODE_BEST, ORBM_TH_HIGH, check_orientation, 0.f);
    ORB_CUDA_TRY(cudaGetLastError());
    std::vector<int> cnt(2 * (size_t)n_jobs, 0);
    for (int k = 0; k < n_jobs; k++) {
        if ((rc = A.fetch(jobs[k].owner, J[k].w.owner, (size_t)jobs[k].cur->n))) return rc;
        if ((rc = A.fetch(&cnt[2 * k], J[k].w.out_cnt, 2))) return rc;
    }
    if ((rc = A.finish())) return rc;
    for (int k = 0; k < n_jobs; k++) jobs[k].n_matches = cnt[2 * k];
    return ORB_OK;
}

extern "C" int orbm_search_for_initialization(const orbm_grid_view* f2, int n1, const uint8_t* desc1, const int* octave1,
                                              const float* angle1, float* prev_xy, int window_size, float nnratio,
                                              int check_orientation, int* matches12, int* n_matches, int device) {
    ORB_REQUIRE(matches12 && n_matches && n1 >= 0, ORB_ERR_ARG, "bad arguments");
    ORB_REQUIRE(n1 == 0 || (desc1 && octave1 && prev_xy && (angle1 || !check_orientation)), ORB_ERR_ARG, "null F1 array");
    int rc = check_device(device);
    if (rc) return rc;
    if ((rc = check_grid(f2, check_orientation != 0))) return rc;
    const int nq = n1, nt = f2->n;
    Arena& A = g_arena;
    if ((rc = A.ensure(device, grid_bytes(f2) + pad((size_t)nq * 8) + 2 * pad((size_t)nq * 4) + pad((size_t)nq * 32) + work_small_bytes(nq, nt),
                       work_scratch_bytes(nq, nt)))) return rc;
    DevGrid G;
    if ((rc = upload_grid(A, f2, &G))) return rc;
    G.uright = nullptr; G.blocked = nullptr;                            // no stereo gate / no pre-claimed features in this search
    const u8* d_desc;
    const float *d_prev, *d_angle = nullptr;
    const int* d_oct;
    if ((rc = upload(A, octave1, (size_t)nq, &d_oct))) return rc;
    if ((rc = upload(A, (const float*)prev_xy, (size_t)nq * 2, &d_prev))) return rc;
    if (check_orientation) { if ((rc = upload(A, angle1, (size_t)nq, &d_angle))) return rc; }
    if ((rc = upload(A, desc1, (size_t)nq * 32, &d_desc))) return rc;
    if ((rc = A.flush())) return rc;
    WinWork w;
    take_work(A, nq, nt, &w);
    if (nq > 0) k_win_init<<<orb_div_up(nq, 256), 256, 0, A.stream>>>(nq, d_oct, d_prev, (float)window_size, w.win);
    if ((rc = run_search(A, G, w, d_desc, d_angle, nq, MODE_INIT, ORBM_TH_LOW, nnratio, check_orientation))) return rc;
    int cnt[2] = {0, 0};
    if ((rc = A.fetch(matches12, w.match, (size_t)nq))) return rc;
    if ((rc = A.fetch(cnt, w.out_cnt, 2))) return rc;
    if ((rc = A.finish())) return rc;
    for (int i = 0; i < nq; i++)                                        // vbPrevMatched update, src/ORBmatcher.cc:516-519
        if (matches12[i] >= 0) { prev_xy[2 * i] = f2->x[matches12[i]]; prev_xy[2 * i + 1] = f2->y[matches12[i]]; }
    *n_matches = cnt[0];
    return ORB_OK;
}

extern "C" int orbm_search_windows(const orbm_grid_view* target, int nq, const uint8_t* active, const float* u, const float* v,
                                   const float* r, const int* min_level, const int* max_level, const uint8_t* desc, const float* angle,
                                   int th_dist, int check_orientation, int* owner, int* n_matches, int device) {
    ORB_REQUIRE(owner && n_matches && nq >= 0 && th_dist >= 0 && th_dist <= 256, ORB_ERR_ARG, "bad arguments");
    ORB_REQUIRE(nq == 0 || (active && u && v && r && min_level && max_level && desc && (angle || !check_orientation)), ORB_ERR_ARG, "null query array");
    int rc = check_device(device);
    if (rc) return rc;
    if ((rc = check_grid(target, check_orientation != 0))) return rc;
    const int nt = target->n;
    Arena& A = g_arena;
    if ((rc = A.ensure(device, grid_bytes(target) + pad(nq) + 6 * pad((size_t)nq * 4) + pad((size_t)nq * 32) + work_small_bytes(nq, nt),
                       work_scratch_bytes(nq, nt)))) return rc;
    DevGrid G;
    if ((rc = upload_grid(A, target, &G))) return rc;
    G.uright = nullptr;                                                 // no stereo gate in these searches
    const u8 *d_active, *d_desc;
    const float *d_u, *d_v, *d_r, *d_angle = nullptr;
    const int *d_minL, *d_maxL;
    if ((rc = upload(A, active, (size_t)nq, &d_active))) return rc;
    if ((rc = upload(A, u, (size_t)nq, &d_u))) return rc;
    if ((rc = upload(A, v, (size_t)nq, &d_v))) return rc;
    if ((rc = upload(A, r, (size_t)nq, &d_r))) return rc;
    if ((rc = upload(A, min_level, (size_t)nq, &d_minL))) return rc;
    if ((rc = upload(A, max_level, (size_t)nq, &d_maxL))) return rc;
    if (check_orientation) { if ((rc = upload(A, angle, (size_t)nq, &d_angle))) return rc; }
    if ((rc = upload(A, desc, (size_t)nq * 32, &d_desc))) return rc;
    if ((rc = A.flush())) return rc;
    WinWork w;
    take_work(A, nq, nt, &w);
    if (nq > 0) k_win_explicit<<<orb_div_up(nq, 256), 256, 0, A.stream>>>(nq, d_active, d_u, d_v, d_r, d_minL, d_maxL, nullptr, WIN_CLAIMS, w.win);
    if ((rc = run_search(A, G, w, d_desc, d_angle, nq, MODE_BEST, th_dist, 0.f, check_orientation))) return rc;
    int cnt[2] = {0, 0};
    if ((rc = A.fetch(owner, w.owner, (size_t)nt))) return rc;
    if ((rc = A.fetch(cnt, w.out_cnt, 2))) return rc;
    if ((rc = A.finish())) return rc;
    *n_matches = cnt[0];
    return ORB_OK;
}

extern "C" int orbm_search_windows_best(const orbm_grid_view* target, int nq, const uint8_t* active, const float* u, const float* v,
                                        const float* r, const int* min_level, const int* max_level, const uint8_t* desc, const float* ur,
                                        const float* inv_level_sigma2, int th_dist, int* best_idx, int device) {
    ORB_REQUIRE(best_idx && nq >= 0 && th_dist >= 0 && th_dist <= 256, ORB_ERR_ARG, "bad arguments");
    ORB_REQUIRE(nq == 0 || (active && u && v && r && min_level && max_level && desc), ORB_ERR_ARG, "null query array");
    ORB_REQUIRE((ur == nullptr) == (inv_level_sigma2 == nullptr), ORB_ERR_ARG, "ur and inv_level_sigma2 go together");
    int rc = check_device(device);
    if (rc) return rc;
    if ((rc = check_grid(target, false))) return rc;
    if (nq == 0) return ORB_OK;
    const int nt = target->n;
    Arena& A = g_arena;
    if ((rc = A.ensure(device, grid_bytes(target) + pad(nq) + 7 * pad((size_t)nq * 4) + pad((size_t)nq * 32) + pad((size_t)target->n_levels * 4) +
                                   work_small_bytes(nq, nt), work_scratch_bytes(nq, nt)))) return rc;
    DevGrid G;
    if ((rc = upload_grid(A, target, &G))) return rc;
    G.blocked = nullptr;                                                // nothing is pre-claimed and nothing claims in these searches
    const u8 *d_active, *d_desc;
    const float *d_u, *d_v, *d_r, *d_ur = nullptr;
    const int *d_minL, *d_maxL;
    if ((rc = upload(A, active, (size_t)nq, &d_active))) return rc;
    if ((rc = upload(A, u, (size_t)nq, &d_u))) return rc;
    if ((rc = upload(A, v, (size_t)nq, &d_v))) return rc;
    if ((rc = upload(A, r, (size_t)nq, &d_r))) return rc;
    if ((rc = upload(A, min_level, (size_t)nq, &d_minL))) return rc;
    if ((rc = upload(A, max_level, (size_t)nq, &d_maxL))) return rc;
    if (ur) {
        if ((rc = upload(A, ur, (size_t)nq, &d_ur))) return rc;
        if ((rc = upload(A, inv_level_sigma2, (size_t)target->n_levels, &G.inv_sigma2))) return rc;
    } else {
        G.uright = nullptr;
    }
    if ((rc = upload(A, desc, (size_t)nq * 32, &d_desc))) return rc;
    if ((rc = A.flush())) return rc;
    WinWork w;
    take_work(A, nq, nt, &w);
    k_win_explicit<<<orb_div_up(nq, 256), 256, 0, A.stream>>>(nq, d_active, d_u, d_v, d_r, d_minL, d_maxL, d_ur, ur ? WIN_CHI2 : 0, w.win);
    if ((rc = run_search(A, G, w, d_desc, nullptr, nq, MODE_BEST, th_dist, 0.f, 0))) return rc;
    if ((rc = A.fetch(best_idx, w.match, (size_t)nq))) return rc;
    return A.finish();
}
