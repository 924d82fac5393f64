// orb_match_common.cuh — pieces shared by the matcher translation units (orb_match.cu, orb_project.cu): the 256-bit Hamming
// distance on 8 x u32, packed top-2 keys, the rotation-histogram helpers and the per-thread device scratch arena.
#pragma once
#include "orb_common.cuh"

#include <algorithm>
#include <vector>

#define KEY_SHIFT 23
#define KEY_IDX_MASK 0x7FFFFFu
#define KEY_NONE 0xFFFFFFFFu             // distance field 511 -> decoded as "no candidate" (256, -1)

__device__ __forceinline__ int ham256(const u32 (&a)[8], const u32 (&b)[8]) {
    return __popc(a[0] ^ b[0]) + __popc(a[1] ^ b[1]) + __popc(a[2] ^ b[2]) + __popc(a[3] ^ b[3]) +
           __popc(a[4] ^ b[4]) + __popc(a[5] ^ b[5]) + __popc(a[6] ^ b[6]) + __popc(a[7] ^ b[7]);
}
__device__ __forceinline__ void load_desc(const u8* p, u32 (&d)[8]) {
    const uint4 lo = __ldg(reinterpret_cast<const uint4*>(p)), hi = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    d[0] = lo.x; d[1] = lo.y; d[2] = lo.z; d[3] = lo.w; d[4] = hi.x; d[5] = hi.y; d[6] = hi.z; d[7] = hi.w;
}
__device__ __forceinline__ int key_dist(u32 k) { return k == KEY_NONE ? 256 : (int)(k >> KEY_SHIFT); }

__device__ __forceinline__ int rot_bin(float a1, float a2) {                   // ORBmatcher.cc:241-246
    const float factor = 1.0f / ORBM_HISTO_LENGTH;
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, factor));
    if (bin == ORBM_HISTO_LENGTH) bin = 0;
    return bin;
}
// ComputeThreeMaxima (ORBmatcher.cc:1604-1645) on bin sizes
__device__ inline void three_maxima_dev(const int* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = histo[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
}

// =====================================================================================================
// Host side
// =====================================================================================================
// Per-thread device scratch arena (grows; process-lifetime cache) with a pinned host mirror.  Small calls (a SearchByBoW of
// two 2000-feature keyframes is ~150 KB in ten arrays) are packed into the mirror and sent with ONE asynchronous copy, and
// their results come back through the mirror with one synchronisation; large calls copy directly.
struct Arena {
    int device = -1;
    u8* base = nullptr;
    u8* hbase = nullptr;
    size_t cap = 0, hcap = 0, used = 0, inEnd = 0;
    bool staged = false;
    cudaStream_t stream = nullptr;
    // `bytes`: inputs + outputs (taken first; mirrored in pinned host memory when small); `scratch`: device-only work space
    // taken after them (candidate lists etc.), never mirrored.
    ~Arena() {                                               // thread exit: give the device / pinned memory and the stream back
        if (device < 0) return;
        if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return; }   // (the runtime may already be gone at process exit)
        if (base) cudaFree(base);
        if (hbase) cudaFreeHost(hbase);
        if (stream) cudaStreamDestroy(stream);
        cudaGetLastError();
    }
    // stageMax: calls whose inputs + outputs fit are packed into the pinned mirror (one copy each way); the batched searches raise it
    int ensure(int dev, size_t bytes, size_t scratch = 0, size_t stageMax = 2u << 20) {
        if (device != dev) {
            if (base) { cudaSetDevice(device); cudaFree(base); base = nullptr; cap = 0; }
            if (hbase) { cudaFreeHost(hbase); hbase = nullptr; hcap = 0; }
            if (stream) { cudaStreamDestroy(stream); stream = nullptr; }
            device = dev;
        }
        ORB_CUDA_TRY(cudaSetDevice(dev));
        if (!stream) ORB_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        bytes += 64 * 256;                                   // per-array alignment slack
        scratch += 16 * 256;
        if (bytes + scratch > cap) {
            if (base) ORB_CUDA_TRY(cudaFree(base));
            base = nullptr;
            cap = orb_align_up(bytes + scratch + ((bytes + scratch) >> 2), 1 << 20);
            ORB_CUDA_TRY(cudaMalloc(&base, cap));
        }
        const size_t hmax = std::max<size_t>(8u << 20, stageMax);
        if (bytes > hcap && bytes <= hmax) {
            if (hbase) ORB_CUDA_TRY(cudaFreeHost(hbase));
            hbase = nullptr;
            hcap = std::min<size_t>(orb_align_up(bytes + (bytes >> 2), 1 << 20), hmax);
            ORB_CUDA_TRY(cudaMallocHost(&hbase, hcap));
        }
        staged = hbase != nullptr && bytes <= hcap && bytes <= stageMax;
        used = 0; inEnd = 0;
        pend.clear();                                        // a call that failed between fetch() and finish() leaves nothing behind
        return ORB_OK;
    }
    template <typename T>
    T* take(size_t count) {
        T* p = reinterpret_cast<T*>(base + used);
        used += orb_align_up(count * sizeof(T), 256);
        return p;
    }
    // send everything put() so far (staged mode); call once, after the last input and before the launch
    int flush(size_t upto = ~(size_t)0) {                    // upto: end of the inputs when outputs have already been taken behind them
        const size_t n = std::min(upto, used);
        if (staged && n > 0) ORB_CUDA_TRY(cudaMemcpyAsync(base, hbase, n, cudaMemcpyHostToDevice, stream));
        inEnd = n;
        return ORB_OK;
    }
    // bring `count` elements at device pointer d back to host pointer h; finish() completes the transfer
    struct Pending { void* h; const u8* d; size_t bytes; };
    std::vector<Pending> pend;
    template <typename T>
    int fetch(T* h, const T* d, size_t count) {
        if (count == 0) return ORB_OK;
        const u8* dp = reinterpret_cast<const u8*>(d);
        if (staged) {
            ORB_CUDA_TRY(cudaMemcpyAsync(hbase + (dp - base), dp, count * sizeof(T), cudaMemcpyDeviceToHost, stream));
            pend.push_back({h, dp, count * sizeof(T)});
        } else {
            ORB_CUDA_TRY(cudaMemcpyAsync(h, dp, count * sizeof(T), cudaMemcpyDeviceToHost, stream));
        }
        return ORB_OK;
    }
    int finish() {
        ORB_CUDA_TRY(cudaStreamSynchronize(stream));
        for (const Pending& p : pend) memcpy(p.h, hbase + (p.d - base), p.bytes);
        pend.clear();
        return ORB_OK;
    }
};
inline thread_local Arena g_arena;

template <typename T>
inline int upload(Arena& A, const T* host, size_t count, const T** dev) {
    T* d = A.take<T>(std::max<size_t>(count, 1));
    if (count) {
        if (A.staged) memcpy(A.hbase + (reinterpret_cast<u8*>(d) - A.base), host, count * sizeof(T));
        else ORB_CUDA_TRY(cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, A.stream));
    }
    *dev = d;
    return ORB_OK;
}
inline size_t pad(size_t bytes) { return orb_align_up(std::max<size_t>(bytes, 1), 256); }

inline int check_device(int device) {
    ORB_REQUIRE(device >= 0 && device < orb_device_count(), ORB_ERR_CUDA, "CUDA device %d not available (no CPU fallback)", device);
    return ORB_OK;
}

