// ref_cli.cpp — TEST INFRASTRUCTURE.  Runs the reference's own ORBextractor (unmodified /root/reference/src/ORBextractor.cc over
// cvshim.hpp) in a process whose heap is MONOTONIC: operator new hands out ever-increasing addresses and nothing is reused.
// Why: DistributeOctTree breaks ties between equal-size nodes by comparing ExtractorNode POINTERS (src/ORBextractor.cc:683), so
// the stock binary's output depends on the allocator's address reuse and changes from call to call.  With a monotonic heap
// "larger pointer" == "created later", which is exactly the canonical rule of the oracle / GPU path (creation sequence number
// in place of the pointer); under it the reference is deterministic and can be compared bit for bit.
//
// usage: ref_extract_cli W H nfeatures scale nlevels ini min image.raw mask.raw|- out.bin
// out.bin: int32 n, then n x 7 float32 (x, y, size, angle, response, octave, class_id), then n x 32 bytes.
#include <sys/mman.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include "ORBextractor.h"

static char* g_arena = nullptr;
static size_t g_used = 0;
static const size_t kArena = (size_t)24 << 30;      // address space only (MAP_NORESERVE); pages are touched on demand

static void* bump(size_t n) {
    if (!g_arena) {
        g_arena = (char*)mmap(nullptr, kArena, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (g_arena == (char*)MAP_FAILED) { fprintf(stderr, "arena mmap failed\n"); abort(); }
    }
    n = (n + 15) & ~(size_t)15;
    if (g_used + n > kArena) { fprintf(stderr, "arena exhausted\n"); abort(); }
    void* p = g_arena + g_used;
    g_used += n;
    return p;
}
void* operator new(size_t n) { return bump(n); }
void* operator new[](size_t n) { return bump(n); }
void operator delete(void*) noexcept {}
void operator delete[](void*) noexcept {}
void operator delete(void*, size_t) noexcept {}
void operator delete[](void*, size_t) noexcept {}

int main(int argc, char** argv) {
    if (argc < 11) return 2;
    const int W = atoi(argv[1]), H = atoi(argv[2]), nf = atoi(argv[3]), nl = atoi(argv[5]), ini = atoi(argv[6]), mn = atoi(argv[7]);
    const float sf = (float)atof(argv[4]);
    std::vector<unsigned char> img((size_t)W * H), mask;
    FILE* f = fopen(argv[8], "rb");
    if (!f || fread(img.data(), 1, img.size(), f) != img.size()) return 3;
    fclose(f);
    if (strcmp(argv[9], "-")) {
        mask.resize((size_t)W * H);
        f = fopen(argv[9], "rb");
        if (!f || fread(mask.data(), 1, mask.size(), f) != mask.size()) return 4;
        fclose(f);
    }
    ORB_SLAM2::ORBextractor ex(nf, sf, nl, ini, mn);
    cv::Mat image(H, W, CV_8UC1, img.data(), (size_t)W), m, desc;
    if (!mask.empty()) m = cv::Mat(H, W, CV_8UC1, mask.data(), (size_t)W);
    std::vector<cv::KeyPoint> kps;
    ex(image, m, kps, desc);
    f = fopen(argv[10], "wb");
    if (!f) return 5;
    const int n = (int)kps.size();
    fwrite(&n, 4, 1, f);
    for (int i = 0; i < n; i++) {
        const float o[7] = {kps[i].pt.x, kps[i].pt.y, kps[i].size, kps[i].angle, kps[i].response, (float)kps[i].octave, (float)kps[i].class_id};
        fwrite(o, 4, 7, f);
    }
    for (int i = 0; i < n; i++) fwrite(desc.ptr(i), 1, 32, f);
    fclose(f);
    return 0;
}
