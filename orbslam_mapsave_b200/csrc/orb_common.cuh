// orb_common.cuh — shared helpers for the B200 ORB kernels (error plumbing, small device utilities).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <cstring>

#include "../../include/orb_b200.h"

typedef unsigned char u8;
typedef unsigned short u16;
typedef unsigned int u32;
typedef unsigned long long u64;

void orb_set_error(const char* fmt, ...);

#define ORB_CUDA_TRY(expr)                                                                         \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            orb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return ORB_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

#define ORB_REQUIRE(cond, code, ...)          \
    do {                                      \
        if (!(cond)) {                        \
            orb_set_error(__VA_ARGS__);       \
            return (code);                    \
        }                                     \
    } while (0)

// device status bits (written with atomicOr by kernels, read back by the host)
enum { ORB_DEV_CAND_OVERFLOW = 1, ORB_DEV_OUT_OVERFLOW = 2, ORB_DEV_NODE_OVERFLOW = 4 };

static inline int orb_div_up(int a, int b) { return (a + b - 1) / b; }
static inline size_t orb_align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

__device__ __forceinline__ int dev_reflect101(int p, int n) {
    // n >= 2 guaranteed by the geometry checks (levels are >= 62 px)
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
    return p;
}
