// common.cuh — shared helpers for the prb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/prb200.h"

namespace prb {

extern thread_local char g_err[512];

inline int fail(const char *what, const char *detail)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", what, detail);
    return 1;
}

#define PRB_CUDA(expr)                                                     \
    do {                                                                   \
        cudaError_t e__ = (expr);                                          \
        if (e__ != cudaSuccess) return prb::fail(#expr, cudaGetErrorString(e__)); \
    } while (0)

#define PRB_LAUNCH_CHECK(name)                                             \
    do {                                                                   \
        cudaError_t e__ = cudaGetLastError();                              \
        if (e__ != cudaSuccess) return prb::fail(name, cudaGetErrorString(e__)); \
    } while (0)

#define PRB_REQUIRE(cond, msg)                                             \
    do {                                                                   \
        if (!(cond)) return prb::fail(__func__, msg);                      \
    } while (0)

inline cudaStream_t S(prb_stream_t s) { return (cudaStream_t)s; }

struct DevInfo {
    int sm_count;
    int smem_optin;
};
const DevInfo &devinfo();

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t hmin(int64_t a, int64_t b) { return a < b ? a : b; }
inline int64_t hmax(int64_t a, int64_t b) { return a > b ? a : b; }

constexpr uint32_t ROW_MASK = 0x00FFFFFFu;
// vstream.cu / vell.cu: shift code of a cell whose selected-gene code is 0.  Counter register r
// adds 1 << (code - 32 r) with PTX's clamping shift: any amount >= 32 (also a wrapped negative
// one) contributes 0, and 255 - 32 r >= 32 for the up to five registers of K <= 21.
constexpr int PRB_SX_NONE = 255;

// np.argmax ordering: NaN is maximal; first index wins ties.
__device__ __forceinline__ bool better(double sa, int64_t ia, double sb, int64_t ib)
{
    const bool na = sa != sa, nb = sb != sb;
    if (na || nb) {
        if (na && nb) return ia < ib;
        return na;
    }
    if (sa > sb) return true;
    if (sa < sb) return false;
    return ia < ib;
}

}  // namespace prb
