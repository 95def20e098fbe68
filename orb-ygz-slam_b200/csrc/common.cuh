// Shared host/device helpers of libygz_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include "../../include/ygz_b200.h"

namespace ygz_b200 {

void set_last_error(const std::string& s);

#define YGZ_CUDA_TRY(expr)                                                                  \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ::ygz_b200::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); \
            return YGZ_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define YGZ_ARG_CHECK(cond)                                               \
    do {                                                                  \
        if (!(cond)) {                                                    \
            ::ygz_b200::set_last_error(std::string("bad argument: ") + #cond); \
            return YGZ_ERR_ARG;                                           \
        }                                                                 \
    } while (0)

static inline int round_up(int v, int a) { return (v + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cvRound on the host (cvtss2si, round-half-even under the default rounding mode)
static inline int host_cv_round(float v) { return (int)lrintf(v); }
static inline int host_cv_round(double v) { return (int)lrint(v); }
static inline int host_cv_floor(float v) { int i = (int)v; return i - (v < (float)i); }

#ifdef __CUDACC__
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ int reflect101(int p, int len) {
    // BORDER_REFLECT_101, |overshoot| < len for every use in this library
    if (p < 0) p = -p;
    if (p >= len) p = 2 * len - 2 - p;
    return p;
}
// streaming 16-byte load that does not pollute L1 (read-once image data)
__device__ __forceinline__ uint4 ld_nc_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
#endif

}  // namespace ygz_b200
