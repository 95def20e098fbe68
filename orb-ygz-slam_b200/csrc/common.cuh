// Shared host/device helpers of libygz_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include "../../include/ygz_b200.h"
// NVTX ranges around every stage launch (the reference's only tracing is glog timing lines): header-only NVTX v3,
// no-ops unless a tool (nsys / ncu --nvtx) is attached.
#include <nvtx3/nvToolsExt.h>

namespace ygz_b200 {

void set_last_error(const std::string& s);
#ifdef __CUDACC__
// vocabulary.cu: node ids of an extractor's device-resident descriptors on the caller's stream (see there)
int vocab_node_ids_async(ygz_vocab* v, const void* d_desc, const int* d_counts, int cap, int n_frames, int levelsup,
                         uint32_t* d_node_out, cudaStream_t st);
int vocab_device(const ygz_vocab* v);
#endif

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

// ---- TMA (cp.async.bulk.tensor) plumbing -------------------------------------------------------------------------
// Host: a 3-D u8 tensor map {x: w bytes, y: h rows (stride pitch), z: frames (stride frame_stride)} with a fixed box.
// The driver entry point is fetched through the runtime so the library does not link libcuda.
static inline int make_tensor_map_u8_3d(CUtensorMap* out, const void* base, int w, int h, int frames, size_t pitch,
                                        size_t frame_stride, int box_w, int box_h) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -1;
        fn = (EncodeFn)p;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // make the init visible to the async (TMA) proxy
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, int phase) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    } while (!ok);
}
// true for exactly one lane of a converged warp (elect.sync): the canonical guard of single-thread TMA / mbarrier issue
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    return pred != 0;
}
// one lane: box (x, y, z) of a 3-D tensor map -> shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"((unsigned long long)map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}
#endif

}  // namespace ygz_b200

namespace ygz_b200 {
struct NvtxRange {          // RAII: one named range per stage on the calling (launching) thread
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace ygz_b200
