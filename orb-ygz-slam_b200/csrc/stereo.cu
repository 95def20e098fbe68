// Frame::ComputeStereoMatches (/root/reference/src/Frame.cc:509-682), SURVEY.md §8(f) rank 3, on the device:
// one WARP per left key point.  Lanes stride over the right key points listed for the key's image row (the reference's
// vRowIndices band table, built on the host in the reference's insertion order), Hamming distance of the 256-bit
// descriptors, first-minimum selection by (distance, list position); then the 11 x 11 SAD of the centre-subtracted
// patches for the 11 column offsets on the key's pyramid level (levels are read in place from the two extractor
// handles, zero copy), parabola refinement and the disparity gates on lane 0 with explicit round-to-nearest float ops.
// The median-based outlier cut (:668-681) is a sort of <= N pairs and stays on the host.
#include "common.cuh"
#include "hamming_bf.cuh"
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace {
using namespace ygz_b200;

struct StereoLevels {
    const uint8_t* L[16];
    const uint8_t* R[16];
    int pitchL[16], pitchR[16], w[16], h[16];
    float scale[16], inv_scale[16];
};

// Sub-pixel refinement of one (left key, best right key) association, src/Frame.cc:590-664: 11 x 11 SAD of the
// centre-subtracted patches on the key's pyramid level for the 11 column offsets, parabola fit, disparity gates.  One warp;
// imL / imR are the level images of the key's frame.  Outputs are meaningful on every lane.
__device__ __forceinline__ void stereo_refine(const StereoLevels& lv, const uint8_t* imL, const uint8_t* imR, const ygz_keypoint& kpL,
                                              float uR0, int levelL, float maxD, float mbf, int lane, float& out_u, float& out_d,
                                              int& out_sad) {
    const float uL = kpL.x;
    const float sf = lv.inv_scale[levelL];
    const float scaleduL = roundf(__fmul_rn(kpL.x, sf)), scaledvL = roundf(__fmul_rn(kpL.y, sf));
    const float scaleduR0 = roundf(__fmul_rn(uR0, sf));
    const int r0 = (int)(scaledvL - 5.f), cL0 = (int)(scaleduL - 5.f);
    const float iniu = scaleduR0 + 5.f - 5.f, endu = scaleduR0 + 5.f + 5.f + 1.f;
    // The 11x11 left patch must lie inside the level: extractor keys always do (>= 19 px from the border), caller-supplied
    // keys near a border would make the reference throw from rowRange / colRange (src/Frame.cc:604-606) — skipped here.
    const bool patch_inside = r0 >= 0 && r0 + 10 < lv.h[levelL] && cL0 >= 0 && cL0 + 10 < lv.w[levelL];
    if (patch_inside && !(iniu < 0.f || endu >= (float)lv.w[levelL])) {
        const int pL = lv.pitchL[levelL], pR = lv.pitchR[levelL];
        const int ctrL = imL[(size_t)(r0 + 5) * pL + cL0 + 5];
        // left patch values of this lane's pixels (121 px, 4 per lane), centre-subtracted
        int lpx[4], ly[4], lx[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = lane + 32 * k;
            ly[k] = i / 11; lx[k] = i - ly[k] * 11;
            lpx[k] = i < 121 ? (int)imL[(size_t)(r0 + ly[k]) * pL + cL0 + lx[k]] - ctrL : 0;
        }
        int bestS = INT_MAX, bestinc = 0;
        float vd[11];
        for (int inc = -5; inc <= 5; inc++) {
            const int cr = (int)(scaleduR0 + (float)inc - 5.f);
            const int ctrR = imR[(size_t)(r0 + 5) * pR + cr + 5];
            int s = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (lane + 32 * k < 121) s += abs(lpx[k] - ((int)imR[(size_t)(r0 + ly[k]) * pR + cr + lx[k]] - ctrR));
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float dist = (float)s;                   // cv::norm(..., NORM_L1): integer sum, exact
            if (dist < (float)bestS) { bestS = s; bestinc = inc; }
            vd[inc + 5] = dist;
        }
        if (!(bestinc == -5 || bestinc == 5)) {
            const float d1 = vd[5 + bestinc - 1], d2 = vd[5 + bestinc], d3 = vd[5 + bestinc + 1];
            const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
            if (!(deltaR < -1.f || deltaR > 1.f)) {          // NaN (flat patch: 0/0) passes this test in the reference too
                float bestuR = __fmul_rn(lv.scale[levelL], __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
                float disparity = __fsub_rn(uL, bestuR);
                if (disparity >= 0.f && disparity < maxD) {
                    if (disparity <= 0.f) { disparity = 0.01f; bestuR = __double2float_rn(__dsub_rn((double)uL, 0.01)); }
                    out_d = __fdiv_rn(mbf, disparity);
                    out_u = bestuR;
                    out_sad = bestS;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_stereo_match(StereoLevels lv, const ygz_keypoint* __restrict__ keysL,
                                                      const uint4* __restrict__ descL, int N,
                                                      const ygz_keypoint* __restrict__ keysR, const uint4* __restrict__ descR,
                                                      const int* __restrict__ row_off, const int* __restrict__ row_idx, int nRows,
                                                      float maxD, float mbf, float* __restrict__ uRight, float* __restrict__ depth,
                                                      int* __restrict__ sad_out) {
    const int iL = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (iL >= N) return;
    const ygz_keypoint kpL = keysL[iL];
    float out_u = -1.f, out_d = -1.f;
    int out_sad = -1;
    const float vL = kpL.y, uL = kpL.x;
    const int levelL = kpL.octave;
    int best = 100, bestpos = INT_MAX;                 // ORBmatcher::TH_HIGH
    const float minU = __fsub_rn(uL, maxD), maxU = uL; // minD = 0
    bool searched = false;
    int c0 = 0, c1 = 0;
    if (vL >= 0.f && (int)vL < nRows && !(maxU < 0.f)) {
        c0 = row_off[(int)vL]; c1 = row_off[(int)vL + 1];
        searched = c1 > c0;
    }
    if (searched) {
        const uint4 a0 = descL[2 * iL], a1 = descL[2 * iL + 1];
        for (int p = c0 + lane; p < c1; p += 32) {
            const int iR = row_idx[p];
            const ygz_keypoint kpR = keysR[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            if (!(kpR.x >= minU && kpR.x <= maxU)) continue;
            const uint4 b0 = descR[2 * iR], b1 = descR[2 * iR + 1];
            const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                          __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
            if (d < best) { best = d; bestpos = p; }   // positions ascend per lane: strict < keeps the first minimum
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, best, o), op = __shfl_xor_sync(0xffffffffu, bestpos, o);
            if (ob < best || (ob == best && op < bestpos)) { best = ob; bestpos = op; }
        }
    }
    if (searched && best < 75 && bestpos != INT_MAX) {   // thOrbDist = (TH_HIGH + TH_LOW) / 2
        const int iR = row_idx[bestpos];
        stereo_refine(lv, lv.L[levelL], lv.R[levelL], kpL, keysR[iR].x, levelL, maxD, mbf, lane, out_u, out_d, out_sad);
    }
    if (lane == 0) { uRight[iL] = out_u; depth[iL] = out_d; sad_out[iL] = out_sad; }
}


// ------------------------------------------------------------------------------------------------ batched, device resident
// The same association for a BATCH of stereo pairs whose key points / descriptors / pyramids are the two extractors' own device
// outputs (frame f of the left handle with frame f of the right handle): no host row table — every warp scans all right keys of
// its frame in index order against the row band [floor(y - r), ceil(y + r)], r = 2 * scale[octave] (:524-536), which visits
// the candidates of vRowIndices[vL] in the reference's insertion order, so the first-minimum selection is unchanged.
struct StereoBatch {
    StereoLevels lv;                    // level base pointers of frame 0
    size_t fsL[16], fsR[16];            // frame strides per level
    const ygz_keypoint *kL, *kR;        // [frames][cap]
    const uint4 *dL, *dR;
    const int *cntL, *cntR;
    int capL, capR, nRows;
    float maxD, mbf;
    float *u, *z;                       // [frames][capL]
    int* sad;
};
struct RightKey { float x; short minr, maxr; int octave; };
constexpr int SB_KEYS_PER_WARP = 4;

__global__ void __launch_bounds__(256) k_stereo_match_batch(const __grid_constant__ StereoBatch A) {
    extern __shared__ RightKey rk[];
    const int f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nL = min(A.cntL[f], A.capL), nR = min(A.cntR[f], A.capR);
    const int first = blockIdx.x * 8 * SB_KEYS_PER_WARP;
    const size_t oL = (size_t)f * A.capL, oR = (size_t)f * A.capR;
    if (first >= nL) {                   // slots past the frame's count: "no match"
        for (int i = first + threadIdx.x; i < min(first + 8 * SB_KEYS_PER_WARP, A.capL); i += blockDim.x) {
            A.u[oL + i] = -1.f; A.z[oL + i] = -1.f; A.sad[oL + i] = -1;
        }
        return;
    }
    for (int i = threadIdx.x; i < nR; i += blockDim.x) {
        const ygz_keypoint k = A.kR[oR + i];
        const float r = __fmul_rn(2.0f, A.lv.scale[k.octave]);
        RightKey q;
        q.x = k.x;
        q.maxr = (short)min((int)ceilf(__fadd_rn(k.y, r)), A.nRows - 1);
        q.minr = (short)max((int)floorf(__fsub_rn(k.y, r)), 0);
        q.octave = k.octave;
        rk[i] = q;
    }
    __syncthreads();
    for (int kk = 0; kk < SB_KEYS_PER_WARP; kk++) {
        const int iL = first + warp * SB_KEYS_PER_WARP + kk;
        if (iL >= A.capL) break;
        float out_u = -1.f, out_d = -1.f;
        int out_sad = -1;
        if (iL < nL) {
            const ygz_keypoint kpL = A.kL[oL + iL];
            const float vL = kpL.y, uL = kpL.x;
            const int levelL = kpL.octave;
            const float minU = __fsub_rn(uL, A.maxD), maxU = uL;
            if (vL >= 0.f && (int)vL < A.nRows && !(maxU < 0.f)) {
                const int row = (int)vL;
                int best = 100, bestpos = INT_MAX;
                const uint4 a0 = A.dL[(oL + iL) * 2], a1 = A.dL[(oL + iL) * 2 + 1];
                for (int p = lane; p < nR; p += 32) {
                    const RightKey q = rk[p];
                    if (row < q.minr || row > q.maxr) continue;
                    if (q.octave < levelL - 1 || q.octave > levelL + 1) continue;
                    if (!(q.x >= minU && q.x <= maxU)) continue;
                    const uint4 b0 = A.dR[(oR + p) * 2], b1 = A.dR[(oR + p) * 2 + 1];
                    const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                                  __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
                    if (d < best) { best = d; bestpos = p; }
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const int ob = __shfl_xor_sync(0xffffffffu, best, o), op = __shfl_xor_sync(0xffffffffu, bestpos, o);
                    if (ob < best || (ob == best && op < bestpos)) { best = ob; bestpos = op; }
                }
                if (best < 75 && bestpos != INT_MAX)
                    stereo_refine(A.lv, A.lv.L[levelL] + A.fsL[levelL] * f, A.lv.R[levelL] + A.fsR[levelL] * f, kpL, rk[bestpos].x, levelL,
                                  A.maxD, A.mbf, lane, out_u, out_d, out_sad);
            }
        }
        if (lane == 0) { A.u[oL + iL] = out_u; A.z[oL + iL] = out_d; A.sad[oL + iL] = out_sad; }
    }
}

// The median-based outlier cut of :668-681 per frame: sort the SAD values of the matched keys, thDist = 1.5f * 1.4f * median,
// drop every association with dist >= thDist.  One CTA per frame, bitonic sort in shared memory (cap <= 4096).
constexpr int SM_SORT = 4096;
__global__ void __launch_bounds__(1024) k_stereo_median_cut(float* __restrict__ u, float* __restrict__ z, const int* __restrict__ sad,
                                                            const int* __restrict__ cntL, int capL, int* __restrict__ n_out) {
    __shared__ int v[SM_SORT];
    __shared__ int m_s, removed_s;
    const int f = blockIdx.x, n = min(cntL[f], capL);
    const size_t o = (size_t)f * capL;
    if (threadIdx.x == 0) { m_s = 0; removed_s = 0; }
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < SM_SORT; i += blockDim.x) {
        const int sv = i < n ? sad[o + i] : -1;
        v[i] = sv >= 0 ? sv : INT_MAX;
        mine += sv >= 0;
    }
    if (mine) atomicAdd(&m_s, mine);
    for (int k = 2; k <= SM_SORT; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < SM_SORT; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const int a = v[i], b = v[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { v[i] = b; v[ixj] = a; }
                }
            }
        }
    __syncthreads();
    const int m = m_s;
    if (m == 0) { if (threadIdx.x == 0) n_out[f] = 0; return; }
    const float median = (float)v[m / 2];
    const float thDist = __fmul_rn(__fmul_rn(1.5f, 1.4f), median);
    int removed = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int sv = sad[o + i];
        if (sv >= 0 && !((float)sv < thDist)) { u[o + i] = -1.f; z[o + i] = -1.f; removed++; }
    }
    if (removed) atomicAdd(&removed_s, removed);
    __syncthreads();
    if (threadIdx.x == 0) n_out[f] = m - removed_s;
}

// C5's "brute-force matcher": dense best / second-best right descriptor for every left descriptor of every pair
__global__ void __launch_bounds__(BF_WARPS * 32) k_hamming_bf_top2_batch(const uint32_t* __restrict__ q, const int* __restrict__ cntq, int capq,
                                                                         const uint32_t* __restrict__ t, const int* __restrict__ cntt, int capt,
                                                                         int* __restrict__ best_idx, int* __restrict__ best_d,
                                                                         int* __restrict__ second_d) {
    const int f = blockIdx.y;
    const int nq = min(cntq[f], capq), nt = min(cntt[f], capt);
    if ((int)blockIdx.x * BF_WARPS * BF_QPW >= nq) return;
    const size_t oq = (size_t)f * capq;
    bf_top2_body(q + oq * 8, nq, t + (size_t)f * capt * 8, nt, best_idx + oq, best_d + oq, second_d + oq, blockIdx.x);
}

}  // namespace

struct ygz_stereo {
    int device = 0;
    cudaStream_t stream = nullptr;
    int cap_kL = 0, cap_dL = 0, cap_u = 0, cap_z = 0, cap_sad = 0, cap_kR = 0, cap_dR = 0, cap_rows = 0, cap_idx = 0;
    ygz_keypoint *d_kL = nullptr, *d_kR = nullptr;
    uint8_t *d_dL = nullptr, *d_dR = nullptr;
    int *d_row_off = nullptr, *d_row_idx = nullptr, *d_sad = nullptr;
    float *d_u = nullptr, *d_z = nullptr;
    // batched form: [frames][capL] outputs and per-frame counts
    int cap_bu = 0, cap_bz = 0, cap_bsad = 0, cap_bn = 0, cap_bfi = 0, cap_bfd = 0, cap_bfs = 0;
    float *d_bu = nullptr, *d_bz = nullptr;
    int *d_bsad = nullptr, *d_bn = nullptr, *d_bfi = nullptr, *d_bfd = nullptr, *d_bfs = nullptr;
    cudaEvent_t ev_right = nullptr, ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
};

static std::map<std::pair<int, const void*>, ygz_stereo*> g_stereo;   // scratch (buffers + stream) per device for the per-frame entry point, per
                                                                       // left extractor handle for the batched one (several batches may be in flight)
static std::mutex g_stereo_mutex;          // calls are serialised: the scratch is shared by every caller of a device

template <class T>
static int grow(T** p, int* cap, int need) {
    if (need <= *cap) return YGZ_OK;
    cudaFree(*p); *p = nullptr;
    const int old = *cap;
    *cap = 0;                                   // nothing is held if the allocation below fails
    const int nc = std::max(need, 2 * old);
    YGZ_CUDA_TRY(cudaMalloc(p, (size_t)nc * sizeof(T)));
    *cap = nc;
    return YGZ_OK;
}

namespace {

// per-device scratch of the stereo entry points (created on first use; the caller holds g_stereo_mutex)
int stereo_scratch(int dev, const void* key, ygz_stereo** out) {
    ygz_stereo*& S = g_stereo[std::make_pair(dev, key)];
    if (!S) {
        ygz_stereo* fresh = new ygz_stereo();
        fresh->device = dev;
        if (cudaStreamCreateWithFlags(&fresh->stream, cudaStreamNonBlocking) != cudaSuccess) {     // never keep a scratch without its stream
            delete fresh;
            YGZ_CUDA_TRY(cudaGetLastError());
            return YGZ_ERR_CUDA;
        }
        S = fresh;
    }
    if (!S->ev_right) {
        YGZ_CUDA_TRY(cudaEventCreateWithFlags(&S->ev_right, cudaEventDisableTiming));
        YGZ_CUDA_TRY(cudaEventCreate(&S->ev0)); YGZ_CUDA_TRY(cudaEventCreate(&S->ev1));
        YGZ_CUDA_TRY(cudaEventCreate(&S->ev2)); YGZ_CUDA_TRY(cudaEventCreate(&S->ev3));
    }
    *out = S;
    return YGZ_OK;
}

int device_of(ygz_orb* h, int* dev) {
    void* p = nullptr; int pit; size_t fs;
    ygz_orb_device_level(h, 0, &p, &pit, &fs);
    cudaPointerAttributes at;
    YGZ_CUDA_TRY(cudaPointerGetAttributes(&at, p));
    *dev = at.device;
    return YGZ_OK;
}

// kernel arguments of the batched association from the two handles' current device buffers
int stereo_batch_args(ygz_orb* left, ygz_orb* right, ygz_stereo* S, int n_frames, float mb, float mbf, StereoBatch* A) {
    float scale[16], inv_scale[16], sg[16], isg[16];
    int fpl[16];
    int rc = ygz_orb_tables(left, scale, inv_scale, sg, isg, fpl);
    if (rc) return rc;
    std::memset(A, 0, sizeof(*A));
    int nlevels = 0;
    for (int l = 0; l < 16; l++) {
        int w = 0, hh = 0, w2 = 0, h2 = 0;
        if (ygz_orb_level_size(left, l, &w, &hh) != YGZ_OK) break;
        if (ygz_orb_level_size(right, l, &w2, &h2) != YGZ_OK || w2 != w || h2 != hh) { set_last_error("left / right extractors differ"); return YGZ_ERR_ARG; }
        void *pl = nullptr, *pr = nullptr; int pitl = 0, pitr = 0; size_t fsl = 0, fsr = 0;
        ygz_orb_device_level(left, l, &pl, &pitl, &fsl);
        ygz_orb_device_level(right, l, &pr, &pitr, &fsr);
        A->lv.L[l] = (const uint8_t*)pl; A->lv.R[l] = (const uint8_t*)pr;
        A->fsL[l] = fsl; A->fsR[l] = fsr;
        A->lv.pitchL[l] = pitl; A->lv.pitchR[l] = pitr; A->lv.w[l] = w; A->lv.h[l] = hh;
        A->lv.scale[l] = scale[l]; A->lv.inv_scale[l] = inv_scale[l];
        nlevels = l + 1;
    }
    YGZ_ARG_CHECK(nlevels >= 1);
    void *kL, *dL, *cL, *kR, *dR, *cR;
    ygz_orb_device_outputs(left, &kL, &dL, &cL, &A->capL);
    ygz_orb_device_outputs(right, &kR, &dR, &cR, &A->capR);
    if (A->capL > SM_SORT || A->capR > 16384) { set_last_error("batched stereo: more than 4096 left / 16384 right key slots per frame"); return YGZ_ERR_CAPACITY; }
    A->kL = (const ygz_keypoint*)kL; A->kR = (const ygz_keypoint*)kR;
    A->dL = (const uint4*)dL; A->dR = (const uint4*)dR;
    A->cntL = (const int*)cL; A->cntR = (const int*)cR;
    A->nRows = A->lv.h[0];
    A->maxD = mbf / mb;                                 // minZ = mb, :541-543
    A->mbf = mbf;
    const int n = n_frames * A->capL;
    if ((rc = grow(&S->d_bu, &S->cap_bu, n)) || (rc = grow(&S->d_bz, &S->cap_bz, n)) || (rc = grow(&S->d_bsad, &S->cap_bsad, n)) ||
        (rc = grow(&S->d_bn, &S->cap_bn, n_frames)) || (rc = grow(&S->d_bfi, &S->cap_bfi, n)) || (rc = grow(&S->d_bfd, &S->cap_bfd, n)) ||
        (rc = grow(&S->d_bfs, &S->cap_bfs, n))) return rc;
    A->u = S->d_bu; A->z = S->d_bz; A->sad = S->d_bsad;
    return YGZ_OK;
}

bool g_stereo_smem_set = false;

// association + median cut (+ brute force) of n_frames pairs on stream st; results stay in the scratch
int stereo_batch_launch(ygz_stereo* S, const StereoBatch& A, int n_frames, int with_bf, cudaStream_t st, cudaEvent_t after_match,
                        cudaEvent_t after_cut) {
    const size_t smem = (size_t)A.capR * sizeof(RightKey);
    if (!g_stereo_smem_set) {
        YGZ_CUDA_TRY(cudaFuncSetAttribute(k_stereo_match_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * (int)sizeof(RightKey)));
        g_stereo_smem_set = true;
    }
    const dim3 grid(ceil_div(A.capL, 8 * SB_KEYS_PER_WARP), n_frames);
    { NvtxRange nvtx_range("ygz/ComputeStereoMatches batch");
    k_stereo_match_batch<<<grid, 256, smem, st>>>(A);
    }
    YGZ_CUDA_TRY(cudaGetLastError());
    if (after_match) YGZ_CUDA_TRY(cudaEventRecord(after_match, st));
    k_stereo_median_cut<<<n_frames, 1024, 0, st>>>(S->d_bu, S->d_bz, S->d_bsad, A.cntL, A.capL, S->d_bn);
    YGZ_CUDA_TRY(cudaGetLastError());
    if (after_cut) YGZ_CUDA_TRY(cudaEventRecord(after_cut, st));
    if (with_bf) {
        YGZ_CUDA_TRY(cudaMemsetAsync(S->d_bfi, 0xFF, (size_t)n_frames * A.capL * 4, st));
        const dim3 g2(ceil_div(A.capL, BF_WARPS * BF_QPW), n_frames);
        k_hamming_bf_top2_batch<<<g2, BF_WARPS * 32, 0, st>>>((const uint32_t*)A.dL, A.cntL, A.capL, (const uint32_t*)A.dR, A.cntR, A.capR,
                                                            S->d_bfi, S->d_bfd, S->d_bfs);
        YGZ_CUDA_TRY(cudaGetLastError());
    }
    return YGZ_OK;
}

}  // namespace

extern "C" {

/* Frame::ComputeStereoMatches, src/Frame.cc:509-682.  The level images are read in place from the two extractor handles
 * (frame `frame` of their last batch: ComputePyramid / extract must have run on both); key points and descriptors are the
 * extractors' host outputs.  Outputs mvuRight / mvDepth (-1 = no match); returns the number of stereo matches (>= 0) or a
 * negative ygz_status. */
int ygz_compute_stereo_matches(ygz_orb* left, ygz_orb* right, int frame, const ygz_keypoint* keysL, const uint8_t* descL, int N,
                               const ygz_keypoint* keysR, const uint8_t* descR, int Nr, float mb, float mbf,
                               float* mvuRight, float* mvDepth) {
    YGZ_ARG_CHECK(left && right && frame >= 0 && N >= 0 && Nr >= 0 && mvuRight && mvDepth && mb > 0.f);
    YGZ_ARG_CHECK(frame < ygz_orb_max_batch(left) && frame < ygz_orb_max_batch(right));
    YGZ_ARG_CHECK((N == 0 || (keysL && descL)) && (Nr == 0 || (keysR && descR)));
    for (int i = 0; i < N; i++) { mvuRight[i] = -1.0f; mvDepth[i] = -1.0f; }
    if (N == 0 || Nr == 0) return 0;
    float scale[16], inv_scale[16], sg[16], isg[16];
    int fpl[16];
    int rc = ygz_orb_tables(left, scale, inv_scale, sg, isg, fpl);
    if (rc) return rc;
    int nlevels = 0;
    StereoLevels lv;
    std::memset(&lv, 0, sizeof(lv));
    for (int l = 0; l < 16; l++) {
        int w = 0, hh = 0, w2 = 0, h2 = 0;
        if (ygz_orb_level_size(left, l, &w, &hh) != YGZ_OK) break;
        if (ygz_orb_level_size(right, l, &w2, &h2) != YGZ_OK || w2 != w || h2 != hh) { set_last_error("left / right extractors differ"); return YGZ_ERR_ARG; }
        void *pl = nullptr, *pr = nullptr; int pitl = 0, pitr = 0; size_t fsl = 0, fsr = 0;
        ygz_orb_device_level(left, l, &pl, &pitl, &fsl);
        ygz_orb_device_level(right, l, &pr, &pitr, &fsr);
        lv.L[l] = (const uint8_t*)pl + fsl * frame; lv.R[l] = (const uint8_t*)pr + fsr * frame;
        lv.pitchL[l] = pitl; lv.pitchR[l] = pitr; lv.w[l] = w; lv.h[l] = hh;
        lv.scale[l] = scale[l]; lv.inv_scale[l] = inv_scale[l];
        nlevels = l + 1;
    }
    YGZ_ARG_CHECK(nlevels >= 1);
    for (int i = 0; i < N; i++) YGZ_ARG_CHECK(keysL[i].octave >= 0 && keysL[i].octave < nlevels);
    for (int i = 0; i < Nr; i++) YGZ_ARG_CHECK(keysR[i].octave >= 0 && keysR[i].octave < nlevels);
    // ---- vRowIndices (:518-537) as CSR, insertion order = iR ascending per row
    const int nRows = lv.h[0];
    std::vector<int> off(nRows + 1, 0);
    std::vector<std::pair<int, int>> band(Nr);
    for (int iR = 0; iR < Nr; iR++) {
        const float kpY = keysR[iR].y;
        const float r = 2.0f * scale[keysR[iR].octave];
        const int maxr = std::min((int)std::ceil(kpY + r), nRows - 1), minr = std::max((int)std::floor(kpY - r), 0);
        band[iR] = std::make_pair(minr, maxr);
        for (int yi = minr; yi <= maxr; yi++) off[yi + 1]++;
    }
    for (int y = 0; y < nRows; y++) off[y + 1] += off[y];
    std::vector<int> idx(std::max(off[nRows], 1)), fill(off.begin(), off.end() - 1);
    for (int iR = 0; iR < Nr; iR++)
        for (int yi = band[iR].first; yi <= band[iR].second; yi++) idx[fill[yi]++] = iR;

    int dev = 0;
    {
        void* p = nullptr; int pit; size_t fs;
        ygz_orb_device_level(left, 0, &p, &pit, &fs);
        cudaPointerAttributes at;
        YGZ_CUDA_TRY(cudaPointerGetAttributes(&at, p));
        dev = at.device;
    }
    YGZ_CUDA_TRY(cudaSetDevice(dev));
    std::lock_guard<std::mutex> lock(g_stereo_mutex);
    ygz_stereo* S = nullptr;
    if ((rc = stereo_scratch(dev, nullptr, &S))) return rc;
    if ((rc = grow(&S->d_kL, &S->cap_kL, N)) || (rc = grow(&S->d_dL, &S->cap_dL, N * 32)) || (rc = grow(&S->d_u, &S->cap_u, N)) ||
        (rc = grow(&S->d_z, &S->cap_z, N)) || (rc = grow(&S->d_sad, &S->cap_sad, N)) ||
        (rc = grow(&S->d_kR, &S->cap_kR, Nr)) || (rc = grow(&S->d_dR, &S->cap_dR, Nr * 32))) return rc;
    if ((rc = grow(&S->d_row_off, &S->cap_rows, nRows + 1)) || (rc = grow(&S->d_row_idx, &S->cap_idx, (int)idx.size()))) return rc;
    // both extractors must have finished writing their pyramids
    YGZ_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)ygz_orb_stream(left)));
    YGZ_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)ygz_orb_stream(right)));
    YGZ_CUDA_TRY(cudaMemcpyAsync(S->d_kL, keysL, (size_t)N * sizeof(ygz_keypoint), cudaMemcpyHostToDevice, S->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(S->d_dL, descL, (size_t)N * 32, cudaMemcpyHostToDevice, S->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(S->d_kR, keysR, (size_t)Nr * sizeof(ygz_keypoint), cudaMemcpyHostToDevice, S->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(S->d_dR, descR, (size_t)Nr * 32, cudaMemcpyHostToDevice, S->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(S->d_row_off, off.data(), (size_t)(nRows + 1) * 4, cudaMemcpyHostToDevice, S->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(S->d_row_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, S->stream));
    const float maxD = mbf / mb;                        // minZ = mb, :541-543
    { NvtxRange nvtx_range("ygz/ComputeStereoMatches");
    k_stereo_match<<<ceil_div(N, 8), 256, 0, S->stream>>>(lv, S->d_kL, reinterpret_cast<const uint4*>(S->d_dL), N, S->d_kR,
                                                         reinterpret_cast<const uint4*>(S->d_dR), S->d_row_off, S->d_row_idx, nRows,
                                                         maxD, mbf, S->d_u, S->d_z, S->d_sad);
    }
    YGZ_CUDA_TRY(cudaGetLastError());
    std::vector<int> sad(N);
    YGZ_CUDA_TRY(cudaMemcpyAsync(mvuRight, S->d_u, (size_t)N * 4, cudaMemcpyDeviceToHost, S->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(mvDepth, S->d_z, (size_t)N * 4, cudaMemcpyDeviceToHost, S->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(sad.data(), S->d_sad, (size_t)N * 4, cudaMemcpyDeviceToHost, S->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(S->stream));
    // ---- median cut (:668-681)
    std::vector<std::pair<int, int>> vDistIdx;
    for (int i = 0; i < N; i++) if (sad[i] >= 0) vDistIdx.push_back(std::make_pair(sad[i], i));
    if (vDistIdx.empty()) return 0;
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = (float)vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    int n = (int)vDistIdx.size();
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if (vDistIdx[i].first < thDist) break;
        mvuRight[vDistIdx[i].second] = -1;
        mvDepth[vDistIdx[i].second] = -1;
        n--;
    }
    return n;
}

/* Batched, device-resident Frame::ComputeStereoMatches (+ the dense brute-force matcher of BASELINE C5): see ygz_b200.h */
int ygz_stereo_batch_async(ygz_orb* left, ygz_orb* right, int n_frames, float mb, float mbf, float* mvuRight, float* mvDepth,
                           int* n_matches, int* bf_best, int* bf_dist, int* bf_second) {
    YGZ_ARG_CHECK(left && right && n_frames >= 1 && mb > 0.f);
    YGZ_ARG_CHECK(n_frames <= ygz_orb_max_batch(left) && n_frames <= ygz_orb_max_batch(right));
    int dev = 0, dev_r = 0, rc;
    if ((rc = device_of(left, &dev)) || (rc = device_of(right, &dev_r))) return rc;
    YGZ_ARG_CHECK(dev == dev_r);
    YGZ_CUDA_TRY(cudaSetDevice(dev));
    std::lock_guard<std::mutex> lock(g_stereo_mutex);
    ygz_stereo* S = nullptr;
    if ((rc = stereo_scratch(dev, left, &S))) return rc;
    StereoBatch A;
    if ((rc = stereo_batch_args(left, right, S, n_frames, mb, mbf, &A))) return rc;
    cudaStream_t st = (cudaStream_t)ygz_orb_stream(left);
    // the right extractor's stream produced half of the inputs: order the left stream after it
    YGZ_CUDA_TRY(cudaEventRecord(S->ev_right, (cudaStream_t)ygz_orb_stream(right)));
    YGZ_CUDA_TRY(cudaStreamWaitEvent(st, S->ev_right, 0));
    const int with_bf = bf_best || bf_dist || bf_second;
    if ((rc = stereo_batch_launch(S, A, n_frames, with_bf, st, nullptr, nullptr))) return rc;
    const size_t bytes = (size_t)n_frames * A.capL * 4;
    if (mvuRight) YGZ_CUDA_TRY(cudaMemcpyAsync(mvuRight, S->d_bu, bytes, cudaMemcpyDeviceToHost, st));
    if (mvDepth) YGZ_CUDA_TRY(cudaMemcpyAsync(mvDepth, S->d_bz, bytes, cudaMemcpyDeviceToHost, st));
    if (n_matches) YGZ_CUDA_TRY(cudaMemcpyAsync(n_matches, S->d_bn, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, st));
    if (bf_best) YGZ_CUDA_TRY(cudaMemcpyAsync(bf_best, S->d_bfi, bytes, cudaMemcpyDeviceToHost, st));
    if (bf_dist) YGZ_CUDA_TRY(cudaMemcpyAsync(bf_dist, S->d_bfd, bytes, cudaMemcpyDeviceToHost, st));
    if (bf_second) YGZ_CUDA_TRY(cudaMemcpyAsync(bf_second, S->d_bfs, bytes, cudaMemcpyDeviceToHost, st));
    return YGZ_OK;
}

int ygz_stereo_batch(ygz_orb* left, ygz_orb* right, int n_frames, float mb, float mbf, float* mvuRight, float* mvDepth, int* n_matches,
                     int* bf_best, int* bf_dist, int* bf_second) {
    int rc = ygz_stereo_batch_async(left, right, n_frames, mb, mbf, mvuRight, mvDepth, n_matches, bf_best, bf_dist, bf_second);
    if (rc) return rc;
    return ygz_orb_wait(left);
}

/* bench.py (C5 `value`): iters x {both extractors' kernels on their own streams, then the batched association + median cut +
 * brute force on the left stream}, inputs resident, CUDA events on the left stream.  stage_ms[4] (optional, last iteration):
 * {extractors (L || R), association, median cut, brute force}. */
int ygz_stereo_frontend_time(ygz_orb* left, ygz_orb* right, int n_frames, int iters, float mb, float mbf, int with_bf, float* ms,
                             float* stage_ms) {
    YGZ_ARG_CHECK(left && right && n_frames >= 1 && iters >= 1 && ms && mb > 0.f);
    YGZ_ARG_CHECK(n_frames <= ygz_orb_max_batch(left) && n_frames <= ygz_orb_max_batch(right));
    int dev = 0, dev_r = 0, rc;
    if ((rc = device_of(left, &dev)) || (rc = device_of(right, &dev_r))) return rc;
    YGZ_ARG_CHECK(dev == dev_r);
    YGZ_CUDA_TRY(cudaSetDevice(dev));
    std::lock_guard<std::mutex> lock(g_stereo_mutex);
    ygz_stereo* S = nullptr;
    if ((rc = stereo_scratch(dev, left, &S))) return rc;
    StereoBatch A;
    if ((rc = stereo_batch_args(left, right, S, n_frames, mb, mbf, &A))) return rc;
    cudaStream_t sl = (cudaStream_t)ygz_orb_stream(left), sr = (cudaStream_t)ygz_orb_stream(right);
    cudaEvent_t t0 = nullptr, t1 = nullptr, fork = nullptr;
    YGZ_CUDA_TRY(cudaEventCreate(&t0)); YGZ_CUDA_TRY(cudaEventCreate(&t1));
    YGZ_CUDA_TRY(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    YGZ_CUDA_TRY(cudaStreamSynchronize(sl)); YGZ_CUDA_TRY(cudaStreamSynchronize(sr));
    YGZ_CUDA_TRY(cudaEventRecord(t0, sl));
    for (int it = 0; it < iters && !rc; it++) {
        const bool last = it == iters - 1;
        YGZ_CUDA_TRY(cudaEventRecord(fork, sl));                   // the right stream starts with the left one
        YGZ_CUDA_TRY(cudaStreamWaitEvent(sr, fork, 0));
        if (last) YGZ_CUDA_TRY(cudaEventRecord(S->ev0, sl));
        if ((rc = ygz_orb_run(left, n_frames)) || (rc = ygz_orb_run(right, n_frames))) break;
        YGZ_CUDA_TRY(cudaEventRecord(S->ev_right, sr));
        YGZ_CUDA_TRY(cudaStreamWaitEvent(sl, S->ev_right, 0));
        if (last) YGZ_CUDA_TRY(cudaEventRecord(S->ev1, sl));
        rc = stereo_batch_launch(S, A, n_frames, with_bf, sl, last ? S->ev2 : nullptr, last ? S->ev3 : nullptr);
    }
    if (!rc) {
        YGZ_CUDA_TRY(cudaEventRecord(t1, sl));
        YGZ_CUDA_TRY(cudaEventSynchronize(t1));
        YGZ_CUDA_TRY(cudaEventElapsedTime(ms, t0, t1));
        if (stage_ms) {
            YGZ_CUDA_TRY(cudaEventElapsedTime(&stage_ms[0], S->ev0, S->ev1));
            YGZ_CUDA_TRY(cudaEventElapsedTime(&stage_ms[1], S->ev1, S->ev2));
            YGZ_CUDA_TRY(cudaEventElapsedTime(&stage_ms[2], S->ev2, S->ev3));
            YGZ_CUDA_TRY(cudaEventElapsedTime(&stage_ms[3], S->ev3, t1));
        }
    }
    cudaEventDestroy(t0); cudaEventDestroy(t1); cudaEventDestroy(fork);
    return rc;
}

}  // extern "C"

