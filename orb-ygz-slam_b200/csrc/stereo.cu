// Frame::ComputeStereoMatches (/root/reference/src/Frame.cc:509-682), SURVEY.md §8(f) rank 3, on the device:
// one WARP per left key point.  Lanes stride over the right key points listed for the key's image row (the reference's
// vRowIndices band table, built on the host in the reference's insertion order), Hamming distance of the 256-bit
// descriptors, first-minimum selection by (distance, list position); then the 11 x 11 SAD of the centre-subtracted
// patches for the 11 column offsets on the key's pyramid level (levels are read in place from the two extractor
// handles, zero copy), parabola refinement and the disparity gates on lane 0 with explicit round-to-nearest float ops.
// The median-based outlier cut (:668-681) is a sort of <= N pairs and stays on the host.
#include "common.cuh"
#include <algorithm>
#include <climits>
#include <cmath>
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
        const float uR0 = keysR[iR].x;
        const float sf = lv.inv_scale[levelL];
        const float scaleduL = roundf(__fmul_rn(kpL.x, sf)), scaledvL = roundf(__fmul_rn(kpL.y, sf));
        const float scaleduR0 = roundf(__fmul_rn(uR0, sf));
        const int r0 = (int)(scaledvL - 5.f), cL0 = (int)(scaleduL - 5.f);
        const float iniu = scaleduR0 + 5.f - 5.f, endu = scaleduR0 + 5.f + 5.f + 1.f;
        if (!(iniu < 0.f || endu >= (float)lv.w[levelL])) {
            const uint8_t* imL = lv.L[levelL];
            const uint8_t* imR = lv.R[levelL];
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
    if (lane == 0) { uRight[iL] = out_u; depth[iL] = out_d; sad_out[iL] = out_sad; }
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
};

static ygz_stereo* g_stereo[16] = {};     // per-device scratch (buffers + stream), created on first use
static std::mutex g_stereo_mutex;          // calls are serialised: the scratch is shared by every caller of a device

template <class T>
static int grow(T** p, int* cap, int need) {
    if (need <= *cap) return YGZ_OK;
    cudaFree(*p); *p = nullptr;
    const int nc = std::max(need, 2 * *cap);
    YGZ_CUDA_TRY(cudaMalloc(p, (size_t)nc * sizeof(T)));
    *cap = nc;
    return YGZ_OK;
}

extern "C" {

/* Frame::ComputeStereoMatches, src/Frame.cc:509-682.  The level images are read in place from the two extractor handles
 * (frame `frame` of their last batch: ComputePyramid / extract must have run on both); key points and descriptors are the
 * extractors' host outputs.  Outputs mvuRight / mvDepth (-1 = no match); returns the number of stereo matches (>= 0) or a
 * negative ygz_status. */
int ygz_compute_stereo_matches(ygz_orb* left, ygz_orb* right, int frame, const ygz_keypoint* keysL, const uint8_t* descL, int N,
                               const ygz_keypoint* keysR, const uint8_t* descR, int Nr, float mb, float mbf,
                               float* mvuRight, float* mvDepth) {
    YGZ_ARG_CHECK(left && right && frame >= 0 && N >= 0 && Nr >= 0 && mvuRight && mvDepth && mb > 0.f);
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
    ygz_stereo*& S = g_stereo[dev & 15];
    if (!S) {
        S = new ygz_stereo();
        S->device = dev;
        YGZ_CUDA_TRY(cudaStreamCreateWithFlags(&S->stream, cudaStreamNonBlocking));
    }
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
    k_stereo_match<<<ceil_div(N, 8), 256, 0, S->stream>>>(lv, S->d_kL, reinterpret_cast<const uint4*>(S->d_dL), N, S->d_kR,
                                                         reinterpret_cast<const uint4*>(S->d_dR), S->d_row_off, S->d_row_idx, nRows,
                                                         maxD, mbf, S->d_u, S->d_z, S->d_sad);
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

}  // extern "C"
