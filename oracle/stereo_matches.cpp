// ORACLE — TEST INFRASTRUCTURE ONLY (see ygz_oracle.h).  CPU restatement of Frame::ComputeStereoMatches
// (/root/reference/src/Frame.cc:509-682), SURVEY.md §8(f) rank 3: row-band candidate table, Hamming search over the
// right key points of a compatible octave inside the disparity range, 11x11 SAD sliding window (+-5 px) on the key
// point's pyramid level, parabola refinement, and the final median-based outlier cut.
// cv::norm(IL, IR, NORM_L1) on CV_32F sums |a-b| in double (integers here, exact) and the result is stored to float.
// std::round is half-away-from-zero.  Oracle rule: with no accepted match the reference indexes an empty vector
// (:670-671); the oracle returns with nothing to cut.
// PARITY UNPINNED (no reference test or vector for this function).
#include "ygz_oracle.h"
#include <algorithm>
#include <climits>
#include <cmath>
#include <utility>
#include <vector>

static inline int hamming256(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; i++) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

extern "C" int orc_compute_stereo_matches(const orc_direct_frame* pyrL, const orc_direct_frame* pyrR, int nlevels,
                                          const orc_keypoint* keysL, const uint8_t* descL, int N,
                                          const orc_keypoint* keysR, const uint8_t* descR, int Nr,
                                          const float* mvScaleFactors, const float* mvInvScaleFactors, float mb, float mbf,
                                          float* mvuRight, float* mvDepth) {
    const int TH_HIGH = 100, TH_LOW = 50;
    for (int i = 0; i < N; i++) { mvuRight[i] = -1.0f; mvDepth[i] = -1.0f; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const int nRows = pyrL->h[0];
    std::vector<std::vector<size_t>> vRowIndices(nRows);
    for (int iR = 0; iR < Nr; iR++) {
        const float kpY = keysR[iR].y;
        const float r = 2.0f * mvScaleFactors[keysR[iR].octave];
        const int maxr = (int)std::ceil(kpY + r);
        const int minr = (int)std::floor(kpY - r);
        for (int yi = minr; yi <= maxr; yi++)
            if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);      // (the reference indexes unchecked)
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    std::vector<std::pair<int, int>> vDistIdx;
    for (int iL = 0; iL < N; iL++) {
        const orc_keypoint& kpL = keysL[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        if (vL < 0 || (size_t)vL >= (size_t)nRows) continue;
        const std::vector<size_t>& vCandidates = vRowIndices[(size_t)vL];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        size_t bestIdxR = 0;
        const uint8_t* dL = descL + (size_t)iL * 32;
        for (size_t iC = 0; iC < vCandidates.size(); iC++) {
            const size_t iR = vCandidates[iC];
            const orc_keypoint& kpR = keysR[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = hamming256(dL, descR + iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = keysR[bestIdxR].x;
            const float scaleFactor = mvInvScaleFactors[kpL.octave];
            const float scaleduL = std::round(kpL.x * scaleFactor);
            const float scaledvL = std::round(kpL.y * scaleFactor);
            const float scaleduR0 = std::round(uR0 * scaleFactor);
            const int w = 5;
            const uint8_t* imL = pyrL->level[kpL.octave];
            const uint8_t* imR = pyrR->level[kpL.octave];
            const int pL = pyrL->pitch[kpL.octave], pR = pyrR->pitch[kpL.octave];
            const int r0 = (int)(scaledvL - w), c0 = (int)(scaleduL - w);
            float IL[11][11];
            {
                const float ctr = (float)imL[(r0 + w) * pL + c0 + w];
                for (int y = 0; y < 11; y++)
                    for (int x = 0; x < 11; x++) IL[y][x] = (float)imL[(r0 + y) * pL + c0 + x] - ctr * 1.0f;
            }
            int bestDistS = INT_MAX;
            int bestincR = 0;
            const int L = 5;
            float vDists[11];
            const float iniu = scaleduR0 + L - w;
            const float endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= pyrR->w[kpL.octave]) continue;
            for (int incR = -L; incR <= +L; incR++) {
                const int cr = (int)(scaleduR0 + incR - w);
                const float ctr = (float)imR[(r0 + w) * pR + cr + w];
                double s = 0;
                for (int y = 0; y < 11; y++)
                    for (int x = 0; x < 11; x++) s += std::fabs((double)(IL[y][x] - ((float)imR[(r0 + y) * pR + cr + x] - ctr * 1.0f)));
                const float dist = (float)s;
                if (dist < bestDistS) { bestDistS = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = mvScaleFactors[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01f; bestuR = (float)(uL - 0.01); }
                mvDepth[iL] = mbf / disparity;
                mvuRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(bestDistS, iL));
            }
        }
    }
    if (vDistIdx.empty()) return 0;
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = (float)vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if (vDistIdx[i].first < thDist) break;
        mvuRight[vDistIdx[i].second] = -1;
        mvDepth[vDistIdx[i].second] = -1;
    }
    int n = 0;
    for (int i = 0; i < N; i++) n += mvuRight[i] >= 0 || mvDepth[i] > 0;
    return n;
}
