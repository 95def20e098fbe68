// ORACLE — TEST INFRASTRUCTURE ONLY (see ygz_oracle.h).  CPU restatement of ygz::ORBmatcher
// (/root/reference/src/ORBmatcher.cc) and the Frame grid helpers it uses
// (/root/reference/src/Frame.cc:314-330, 424-493) over flat array views of the pointer graph.
#include "ygz_oracle.h"
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

const int TH_HIGH = 100;      // src/ORBmatcher.cc:36
const int TH_LOW = 50;        // :37
const int HISTO_LENGTH = 30;  // :38
const int GRID_COLS = 64, GRID_ROWS = 48;   // include/Frame.h:38-39

// ---- src/ORBmatcher.cc:1507-1523
inline int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t va, vb;
        std::memcpy(&va, a + 4 * i, 4);
        std::memcpy(&vb, b + 4 * i, 4);
        dist += __builtin_popcount(va ^ vb);
    }
    return dist;
}

// ---- src/Frame.cc:314-330 + :483-493
struct Grid {
    std::vector<int> cell[GRID_COLS][GRID_ROWS];
    float invW, invH;
    explicit Grid(const orc_frame_view* f) {
        invW = static_cast<float>(GRID_COLS) / (f->max_x - f->min_x);
        invH = static_cast<float>(GRID_ROWS) / (f->max_y - f->min_y);
        for (int i = 0; i < f->n; i++) {
            int px = (int)std::round((f->keys[i].x - f->min_x) * invW);
            int py = (int)std::round((f->keys[i].y - f->min_y) * invH);
            if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
            cell[px][py].push_back(i);
        }
    }
};

// ---- src/Frame.cc:424-481
std::vector<int> features_in_area(const orc_frame_view* f, const Grid& g, float x, float y, float r,
                                  int minLevel = -1, int maxLevel = -1) {
    std::vector<int> out;
    const int nMinCellX = std::max(0, (int)std::floor((x - f->min_x - r) * g.invW));
    if (nMinCellX >= GRID_COLS) return out;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - f->min_x + r) * g.invW));
    if (nMaxCellX < 0) return out;
    const int nMinCellY = std::max(0, (int)std::floor((y - f->min_y - r) * g.invH));
    if (nMinCellY >= GRID_ROWS) return out;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - f->min_y + r) * g.invH));
    if (nMaxCellY < 0) return out;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
            for (int idx : g.cell[ix][iy]) {
                const orc_keypoint& kp = f->keys[idx];
                if (bCheckLevels) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                const float distx = kp.x - x, disty = kp.y - y;
                if (std::fabs(distx) < r && std::fabs(disty) < r) out.push_back(idx);
            }
    return out;
}

// ---- src/ORBmatcher.cc:1471-1502
void three_maxima(const int* sz, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = sz[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

inline int rot_bin(float a1, float a2) {
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)std::round(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

inline bool mp_bad(const orc_map_points* m, int i) { return m->bad && m->bad[i]; }
inline int mp_obs(const orc_map_points* m, int i) { return m->observations ? m->observations[i] : 1; }
inline float uright(const orc_frame_view* f, int i) { return f->u_right ? f->u_right[i] : -1.f; }

}  // namespace

extern "C" {

int orc_hamming256(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

void orc_three_maxima(const int* sz, int L, int* i1, int* i2, int* i3) {
    int a = -1, b = -1, c = -1;
    three_maxima(sz, L, a, b, c);
    *i1 = a; *i2 = b; *i3 = c;
}

int orc_features_in_area(const orc_frame_view* f, float x, float y, float r, int minL, int maxL, int* out, int cap) {
    Grid g(f);
    std::vector<int> v = features_in_area(f, g, x, y, r, minL, maxL);
    if ((int)v.size() > cap) return -(int)v.size();
    std::copy(v.begin(), v.end(), out);
    return (int)v.size();
}

void orc_bf_top2(const uint8_t* q, int nq, const uint8_t* t, int nt, int* bi, int* bd, int* sd) {
    for (int i = 0; i < nq; i++) {
        int best = 256, second = 256, idx = -1;
        for (int j = 0; j < nt; j++) {
            int d = descriptor_distance(q + 32 * (size_t)i, t + 32 * (size_t)j);
            if (d < best) { second = best; best = d; idx = j; }
            else if (d < second) second = d;
        }
        bi[i] = idx; bd[i] = best; sd[i] = second;
    }
}

// ---- src/ORBmatcher.cc:155-263
int orc_search_by_bow(float nnratio, int check_ori, const orc_frame_view* kf, const orc_map_points* mps,
                      const orc_frame_view* f, int* out) {
    for (int i = 0; i < f->n; i++) out[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    int ki = 0, fi = 0;
    while (ki < kf->fv_nodes && fi < f->fv_nodes) {
        const uint32_t kn = kf->fv_node_id[ki], fn = f->fv_node_id[fi];
        if (kn == fn) {
            for (int a = kf->fv_offset[ki]; a < kf->fv_offset[ki + 1]; a++) {
                const int realIdxKF = (int)kf->fv_index[a];
                const int pMP = kf->map_point[realIdxKF];
                if (pMP < 0) continue;
                if (mp_bad(mps, pMP)) continue;
                const uint8_t* dKF = kf->desc + 32 * (size_t)realIdxKF;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (int b = f->fv_offset[fi]; b < f->fv_offset[fi + 1]; b++) {
                    const int realIdxF = (int)f->fv_index[b];
                    if (out[realIdxF] >= 0) continue;
                    const int dist = descriptor_distance(dKF, f->desc + 32 * (size_t)realIdxF);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 <= TH_LOW) {
                    if (static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
                        out[bestIdxF] = pMP;
                        if (check_ori) rotHist[rot_bin(kf->keys[realIdxKF].angle, f->keys[bestIdxF].angle)].push_back(bestIdxF);
                        nmatches++;
                    }
                }
            }
            ki++; fi++;
        } else if (kn < fn) {
            while (ki < kf->fv_nodes && kf->fv_node_id[ki] < fn) ki++;     // lower_bound(Fit->first)
        } else {
            while (fi < f->fv_nodes && f->fv_node_id[fi] < kn) fi++;
        }
    }
    if (check_ori) {
        int sz[HISTO_LENGTH], i1 = -1, i2 = -1, i3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) sz[i] = (int)rotHist[i].size();
        three_maxima(sz, HISTO_LENGTH, i1, i2, i3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int idx : rotHist[i]) { out[idx] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// ---- src/ORBmatcher.cc:480-595  SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)
int orc_search_by_bow_kf(float nnratio, int check_ori, const orc_frame_view* kf1, const orc_frame_view* kf2,
                         const orc_map_points* mps, int* matches12) {
    for (int i = 0; i < kf1->n; i++) matches12[i] = -1;
    std::vector<char> vbMatched2(kf2->n, 0);
    std::vector<int> rotHist[HISTO_LENGTH];
    int nmatches = 0;
    int i1n = 0, i2n = 0;
    while (i1n < kf1->fv_nodes && i2n < kf2->fv_nodes) {
        const uint32_t n1 = kf1->fv_node_id[i1n], n2 = kf2->fv_node_id[i2n];
        if (n1 == n2) {
            for (int a = kf1->fv_offset[i1n]; a < kf1->fv_offset[i1n + 1]; a++) {
                const int idx1 = (int)kf1->fv_index[a];
                const int pMP1 = kf1->map_point[idx1];
                if (pMP1 < 0) continue;
                if (mp_bad(mps, pMP1)) continue;
                const uint8_t* d1 = kf1->desc + 32 * (size_t)idx1;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int b = kf2->fv_offset[i2n]; b < kf2->fv_offset[i2n + 1]; b++) {
                    const int idx2 = (int)kf2->fv_index[b];
                    const int pMP2 = kf2->map_point[idx2];
                    if (vbMatched2[idx2] || pMP2 < 0) continue;
                    if (mp_bad(mps, pMP2)) continue;
                    const int dist = descriptor_distance(d1, kf2->desc + 32 * (size_t)idx2);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < TH_LOW) {
                    if (static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
                        matches12[idx1] = kf2->map_point[bestIdx2];
                        vbMatched2[bestIdx2] = 1;
                        if (check_ori) rotHist[rot_bin(kf1->keys[idx1].angle, kf2->keys[bestIdx2].angle)].push_back(idx1);
                        nmatches++;
                    }
                }
            }
            i1n++; i2n++;
        } else if (n1 < n2) {
            while (i1n < kf1->fv_nodes && kf1->fv_node_id[i1n] < n2) i1n++;
        } else {
            while (i2n < kf2->fv_nodes && kf2->fv_node_id[i2n] < n1) i2n++;
        }
    }
    if (check_ori) {
        int sz[HISTO_LENGTH], a1 = -1, a2 = -1, a3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) sz[i] = (int)rotHist[i].size();
        three_maxima(sz, HISTO_LENGTH, a1, a2, a3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == a1 || i == a2 || i == a3) continue;
            for (int idx : rotHist[i]) { matches12[idx] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// ---- src/ORBmatcher.cc:136-153 CheckDistEpipolarLine (F12 row-major 3x3; sigma2 = pKF2->mvLevelSigma2)
static bool check_dist_epipolar_line(const orc_keypoint& kp1, const orc_keypoint& kp2, const float* F12, const float* sigma2) {
    const float a = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
    const float b = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
    const float c = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
    const float num = a * kp2.x + b * kp2.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * sigma2[kp2.octave];
}

// ---- src/ORBmatcher.cc:597-746  SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo).
// ex_ey: epipole of KF1's centre in KF2 (:603-609, computed by the caller from the poses); level_sigma2 = pKF2->mvLevelSigma2.
// pairs: up to kf1->n (idx1, idx2) pairs in ascending idx1.  Returns nmatches.
int orc_search_for_triangulation(int check_ori, const orc_frame_view* kf1, const orc_frame_view* kf2, const float* F12,
                                 float ex, float ey, const float* level_sigma2, int only_stereo, int* pairs) {
    int nmatches = 0;
    std::vector<char> vbMatched2(kf2->n, 0);
    std::vector<int> vMatches12(kf1->n, -1);
    std::vector<int> rotHist[HISTO_LENGTH];
    int i1n = 0, i2n = 0;
    while (i1n < kf1->fv_nodes && i2n < kf2->fv_nodes) {
        const uint32_t n1 = kf1->fv_node_id[i1n], n2 = kf2->fv_node_id[i2n];
        if (n1 == n2) {
            for (int a = kf1->fv_offset[i1n]; a < kf1->fv_offset[i1n + 1]; a++) {
                const int idx1 = (int)kf1->fv_index[a];
                if (kf1->map_point[idx1] >= 0) continue;                 // already a MapPoint
                const bool bStereo1 = uright(kf1, idx1) >= 0;
                if (only_stereo && !bStereo1) continue;
                const orc_keypoint& kp1 = kf1->keys[idx1];
                const uint8_t* d1 = kf1->desc + 32 * (size_t)idx1;
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int b = kf2->fv_offset[i2n]; b < kf2->fv_offset[i2n + 1]; b++) {
                    const int idx2 = (int)kf2->fv_index[b];
                    if (vbMatched2[idx2] || kf2->map_point[idx2] >= 0) continue;
                    const bool bStereo2 = uright(kf2, idx2) >= 0;
                    if (only_stereo && !bStereo2) continue;
                    const int dist = descriptor_distance(d1, kf2->desc + 32 * (size_t)idx2);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const orc_keypoint& kp2 = kf2->keys[idx2];
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ex - kp2.x, distey = ey - kp2.y;
                        if (distex * distex + distey * distey < 100 * kf2->scale_factors[kp2.octave]) continue;
                    }
                    if (check_dist_epipolar_line(kp1, kp2, F12, level_sigma2)) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    vMatches12[idx1] = bestIdx2;
                    nmatches++;
                    if (check_ori) rotHist[rot_bin(kp1.angle, kf2->keys[bestIdx2].angle)].push_back(idx1);
                }
            }
            i1n++; i2n++;
        } else if (n1 < n2) {
            while (i1n < kf1->fv_nodes && kf1->fv_node_id[i1n] < n2) i1n++;
        } else {
            while (i2n < kf2->fv_nodes && kf2->fv_node_id[i2n] < n1) i2n++;
        }
    }
    if (check_ori) {
        int sz[HISTO_LENGTH], a1 = -1, a2 = -1, a3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) sz[i] = (int)rotHist[i].size();
        three_maxima(sz, HISTO_LENGTH, a1, a2, a3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == a1 || i == a2 || i == a3) continue;
            for (int idx : rotHist[i]) { vMatches12[idx] = -1; nmatches--; }
        }
    }
    int np = 0;
    for (int i = 0; i < kf1->n; i++) {
        if (vMatches12[i] < 0) continue;
        pairs[2 * np] = i; pairs[2 * np + 1] = vMatches12[i];
        np++;
    }
    return nmatches;
}

// ---- src/ORBmatcher.cc:43-133
int orc_search_by_projection_points(float nnratio, orc_frame_view* F, const orc_map_points* mps,
                                    const int* query_ids, int nq, float th, int check_level) {
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    Grid g(F);
    for (int q = 0; q < nq; q++) {
        const int pMP = query_ids[q];
        if (!mps->track_in_view[pMP]) continue;
        if (mp_bad(mps, pMP)) continue;
        const int nPredictedLevel = mps->track_scale_level[pMP];
        float r = mps->track_view_cos[pMP] > 0.998 ? 2.5f : 4.0f;
        if (bFactor) r *= th;
        std::vector<int> vIndices = check_level
            ? features_in_area(F, g, mps->track_proj_x[pMP], mps->track_proj_y[pMP], r * F->scale_factors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel)
            : features_in_area(F, g, mps->track_proj_x[pMP], mps->track_proj_y[pMP], r * F->scale_factors[nPredictedLevel], -1, -1);
        if (vIndices.empty()) continue;
        const uint8_t* MPd = mps->desc + 32 * (size_t)pMP;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : vIndices) {
            if (F->map_point[idx] >= 0)
                if (mp_obs(mps, F->map_point[idx]) > 0) continue;
            if (uright(F, idx) > 0) {
                const float er = std::fabs(mps->track_proj_xr[pMP] - uright(F, idx));
                if (er > r * F->scale_factors[nPredictedLevel]) continue;
            }
            const int dist = descriptor_distance(MPd, F->desc + 32 * (size_t)idx);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel;
                bestLevel = F->keys[idx].octave; bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = F->keys[idx].octave; bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            F->map_point[bestIdx] = pMP;
            nmatches++;
        }
    }
    return nmatches;
}

// ---- src/ORBmatcher.cc:1218-1350
int orc_search_by_projection_last(int check_ori, orc_frame_view* C, const orc_frame_view* L,
                                  const orc_map_points* mps, float th, int mono, int check_level) {
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float* R = C->Rcw; const float* t = C->tcw;
    // twc = -Rcw^T tcw ; tlc = Rlw twc + tlw
    float twc[3], tlc[3];
    for (int i = 0; i < 3; i++) twc[i] = -1 * (R[0 * 3 + i] * t[0] + R[1 * 3 + i] * t[1] + R[2 * 3 + i] * t[2]);
    for (int i = 0; i < 3; i++) tlc[i] = L->Rcw[i * 3 + 0] * twc[0] + L->Rcw[i * 3 + 1] * twc[1] + L->Rcw[i * 3 + 2] * twc[2] + L->tcw[i];
    const bool bForward = tlc[2] > C->mb && !mono;
    const bool bBackward = -tlc[2] > C->mb && !mono;
    Grid g(C);
    for (int i = 0; i < L->n; i++) {
        const int pMP = L->map_point[i];
        if (pMP < 0) continue;
        if (L->outlier && L->outlier[i]) continue;
        const float* Pw = mps->world_pos + 3 * (size_t)pMP;
        const float xc = R[0] * Pw[0] + R[1] * Pw[1] + R[2] * Pw[2] + t[0];
        const float yc = R[3] * Pw[0] + R[4] * Pw[1] + R[5] * Pw[2] + t[1];
        const float zc = R[6] * Pw[0] + R[7] * Pw[1] + R[8] * Pw[2] + t[2];
        const float invzc = (float)(1.0 / zc);
        if (invzc < 0) continue;
        float u = C->fx * xc * invzc + C->cx;
        float v = C->fy * yc * invzc + C->cy;
        if (u < C->min_x || u > C->max_x) continue;
        if (v < C->min_y || v > C->max_y) continue;
        const int nLastOctave = L->keys[i].octave;
        const float radius = th * C->scale_factors[nLastOctave];
        std::vector<int> vIndices2;
        if (!check_level) vIndices2 = features_in_area(C, g, u, v, radius);
        else if (bForward) vIndices2 = features_in_area(C, g, u, v, radius, nLastOctave);
        else if (bBackward) vIndices2 = features_in_area(C, g, u, v, radius, 0, nLastOctave);
        else vIndices2 = features_in_area(C, g, u, v, radius, nLastOctave - 1, nLastOctave + 1);
        if (vIndices2.empty()) continue;
        const uint8_t* dMP = mps->desc + 32 * (size_t)pMP;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            if (C->map_point[i2] >= 0)
                if (mp_obs(mps, C->map_point[i2]) > 0) continue;
            if (uright(C, i2) > 0) {
                const float ur = u - C->mbf * invzc;
                const float er = std::fabs(ur - uright(C, i2));
                if (er > radius) continue;
            }
            const int dist = descriptor_distance(dMP, C->desc + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            C->map_point[bestIdx2] = pMP;
            nmatches++;
            if (check_ori) rotHist[rot_bin(L->keys[i].angle, C->keys[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (check_ori) {
        int sz[HISTO_LENGTH], i1 = -1, i2 = -1, i3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) sz[i] = (int)rotHist[i].size();
        three_maxima(sz, HISTO_LENGTH, i1, i2, i3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != i1 && i != i2 && i != i3)
                for (int idx : rotHist[i]) { C->map_point[idx] = -1; nmatches--; }
    }
    return nmatches;
}

// ---- src/ORBmatcher.cc:1352-1469
int orc_search_by_projection_kf(int check_ori, orc_frame_view* C, const orc_frame_view* KF,
                                const orc_map_points* mps, const uint8_t* found, float th, int ORBdist) {
    int nmatches = 0;
    const float* R = C->Rcw; const float* t = C->tcw;
    float Ow[3];
    for (int i = 0; i < 3; i++) Ow[i] = -1 * (R[0 * 3 + i] * t[0] + R[1 * 3 + i] * t[1] + R[2 * 3 + i] * t[2]);
    std::vector<int> rotHist[HISTO_LENGTH];
    Grid g(C);
    for (int i = 0; i < KF->n; i++) {
        const int pMP = KF->map_point[i];
        if (pMP < 0) continue;
        if (mp_bad(mps, pMP) || (found && found[pMP])) continue;
        const float* Pw = mps->world_pos + 3 * (size_t)pMP;
        const float xc = R[0] * Pw[0] + R[1] * Pw[1] + R[2] * Pw[2] + t[0];
        const float yc = R[3] * Pw[0] + R[4] * Pw[1] + R[5] * Pw[2] + t[1];
        const float zc = R[6] * Pw[0] + R[7] * Pw[1] + R[8] * Pw[2] + t[2];
        const float invzc = (float)(1.0 / zc);
        const float u = C->fx * xc * invzc + C->cx;
        const float v = C->fy * yc * invzc + C->cy;
        if (u < C->min_x || u > C->max_x) continue;
        if (v < C->min_y || v > C->max_y) continue;
        const float PO[3] = {Pw[0] - Ow[0], Pw[1] - Ow[1], Pw[2] - Ow[2]};
        const float dist3D = std::sqrt(PO[0] * PO[0] + PO[1] * PO[1] + PO[2] * PO[2]);
        const float maxDistance = 1.2f * mps->max_dist[pMP];     // src/MapPoint.cc:338-341
        const float minDistance = 0.8f * mps->min_dist[pMP];     // :333-336
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        // MapPoint::PredictScale src/MapPoint.cc:359-373
        const float ratio = mps->max_dist[pMP] / dist3D;
        int nPredictedLevel = (int)std::ceil(std::log(ratio) / C->log_scale_factor);
        if (nPredictedLevel < 0) nPredictedLevel = 0;
        else if (nPredictedLevel >= C->n_levels) nPredictedLevel = C->n_levels - 1;
        const float radius = th * C->scale_factors[nPredictedLevel];
        const std::vector<int> vIndices2 = features_in_area(C, g, u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1);
        if (vIndices2.empty()) continue;
        const uint8_t* dMP = mps->desc + 32 * (size_t)pMP;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            if (C->map_point[i2] >= 0) continue;
            const int dist = descriptor_distance(dMP, C->desc + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= ORBdist) {
            C->map_point[bestIdx2] = pMP;
            nmatches++;
            if (check_ori) rotHist[rot_bin(KF->keys[i].angle, C->keys[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (check_ori) {
        int sz[HISTO_LENGTH], i1 = -1, i2 = -1, i3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) sz[i] = (int)rotHist[i].size();
        three_maxima(sz, HISTO_LENGTH, i1, i2, i3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != i1 && i != i2 && i != i3)
                for (int idx : rotHist[i]) { C->map_point[idx] = -1; nmatches--; }
    }
    return nmatches;
}

// ---- src/ORBmatcher.cc:375-478
int orc_search_for_initialization(float nnratio, int check_ori, const orc_frame_view* F1, const orc_frame_view* F2,
                                  float* prev, int* m12, int windowSize) {
    int nmatches = 0;
    for (int i = 0; i < F1->n; i++) m12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<int> vMatchedDistance(F2->n, INT_MAX);
    std::vector<int> vnMatches21(F2->n, -1);
    Grid g(F2);
    for (int i1 = 0; i1 < F1->n; i1++) {
        const int level1 = F1->keys[i1].octave;
        if (level1 > 0) continue;
        std::vector<int> vIndices2 = features_in_area(F2, g, prev[2 * i1], prev[2 * i1 + 1], (float)windowSize, level1, level1);
        if (vIndices2.empty()) continue;
        const uint8_t* d1 = F1->desc + 32 * (size_t)i1;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            const int dist = descriptor_distance(d1, F2->desc + 32 * (size_t)i2);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) { m12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                m12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (check_ori) rotHist[rot_bin(F1->keys[i1].angle, F2->keys[bestIdx2].angle)].push_back(i1);
            }
        }
    }
    if (check_ori) {
        int sz[HISTO_LENGTH], i1 = -1, i2 = -1, i3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) sz[i] = (int)rotHist[i].size();
        three_maxima(sz, HISTO_LENGTH, i1, i2, i3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int idx1 : rotHist[i])
                if (m12[idx1] >= 0) { m12[idx1] = -1; nmatches--; }
        }
    }
    for (int i1 = 0; i1 < F1->n; i1++)
        if (m12[i1] >= 0) { prev[2 * i1] = F2->keys[m12[i1]].x; prev[2 * i1 + 1] = F2->keys[m12[i1]].y; }
    return nmatches;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// The Sim3 family of the back end (loop closing): SearchByProjection(KF, Scw, ...) :265-373, Fuse(KF, Scw, ...) :888-1004,
// SearchBySim3 :1006-1216.  They do their geometry in CV_32F cv::Mat algebra; the restatement follows the arithmetic OpenCV's small
// float matrices get in oracle/ref_mock/mock_cv.hpp (against which the reference's own code is compiled for the pin): a matrix
// product accumulates in double and rounds once, scaling multiplies by a double factor and rounds, elementwise ops are float.
namespace {
inline void mat33_vec(const float* R, const float* p, float* o) {
    for (int i = 0; i < 3; i++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += (double)R[3 * i + k] * p[k];
        o[i] = (float)s;
    }
}
inline float norm3(const float* p) { double s = 0; for (int k = 0; k < 3; k++) s += (double)p[k] * p[k]; return (float)std::sqrt(s); }
inline double dot3(const float* a, const float* b) { double s = 0; for (int k = 0; k < 3; k++) s += (double)a[k] * b[k]; return s; }

struct SimCam { float R[9], t[3], Ow[3]; };
// Decompose Scw (:274-278 / :897-901)
void decompose_scw(const float* S, SimCam& C) {
    double d = 0;
    for (int k = 0; k < 3; k++) d += (double)S[k] * S[k];                         // sRcw.row(0).dot(sRcw.row(0))
    const float scw = (float)std::sqrt(d);
    const double inv = 1.0 / scw;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C.R[3 * i + j] = (float)(S[4 * i + j] * inv);       // sRcw / scw
    for (int i = 0; i < 3; i++) C.t[i] = (float)(S[4 * i + 3] * inv);                                           // Scw.col(3) / scw
    float nRt[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) nRt[3 * i + j] = (float)(C.R[3 * j + i] * -1.0);   // -Rcw.t()
    mat33_vec(nRt, C.t, C.Ow);
}
inline int predict_scale(float max_dist_raw, float dist, const orc_frame_view* kf) {        // MapPoint::PredictScale(dist, pKF)
    const float ratio = max_dist_raw / dist;
    int nScale = (int)std::ceil(std::log(ratio) / kf->log_scale_factor);
    if (nScale < 0) nScale = 0; else if (nScale >= kf->n_levels) nScale = kf->n_levels - 1;
    return nScale;
}
inline bool kf_is_in_image(const orc_frame_view* kf, float x, float y) {                   // KeyFrame::IsInImage, integer bounds
    return x >= (int)kf->min_x && x < (int)kf->max_x && y >= (int)kf->min_y && y < (int)kf->max_y;
}
}  // namespace

extern "C" {

// SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th) :265-373.
// point_ids index mps; matched[kf->n]: map-point index per KF feature (in / out, -1 = NULL).
int orc_search_by_projection_sim3(const orc_frame_view* kf, const float* Scw, const orc_map_points* mps, const int* point_ids, int n_points,
                                  int* matched, int th) {
    SimCam C;
    decompose_scw(Scw, C);
    std::vector<char> already(mps->m, 0);                        // spAlreadyFound(vpMatched) minus NULL
    for (int i = 0; i < kf->n; i++) if (matched[i] >= 0) already[matched[i]] = 1;
    Grid g(kf);
    int nmatches = 0;
    for (int iMP = 0; iMP < n_points; iMP++) {
        const int pMP = point_ids[iMP];
        if (mp_bad(mps, pMP) || already[pMP]) continue;
        const float* Pw = mps->world_pos + 3 * (size_t)pMP;
        float pc[3];
        mat33_vec(C.R, Pw, pc);
        for (int k = 0; k < 3; k++) pc[k] = pc[k] + C.t[k];
        if (pc[2] < 0.0) continue;
        const float invz = 1 / pc[2];
        const float x = pc[0] * invz, y = pc[1] * invz;
        const float u = kf->fx * x + kf->cx, v = kf->fy * y + kf->cy;
        if (!kf_is_in_image(kf, u, v)) continue;
        const float maxDistance = 1.2f * mps->max_dist[pMP], minDistance = 0.8f * mps->min_dist[pMP];
        float PO[3];
        for (int k = 0; k < 3; k++) PO[k] = Pw[k] - C.Ow[k];
        const float dist = norm3(PO);
        if (dist < minDistance || dist > maxDistance) continue;
        const float* Pn = mps->normal + 3 * (size_t)pMP;
        if (dot3(PO, Pn) < 0.5 * dist) continue;
        const int lvl = predict_scale(mps->max_dist[pMP], dist, kf);
        const float radius = th * kf->scale_factors[lvl];
        const std::vector<int> cand = features_in_area(kf, g, u, v, radius);
        if (cand.empty()) continue;
        const uint8_t* dMP = mps->desc + 32 * (size_t)pMP;
        int bestDist = 256, bestIdx = -1;
        for (int idx : cand) {
            if (matched[idx] >= 0) continue;
            const int kpLevel = kf->keys[idx].octave;
            if (kpLevel < lvl - 1 || kpLevel > lvl) continue;
            const int d = descriptor_distance(dMP, kf->desc + 32 * (size_t)idx);
            if (d < bestDist) { bestDist = d; bestIdx = idx; }
        }
        if (bestDist <= TH_LOW) { matched[bestIdx] = pMP; nmatches++; }
    }
    return nmatches;
}

// Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint) :888-1004.
// kf->map_point is pKF's slot table (in / out: AddMapPoint fills free slots); replace_point[n_points] (in / out, the caller
// initialises it, upstream with NULL = -1).
int orc_fuse_sim3(orc_frame_view* kf, const float* Scw, const orc_map_points* mps, const int* point_ids, int n_points, float th,
                  int* replace_point) {
    SimCam C;
    decompose_scw(Scw, C);
    std::vector<char> already(mps->m, 0);                        // spAlreadyFound = pKF->GetMapPoints(): non-null and not bad, at entry
    for (int i = 0; i < kf->n; i++) if (kf->map_point[i] >= 0 && !mp_bad(mps, kf->map_point[i])) already[kf->map_point[i]] = 1;
    Grid g(kf);
    int nFused = 0;
    for (int iMP = 0; iMP < n_points; iMP++) {
        const int pMP = point_ids[iMP];
        if (mp_bad(mps, pMP) || already[pMP]) continue;
        const float* Pw = mps->world_pos + 3 * (size_t)pMP;
        float pc[3];
        mat33_vec(C.R, Pw, pc);
        for (int k = 0; k < 3; k++) pc[k] = pc[k] + C.t[k];
        if (pc[2] < 0.0f) continue;
        const float invz = (float)(1.0 / pc[2]);
        const float x = pc[0] * invz, y = pc[1] * invz;
        const float u = kf->fx * x + kf->cx, v = kf->fy * y + kf->cy;
        if (!kf_is_in_image(kf, u, v)) continue;
        const float maxDistance = 1.2f * mps->max_dist[pMP], minDistance = 0.8f * mps->min_dist[pMP];
        float PO[3];
        for (int k = 0; k < 3; k++) PO[k] = Pw[k] - C.Ow[k];
        const float dist3D = norm3(PO);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const float* Pn = mps->normal + 3 * (size_t)pMP;
        if (dot3(PO, Pn) < 0.5 * dist3D) continue;
        const int lvl = predict_scale(mps->max_dist[pMP], dist3D, kf);
        const float radius = th * kf->scale_factors[lvl];
        const std::vector<int> cand = features_in_area(kf, g, u, v, radius);
        if (cand.empty()) continue;
        const uint8_t* dMP = mps->desc + 32 * (size_t)pMP;
        int bestDist = INT_MAX, bestIdx = -1;
        for (int idx : cand) {
            const int kpLevel = kf->keys[idx].octave;
            if (kpLevel < lvl - 1 || kpLevel > lvl) continue;
            const int d = descriptor_distance(dMP, kf->desc + 32 * (size_t)idx);
            if (d < bestDist) { bestDist = d; bestIdx = idx; }
        }
        if (bestDist <= TH_LOW) {
            const int pMPinKF = kf->map_point[bestIdx];
            if (pMPinKF >= 0) {
                if (!mp_bad(mps, pMPinKF)) replace_point[iMP] = pMPinKF;
            } else {
                kf->map_point[bestIdx] = pMP;                  // pMP->AddObservation(pKF, bestIdx); pKF->AddMapPoint(pMP, bestIdx)
            }
            nFused++;
        }
    }
    return nFused;
}

// The per-point search of Fuse(pKF, vpMapPoints, th) :748-886 (local mapping; Eigen float geometry, stereo-aware chi2 gate): for
// every entry of point_ids (-1 = NULL) the key point of `kf` the loop body would pick — best_idx[i] / best_dist[i] (-1 / 256 when a
// gate fails or no candidate survives), WITHOUT the isBad / IsInKeyFrame skip and without the map mutation at the end of the body
// (those need the observation graph and are replayed by the caller in upstream's order; none of the quantities read here is changed
// by them for a point that is still eligible).  Ow = pKF->GetCameraCenter(), inv_level_sigma2 = pKF->mvInvLevelSigma2.
void orc_fuse_candidates(const orc_frame_view* kf, const float* Ow, const float* inv_level_sigma2, const orc_map_points* mps,
                         const int* point_ids, int n_points, float th, int* best_idx, int* best_dist) {
    Grid g(kf);
    const float* R = kf->Rcw;
    const float* t = kf->tcw;
    for (int i = 0; i < n_points; i++) {
        best_idx[i] = -1;
        best_dist[i] = 256;
        const int pMP = point_ids[i];
        if (pMP < 0) continue;
        const float* Pw = mps->world_pos + 3 * (size_t)pMP;
        float pc[3];
        for (int r = 0; r < 3; r++) {                                       // Rcw * p3Dw + tcw, float, left to right
            float a = R[3 * r] * Pw[0];
            a += R[3 * r + 1] * Pw[1];
            a += R[3 * r + 2] * Pw[2];
            pc[r] = a + t[r];
        }
        if (pc[2] < 0.0f) continue;
        const float invz = 1 / pc[2];
        const float x = pc[0] * invz, y = pc[1] * invz;
        const float u = kf->fx * x + kf->cx, v = kf->fy * y + kf->cy;
        if (!kf_is_in_image(kf, u, v)) continue;
        const float ur = u - kf->mbf * invz;
        const float maxDistance = 1.2f * mps->max_dist[pMP], minDistance = 0.8f * mps->min_dist[pMP];
        float PO[3];
        for (int k = 0; k < 3; k++) PO[k] = Pw[k] - Ow[k];
        float sq = PO[0] * PO[0]; sq += PO[1] * PO[1]; sq += PO[2] * PO[2];
        const float dist3D = std::sqrt(sq);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const float* Pn = mps->normal + 3 * (size_t)pMP;
        float dt = PO[0] * Pn[0]; dt += PO[1] * Pn[1]; dt += PO[2] * Pn[2];
        if (dt < 0.5 * dist3D) continue;                                    // viewing angle below 60 degrees
        const int lvl = predict_scale(mps->max_dist[pMP], dist3D, kf);
        const float radius = th * kf->scale_factors[lvl];
        const std::vector<int> cand = features_in_area(kf, g, u, v, radius);
        const uint8_t* dMP = mps->desc + 32 * (size_t)pMP;
        int bestDist = 256, bestIdx = -1;
        for (int idx : cand) {
            const orc_keypoint& kp = kf->keys[idx];
            const int kpLevel = kp.octave;
            if (kpLevel < lvl - 1 || kpLevel > lvl) continue;
            const float ex = u - kp.x, ey = v - kp.y;
            if (uright(kf, idx) >= 0) {
                const float er = ur - uright(kf, idx);
                const float e2 = ex * ex + ey * ey + er * er;
                if (e2 * inv_level_sigma2[kpLevel] > 7.8) continue;
            } else {
                const float e2 = ex * ex + ey * ey;
                if (e2 * inv_level_sigma2[kpLevel] > 5.99) continue;
            }
            const int d = descriptor_distance(dMP, kf->desc + 32 * (size_t)idx);
            if (d < bestDist) { bestDist = d; bestIdx = idx; }
        }
        best_idx[i] = bestIdx;
        best_dist[i] = bestDist;
    }
}

// SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) :1006-1216.  matches12[kf1->n] in / out (map-point index of KF2 or -1);
// idx_in_kf2[mps->m] = pMP->GetIndexInKeyFrame(pKF2) (-1 = not observed there).
int orc_search_by_sim3(const orc_frame_view* kf1, const orc_frame_view* kf2, const orc_map_points* mps, int* matches12,
                       const int* idx_in_kf2, float s12, const float* R12, const float* t12, float th) {
    const float fx = kf1->fx, fy = kf1->fy, cx = kf1->cx, cy = kf1->cy;            // pKF1's intrinsics in BOTH directions, as upstream
    float sR12[9], sR21[9], t21[3], nsR21[9];
    const double inv_s = 1.0 / s12;
    for (int i = 0; i < 9; i++) sR12[i] = (float)(R12[i] * (double)s12);           // s12 * R12
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) sR21[3 * i + j] = (float)(R12[3 * j + i] * inv_s);   // (1.0 / s12) * R12.t()
    for (int i = 0; i < 9; i++) nsR21[i] = (float)(sR21[i] * -1.0);
    mat33_vec(nsR21, t12, t21);                                                    // -sR21 * t12
    const int N1 = kf1->n, N2 = kf2->n;
    std::vector<char> am1(N1, 0), am2(N2, 0);
    for (int i = 0; i < N1; i++) {
        const int pMP = matches12[i];
        if (pMP >= 0) {
            am1[i] = 1;
            const int idx2 = idx_in_kf2[pMP];
            if (idx2 >= 0 && idx2 < N2) am2[idx2] = 1;
        }
    }
    std::vector<int> vnMatch1(N1, -1), vnMatch2(N2, -1);
    Grid g1(kf1), g2(kf2);
    auto one_way = [&](const orc_frame_view* from, const orc_frame_view* to, const Grid& gto, const std::vector<char>& am, const float* sR,
                       const float* tt, std::vector<int>& out) {
        for (int i = 0; i < from->n; i++) {
            const int pMP = from->map_point[i];
            if (pMP < 0 || am[i]) continue;
            if (mp_bad(mps, pMP)) continue;
            const float* Pw = mps->world_pos + 3 * (size_t)pMP;
            float pa[3], pb[3];
            mat33_vec(from->Rcw, Pw, pa);
            for (int k = 0; k < 3; k++) pa[k] = pa[k] + from->tcw[k];
            mat33_vec(sR, pa, pb);
            for (int k = 0; k < 3; k++) pb[k] = pb[k] + tt[k];
            if (pb[2] < 0.0) continue;
            const float invz = (float)(1.0 / pb[2]);
            const float x = pb[0] * invz, y = pb[1] * invz;
            const float u = fx * x + cx, v = fy * y + cy;
            if (!kf_is_in_image(to, u, v)) continue;
            const float maxDistance = 1.2f * mps->max_dist[pMP], minDistance = 0.8f * mps->min_dist[pMP];
            const float dist3D = norm3(pb);
            if (dist3D < minDistance || dist3D > maxDistance) continue;
            const int lvl = predict_scale(mps->max_dist[pMP], dist3D, to);
            const float radius = th * to->scale_factors[lvl];
            const std::vector<int> cand = features_in_area(to, gto, u, v, radius);
            if (cand.empty()) continue;
            const uint8_t* dMP = mps->desc + 32 * (size_t)pMP;
            int bestDist = INT_MAX, bestIdx = -1;
            for (int idx : cand) {
                const int o = to->keys[idx].octave;
                if (o < lvl - 1 || o > lvl) continue;
                const int d = descriptor_distance(dMP, to->desc + 32 * (size_t)idx);
                if (d < bestDist) { bestDist = d; bestIdx = idx; }
            }
            if (bestDist <= TH_HIGH) out[i] = bestIdx;
        }
    };
    one_way(kf1, kf2, g2, am1, sR21, t21, vnMatch1);
    one_way(kf2, kf1, g1, am2, sR12, t12, vnMatch2);
    int nFound = 0;
    for (int i1 = 0; i1 < N1; i1++) {
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0 && vnMatch2[idx2] == i1) { matches12[i1] = kf2->map_point[idx2]; nFound++; }
    }
    return nFound;
}

}  // extern "C"
