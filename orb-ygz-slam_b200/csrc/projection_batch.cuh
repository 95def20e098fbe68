// Batched, device-resident projection searches (included by matcher.cu):
//
//   MODE 0  ORBmatcher::SearchByProjection(Frame& Current, const Frame& Last, th, bMono, checkLevel)
//           /root/reference/src/ORBmatcher.cc:1218-1350 — the per-frame call of Tracking::TrackWithMotionModel (src/Tracking.cc:1087)
//   MODE 1  ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>&, th, checkLevel)
//           :43-126 — the per-frame call of Tracking::SearchLocalPoints (src/Tracking.cc:1590)
//
// for a batch of frames whose key points / descriptors are the extractor's resident outputs.  One CTA per (query set, frame):
//
//   * the 64 x 48 feature grid of the frame (Frame::AssignFeaturesToGrid / PosInGrid, src/Frame.cc:314-330, 483-493) is built in
//     shared memory, every cell's list sorted by feature index, so Frame::GetFeaturesInArea (:424-481) visits the candidates in the
//     reference's order (its two cell loops, indices ascending inside a cell) and `dist < bestDist` keeps the same first minimum;
//   * the reference loop is SEQUENTIAL through the occupancy of mvpMapPoints (a query only considers features no earlier query took).
//     It is resolved in rounds ("deterministic reservations"): every unresolved query q scans its window over the features not yet
//     taken and registers itself, by atomicMin, on each candidate that can influence its decision; q becomes final when it is the
//     smallest registered query on all of them — then no earlier query can change what q sees, and q decides exactly as the
//     sequential loop would.  The smallest unresolved query always qualifies, so the rounds terminate (typically 2-5 rounds);
//   * MODE 0 ends with the rotation histogram + ComputeThreeMaxima on the finished assignment (:1317-1347).
//
// Assumption of the batched form (the per-call host entry points ygz_search_by_projection_* handle the general case): every query
// map point has Observations() > 0, i.e. a feature it takes blocks later queries (:1293-1295, :90-92).
#pragma once

namespace {

struct ProjArgs {
    const ygz_keypoint* kps; const uint4* desc; const int* counts; int cap;   // extractor outputs [frames][cap]
    const int *q_frame, *t_frame;       // per pair: frame of the queries (MODE 0: the last frame), target (current) frame
    // MODE 0: per LAST-frame feature
    const float* world_pos;             // [frames][cap][3]  mvpMapPoints[i]->GetWorldPos()
    const uint8_t* valid;               // [frames][cap]     map point non-null && !mvbOutlier[i]
    const float* pose;                  // [frames][12]      mTcw: R row-major, t
    // MODE 1: per query map point (M per pair)
    int M;
    const float *proj_x, *proj_y, *proj_xr, *view_cos;      // [pairs][M]  mTrackProjX / Y / XR, mTrackViewCos
    const int* level;                   // [pairs][M]  mnTrackScaleLevel
    const uint8_t* usable;              // [pairs][M]  mbTrackInView && !isBad()
    const uint4* q_desc;                // MODE 1: [pairs][M][2] GetDescriptor(); MODE 0: optional [frames][cap][2] (NULL: the last frame's own)
    const uint8_t* occupied;            // [pairs][cap] (optional): target feature holds a map point with Observations() > 0 before the call
    const float* u_right;               // [frames][cap] (optional) mvuRight of the target frame
    float fx, fy, cx, cy, mb, mbf, min_x, min_y, max_x, max_y;
    float scale[16];
    float th, nnratio;
    int mono, check_level, check_ori;
    int* matches;                       // [pairs][cap]: query index assigned to target feature, -1 = none
    int* nmatches;                      // [pairs]
    int* rounds;                        // [pairs] resolution rounds (diagnostic)
};

constexpr int PJ_THREADS = 256;
constexpr int PJ_CELLS = GRID_COLS * GRID_ROWS;
constexpr int PJ_MAX_CAP = 4096;

inline size_t proj_smem_bytes(int cap, int nq) {
    return (size_t)cap * (32 + 4 * 4 + 4 + 2 + 2 + 1) + (size_t)nq * (4 + 1) + (size_t)(PJ_CELLS + 1) * 4 + 64 * 4 + 64;
}

struct ProjQuery { float x, y, r, gate_ref, gate_r; int minLevel, maxLevel; bool live; };

// GetFeaturesInArea's cell window (src/Frame.cc:428-446); false = empty result
__device__ __forceinline__ bool proj_window(const ProjArgs& A, float invW, float invH, const ProjQuery& Q, int& x0, int& x1, int& y0, int& y1) {
    x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(Q.x, A.min_x), Q.r), invW)));
    if (x0 >= GRID_COLS) return false;
    x1 = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(Q.x, A.min_x), Q.r), invW)));
    if (x1 < 0) return false;
    y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(Q.y, A.min_y), Q.r), invH)));
    if (y0 >= GRID_ROWS) return false;
    y1 = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(Q.y, A.min_y), Q.r), invH)));
    return y1 >= 0;
}

template <int MODE>
__global__ void __launch_bounds__(PJ_THREADS) k_projection_batch(const __grid_constant__ ProjArgs A) {
    extern __shared__ __align__(16) unsigned char pj_smem[];
    const int p = blockIdx.x, tid = threadIdx.x;
    const int ft = A.t_frame[p], fq = A.q_frame[p];
    const int cap = A.cap;
    const int nT = min(A.counts[ft], cap);
    const int nQ = MODE == 0 ? min(A.counts[fq], cap) : A.M;
    const int qcap = MODE == 0 ? cap : A.M;
    // ---- shared-memory carve-up (16-byte items first)
    uint4* sdesc = reinterpret_cast<uint4*>(pj_smem);                               // [cap][2]
    float* sx = reinterpret_cast<float*>(sdesc + 2 * (size_t)cap);
    float* sy = sx + cap;
    float* sang = sy + cap;
    float* sur = sang + cap;
    int* want = reinterpret_cast<int*>(sur + cap);                                  // [cap] smallest unresolved query registered per feature
    int* best = want + cap;                                                         // [qcap] per query: chosen feature, -1 none
    int* cell_end = best + qcap;                                                    // [PJ_CELLS + 1] end offset of each cell's list
    int* hist = cell_end + PJ_CELLS + 1;                                            // [32] rotation histogram; [32..63] flags / scalars
    short* owner = reinterpret_cast<short*>(hist + 64);                            // [cap] query that took the feature, -1 free, -2 occupied before
    unsigned short* cell_idx = reinterpret_cast<unsigned short*>(owner + cap);     // [cap] cell lists
    unsigned char* soct = reinterpret_cast<unsigned char*>(cell_idx + cap);        // [cap]
    unsigned char* state = soct + cap;                                              // [qcap] 0 unresolved, 1 resolved, 2 final in this round
    int* flag = hist + 32;

    const float invW = __fdiv_rn((float)GRID_COLS, __fsub_rn(A.max_x, A.min_x));   // mfGridElementWidthInv, src/Frame.cc:151-152
    const float invH = __fdiv_rn((float)GRID_ROWS, __fsub_rn(A.max_y, A.min_y));
    const size_t oT = (size_t)ft * cap, oQ = (size_t)fq * cap;

    // ---- stage the target frame and build its grid
    for (int c = tid; c <= PJ_CELLS; c += PJ_THREADS) cell_end[c] = 0;
    if (tid < 64) hist[tid] = 0;
    __syncthreads();
    auto cell_of = [&](float x, float y) {
        const int px = (int)roundf(__fmul_rn(__fsub_rn(x, A.min_x), invW));          // PosInGrid :483-493
        const int py = (int)roundf(__fmul_rn(__fsub_rn(y, A.min_y), invH));
        return (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) ? -1 : px * GRID_ROWS + py;
    };
    for (int i = tid; i < nT; i += PJ_THREADS) {
        const ygz_keypoint k = A.kps[oT + i];
        sx[i] = k.x; sy[i] = k.y; sang[i] = k.angle; soct[i] = (unsigned char)k.octave;
        sur[i] = A.u_right ? A.u_right[oT + i] : -1.f;
        sdesc[2 * i] = A.desc[(oT + i) * 2]; sdesc[2 * i + 1] = A.desc[(oT + i) * 2 + 1];
        owner[i] = (A.occupied && A.occupied[(size_t)p * cap + i]) ? (short)-2 : (short)-1;
        const int c = cell_of(k.x, k.y);
        if (c >= 0) atomicAdd(&cell_end[c], 1);
    }
    for (int q = tid; q < nQ; q += PJ_THREADS) { state[q] = 0; best[q] = -1; }
    __syncthreads();
    {   // exclusive scan of the 3072 cell counts (12 cells per thread), cell_end[c] <- start of cell c
        constexpr int PER = PJ_CELLS / PJ_THREADS;
        static_assert(PER * PJ_THREADS == PJ_CELLS, "cells per thread");
        int loc[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) { loc[k] = cell_end[tid * PER + k]; sum += loc[k]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += v; }
        __shared__ int warp_tot[PJ_THREADS / 32];
        if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
        __syncthreads();
        int base = incl - sum;
        for (int w = 0; w < (tid >> 5); w++) base += warp_tot[w];
#pragma unroll
        for (int k = 0; k < PER; k++) { cell_end[tid * PER + k] = base; base += loc[k]; }
    }
    __syncthreads();
    for (int i = tid; i < nT; i += PJ_THREADS) {       // fill (cursor = cell_end[c], ends up at the cell's end offset)
        const int c = cell_of(sx[i], sy[i]);
        if (c >= 0) cell_idx[atomicAdd(&cell_end[c], 1)] = (unsigned short)i;
    }
    __syncthreads();
    for (int c = tid; c < PJ_CELLS; c += PJ_THREADS) { // ascending feature index inside each cell (AssignFeaturesToGrid pushes in index order)
        const int b = c ? cell_end[c - 1] : 0, e = cell_end[c];
        for (int i = b + 1; i < e; i++) {
            const unsigned short v = cell_idx[i];
            int j = i - 1;
            while (j >= b && cell_idx[j] > v) { cell_idx[j + 1] = cell_idx[j]; j--; }
            cell_idx[j + 1] = v;
        }
    }
    // ---- per-pair scalars of MODE 0: forward / backward motion (:1228-1240)
    if (MODE == 0 && tid == 0) {
        const float* Rc = A.pose + 12 * (size_t)ft; const float* tc = Rc + 9;
        const float* Rl = A.pose + 12 * (size_t)fq; const float* tl = Rl + 9;
        float twc[3];
        for (int i = 0; i < 3; i++)
            twc[i] = -__fadd_rn(__fadd_rn(__fmul_rn(Rc[0 * 3 + i], tc[0]), __fmul_rn(Rc[1 * 3 + i], tc[1])), __fmul_rn(Rc[2 * 3 + i], tc[2]));
        const float tlc2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Rl[6], twc[0]), __fmul_rn(Rl[7], twc[1])), __fmul_rn(Rl[8], twc[2])), tl[2]);
        flag[1] = (tlc2 > A.mb && !A.mono) ? 1 : 0;        // bForward
        flag[2] = (-tlc2 > A.mb && !A.mono) ? 1 : 0;       // bBackward
    }
    __syncthreads();
    const bool bForward = MODE == 0 && flag[1], bBackward = MODE == 0 && flag[2];

    // query q -> search window; Q.live == false: the reference `continue`s before searching
    auto make_query = [&](int q, uint4& d0, uint4& d1) {
        ProjQuery Q;
        Q.live = false;
        if (MODE == 0) {
            if (!A.valid[oQ + q]) return Q;
            const float* Pw = A.world_pos + 3 * (oQ + q);
            const float* R = A.pose + 12 * (size_t)ft; const float* t = R + 9;
            const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[0], Pw[0]), __fmul_rn(R[1], Pw[1])), __fmul_rn(R[2], Pw[2])), t[0]);
            const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3], Pw[0]), __fmul_rn(R[4], Pw[1])), __fmul_rn(R[5], Pw[2])), t[1]);
            const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[6], Pw[0]), __fmul_rn(R[7], Pw[1])), __fmul_rn(R[8], Pw[2])), t[2]);
            const float invzc = __double2float_rn(__ddiv_rn(1.0, (double)zc));       // const float invzc = 1.0 / x3Dc[2]
            if (invzc < 0) return Q;
            const float u = __fadd_rn(__fmul_rn(__fmul_rn(A.fx, xc), invzc), A.cx);
            const float v = __fadd_rn(__fmul_rn(__fmul_rn(A.fy, yc), invzc), A.cy);
            if (u < A.min_x || u > A.max_x || v < A.min_y || v > A.max_y) return Q;
            const int oct = A.kps[oQ + q].octave;
            Q.x = u; Q.y = v; Q.r = __fmul_rn(A.th, A.scale[oct]);
            if (!A.check_level) { Q.minLevel = -1; Q.maxLevel = -1; }
            else if (bForward) { Q.minLevel = oct; Q.maxLevel = -1; }
            else if (bBackward) { Q.minLevel = 0; Q.maxLevel = oct; }
            else { Q.minLevel = oct - 1; Q.maxLevel = oct + 1; }
            Q.gate_ref = __fsub_rn(u, __fmul_rn(A.mbf, invzc));                       // ur = u - mbf * invzc
            Q.gate_r = Q.r;
            const uint4* dq = A.q_desc ? A.q_desc : A.desc;
            d0 = dq[(oQ + q) * 2]; d1 = dq[(oQ + q) * 2 + 1];
        } else {
            const size_t o = (size_t)p * A.M + q;
            if (!A.usable[o]) return Q;
            const int lvl = A.level[o];
            float r = A.view_cos[o] > 0.998f ? 2.5f : 4.0f;                           // RadiusByViewingCos :128-133
            if (A.th != 1.0f) r = __fmul_rn(r, A.th);
            Q.x = A.proj_x[o]; Q.y = A.proj_y[o]; Q.r = __fmul_rn(r, A.scale[lvl]);
            if (A.check_level) { Q.minLevel = lvl - 1; Q.maxLevel = lvl; } else { Q.minLevel = -1; Q.maxLevel = -1; }
            Q.gate_ref = A.proj_xr[o];
            Q.gate_r = Q.r;
            d0 = A.q_desc[o * 2]; d1 = A.q_desc[o * 2 + 1];
        }
        Q.live = true;
        return Q;
    };
    // static tests of GetFeaturesInArea + the stereo gate for target feature i2
    auto passes = [&](const ProjQuery& Q, int i2) {
        const bool bCheckLevels = (Q.minLevel > 0) || (Q.maxLevel >= 0);
        if (bCheckLevels) {
            const int o = soct[i2];
            if (o < Q.minLevel) return false;
            if (Q.maxLevel >= 0 && o > Q.maxLevel) return false;
        }
        if (!(fabsf(__fsub_rn(sx[i2], Q.x)) < Q.r && fabsf(__fsub_rn(sy[i2], Q.y)) < Q.r)) return false;
        return true;
    };
    auto gate = [&](const ProjQuery& Q, int i2) {          // false: rejected by the mvuRight test (:93-97, :1297-1302)
        if (sur[i2] > 0.f) { if (fabsf(__fsub_rn(Q.gate_ref, sur[i2])) > Q.gate_r) return false; }
        return true;
    };

    int rounds = 0;
    for (;;) {
        for (int i = tid; i < nT; i += PJ_THREADS) want[i] = INT_MAX;
        if (tid == 0) flag[0] = 0;
        __syncthreads();
        // ---- phase B: every unresolved query scans its window over the free features and registers itself
        for (int q = tid; q < nQ; q += PJ_THREADS) {
            if (state[q]) continue;
            uint4 d0, d1;
            const ProjQuery Q = make_query(q, d0, d1);
            int x0, x1, y0, y1;
            if (!Q.live || !proj_window(A, invW, invH, Q, x0, x1, y0, y1)) { state[q] = 1; best[q] = -1; continue; }
            int bestDist = 256, bestIdx = -1, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1;
            for (int ix = x0; ix <= x1; ix++)
                for (int iy = y0; iy <= y1; iy++) {
                    const int c = ix * GRID_ROWS + iy;
                    for (int j = c ? cell_end[c - 1] : 0, e = cell_end[c]; j < e; j++) {
                        const int i2 = cell_idx[j];
                        if (!passes(Q, i2)) continue;
                        if (owner[i2] != -1) continue;                                 // taken (finally) or occupied before the call
                        if (!gate(Q, i2)) continue;
                        const int d = hamming256(d0, d1, sdesc[2 * i2], sdesc[2 * i2 + 1]);
                        if (MODE == 0) { if (d <= TH_HIGH) atomicMin(&want[i2], q); }
                        else atomicMin(&want[i2], q);
                        if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestLevel2 = bestLevel; bestLevel = soct[i2]; bestIdx = i2; }
                        else if (d < bestDist2) { bestLevel2 = soct[i2]; bestDist2 = d; }
                    }
                }
            if (MODE == 0) best[q] = bestDist <= TH_HIGH ? bestIdx : -1;
            else {
                const bool accept = bestDist <= TH_HIGH && !(bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(A.nnratio, (float)bestDist2));
                best[q] = accept ? bestIdx : -1 - (bestIdx >= 0 ? 1 : 0);             // -1: no candidate at all, -2: candidates but no match
            }
        }
        __syncthreads();
        // ---- phase C: a query is final when no smaller unresolved query is registered on the candidates that matter to it
        for (int q = tid; q < nQ; q += PJ_THREADS) {
            if (state[q]) continue;
            const int b = best[q];
            bool final_ = false;
            if (MODE == 0) {
                final_ = b < 0 || want[b] == q;
            } else if (b == -1) {
                final_ = true;
            } else {          // the ratio test depends on the second best too: every free candidate must be q's
                uint4 d0, d1;
                const ProjQuery Q = make_query(q, d0, d1);
                int x0, x1, y0, y1;
                proj_window(A, invW, invH, Q, x0, x1, y0, y1);
                final_ = true;
                for (int ix = x0; ix <= x1 && final_; ix++)
                    for (int iy = y0; iy <= y1 && final_; iy++) {
                        const int c = ix * GRID_ROWS + iy;
                        for (int j = c ? cell_end[c - 1] : 0, e = cell_end[c]; j < e; j++) {
                            const int i2 = cell_idx[j];
                            if (!passes(Q, i2) || owner[i2] != -1 || !gate(Q, i2)) continue;
                            if (want[i2] != q) { final_ = false; break; }
                        }
                    }
            }
            if (final_) state[q] = 2;                     // applied below: phase C of other queries still reads `owner` as of this round's start
            else flag[0] = 1;
        }
        __syncthreads();
        for (int q = tid; q < nQ; q += PJ_THREADS)
            if (state[q] == 2) {
                state[q] = 1;
                if (best[q] >= 0) owner[best[q]] = (short)q;   // distinct final queries choose distinct features (each is the minimum on its own)
            }
        __syncthreads();
        rounds++;
        if (!flag[0]) break;
        __syncthreads();
    }

    // ---- rotation consistency (:1317-1347) and output
    __shared__ int n_taken;
    if (tid == 0) n_taken = 0;
    __syncthreads();
    auto rot_bin_of = [&](int i2) {
        float rot = __fsub_rn(A.kps[oQ + owner[i2]].angle, sang[i2]);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
        if (bin == HISTO_LENGTH) bin = 0;
        return bin;
    };
    int mine = 0;
    for (int i = tid; i < nT; i += PJ_THREADS)
        if (owner[i] >= 0) { mine++; if (MODE == 0 && A.check_ori) atomicAdd(&hist[rot_bin_of(i)], 1); }
    if (mine) atomicAdd(&n_taken, mine);
    __syncthreads();
    if (MODE == 0 && A.check_ori) {
        if (tid == 0) {       // ComputeThreeMaxima :1471-1502
            int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
            for (int i = 0; i < HISTO_LENGTH; i++) {
                const int s = hist[i];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
                else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
                else if (s > max3) { max3 = s; ind3 = i; }
            }
            if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
            else if (max3 < 0.1f * (float)max1) ind3 = -1;
            flag[3] = ind1; flag[4] = ind2; flag[5] = ind3;
        }
        __syncthreads();
        int dropped = 0;
        for (int i = tid; i < nT; i += PJ_THREADS)
            if (owner[i] >= 0) {
                const int b = rot_bin_of(i);
                if (b != flag[3] && b != flag[4] && b != flag[5]) { owner[i] = -1; dropped++; }
            }
        if (dropped) atomicSub(&n_taken, dropped);
        __syncthreads();
    }
    int* out = A.matches + (size_t)p * cap;
    for (int i = tid; i < cap; i += PJ_THREADS) out[i] = (i < nT && owner[i] >= 0) ? (int)owner[i] : -1;
    if (tid == 0) { A.nmatches[p] = n_taken; if (A.rounds) A.rounds[p] = rounds; }
}

}  // namespace
