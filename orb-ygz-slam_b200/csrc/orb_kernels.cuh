// Device kernels of the batched ORB extractor (sm_100a).  Each kernel cites the reference lines whose
// arithmetic it reproduces bit-exactly (/root/reference/src/ORBextractor.cc) and the OpenCV primitive it
// stands in for (behaviour pinned to cv2 4.13).
#pragma once
#include "common.cuh"

namespace ygz_b200 {

__host__ __device__ inline int round4(int v) { return (v + 3) & ~3; }

constexpr int MAX_LEVELS = 16;
constexpr int MAX_ROOTS = 16;
constexpr int EDGE_THRESHOLD = 19;   // src/ORBextractor.cc:75
constexpr int HALF_PATCH = 15;       // :74
constexpr int MIN_BORDER = EDGE_THRESHOLD - 3;   // :731

struct ResizeTabX { int ofs; short a0, a1; };          // xofs / ialpha of cv::resize
struct ResizeTabY { short y0, y1, b0, b1; };           // clipped source rows / ibeta

struct LevelDev {
    uint8_t* base;          // [batch][h][pitch]
    uint8_t* blur;          // GaussianBlur(level), same geometry (written by k_blur_levels)
    size_t frame_stride;
    int w, h, pitch;
    float scale;            // mvScaleFactor[l]
    float size;             // (float)(int)(31 * scale), src/ORBextractor.cc:789,797
    // FAST cells of this level (src/ORBextractor.cc:739-762)
    int cell_begin, cell_end;         // range in the global cell list
    int key_off;                      // offset of this level's compacted keys in the per-frame key array
    int key_cap;
    // DistributeOctTree parameters (src/ORBextractor.cc:533-556)
    int n_target;                     // mnFeaturesPerLevel[l]
    int n_roots;                      // nIni
    float hX;
    short root_x0[MAX_ROOTS], root_x1[MAX_ROOTS];
    int root_h;                       // maxY - minY
    int node_cap;                     // max list length
    int sel_off;                      // offset of this level's selected keys in the per-frame sel array
    const ResizeTabX* tx;             // tables to build this level from the previous one
    const ResizeTabY* ty;
};

struct CellDesc {          // one cv::FAST call of ComputeKeyPointsOctTree (:756-770)
    short level;
    short x0, y0;          // iniX, iniY (level coordinates of the cell image origin)
    short cw, ch;          // cell image size (maxX-iniX, maxY-iniY)
    short offx, offy;      // j*wCell, i*hCell  (:774-775)
    short cap;             // max NMS survivors = ceil(dw/2)*ceil(dh/2)
    int out_off;           // slot offset in the per-frame candidate pool
};

struct OrbParams {
    int nlevels;
    int ini_th, min_th;
    int n_cells;
    int pool_size;         // u32 slots per frame in the candidate pool
    int key_total;         // compacted keys capacity per frame (sum over levels)
    int sel_total;         // selected-key capacity per frame (sum of node_cap)
    int out_cap;           // keypoints per frame capacity of the output arrays
    int umax[16];
    int vlim[16];          // vlim[|u|] = largest |v| with |u| <= umax[|v|] (umax is non-increasing: rows of a column are contiguous)
    LevelDev lv[MAX_LEVELS];
};

// packed candidate: x[0:12) y[12:24) score[24:32), coordinates relative to (16,16)
__device__ __forceinline__ uint32_t pack_cand(int x, int y, int s) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)s << 24); }
__device__ __forceinline__ int cand_x(uint32_t c) { return c & 0xFFF; }
__device__ __forceinline__ int cand_y(uint32_t c) { return (c >> 12) & 0xFFF; }
__device__ __forceinline__ int cand_s(uint32_t c) { return c >> 24; }

// ------------------------------------------------------------------------------------------------
// K1  pyramid level from the previous level: cv::resize(INTER_LINEAR) on CV_8UC1, 11-bit fixed point
//     (src/ORBextractor.cc:1139).  4 output pixels per thread, one 32-bit store.
//     (Measured alternative: horizontal pass of the whole staged window into shared int32 rows, then a vertical
//     pass, 256 x 16 tiles — 0.372 vs 0.381 ms for the 7 levels, not worth a second kernel.)
// ------------------------------------------------------------------------------------------------
constexpr int PYR_ROWS = 8;          // output rows per CTA / thread
constexpr int PYR_THREADS = 64;      // 4 output columns per thread -> 256 columns per CTA
constexpr int PYR_SROWS = 12;        // source rows staged per CTA (8 output rows at scale 1.2 span <= 11 source rows)
constexpr int PYR_SPITCH = 352;      // staged bytes per source row (256 * 1.2 + slack, multiple of 16)

// The CTA stages the source window of its 256 x 8 output tile in shared memory with coalesced 16-byte streaming loads
// (every source byte is read from HBM/L2 once per CTA), then each thread produces a 4 x 8 block: horizontal taps from
// shared memory with the column tables (xofs / ialpha of cv::resize) held in registers, source rows reused between
// consecutive output rows, one aligned 32-bit store per row.  Falls back to direct global reads when the window does
// not fit the staging buffer (level-to-level scale outside ~[1.0, 1.35]).
__global__ void __launch_bounds__(PYR_THREADS) k_pyr_resize(const uint8_t* __restrict__ src, int spitch, size_t sframe,
                                                            int sw, int sh, uint8_t* __restrict__ dst, int dpitch, size_t dframe,
                                                            int dw, int dh, const ResizeTabX* __restrict__ tx,
                                                            const ResizeTabY* __restrict__ ty) {
    __shared__ __align__(16) uint8_t stage[PYR_SROWS * PYR_SPITCH];
    const int xg0 = blockIdx.x * PYR_THREADS * 4;              // first output column of the CTA
    const int x4 = xg0 + threadIdx.x * 4;
    const int dy0 = blockIdx.y * PYR_ROWS;
    const int dy1 = min(dy0 + PYR_ROWS, dh) - 1;
    const uint8_t* S = src + (size_t)blockIdx.z * sframe;
    uint8_t* D = dst + (size_t)blockIdx.z * dframe;
    // source window of the tile
    const int xlast = min(xg0 + PYR_THREADS * 4, dw) - 1;
    const int sx_lo = __ldg(reinterpret_cast<const int2*>(tx) + xg0).x & ~15;
    const int sx_hi = min(__ldg(reinterpret_cast<const int2*>(tx) + xlast).x + 1, sw - 1);
    // the row tables of the tile are fetched up front (independent loads) instead of one dependent load per row
    ResizeTabY yt[PYR_ROWS];
#pragma unroll
    for (int r = 0; r < PYR_ROWS; r++) {
        const int2 v = __ldg(reinterpret_cast<const int2*>(ty) + min(dy0 + r, dh - 1));
        yt[r].y0 = (short)(v.x & 0xFFFF); yt[r].y1 = (short)(v.x >> 16);
        yt[r].b0 = (short)(v.y & 0xFFFF); yt[r].b1 = (short)(v.y >> 16);
    }
    const int sy_lo = yt[0].y0, sy_hi = ty[dy1].y1;
    const int nrows = sy_hi - sy_lo + 1;
    const int nvec = (sx_hi - sx_lo) / 16 + 1;                 // 16-byte vectors per staged row (reads stay inside the pitch)
    const bool staged = (nrows <= PYR_SROWS) && (nvec * 16 <= PYR_SPITCH) && (sx_lo + nvec * 16 <= spitch);
    if (staged) {
        for (int i = threadIdx.x; i < nrows * nvec; i += PYR_THREADS) {
            const int r = i / nvec, c = i - r * nvec;
            const uint4 v = ld_nc_u4(S + (size_t)(sy_lo + r) * spitch + sx_lo + c * 16);
            *reinterpret_cast<uint4*>(stage + r * PYR_SPITCH + c * 16) = v;
        }
    }
    __syncthreads();
    if (x4 >= dw) return;
    int ofs0[4], ofs1[4], a0[4], a1[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int dx = min(x4 + i, dw - 1);
        const int2 tv = __ldg(reinterpret_cast<const int2*>(tx) + dx);
        ofs0[i] = tv.x;
        ofs1[i] = min(tv.x + 1, sw - 1);
        a0[i] = (short)(tv.y & 0xFFFF);
        a1[i] = (short)(tv.y >> 16);
    }
    int hA[4], hB[4];                // horizontally interpolated source rows yA (= y0 of the current row) and yB (= y1)
    int yA = -1, yB = -1;
    auto hrow = [&](int y, int* h) {
        if (staged) {
            const uint8_t* R = stage + (y - sy_lo) * PYR_SPITCH - sx_lo;
#pragma unroll
            for (int i = 0; i < 4; i++) h[i] = R[ofs0[i]] * a0[i] + R[ofs1[i]] * a1[i];
        } else {
            const uint8_t* R = S + (size_t)y * spitch;
#pragma unroll
            for (int i = 0; i < 4; i++) h[i] = R[ofs0[i]] * a0[i] + R[ofs1[i]] * a1[i];
        }
    };
#pragma unroll
    for (int r = 0; r < PYR_ROWS; r++) {
        const int dy = dy0 + r;
        if (dy > dy1) break;
        const ResizeTabY yy = yt[r];
        if (yy.y0 == yB) {           // the usual case when downscaling: this row's upper source row is the previous lower one
#pragma unroll
            for (int i = 0; i < 4; i++) hA[i] = hB[i];
            yA = yB;
        } else if (yy.y0 != yA) { hrow(yy.y0, hA); yA = yy.y0; }
        if (yy.y1 == yA) {
#pragma unroll
            for (int i = 0; i < 4; i++) hB[i] = hA[i];
            yB = yA;
        } else if (yy.y1 != yB) { hrow(yy.y1, hB); yB = yy.y1; }
        uint32_t outw = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int v = (((yy.b0 * (hA[i] >> 4)) >> 16) + ((yy.b1 * (hB[i] >> 4)) >> 16) + 2) >> 2;
            outw |= (uint32_t)(v & 0xFF) << (8 * i);
        }
        *reinterpret_cast<uint32_t*>(D + (size_t)dy * dpitch + x4) = outw;   // pitch is a multiple of 16; pad bytes are scratch
    }
}

// ------------------------------------------------------------------------------------------------
// K2  FAST-9/16 per ~30x30 cell with the 20 -> 7 threshold fallback: the cv::FAST(cell, iniTh, true) /
//     cv::FAST(cell, minTh, true) pair of src/ORBextractor.cc:765-770 as ONE dense pass
//     (score map S = best-1, NMS masked to the cell's detection window).
//     One CTA per (cell, frame).
// ------------------------------------------------------------------------------------------------
constexpr int FAST_WARPS = 8;                       // one warp per cell, FAST_WARPS cells per CTA, no block barrier
constexpr int FAST_THREADS = FAST_WARPS * 32;

// FAST-9 score of one pixel, both polarities at once in the two 16-bit halves of a register (VIMNMX3.S16x2):
//   low  half: c - r_k + 256  (darker ring),   high half: r_k - c + 255 = 511 - low  (brighter ring), all in [0, 511],
//   i.e. E_k = 0x01FF0000 - low * 0xFFFF = r_k * 0xFFFF + K2 with K2 = 0x01FF0000 - (c + 256) * 0xFFFF: one IMAD per ring pixel.
// min over 9 contiguous = min3 of three min3 (32 instructions for all 16 arcs), then a max3 tree.
// No negated operand feeds a min/max (ptxas 12.9 drops the negation when fusing into VIMNMX3).
__device__ __forceinline__ int fast9_best_packed(int c, const int* r) {
    unsigned K2 = 0x01FF0000u - (unsigned)(c + 256) * 0xFFFFu;
    asm volatile("" : "+r"(K2));        // opaque: otherwise ptxas re-associates to (r_k - c) * 0xFFFF + const, TWO IMADs per ring pixel (r2 profile)
    unsigned E[16], e3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) E[k] = (unsigned)r[k] * 0xFFFFu + K2;
#pragma unroll
    for (int k = 0; k < 16; k++) e3[k] = __vimin3_s16x2(E[k], E[(k + 1) & 15], E[(k + 2) & 15]);
    unsigned m[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m[k] = __vimin3_s16x2(e3[k], e3[(k + 3) & 15], e3[(k + 6) & 15]);
    const unsigned a0 = __vimax3_s16x2(m[0], m[1], m[2]), a1 = __vimax3_s16x2(m[3], m[4], m[5]);
    const unsigned a2 = __vimax3_s16x2(m[6], m[7], m[8]), a3 = __vimax3_s16x2(m[9], m[10], m[11]);
    const unsigned a4 = __vimax3_s16x2(m[12], m[13], m[14]);
    const unsigned b0 = __vimax3_s16x2(a0, a1, a2), b1 = __vimax3_s16x2(a3, a4, m[15]);
    const unsigned bb = __vmaxs2(b0, b1);
    const int dark = (int)(bb & 0xFFFFu) - 256, bright = (int)(bb >> 16) - 255;
    return max(dark, bright);
}

// shared-memory bytes one warp needs for a cell image of cw x ch (x0 & 3 = worst-case 3 extra columns)
__host__ __device__ inline int fast_warp_smem(int cw, int ch) {
    const int dw = cw - 6, dh = ch - 6;
    const int tp = round4(cw + 3);
    return round4((ch + 1) * tp) + round4((dw + 2) * (dh + 2)) + round4(dw * dh) + round4(2 * dw * dh);
}

// (A compile-time tile pitch for the common <= 37 px cells was measured: same 1.09 ms, so only this flavour is kept.)
// `ids` (optional): the cells this launch covers — the generic kernel handles the cells k_fast_cells_v2 does not take
// (detection window wider than 32 columns, cell image taller than its TMA box, or a level without a tensor map).
__global__ void __launch_bounds__(FAST_THREADS) k_fast_cells(OrbParams P, const CellDesc* __restrict__ cells,
                                                             const int* __restrict__ ids, int n_ids,
                                                             uint32_t* __restrict__ pool,
                                                             uint16_t* __restrict__ cell_count, int warp_smem) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * FAST_WARPS + warp;
    if (slot >= n_ids) return;
    const int cell = ids ? ids[slot] : slot;
    const int frame = blockIdx.y;
    const CellDesc cd = cells[cell];
    const LevelDev& L = P.lv[cd.level];
    const int cw = cd.cw, ch = cd.ch;
    const int dw = cw - 6, dh = ch - 6;               // detection window [3, cw-3) x [3, ch-3)
    uint16_t* cnt_out = cell_count + (size_t)frame * P.n_cells + cell;
    if (dw <= 0 || dh <= 0) { if (lane == 0) *cnt_out = 0; return; }
    const int npx = dw * dh;
    const int sw2 = dw + 2;
    const int xa = cd.x0 & ~3, shift = cd.x0 - xa;    // tile is loaded as aligned 32-bit words
    const int TP = round4(cw + shift);
    uint8_t* tile = smem_raw + (size_t)warp * warp_smem;               // (ch + 1) x TP (one spare row: the packed pre-test reads row pairs)
    uint8_t* sc = tile + round4((ch + 1) * TP);                        // (dh+2) x sw2 scores, zero border
    uint8_t* flag = sc + round4(sw2 * (dh + 2));                       // per item: score if NMS survivor else 0
    uint16_t* items = reinterpret_cast<uint16_t*>(flag + round4(npx)); // pre-test survivors: x | y<<6

    {   // ---- tile load (coalesced words) + score tile clear
        const int wpr = TP >> 2;
        const uint8_t* g0 = L.base + (size_t)frame * L.frame_stride + (size_t)cd.y0 * L.pitch + xa;
        // (row, word) of this lane's element advanced incrementally: wpr is a run-time 9..18, a division per element
        // cost 13 % of the kernel's instructions in the round-1 profile
        int r = 0, c = lane;
        while (c >= wpr) { c -= wpr; r++; }
        while (r < ch) {
            reinterpret_cast<uint32_t*>(tile + r * TP)[c] = __ldg(reinterpret_cast<const uint32_t*>(g0 + (size_t)r * L.pitch) + c);
            c += 32;
            while (c >= wpr) { c -= wpr; r++; }
        }
        const int nsw = round4(sw2 * (dh + 2)) >> 2;
        for (int i = lane; i < nsw; i += 32) reinterpret_cast<uint32_t*>(sc)[i] = 0u;
    }
    __syncwarp();

    // ring offsets in the tile (cv::FAST offsets16)
    const int o0 = 3 * TP, o1 = 3 * TP + 1, o2 = 2 * TP + 2, o3 = TP + 3, o4 = 3, o5 = -TP + 3, o6 = -2 * TP + 2,
              o7 = -3 * TP + 1, o8 = -3 * TP, o9 = -3 * TP - 1, o10 = -2 * TP - 2, o11 = -TP - 3, o12 = -3,
              o13 = TP - 3, o14 = 2 * TP - 2, o15 = 3 * TP - 1;
    const uint8_t* t0 = tile + 3 * TP + shift + 3;     // pixel (0,0) of the detection window

    // The reference runs cv::FAST at iniTh and, only when that returns nothing, again at minTh.  The score of a pixel
    // (largest threshold at which it is still a corner) does not depend on the call's threshold, and a pixel that is not
    // a corner at t can never suppress one that is (its score is < t), so pass 0 below restricted to the iniTh
    // pre-test survivors equals the iniTh call; pass 1 (rare: flat cells) equals the minTh call.
    int total = 0;
    for (int pass = 0; pass < 2; pass++) {
        const int tmin = pass ? P.min_th : P.ini_th;          // threshold of this cv::FAST call
        // ---- phase A: opposite-pair pre-test at the call's threshold (necessary for 9 contiguous): a brighter arc
        //      needs max(a,b) > c + t for every opposite pair (a,b), a darker arc min(a,b) < c - t.
        //      Straight-line (no early exits: a warp almost always has a surviving lane); ordered compaction keeps
        //      raster order.
        total = 0;
        if (dw <= 32) {
            // two detection rows per step, packed in the 16-bit halves of a register (VIMNMX.U16x2): lo = row y, hi = row y+1,
            // on the RAW ring values.  With K3 = 0x7FFF - t per half, bit 15 / 31 of (bmin + K3 - c) is "bmin > c + t" and
            // of (c + K3 - dmax) is "dmax < c - t"; every half stays inside [0, 0xFFFF], so plain 32-bit adds never carry
            // between the halves.  Lanes >= dw compute on the neighbouring bytes and are masked afterwards.
            const unsigned K3 = (0x7FFFu - (unsigned)tmin) * 0x00010001u;
            const unsigned lane_mask = lane < dw ? 0x80008000u : 0u;
            for (int y = 0; y < dh; y += 2) {
                const uint8_t* p = t0 + y * TP + lane;
#define LD2(o) ((unsigned)p[(o)] | ((unsigned)p[(o) + TP] << 16))
                const unsigned c2 = LD2(0);
                const unsigned A0 = LD2(o0), A8 = LD2(o8), A4 = LD2(o4), A12 = LD2(o12);
                const unsigned A2 = LD2(o2), A10 = LD2(o10), A6 = LD2(o6), A14 = LD2(o14);
#undef LD2
                const unsigned bmin = __vminu2(__vminu2(__vmaxu2(A0, A8), __vmaxu2(A4, A12)), __vminu2(__vmaxu2(A2, A10), __vmaxu2(A6, A14)));
                const unsigned dmax = __vmaxu2(__vmaxu2(__vminu2(A0, A8), __vminu2(A4, A12)), __vmaxu2(__vminu2(A2, A10), __vminu2(A6, A14)));
                unsigned z = ((bmin + K3 - c2) | (c2 + K3 - dmax)) & lane_mask;
                if (y + 1 >= dh) z &= 0xFFFFu;
                const int p0 = z & 0x8000u, p1 = z >> 31;
                const unsigned bal0 = __ballot_sync(0xffffffffu, p0);
                const unsigned bal1 = __ballot_sync(0xffffffffu, p1);
                if (p0) items[total + __popc(bal0 & lanemask_lt())] = (uint16_t)(lane | (y << 6));
                total += __popc(bal0);
                if (p1) items[total + __popc(bal1 & lanemask_lt())] = (uint16_t)(lane | ((y + 1) << 6));
                total += __popc(bal1);
            }
        } else {
            for (int y = 0; y < dh; y++) {
                for (int xb = 0; xb < dw; xb += 32) {
                    const int x = xb + lane;
                    bool pass_px = false;
                    if (x < dw) {
                        const uint8_t* p = t0 + y * TP + x;
                        const int c = p[0];
                        const int d0 = p[o0] - c, d8 = p[o8] - c, d4 = p[o4] - c, d12 = p[o12] - c;
                        const int d2 = p[o2] - c, d10 = p[o10] - c, d6 = p[o6] - c, d14 = p[o14] - c;
                        const int bmin = min(min(max(d0, d8), max(d4, d12)), min(max(d2, d10), max(d6, d14)));
                        const int dmax = max(max(min(d0, d8), min(d4, d12)), max(min(d2, d10), min(d6, d14)));
                        pass_px = (bmin > tmin) | (dmax + tmin < 0);
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, pass_px);
                    if (pass_px) items[total + __popc(bal & lanemask_lt())] = (uint16_t)(x | (y << 6));
                    total += __popc(bal);
                }
            }
        }
        __syncwarp();

        // ---- phase B: FAST-9 score of the survivors (both polarities packed in the 16-bit halves, fast9_best_packed).
        //      Measured alternative (two single-polarity items per lane, one per half): 1.55 ms vs 1.45 ms — the extra packing
        //      and the halved lane utilisation of the last iteration cost more than the shared min/max network saves.
        for (int g = lane; g < total; g += 32) {
            const int it = items[g];
            const int x = it & 63, y = it >> 6;
            const uint8_t* p = t0 + y * TP + x;
            const int r[16] = {p[o0], p[o1], p[o2], p[o3], p[o4], p[o5], p[o6], p[o7],
                               p[o8], p[o9], p[o10], p[o11], p[o12], p[o13], p[o14], p[o15]};
            const int best = fast9_best_packed(p[0], r);
            if (best > tmin) sc[(y + 1) * sw2 + x + 1] = (uint8_t)(best - 1);   // corner at this threshold; response = best-1
        }
        __syncwarp();

        // ---- phase C: 3x3 non-max suppression inside the detection window (strict >, zeros outside)
        int any_keep = 0;
        for (int g = lane; g < total; g += 32) {
            const int it = items[g];
            const int x = it & 63, y = it >> 6;
            const uint8_t* q = sc + (y + 1) * sw2 + x + 1;
            const int s = q[0];
            int keep = 0;
            if (s > 0 && s >= tmin) {
                keep = (s > q[-1]) & (s > q[1]) & (s > q[-sw2 - 1]) & (s > q[-sw2]) & (s > q[-sw2 + 1]) &
                       (s > q[sw2 - 1]) & (s > q[sw2]) & (s > q[sw2 + 1]);
            }
            flag[g] = keep ? (uint8_t)s : 0;
            any_keep |= keep;
        }
        __syncwarp();
        // vKeysCell.empty() after the iniTh call decides the fallback to minTh (src/ORBextractor.cc:765-770)
        if (__any_sync(0xffffffffu, any_keep)) break;
    }

    // ---- phase D: ordered emission (raster order inside the cell), coordinates relative to (16,16)
    uint32_t* out = pool + (size_t)frame * P.pool_size + cd.out_off;
    int base = 0;
    for (int g0 = 0; g0 < total; g0 += 32) {
        const int g = g0 + lane;
        int s = 0, it = 0;
        if (g < total) { s = flag[g]; it = items[g]; }
        const bool keep = s > 0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = base + __popc(bal & lanemask_lt());
            if (pos < cd.cap) out[pos] = pack_cand((it & 63) + 3 + cd.offx, (it >> 6) + 3 + cd.offy, s);
        }
        base += __popc(bal);
    }
    if (lane == 0) *cnt_out = (uint16_t)min(base, (int)cd.cap);
}

// ------------------------------------------------------------------------------------------------
// K3  DistributeOctTree (src/ORBextractor.cc:533-723, DivideNode :479-531) as a data-parallel
//     quadtree refinement.  One CTA per (frame, level).  Keys never move: each key carries the list
//     position of its node; one pass = count keys per child (shared-memory atomics), choose the parents
//     to split in the reference's processing order (list order, or largest-first with the oracle's
//     creation-order tie-break), build the next list generation, relabel keys.
//     List order is reproduced exactly: new list = reverse(creation order of the children) ++ untouched
//     nodes in their old order (push_front + erase semantics of :605-640 and :662-692).
// ------------------------------------------------------------------------------------------------
constexpr int OCT_THREADS = 256;     // measured: 512 threads 0.357 ms vs 0.302 ms (fewer resident CTAs to hide the dependent global loads)

// in-place inclusive scan of a[0..n) in shared memory by the whole CTA; returns the total
__device__ inline int block_scan_inclusive(int* a, int n, int* tmp /* >= OCT_THREADS/32 + 1 ints */) {
    const int T = blockDim.x;
    const int ipt = (n + T - 1) / T;
    const int b = threadIdx.x * ipt, e = min(n, b + ipt);
    int sum = 0;
    for (int i = b; i < e; i++) { sum += a[i]; a[i] = sum; }
    // exclusive scan of per-thread sums
    int v = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane == 31) tmp[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int w = lane < (T >> 5) ? tmp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        if (lane < (T >> 5)) tmp[lane] = w;        // inclusive warp totals
    }
    __syncthreads();
    const int offset = (v - sum) + (warp ? tmp[warp - 1] : 0);
    for (int i = b; i < e; i++) a[i] += offset;
    const int total = tmp[(T >> 5) - 1];
    __syncthreads();
    return total;
}

struct OctSmem {      // all arrays sized node_cap (NC); carved from dynamic shared memory
    short *x0[2], *x1[2], *y0[2], *y1[2];
    int* cnt[2];
    int* qcnt;        // NC*4
    short* childpos;  // NC*4
    short* newpos;    // NC
    short* ord;       // NC
    unsigned char* processed;   // NC
    int* scanA;       // NC
    int* scanB;       // NC
    unsigned* best;   // NC
};

__host__ __device__ inline size_t oct_smem_bytes(int NC) {
    // 2 gens * (4 shorts + 1 int) + qcnt 16 + childpos 8 + newpos 2 + ord 2 + processed 1(+pad) + scanA 4 + scanB 4 + best 4
    return (size_t)NC * (2 * (8 + 4) + 16 + 8 + 2 + 2 + 4 + 4 + 4 + 4) + 64;
}

__device__ __forceinline__ int quadrant_of(int x, int y, int x0, int x1, int y0, int y1) {
    const int hx = x0 + ((x1 - x0 + 1) >> 1);      // ceil((UR.x-UL.x)/2), :480
    const int hy = y0 + ((y1 - y0 + 1) >> 1);      // :481
    return (x < hx ? 0 : 1) + (y < hy ? 0 : 2);    // n1 UL, n2 UR, n3 BL, n4 BR (:511-519)
}

template <int OCT_U>       // keys per thread in flight in the vote / relabel loops
__global__ void __launch_bounds__(OCT_THREADS, 4) k_octree(OrbParams P, const CellDesc* __restrict__ cells,
                                                        const uint32_t* __restrict__ pool,
                                                        const uint16_t* __restrict__ cell_count,
                                                        uint32_t* __restrict__ keys_all, uint16_t* __restrict__ keynode_all,
                                                        int* __restrict__ nkeys_all, uint32_t* __restrict__ sel_all,
                                                        int* __restrict__ nsel_all, int node_smem_bytes, int key_cache) {
    extern __shared__ __align__(16) uint8_t smem_oct[];
    // grid = (frames, levels): the heavy level-0 problems of ALL frames are scheduled first, the cheap coarse levels fill the tail
    const int level = blockIdx.y, frame = blockIdx.x;
    const LevelDev& L = P.lv[level];
    const int NC = L.node_cap;
    const int N = L.n_target;
    const int tid = threadIdx.x;
    __shared__ int s_tmp[OCT_THREADS / 32 + 1];
    __shared__ int s_misc[8];

    // ---- carve shared memory
    OctSmem S;
    {
        uint8_t* p = smem_oct;
        S.cnt[0] = (int*)p; p += 4 * NC;
        S.cnt[1] = (int*)p; p += 4 * NC;
        S.qcnt = (int*)p; p += 16 * NC;
        S.scanA = (int*)p; p += 4 * NC;
        S.scanB = (int*)p; p += 4 * NC;
        S.best = (unsigned*)p; p += 4 * NC;
        for (int g = 0; g < 2; g++) {
            S.x0[g] = (short*)p; p += 2 * NC; S.x1[g] = (short*)p; p += 2 * NC;
            S.y0[g] = (short*)p; p += 2 * NC; S.y1[g] = (short*)p; p += 2 * NC;
        }
        S.childpos = (short*)p; p += 8 * NC;
        S.newpos = (short*)p; p += 2 * NC;
        S.ord = (short*)p; p += 2 * NC;
        S.processed = (unsigned char*)p;
    }

    uint32_t* keys = keys_all + (size_t)frame * P.key_total + L.key_off;
    uint16_t* keynode = keynode_all + (size_t)frame * P.key_total + L.key_off;
    // shared-memory cache of the key list (packed candidates + node labels): the refinement passes re-read every key
    // ~20 times with no reuse distance an L1 would catch; levels whose key count exceeds the cache stay in global memory
    uint32_t* skeys = reinterpret_cast<uint32_t*>(smem_oct + node_smem_bytes);
    uint16_t* snode = reinterpret_cast<uint16_t*>(skeys + key_cache);
    uint32_t* sel = sel_all + (size_t)frame * P.sel_total + L.sel_off;

    // ---- step 0: compact this level's per-cell candidate slots into one ordered key list
    const int ncell = L.cell_end - L.cell_begin;
    int n = 0;
    {
        // scan of cell counts in chunks of NC (scanA as buffer)
        int done = 0;
        int running = 0;
        while (done < ncell) {
            const int m = min(NC, ncell - done);
            for (int i = tid; i < m; i += OCT_THREADS)
                S.scanA[i] = cell_count[(size_t)frame * P.n_cells + L.cell_begin + done + i];
            __syncthreads();
            const int tot = block_scan_inclusive(S.scanA, m, s_tmp);
            // one warp per cell copies its slots
            for (int c = (tid >> 5); c < m; c += OCT_THREADS / 32) {
                const int cntc = S.scanA[c] - (c ? S.scanA[c - 1] : 0);
                const int dst = running + S.scanA[c] - cntc;
                const uint32_t* src = pool + (size_t)frame * P.pool_size + cells[L.cell_begin + done + c].out_off;
                for (int i = (tid & 31); i < cntc; i += 32) keys[dst + i] = src[i];
            }
            running += tot;
            done += m;
            __syncthreads();
        }
        n = running;
    }
    if (tid == 0) nkeys_all[(size_t)frame * P.nlevels + level] = n;
    __syncthreads();     // keys[] written by this CTA are visible to it after the barrier
    if (n <= key_cache) {
        for (int k = tid; k < n; k += OCT_THREADS) skeys[k] = keys[k];
        keys = skeys;
        keynode = snode;
        __syncthreads();
    }

    int cur = 0;
    int Lsz = 0;
    // ---- roots (:537-574)
    if (L.n_roots < 1 || n == 0) {
        if (tid == 0) nsel_all[(size_t)frame * P.nlevels + level] = 0;
        return;
    }
    for (int i = tid; i < L.n_roots; i += OCT_THREADS) S.scanA[i] = 0;
    __syncthreads();
    for (int k = tid; k < n; k += OCT_THREADS) {
        const int x = cand_x(keys[k]);
        int r = (int)__fdiv_rn((float)x, L.hX);          // vpIniNodes[kp.pt.x / hX], :561
        r = min(r, L.n_roots - 1);
        keynode[k] = (uint16_t)r;                        // provisional: root index
        atomicAdd(&S.scanA[r], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int pos = 0;
        for (int r = 0; r < L.n_roots; r++) {
            const int c = S.scanA[r];
            if (c > 0) {                                  // empty roots are erased (:570-571)
                S.x0[0][pos] = L.root_x0[r]; S.x1[0][pos] = L.root_x1[r];
                S.y0[0][pos] = 0; S.y1[0][pos] = (short)L.root_h;
                S.cnt[0][pos] = c;
                S.scanB[r] = pos;
                pos++;
            }
        }
        s_misc[0] = pos;
    }
    __syncthreads();
    Lsz = s_misc[0];
    for (int k = tid; k < n; k += OCT_THREADS) keynode[k] = (uint16_t)S.scanB[keynode[k]];
    __syncthreads();

    // ---- refinement passes
    bool careful = false;
    for (int guard = 0; guard < 64; guard++) {
        const short* X0 = S.x0[cur]; const short* X1 = S.x1[cur];
        const short* Y0 = S.y0[cur]; const short* Y1 = S.y1[cur];
        const int* CNT = S.cnt[cur];
        const int nxt = cur ^ 1;
        // P1: clear child counters
        for (int i = tid; i < Lsz * 4; i += OCT_THREADS) S.qcnt[i] = 0;
        for (int i = tid; i < Lsz; i += OCT_THREADS) S.processed[i] = 0;
        __syncthreads();
        // P2: keys of expandable nodes vote for their quadrant
        // (node label and packed key are fetched together, OCT_U keys per thread in flight: the loop is bound by the latency of these
        //  two global loads, and fetching the key only after the label's node test doubled the dependent chain)
        for (int k0 = tid; k0 < n; k0 += OCT_THREADS * OCT_U) {
            int nid[OCT_U]; uint32_t c[OCT_U];
#pragma unroll
            for (int u = 0; u < OCT_U; u++) {
                const int k = k0 + u * OCT_THREADS;
                if (k < n) { nid[u] = keynode[k]; c[u] = keys[k]; }
            }
#pragma unroll
            for (int u = 0; u < OCT_U; u++) {
                const int k = k0 + u * OCT_THREADS;
                if (k < n && CNT[nid[u]] > 1) {
                    const int q = quadrant_of(cand_x(c[u]), cand_y(c[u]), X0[nid[u]], X1[nid[u]], Y0[nid[u]], Y1[nid[u]]);
                    atomicAdd(&S.qcnt[nid[u] * 4 + q], 1);
                }
            }
        }
        __syncthreads();
        // P3: processing order of the expandable nodes
        int E;
        if (!careful) {
            for (int i = tid; i < Lsz; i += OCT_THREADS) S.scanA[i] = CNT[i] > 1;
            __syncthreads();
            E = block_scan_inclusive(S.scanA, Lsz, s_tmp);
            for (int i = tid; i < Lsz; i += OCT_THREADS)
                if (CNT[i] > 1) S.ord[S.scanA[i] - 1] = (short)i;      // list order (:594-641)
        } else {
            // sort(vPrevSizeAndPointerToNode) ascending by (count, creation), visited from the back (:656-657):
            // larger count first; equal counts: later-created first == earlier in the list
            for (int i = tid; i < Lsz; i += OCT_THREADS) S.scanA[i] = CNT[i] > 1;
            __syncthreads();
            E = block_scan_inclusive(S.scanA, Lsz, s_tmp);
            for (int i = tid; i < Lsz; i += OCT_THREADS) {
                const int ci = CNT[i];
                if (ci > 1) {
                    int rank = 0;
                    for (int j = 0; j < Lsz; j++) {
                        const int cj = CNT[j];
                        rank += (cj > 1) && (cj > ci || (cj == ci && j < i));
                    }
                    S.ord[rank] = (short)i;
                }
            }
        }
        __syncthreads();
        // P4: children per parent in processing order; prefix sums of created nodes and of list growth
        for (int r = tid; r < E; r += OCT_THREADS) {
            const int i = S.ord[r];
            const int cc = (S.qcnt[i * 4] > 0) + (S.qcnt[i * 4 + 1] > 0) + (S.qcnt[i * 4 + 2] > 0) + (S.qcnt[i * 4 + 3] > 0);
            S.scanA[r] = cc;
        }
        __syncthreads();
        block_scan_inclusive(S.scanA, E, s_tmp);       // scanA[r] = children created by parents 0..r
        // cutoff: careful mode stops as soon as the list reaches N (:693-694)
        if (tid == 0) s_misc[1] = E;
        __syncthreads();
        if (careful) {
            for (int r = tid; r < E; r += OCT_THREADS) {
                const int size_after = Lsz + S.scanA[r] - (r + 1);
                const int size_before = r ? Lsz + S.scanA[r - 1] - r : Lsz;
                if (size_after >= N && size_before < N) atomicMin(&s_misc[1], r + 1);
            }
            __syncthreads();
        }
        const int K = s_misc[1];
        const int C = K ? S.scanA[K - 1] : 0;
        const int newL = Lsz + C - K;
        __syncthreads();
        if (tid == 0) s_misc[2] = 0;     // nToExpand
        // P5: create the children (new list front, reversed creation order) ...
        for (int r = tid; r < K; r += OCT_THREADS) {
            const int i = S.ord[r];
            S.processed[i] = 1;
            const int x0 = X0[i], x1 = X1[i], y0 = Y0[i], y1 = Y1[i];
            const int hx = x0 + ((x1 - x0 + 1) >> 1), hy = y0 + ((y1 - y0 + 1) >> 1);
            int t = r ? S.scanA[r - 1] : 0;
            int nexp = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int c = S.qcnt[i * 4 + q];
                if (c > 0) {
                    const int pos = C - 1 - t;
                    t++;
                    S.x0[nxt][pos] = (short)((q & 1) ? hx : x0);
                    S.x1[nxt][pos] = (short)((q & 1) ? x1 : hx);
                    S.y0[nxt][pos] = (short)((q & 2) ? hy : y0);
                    S.y1[nxt][pos] = (short)((q & 2) ? y1 : hy);
                    S.cnt[nxt][pos] = c;
                    S.childpos[i * 4 + q] = (short)pos;
                    nexp += c > 1;
                }
            }
            if (nexp) atomicAdd(&s_misc[2], nexp);
        }
        __syncthreads();
        // ... followed by the untouched nodes in their old order
        for (int i = tid; i < Lsz; i += OCT_THREADS) S.scanB[i] = S.processed[i] ? 0 : 1;
        __syncthreads();
        block_scan_inclusive(S.scanB, Lsz, s_tmp);
        for (int i = tid; i < Lsz; i += OCT_THREADS) {
            if (!S.processed[i]) {
                const int pos = C + S.scanB[i] - 1;
                S.newpos[i] = (short)pos;
                S.x0[nxt][pos] = X0[i]; S.x1[nxt][pos] = X1[i];
                S.y0[nxt][pos] = Y0[i]; S.y1[nxt][pos] = Y1[i];
                S.cnt[nxt][pos] = CNT[i];
            }
        }
        __syncthreads();
        // P6: relabel keys
        for (int k0 = tid; k0 < n; k0 += OCT_THREADS * OCT_U) {
            int nid[OCT_U]; uint32_t c[OCT_U];
#pragma unroll
            for (int u = 0; u < OCT_U; u++) {
                const int k = k0 + u * OCT_THREADS;
                if (k < n) { nid[u] = keynode[k]; c[u] = keys[k]; }
            }
#pragma unroll
            for (int u = 0; u < OCT_U; u++) {
                const int k = k0 + u * OCT_THREADS;
                if (k < n) {
                    if (S.processed[nid[u]]) {
                        const int q = quadrant_of(cand_x(c[u]), cand_y(c[u]), X0[nid[u]], X1[nid[u]], Y0[nid[u]], Y1[nid[u]]);
                        keynode[k] = (uint16_t)S.childpos[nid[u] * 4 + q];
                    } else {
                        keynode[k] = (uint16_t)S.newpos[nid[u]];
                    }
                }
            }
        }
        const int nToExpand = s_misc[2];
        __syncthreads();
        const int prevSize = Lsz;
        Lsz = newL;
        cur = nxt;
        // termination (:645-701)
        if (Lsz >= N || Lsz == prevSize) break;
        if (!careful && (Lsz + nToExpand * 3) > N) careful = true;
    }

    // ---- retain the best point of each node (:704-720): max response, first in key order on ties
    for (int i = tid; i < Lsz; i += OCT_THREADS) S.best[i] = 0;
    __syncthreads();
    for (int k = tid; k < n; k += OCT_THREADS)
        atomicMax(&S.best[keynode[k]], ((uint32_t)cand_s(keys[k]) << 24) | (uint32_t)(0xFFFFFF - k));
    __syncthreads();
    for (int i = tid; i < Lsz; i += OCT_THREADS) sel[i] = keys[0xFFFFFF - (S.best[i] & 0xFFFFFF)];
    if (tid == 0) nsel_all[(size_t)frame * P.nlevels + level] = Lsz;
}

// ------------------------------------------------------------------------------------------------
// K4  GaussianBlur 7x7 sigma 2 of every pyramid level (src/ORBextractor.cc:1010,1083): OpenCV's 8-bit path is
//     8.8 fixed point, separable, weights [18,34,48,56,48,34,18]/256 per pass, BORDER_REFLECT_101 on the cloned
//     level, 16-bit intermediate, final (v + 2^15) >> 16.
//     One CTA per 128 x 32 output tile of one level of one frame.  The raw (128+32) x (32+6) box comes in with ONE
//     TMA bulk tensor copy (cp.async.bulk.tensor.3d; the box starts 16 columns left of the tile so its global address
//     stays 16-byte aligned; out-of-image elements are zero-filled by the TMA unit and then overwritten by the
//     REFLECT_101 fix-up, whose sources are always inside the box).  Horizontal pass: 2 x IDP.4A per output
//     (bytes x weight bytes), stored as VERTICAL u16 pairs so that the vertical pass is 4 x IDP.2A per output
//     with the rounding constant as the accumulator seed; no unpacking, no 32-bit multiplies.
// ------------------------------------------------------------------------------------------------
constexpr int BLUR_TW = 128, BLUR_TH = 32;
constexpr int BLUR_THREADS = 256;
constexpr int BLUR_XPAD = 16;                          // left halo in the box (3 needed, 16 keeps the TMA start aligned)
constexpr int BLUR_BOXW = BLUR_TW + 2 * BLUR_XPAD;     // 160
constexpr int BLUR_BOXH = BLUR_TH + 6;                 // 38
constexpr int BLUR_NPAIR = BLUR_BOXH / 2;              // 19 vertical pairs of H rows

struct BlurTile { short level, tx, ty, pad; };
struct TmaMaps {                       // per level: u8 [frames][h][pitch], box BLUR_BOXW x BLUR_BOXH x 1
    CUtensorMap level[MAX_LEVELS];
    unsigned ok;                       // bit l: level l is loaded by TMA (else: manual loader)
};

__global__ void __launch_bounds__(BLUR_THREADS) k_blur_levels(const __grid_constant__ OrbParams P,
                                                              const __grid_constant__ TmaMaps maps,
                                                              const BlurTile* __restrict__ tiles) {
    __shared__ __align__(128) uint8_t raw[BLUR_BOXH * BLUR_BOXW];          // 6080 B
    __shared__ __align__(16) uint32_t Hp[BLUR_NPAIR * BLUR_TW];            // 9728 B: H[2p][c] | H[2p+1][c] << 16
    __shared__ __align__(8) unsigned long long bar;
    const BlurTile t = tiles[blockIdx.x];
    const int frame = blockIdx.y;
    const LevelDev& L = P.lv[t.level];
    const int x0 = t.tx * BLUR_TW, y0 = t.ty * BLUR_TH;
    const int gx0 = x0 - BLUR_XPAD, gy0 = y0 - 3;          // image coordinates of raw[0][0]
    const int tid = threadIdx.x;
    // columns of the box the passes read: 13 .. 146 (x0-3 .. x0+130)
    if ((maps.ok >> t.level) & 1u) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            mbar_expect_tx(&bar, BLUR_BOXH * BLUR_BOXW);
            tma_load_3d(raw, &maps.level[t.level], gx0, gy0, frame, &bar);
        }
        __syncthreads();
        mbar_wait(&bar, 0);
        const bool edge = x0 == 0 || x0 + BLUR_TW + 3 > L.w || y0 == 0 || y0 + BLUR_TH + 3 > L.h;   // CTA-uniform
        if (edge) {
            // REFLECT_101: the three columns left / right of the image first, then the three rows above / below
            {
                const int r = tid >> 2, k = tid & 3;                          // 64 rows x 4, k < 3 used
                if (r < BLUR_BOXH && k < 3) {
                    uint8_t* row = raw + r * BLUR_BOXW;
                    if (x0 == 0) row[BLUR_XPAD - 1 - k] = row[BLUR_XPAD + 1 + k];          // x = -1-k  <-  x = 1+k
                    const int c = L.w + k - gx0;                                           // x = w+k  <-  x = w-2-k
                    if (c <= 146) row[c] = row[L.w - 2 - k - gx0];
                }
            }
            __syncthreads();
            if (tid < 136) {
                uint8_t* col = raw + 12 + tid;
                if (y0 == 0) {
#pragma unroll
                    for (int k = 0; k < 3; k++) col[(2 - k) * BLUR_BOXW] = col[(4 + k) * BLUR_BOXW];   // y = -1-k <- y = 1+k
                }
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const int r = L.h + k - gy0;                                                      // y = h+k <- y = h-2-k
                    if (r < BLUR_BOXH) col[r * BLUR_BOXW] = col[(L.h - 2 - k - gy0) * BLUR_BOXW];
                }
            }
            __syncthreads();
        }
    } else {
        const uint8_t* img = L.base + (size_t)frame * L.frame_stride;
        for (int i = tid; i < BLUR_BOXH * 136; i += BLUR_THREADS) {
            const int r = i / 136, c = 12 + (i - r * 136);
            const int x = min(max(gx0 + c, -3), L.w + 2), y = min(max(gy0 + r, -3), L.h + 2);
            raw[r * BLUR_BOXW + c] = img[(size_t)reflect101(y, L.h) * L.pitch + reflect101(x, L.w)];
        }
        __syncthreads();
    }
    // ---- horizontal pass: item = (row pair p, 4-column group g); b[k] = raw column 4g+13+k
    for (int i = tid; i < BLUR_NPAIR * (BLUR_TW / 4); i += BLUR_THREADS) {
        const int p = i >> 5, g = i & 31;
        unsigned h[2][4];
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(raw + (2 * p + rr) * BLUR_BOXW) + g + 3;
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t A = j == 3 ? w1 : __funnelshift_r(w0, w1, 8 * (j + 1));     // b[j .. j+3]
                const uint32_t Bq = j == 3 ? w2 : __funnelshift_r(w1, w2, 8 * (j + 1));    // b[j+4 .. j+7]
                h[rr][j] = __dp4a(Bq, 0x00122230u, __dp4a(A, 0x38302212u, 0u));            // 18,34,48,56 | 48,34,18,0
            }
        }
        uint4 o;
        o.x = h[0][0] | (h[1][0] << 16); o.y = h[0][1] | (h[1][1] << 16);
        o.z = h[0][2] | (h[1][2] << 16); o.w = h[0][3] | (h[1][3] << 16);
        *reinterpret_cast<uint4*>(&Hp[p * BLUR_TW + 4 * g]) = o;
    }
    __syncthreads();
    // ---- vertical pass: warp w -> output rows 4w..4w+3, lane g -> columns 4g..4g+3.  Output row o uses H rows
    //      o..o+6 = pairs (o/2 .. o/2+3) with weights (18,34)(48,56)(48,34)(18,0) for even o, (0,18)(34,48)(56,48)(34,18) for odd.
    {
        const int w = tid >> 5, g = tid & 31;
        if (x0 + 4 * g >= L.pitch) return;
        const uint32_t* hp = Hp + (2 * w) * BLUR_TW + 4 * g;
        uint4 q[5];
#pragma unroll
        for (int k = 0; k < 5; k++) q[k] = *reinterpret_cast<const uint4*>(hp + k * BLUR_TW);
        uint8_t* dst = L.blur + (size_t)frame * L.frame_stride + (size_t)(y0 + 4 * w) * L.pitch + x0 + 4 * g;
#pragma unroll
        for (int o = 0; o < 4; o++) {
            if (y0 + 4 * w + o >= L.h) break;
            const uint4 &p0 = q[o >> 1], &p1 = q[(o >> 1) + 1], &p2 = q[(o >> 1) + 2], &p3 = q[(o >> 1) + 3];
            const unsigned c0 = (o & 1) ? 0x1200u : 0x2212u, c1 = (o & 1) ? 0x3022u : 0x3830u,
                           c2 = (o & 1) ? 0x3038u : 0x2230u, c3 = (o & 1) ? 0x1222u : 0x0012u;
            const unsigned a0 = __dp2a_lo(p3.x, c3, __dp2a_lo(p2.x, c2, __dp2a_lo(p1.x, c1, __dp2a_lo(p0.x, c0, 1u << 15))));
            const unsigned a1 = __dp2a_lo(p3.y, c3, __dp2a_lo(p2.y, c2, __dp2a_lo(p1.y, c1, __dp2a_lo(p0.y, c0, 1u << 15))));
            const unsigned a2 = __dp2a_lo(p3.z, c3, __dp2a_lo(p2.z, c2, __dp2a_lo(p1.z, c1, __dp2a_lo(p0.z, c0, 1u << 15))));
            const unsigned a3 = __dp2a_lo(p3.w, c3, __dp2a_lo(p2.w, c2, __dp2a_lo(p1.w, c1, __dp2a_lo(p0.w, c0, 1u << 15))));
            // byte 2 of each accumulator = (acc + 2^15) >> 16  (acc < 2^24)
            const unsigned lo = __byte_perm(a0, a1, 0x0062), hi = __byte_perm(a2, a3, 0x0062);
            *reinterpret_cast<uint32_t*>(dst + (size_t)o * L.pitch) = __byte_perm(lo, hi, 0x5410);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K5  orientation + rBRIEF, one warp per keypoint:
//       IC_Angle              src/ORBextractor.cc:77-101  (raw level, int moments, cv::fastAtan2)
//       computeOrbDescriptor  src/ORBextractor.cc:105-149 (blurred level; float, no fma, round-half-even)
//     Two TMA bulk tensor copies per keypoint on one mbarrier: the 31-row raw box (orientation) and the 37-row blurred
//     box (512 rotated samples), each 64 columns wide with the start rounded down to 16 bytes.  Everything after the
//     wait is shared-memory reads: 32 random global bytes per instruction cost ~20 L1 wavefronts, 32 random shared
//     bytes ~3.  Measured (C1, B = 256): global-memory IC loop 0.37 ms, this version 0.22 ms.
// ------------------------------------------------------------------------------------------------
constexpr int DESC_WARPS = 8;           // one warp per keypoint, DESC_WARPS keypoints per CTA, no block barrier
constexpr int DESC_THREADS = DESC_WARPS * 32;
constexpr int PR = 18;                  // reach of the rotated pattern (|coordinate| <= 13 => <= 18.4 after rotation)
constexpr int PW = 2 * PR + 1;          // 37 rows / columns
constexpr int PBOX = 64;                // TMA box width: 37 columns + up to 15 of alignment slack
constexpr int ICW = 2 * HALF_PATCH + 1; // 31 rows / columns of the orientation patch
constexpr int DESC_RAW_OFF = 2432;      // blurred box 37 * 64 = 2368 B, rounded to 128 B (TMA destination alignment)
constexpr int DESC_WARP_SMEM = 4480;    // + raw box 31 * 64 = 1984 B, rounded to 128 B
struct PatchMaps {                      // per level: blurred u8 [frames][h][pitch], box PBOX x PW x 1; raw, box PBOX x ICW x 1
    CUtensorMap blur[MAX_LEVELS];
    CUtensorMap raw[MAX_LEVELS];
    unsigned ok;
};

// bit_pattern_31_ (data, src/ORBextractor.cc:152-410) lives in GLOBAL memory as float4 (x0,y0,x1,y1) [test j of a byte][lane = byte]:
// every lane reads its own eight tests (lane-divergent addresses would serialise on the constant cache), and the
// transposed order makes each of the eight loads one contiguous 512-byte warp access.

__device__ __forceinline__ float fast_atan2_dev(float y, float x) {
    // cv::fastAtan2 (degrees), evaluated without fma contraction
    const float scale = (float)(180.0 / 3.141592653589793238462643383279502884197169399375);
    const float p1 = __fmul_rn(0.9997878412794807f, scale);
    const float p3 = __fmul_rn(-0.3258083974640975f, scale);
    const float p5 = __fmul_rn(0.1555786518463281f, scale);
    const float p7 = __fmul_rn(-0.04432655554792128f, scale);
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// cosf/sinf of the reference (glibc) — see trig_glibc.cuh
__device__ void ref_cos_sin(float angle_rad, float* c, float* s);

struct ExistingKey { float x, y, angle; int octave; };   // pt * mvInvScaleFactor[octave], :1094-1095

// mode 0: keys selected by k_octree (angle computed);  mode 1: pre-existing keys (angle kept, :1088-1099);
// mode 2: level-0 keys of ComputeKeyPointsDSOSingleLevel (flat list of x | y << 12, angle computed, size 7, :1363-1367);
// mode 3: pre-existing keys whose angle that function refreshes (:1377-1380) — only the angle is written back
// lane = patch column for the IC moments (31 coalesced bytes per row), then lane j = descriptor byte j: 16 gathers
// from the blurred level, all independent, so the memory system sees 512 loads in flight per warp.
__global__ void __launch_bounds__(DESC_THREADS, 6) k_orient_describe(const __grid_constant__ OrbParams P,
                                                                  const __grid_constant__ PatchMaps maps,
                                                                  int mode, int n_existing, int n_items,
                                                                  const float4* __restrict__ pattern,
                                                                  const uint32_t* __restrict__ sel_all,
                                                                  const int* __restrict__ nsel_all,
                                                                  const ExistingKey* __restrict__ existing,
                                                                  ygz_keypoint* __restrict__ kps_out,
                                                                  uint8_t* __restrict__ desc_out,
                                                                  int* __restrict__ counts_out,
                                                                  const int4* __restrict__ per_frame = nullptr) {
    __shared__ __align__(128) uint8_t smem[DESC_WARPS * DESC_WARP_SMEM];
    __shared__ __align__(8) unsigned long long tma_bar[DESC_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * DESC_WARPS + warp;          // (level, slot) index or existing-key index
    const int frame = blockIdx.y;
    if (item >= n_items) return;
    uint8_t* G = smem + warp * DESC_WARP_SMEM;                 // blurred patch, PW rows x PBOX
    int level, cx, cy, out_idx, score = 0;
    float angle = 0.f;
    if (mode == 0) {
        const int s = item;
        // lane l holds level l's slot offset and selected-key count; the level of slot s and the number of keys of
        // the levels before it come from one ballot and one shuffle reduction
        const int* nsel = nsel_all + (size_t)frame * P.nlevels;
        const int my_off = lane < P.nlevels ? P.lv[lane].sel_off : 0x7fffffff;
        const int my_n = lane < P.nlevels ? nsel[lane] : 0;
        level = __popc(__ballot_sync(0xffffffffu, s >= my_off)) - 1;
        const int slot = s - __shfl_sync(0xffffffffu, my_off, level);
        int before = lane < level ? my_n : 0, tot = my_n;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            before += __shfl_xor_sync(0xffffffffu, before, o);
            tot += __shfl_xor_sync(0xffffffffu, tot, o);
        }
        before += n_existing;
        if (s == 0 && lane == 0) counts_out[frame] = tot + n_existing;
        const int n_level = __shfl_sync(0xffffffffu, my_n, level);
        if (slot >= n_level) return;
        const uint32_t c = sel_all[(size_t)frame * P.sel_total + s];
        cx = cand_x(c) + MIN_BORDER;          // :794-795
        cy = cand_y(c) + MIN_BORDER;
        score = cand_s(c);
        out_idx = before + slot;
    } else if (mode == 2) {
        // batched DSO selection (per_frame = DsoFrameState {g, cnt, done, n_existing} per frame): n_items is the stride of the key
        // lists, n_existing the stride of the existing-key lists; the frame's own counts come from its state
        int ne = n_existing, ni = n_items;
        if (per_frame) { const int4 st = per_frame[frame]; ne = st.w; ni = st.y; if (item >= ni) return; }
        const uint32_t c = sel_all[(per_frame ? (size_t)frame * n_items : 0) + item];
        level = 0;
        cx = (int)(c & 0xFFFu);
        cy = (int)((c >> 12) & 0xFFFu);
        out_idx = ne + item;
        if (item == 0 && lane == 0) counts_out[frame] = ne + ni;
    } else {
        if (per_frame && item >= per_frame[frame].w) return;
        const ExistingKey ek = existing[(size_t)frame * n_existing + item];
        level = ek.octave;
        cx = __float2int_rn(ek.x);             // cvRound(kpt.pt.x), :111
        cy = __float2int_rn(ek.y);
        angle = ek.angle;
        out_idx = item;
    }
    const LevelDev& L = P.lv[level];
    cx = min(max(cx, 0), L.w - 1);
    cy = min(max(cy, 0), L.h - 1);
    // ---- blurred patch rows cy-18..cy+18, columns cx-18..cx+18 -> shared
    const int x0 = cx - PR, y0 = cy - PR;
    const bool inside = x0 >= 0 && y0 >= 0 && y0 + PW <= L.h && x0 + PW <= L.w;
    const bool by_tma = inside && ((maps.ok >> level) & 1u);
    int shift;                                                  // column of x0 inside the shared row
    if (by_tma) {
        // one elected lane issues the bulk tensor copy (SASS: UTMALDG.3D); the box start must be 16-byte aligned in
        // global memory, so x is rounded down and `shift` keeps the offset.  Columns past the level width are
        // zero-filled by the TMA unit and never read.
        const int xa16 = x0 & ~15;
        shift = x0 - xa16;
        const bool with_raw = mode != 1;                           // raw 31-row box for IC_Angle rides on the same barrier
        if (elect_one()) {
            mbar_init(&tma_bar[warp], 1);
            mbar_expect_tx(&tma_bar[warp], with_raw ? (PW + ICW) * PBOX : PW * PBOX);
            if (with_raw) tma_load_3d(G + DESC_RAW_OFF, &maps.raw[level], (cx - HALF_PATCH) & ~15, cy - HALF_PATCH, frame, &tma_bar[warp]);
            tma_load_3d(G, &maps.blur[level], xa16, y0, frame, &tma_bar[warp]);
        }
        __syncwarp();
    } else {                                                    // small level / patch over the border: byte path, REFLECT_101
        const uint8_t* img = L.blur + (size_t)frame * L.frame_stride;
        shift = 0;
        for (int i = lane; i < PW * PW; i += 32) {
            const int r = i / PW, c = i - r * PW;
            G[r * PBOX + c] = img[(size_t)reflect101(y0 + r, L.h) * L.pitch + reflect101(x0 + c, L.w)];
        }
        __syncwarp();
    }
    // ---- IC_Angle on the raw level: lane = column u, loop over rows v (keys of modes 0 and 2 are >= 19 px inside)
    if (mode != 1) {
        const uint8_t* c0 = L.base + (size_t)frame * L.frame_stride + (size_t)cy * L.pitch + cx;
        int m10 = 0, m01 = 0;
        const int u = lane - HALF_PATCH;                        // lanes 0..30 <-> u = -15..15
        if (by_tma) {
            // raw 31-row box already on its way by TMA: column u of the circular patch covers rows |v| <= vlim[|u|]
            mbar_wait(&tma_bar[warp], 0);
            const int vl = lane < 31 ? P.vlim[abs(u)] : -1;
            const uint8_t* pc = G + DESC_RAW_OFF + ((cx - HALF_PATCH) & 15) + lane;
            int colsum = 0;
#pragma unroll
            for (int k = 0; k < ICW; k++) {
                const int v = k - HALF_PATCH;
                const int I = abs(v) <= vl ? (int)pc[k * PBOX] : 0;
                colsum += I;
                m01 += v * I;
            }
            m10 = u * colsum;
            // (measured, C1 B=256: lane = patch ROW with the row read as three LDS.128, funnel shifts and IDP.4A against per-|v| weight rows
            //  is ~230 instructions shorter per key point but SLOWER, 0.265 ms against 0.227: rows are 16 words apart, so every 128-bit
            //  access is a 4-way bank conflict and the weight rows cost four more L1 requests per lane; nine LDS.32 per lane: 0.347 ms.)
        } else if (lane < 31) {                                   // small level, or a pre-existing key near the border (REFLECT_101)
            const int au = abs(u);
            const bool in15 = cx >= HALF_PATCH && cy >= HALF_PATCH && cx + HALF_PATCH < L.w && cy + HALF_PATCH < L.h;
            const uint8_t* img0 = L.base + (size_t)frame * L.frame_stride;
#pragma unroll 4
            for (int v = -HALF_PATCH; v <= HALF_PATCH; v++) {
                if (au <= P.umax[abs(v)]) {
                    const int I = in15 ? (int)__ldg(c0 + v * L.pitch + u)
                                       : (int)img0[(size_t)reflect101(cy + v, L.h) * L.pitch + reflect101(cx + u, L.w)];
                    m10 += u * I;
                    m01 += v * I;
                }
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            m10 += __shfl_xor_sync(0xffffffffu, m10, o);
            m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        }
        angle = fast_atan2_dev((float)m01, (float)m10);        // every lane computes the same value
    }
    float a, b;
    {
        const float factorPI = (float)(3.141592653589793238462643383279502884197169399375 / 180.0);   // (float)(CV_PI/180.f)
        float ca = 0.f, sb = 0.f;
        if (lane == 0) ref_cos_sin(__fmul_rn(angle, factorPI), &ca, &sb);
        a = __shfl_sync(0xffffffffu, ca, 0);
        b = __shfl_sync(0xffffffffu, sb, 0);
    }
    if (by_tma && mode == 1) mbar_wait(&tma_bar[warp], 0);
    // ---- lane j evaluates tests 8j..8j+7 = descriptor byte j (:118-145) on the blurred patch
    {
        const uint8_t* ctr = G + PR * PBOX + PR + shift;
        // cvRound = cvtss2si = F2I.RN (the magic-add alternative, v + 1.5 * 2^23, measured no faster)
        auto rnd = [](float v) -> int { return __float2int_rn(v); };
        auto sample = [&](float xf, float yf) -> int {
            const int dy = rnd(__fadd_rn(__fmul_rn(xf, b), __fmul_rn(yf, a)));   // cvRound(x*b + y*a)
            const int dx = rnd(__fsub_rn(__fmul_rn(xf, a), __fmul_rn(yf, b)));   // cvRound(x*a - y*b)
            return ctr[dy * PBOX + dx];
        };
        int byte = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float4 pt = __ldg(pattern + j * 32 + lane);         // transposed table: one instruction reads 512 contiguous bytes
            const int t0 = sample(pt.x, pt.y);
            const int t1 = sample(pt.z, pt.w);
            byte |= (t0 < t1) << j;
        }
        desc_out[((size_t)frame * P.out_cap + out_idx) * 32 + lane] = (uint8_t)byte;
    }
    if (lane == 0 && mode == 3) kps_out[(size_t)frame * P.out_cap + out_idx].angle = angle;
    if (lane == 0 && mode == 2) {
        ygz_keypoint kp;           // cv::KeyPoint(Point2f(x, y), 7): response 0, octave 0, class_id -1
        kp.x = (float)cx; kp.y = (float)cy; kp.size = 7.f; kp.angle = angle; kp.response = 0.f; kp.octave = 0; kp.class_id = -1;
        kps_out[(size_t)frame * P.out_cap + out_idx] = kp;
    }
    if (lane == 0 && mode == 0) {
        ygz_keypoint kp;
        if (level != 0) {            // keypoint->pt *= scale, :1019-1024
            kp.x = __fmul_rn((float)cx, L.scale);
            kp.y = __fmul_rn((float)cy, L.scale);
        } else {
            kp.x = (float)cx; kp.y = (float)cy;
        }
        kp.size = L.size;
        kp.angle = angle;
        kp.response = (float)score;
        kp.octave = level;
        kp.class_id = -1;
        kps_out[(size_t)frame * P.out_cap + out_idx] = kp;
    }
}

}  // namespace ygz_b200
