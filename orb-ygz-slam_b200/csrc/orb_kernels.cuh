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
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pyr_resize(const uint8_t* __restrict__ src, int spitch, size_t sframe,
                                                    int sw, uint8_t* __restrict__ dst, int dpitch, size_t dframe,
                                                    int dw, int dh, const ResizeTabX* __restrict__ tx,
                                                    const ResizeTabY* __restrict__ ty) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int dy = blockIdx.y * blockDim.y + threadIdx.y;
    if (x4 >= dw || dy >= dh) return;
    const uint8_t* S = src + (size_t)blockIdx.z * sframe;
    uint8_t* D = dst + (size_t)blockIdx.z * dframe + (size_t)dy * dpitch;
    const ResizeTabY yy = ty[dy];
    const uint8_t* S0 = S + (size_t)yy.y0 * spitch;
    const uint8_t* S1 = S + (size_t)yy.y1 * spitch;
    uint32_t outw = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int dx = x4 + i;
        int v = 0;
        if (dx < dw) {
            const int2 tv = __ldg(reinterpret_cast<const int2*>(tx) + dx);
            ResizeTabX t; t.ofs = tv.x; t.a0 = (short)(tv.y & 0xFFFF); t.a1 = (short)(tv.y >> 16);
            const int sx1 = min(t.ofs + 1, sw - 1);
            const int r0 = S0[t.ofs] * t.a0 + S0[sx1] * t.a1;
            const int r1 = S1[t.ofs] * t.a0 + S1[sx1] * t.a1;
            v = (((yy.b0 * (r0 >> 4)) >> 16) + ((yy.b1 * (r1 >> 4)) >> 16) + 2) >> 2;
        }
        outw |= (uint32_t)(v & 0xFF) << (8 * i);
    }
    if (x4 + 3 < dpitch) *reinterpret_cast<uint32_t*>(D + x4) = outw;   // pitch is a multiple of 4; pad bytes are scratch
}

// ------------------------------------------------------------------------------------------------
// K2  FAST-9/16 per ~30x30 cell with the 20 -> 7 threshold fallback: the cv::FAST(cell, iniTh, true) /
//     cv::FAST(cell, minTh, true) pair of src/ORBextractor.cc:765-770 as ONE dense pass
//     (score map S = best-1, NMS masked to the cell's detection window).
//     One CTA per (cell, frame).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int fast9_best(int c, const int* r) {
    // best = max over the 16 arcs of 9 contiguous ring pixels of min(c - r) and of min(r - c).
    // Both polarities are evaluated as MIN networks over explicit differences: ptxas 12.9 (sm_100a) drops the
    // integer negation when it fuses max(max(a, -b), c) into VIMNMX3 (observed: scores inflated to +max(d)),
    // so no negated operand may feed a min/max here.
    int dk[16], br[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { dk[k] = c - r[k]; br[k] = r[k] - c; }
    int d2[16], b2[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { d2[k] = min(dk[k], dk[(k + 1) & 15]); b2[k] = min(br[k], br[(k + 1) & 15]); }
    int d4[16], b4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { d4[k] = min(d2[k], d2[(k + 2) & 15]); b4[k] = min(b2[k], b2[(k + 2) & 15]); }
    int best = -256;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int d9 = min(min(d4[k], d4[(k + 4) & 15]), dk[(k + 8) & 15]);    // min of (c - r) over arc k..k+8
        const int b9 = min(min(b4[k], b4[(k + 4) & 15]), br[(k + 8) & 15]);    // min of (r - c) over arc
        best = max(best, max(d9, b9));
    }
    return best;
}

constexpr int FAST_THREADS = 128;

__global__ void __launch_bounds__(FAST_THREADS) k_fast_cells(OrbParams P, const CellDesc* __restrict__ cells,
                                                             uint32_t* __restrict__ pool,
                                                             uint16_t* __restrict__ cell_count) {
    extern __shared__ uint8_t smem_raw[];
    const CellDesc cd = cells[blockIdx.x];
    const int frame = blockIdx.y;
    const LevelDev& L = P.lv[cd.level];
    const int cw = cd.cw, ch = cd.ch;
    const int dw = cw - 6, dh = ch - 6;               // detection window [3, cw-3) x [3, ch-3)
    uint16_t* cnt_out = cell_count + (size_t)frame * P.n_cells + blockIdx.x;
    if (dw <= 0 || dh <= 0) { if (threadIdx.x == 0) *cnt_out = 0; return; }
    const int npx = dw * dh;
    const int sw2 = dw + 2;
    // shared layout
    uint8_t* raw = smem_raw;                                      // ch x cw
    uint8_t* sc = raw + round4(cw * ch);                          // (dh+2) x (dw+2) scores, zero border
    uint8_t* flag = sc + round4((dw + 2) * (dh + 2));             // per item: score if NMS survivor else 0
    uint16_t* items = reinterpret_cast<uint16_t*>(flag + round4(npx));   // pre-test survivors, raster order
    __shared__ int warp_cnt[FAST_THREADS / 32];
    __shared__ int warp_off[FAST_THREADS / 32 + 1];

    const uint8_t* img = L.base + (size_t)frame * L.frame_stride + (size_t)cd.y0 * L.pitch + cd.x0;
    for (int i = threadIdx.x; i < cw * ch; i += FAST_THREADS) {
        const int y = i / cw, x = i - y * cw;
        raw[i] = img[(size_t)y * L.pitch + x];
    }
    for (int i = threadIdx.x; i < (dw + 2) * (dh + 2); i += FAST_THREADS) sc[i] = 0;
    __syncthreads();

    const int tmin = P.min_th;
    // ring offsets in the raw tile (cv::FAST offsets16)
    const int o0 = 3 * cw, o1 = 3 * cw + 1, o2 = 2 * cw + 2, o3 = cw + 3, o4 = 3, o5 = -cw + 3, o6 = -2 * cw + 2,
              o7 = -3 * cw + 1, o8 = -3 * cw, o9 = -3 * cw - 1, o10 = -2 * cw - 2, o11 = -cw - 3, o12 = -3,
              o13 = cw - 3, o14 = 2 * cw - 2, o15 = 3 * cw - 1;

    // ---- phase A: opposite-pair pre-test at minTh (necessary for 9 contiguous), ordered per-warp compaction
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = FAST_THREADS / 32;
    const int chunk = ((npx + nwarps - 1) / nwarps + 31) & ~31;
    {
        const int start = warp * chunk, end = min(npx, start + chunk);
        int cnt = 0;
        for (int base = start; base < end; base += 32) {
            const int idx = base + lane;
            bool pass = false;
            if (idx < end) {
                const int y = idx / dw, x = idx - y * dw;
                const uint8_t* p = raw + (y + 3) * cw + x + 3;
                const int c = p[0];
                const int hi = c + tmin, lo = c - tmin;
                int a = p[o0], b = p[o8];
                int m = ((a > hi) | (b > hi)) * 2 + ((a < lo) | (b < lo));
                if (m) {
                    a = p[o4]; b = p[o12];
                    m &= ((a > hi) | (b > hi)) * 2 + ((a < lo) | (b < lo));
                    if (m) {
                        a = p[o2]; b = p[o10];
                        m &= ((a > hi) | (b > hi)) * 2 + ((a < lo) | (b < lo));
                        a = p[o6]; b = p[o14];
                        m &= ((a > hi) | (b > hi)) * 2 + ((a < lo) | (b < lo));
                    }
                }
                pass = m != 0;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            if (pass) items[start + cnt + __popc(bal & lanemask_lt())] = (uint16_t)idx;
            cnt += __popc(bal);
        }
        if (lane == 0) warp_cnt[warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < nwarps; w++) { warp_off[w] = s; s += warp_cnt[w]; }
        warp_off[nwarps] = s;
    }
    __syncthreads();
    const int total = warp_off[nwarps];
    auto item_at = [&](int g) -> int {
        int w = 0;
#pragma unroll
        for (int k = 1; k < nwarps; k++) w += (g >= warp_off[k]);
        return items[w * chunk + (g - warp_off[w])];
    };

    // ---- phase B: full FAST-9 score of the pre-test survivors
    for (int g = threadIdx.x; g < total; g += FAST_THREADS) {
        const int idx = item_at(g);
        const int y = idx / dw, x = idx - y * dw;
        const uint8_t* p = raw + (y + 3) * cw + x + 3;
        int r[16] = {p[o0], p[o1], p[o2], p[o3], p[o4], p[o5], p[o6], p[o7],
                     p[o8], p[o9], p[o10], p[o11], p[o12], p[o13], p[o14], p[o15]};
        const int best = fast9_best(p[0], r);
        if (best > tmin) sc[(y + 1) * sw2 + x + 1] = (uint8_t)(best - 1);   // corner at minTh; response = best-1
    }
    __syncthreads();

    // ---- phase C: 3x3 non-max suppression inside the detection window (strict >, zeros outside)
    int any_hi = 0;
    for (int g = threadIdx.x; g < total; g += FAST_THREADS) {
        const int idx = item_at(g);
        const int y = idx / dw, x = idx - y * dw;
        const uint8_t* q = sc + (y + 1) * sw2 + x + 1;
        const int s = q[0];
        int keep = 0;
        // a corner at minTh has s >= minTh; s == 0 can only be a corner when minTh == 0 and is never kept by cv either
        if (s > 0 && s >= tmin) {
            keep = (s > q[-1]) & (s > q[1]) & (s > q[-sw2 - 1]) & (s > q[-sw2]) & (s > q[-sw2 + 1]) &
                   (s > q[sw2 - 1]) & (s > q[sw2]) & (s > q[sw2 + 1]);
        }
        flag[g] = keep ? (uint8_t)s : 0;
        any_hi |= keep && (s >= P.ini_th);
    }
    // vKeysCell.empty() after the iniTh call decides the fallback (src/ORBextractor.cc:767)
    const int use_hi = __syncthreads_or(any_hi);
    const int T = use_hi ? P.ini_th : tmin;

    // ---- phase D: ordered emission (raster order inside the cell), coordinates relative to (16,16)
    uint32_t* out = pool + (size_t)frame * P.pool_size + cd.out_off;
    int base = 0;
    for (int g0 = 0; g0 < total; g0 += FAST_THREADS) {
        const int g = g0 + threadIdx.x;
        int s = 0, idx = 0;
        if (g < total) { s = flag[g]; idx = item_at(g); }
        const bool keep = s > 0 && s >= T;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; w++) off += warp_cnt[w];
        if (keep) {
            const int y = idx / dw, x = idx - y * dw;
            const int pos = off + __popc(bal & lanemask_lt());
            if (pos < cd.cap) out[pos] = pack_cand(x + 3 + cd.offx, y + 3 + cd.offy, s);
        }
        for (int w = 0; w < nwarps; w++) base += warp_cnt[w];
        __syncthreads();
    }
    if (threadIdx.x == 0) *cnt_out = (uint16_t)min(base, (int)cd.cap);
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
constexpr int OCT_THREADS = 256;

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

__global__ void __launch_bounds__(OCT_THREADS) k_octree(OrbParams P, const CellDesc* __restrict__ cells,
                                                        const uint32_t* __restrict__ pool,
                                                        const uint16_t* __restrict__ cell_count,
                                                        uint32_t* __restrict__ keys_all, uint16_t* __restrict__ keynode_all,
                                                        int* __restrict__ nkeys_all, uint32_t* __restrict__ sel_all,
                                                        int* __restrict__ nsel_all) {
    extern __shared__ __align__(16) uint8_t smem_oct[];
    const int level = blockIdx.x, frame = blockIdx.y;
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
        for (int k = tid; k < n; k += OCT_THREADS) {
            const int nid = keynode[k];
            if (CNT[nid] > 1) {
                const uint32_t c = keys[k];
                const int q = quadrant_of(cand_x(c), cand_y(c), X0[nid], X1[nid], Y0[nid], Y1[nid]);
                atomicAdd(&S.qcnt[nid * 4 + q], 1);
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
        for (int k = tid; k < n; k += OCT_THREADS) {
            const int nid = keynode[k];
            if (S.processed[nid]) {
                const uint32_t c = keys[k];
                const int q = quadrant_of(cand_x(c), cand_y(c), X0[nid], X1[nid], Y0[nid], Y1[nid]);
                keynode[k] = (uint16_t)S.childpos[nid * 4 + q];
            } else {
                keynode[k] = (uint16_t)S.newpos[nid];
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
// K4  orientation + rBRIEF for one keypoint per CTA, with the 7x7 sigma-2 Gaussian evaluated on the fly
//     from the RAW level (no blurred pyramid in HBM):
//       IC_Angle              src/ORBextractor.cc:77-101  (raw level, int moments, cv::fastAtan2)
//       GaussianBlur 7x7 s2   src/ORBextractor.cc:1010,1083 (8.8 fixed point [18,34,48,56,48,34,18], REFLECT_101)
//       computeOrbDescriptor  src/ORBextractor.cc:105-149 (float, no fma, round-half-even)
// ------------------------------------------------------------------------------------------------
constexpr int DESC_THREADS = 256;
constexpr int PR = 21;                 // patch radius: 18 (pattern reach) + 3 (blur)
constexpr int PW = 2 * PR + 1;         // 43
constexpr int RAWP = 44;               // raw pitch
constexpr int HW = 37;                 // horizontally blurred columns (-18..18)
constexpr int HP = 38;                 // pitch of H

__constant__ int8_t c_pattern[256 * 4];     // bit_pattern_31_ (data), src/ORBextractor.cc:152-410

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

// mode 0: keys selected by k_octree (angle computed);  mode 1: pre-existing keys (angle kept, :1088-1099)
__global__ void __launch_bounds__(DESC_THREADS) k_orient_describe(OrbParams P, int mode, int n_existing,
                                                                  const uint32_t* __restrict__ sel_all,
                                                                  const int* __restrict__ nsel_all,
                                                                  const ExistingKey* __restrict__ existing,
                                                                  ygz_keypoint* __restrict__ kps_out,
                                                                  uint8_t* __restrict__ desc_out,
                                                                  int* __restrict__ counts_out) {
    __shared__ uint8_t raw[PW * RAWP];
    __shared__ uint16_t H[PW * HP];
    __shared__ int s_m[2][DESC_THREADS / 32];
    __shared__ float s_ab[3];
    const int tid = threadIdx.x;
    const int frame = blockIdx.y;
    int level, cx, cy, out_idx, score = 0;
    float angle = 0.f;
    if (mode == 0) {
        // blockIdx.x enumerates (level, slot) over the per-frame sel array
        const int s = blockIdx.x;
        level = 0;
        while (level + 1 < P.nlevels && s >= P.lv[level + 1].sel_off) level++;
        const int slot = s - P.lv[level].sel_off;
        const int* nsel = nsel_all + (size_t)frame * P.nlevels;
        int before = n_existing;
        for (int l = 0; l < level; l++) before += nsel[l];
        if (s == 0 && tid == 0) {
            int tot = n_existing;
            for (int l = 0; l < P.nlevels; l++) tot += nsel[l];
            counts_out[frame] = tot;
        }
        if (slot >= nsel[level]) return;
        const uint32_t c = sel_all[(size_t)frame * P.sel_total + s];
        cx = cand_x(c) + MIN_BORDER;          // :794-795
        cy = cand_y(c) + MIN_BORDER;
        score = cand_s(c);
        out_idx = before + slot;
    } else {
        const ExistingKey ek = existing[(size_t)frame * n_existing + blockIdx.x];
        level = ek.octave;
        cx = __float2int_rn(ek.x);             // cvRound(kpt.pt.x), :111
        cy = __float2int_rn(ek.y);
        angle = ek.angle;
        out_idx = blockIdx.x;
    }
    const LevelDev& L = P.lv[level];
    cx = min(max(cx, 0), L.w - 1);
    cy = min(max(cy, 0), L.h - 1);
    const uint8_t* img = L.base + (size_t)frame * L.frame_stride;
    // ---- raw patch with REFLECT_101 at the level border
    for (int i = tid; i < PW * PW; i += DESC_THREADS) {
        const int r = i / PW, c = i - r * PW;
        const int yy = reflect101(cy - PR + r, L.h), xx = reflect101(cx - PR + c, L.w);
        raw[r * RAWP + c] = img[(size_t)yy * L.pitch + xx];
    }
    __syncthreads();
    // ---- IC_Angle on the raw patch
    if (mode == 0) {
        int m10 = 0, m01 = 0;
        for (int i = tid; i < 31 * 31; i += DESC_THREADS) {
            const int v = i / 31 - HALF_PATCH, u = i - (i / 31) * 31 - HALF_PATCH;
            if (abs(u) <= P.umax[abs(v)]) {
                const int I = raw[(PR + v) * RAWP + PR + u];
                m10 += u * I;
                m01 += v * I;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            m10 += __shfl_xor_sync(0xffffffffu, m10, o);
            m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        }
        if ((tid & 31) == 0) { s_m[0][tid >> 5] = m10; s_m[1][tid >> 5] = m01; }
    }
    // ---- horizontal 7-tap pass (8.8 fixed point), columns -18..18 of all 43 rows
    for (int i = tid; i < PW * HW; i += DESC_THREADS) {
        const int r = i / HW, c = i - r * HW;            // output column c <-> raw columns c .. c+6
        const uint8_t* p = raw + r * RAWP + c;
        H[r * HP + c] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
    }
    __syncthreads();
    if (tid == 0) {
        if (mode == 0) {
            int m10 = 0, m01 = 0;
            for (int w = 0; w < DESC_THREADS / 32; w++) { m10 += s_m[0][w]; m01 += s_m[1][w]; }
            angle = fast_atan2_dev((float)m01, (float)m10);
        }
        const float factorPI = (float)(3.141592653589793238462643383279502884197169399375 / 180.0);   // (float)(CV_PI/180.f)
        float a, b;
        ref_cos_sin(__fmul_rn(angle, factorPI), &a, &b);
        s_ab[0] = a; s_ab[1] = b; s_ab[2] = angle;
    }
    __syncthreads();
    const float a = s_ab[0], b = s_ab[1];
    // ---- 256 binary tests, one per thread; vertical 7-tap evaluated at the two sample points
    {
        const int px0 = c_pattern[tid * 4 + 0], py0 = c_pattern[tid * 4 + 1];
        const int px1 = c_pattern[tid * 4 + 2], py1 = c_pattern[tid * 4 + 3];
        auto sample = [&](int px, int py) -> int {
            const float xf = (float)px, yf = (float)py;
            const int dy = __float2int_rn(__fadd_rn(__fmul_rn(xf, b), __fmul_rn(yf, a)));   // cvRound(x*b + y*a)
            const int dx = __float2int_rn(__fsub_rn(__fmul_rn(xf, a), __fmul_rn(yf, b)));   // cvRound(x*a - y*b)
            const uint16_t* h = H + (PR + dy - 3) * HP + (dx + 18);
            const uint32_t acc = 18u * (h[0] + h[6 * HP]) + 34u * (h[HP] + h[5 * HP]) + 48u * (h[2 * HP] + h[4 * HP]) + 56u * h[3 * HP];
            return (int)((acc + (1u << 15)) >> 16);
        };
        const int t0 = sample(px0, py0), t1 = sample(px1, py1);
        const unsigned bal = __ballot_sync(0xffffffffu, t0 < t1);      // bit j of byte i = test 8i+j (:118-145)
        if ((tid & 31) == 0)
            *reinterpret_cast<uint32_t*>(desc_out + ((size_t)frame * P.out_cap + out_idx) * 32 + (tid >> 5) * 4) = bal;
    }
    if (tid == 0 && mode == 0) {
        ygz_keypoint kp;
        if (level != 0) {            // keypoint->pt *= scale, :1019-1024
            kp.x = __fmul_rn((float)cx, L.scale);
            kp.y = __fmul_rn((float)cy, L.scale);
        } else {
            kp.x = (float)cx; kp.y = (float)cy;
        }
        kp.size = L.size;
        kp.angle = s_ab[2];
        kp.response = (float)score;
        kp.octave = level;
        kp.class_id = -1;
        kps_out[(size_t)frame * P.out_cap + out_idx] = kp;
    }
}

}  // namespace ygz_b200
