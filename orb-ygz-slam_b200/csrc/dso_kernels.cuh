// YGZ "DSO" key-point selection, device side:
//   ORBextractor::ComputeKeyPointsDSOSingleLevel   /root/reference/src/ORBextractor.cc:1277-1381
//   ORBextractor::ShiTomasiScore                   :1152-1186
//   fast::fast_corner_detect_10 / _sse2            /root/reference/Thirdparty/fast/src/{fast_10,faster_corner_10_sse}.cpp
// One WARP per grid cell: FAST-10 at threshold 20 over the cell's detection region (whole g x g cell below 22 px — the
// plain detector the SSE2 entry point falls back to tests every pixel — else the interior [3, g-3)^2), threshold 5 when
// that finds nothing; 20-px image border and occupancy filters; Shi-Tomasi score of every surviving corner; the three
// best by (score desc, raster order) — the oracle's stable-sort reading of the reference's std::sort.
#pragma once
#include "common.cuh"

namespace ygz_b200 {

constexpr int DSO_WARPS = 4;
constexpr int DSO_GMAX = 96;                     // largest supported cell size (tile (g+6)^2 bytes per warp in shared memory)

__host__ __device__ inline int dso_tile_pitch(int g) { return (g + 6 + 3) & ~3; }
__host__ __device__ inline int dso_warp_smem(int g) { return (dso_tile_pitch(g) * (g + 6) + 15) & ~15; }

// >= 10 contiguous ring pixels all brighter than c + b or all darker than c - b
__device__ __forceinline__ bool fast10_corner(const uint8_t* p, int tp, int b) {
    const int c = p[0];
    const int hi = c + b, lo = c - b;
    unsigned br = 0, dk = 0;
#define RING(k, dx, dy) { const int v = p[(dx) + (dy) * tp]; br |= (unsigned)(v > hi) << (k); dk |= (unsigned)(v < lo) << (k); }
    RING(0, 0, 3) RING(1, 1, 3) RING(2, 2, 2) RING(3, 3, 1) RING(4, 3, 0) RING(5, 3, -1) RING(6, 2, -2) RING(7, 1, -3)
    RING(8, 0, -3) RING(9, -1, -3) RING(10, -2, -2) RING(11, -3, -1) RING(12, -3, 0) RING(13, -3, 1) RING(14, -2, 2) RING(15, -1, 3)
#undef RING
    br |= br << 16; dk |= dk << 16;
    unsigned rb = br, rd = dk;
#pragma unroll
    for (int i = 1; i < 10; i++) { rb &= br >> i; rd &= dk >> i; }
    return ((rb | rd) & 0xFFFFu) != 0;
}

// ShiTomasiScore(img, u, v), src/ORBextractor.cc:1152-1186 (integer sums are exact in float; no fma)
__device__ inline float shi_tomasi(const uint8_t* img, int w, int h, int pitch, int u, int v) {
    const int x_min = u - 4, x_max = u + 4, y_min = v - 4, y_max = v + 4;
    if (x_min < 1 || x_max >= w - 1 || y_min < 1 || y_max >= h - 1) return 0.f;
    int sxx = 0, syy = 0, sxy = 0;
    for (int y = y_min; y < y_max; y++) {
        const uint8_t* r = img + (size_t)y * pitch + x_min;
#pragma unroll
        for (int x = 0; x < 8; x++) {
            const int dx = (int)r[x + 1] - (int)r[x - 1];
            const int dy = (int)r[x + pitch] - (int)r[x - pitch];
            sxx += dx * dx; syy += dy * dy; sxy += dx * dy;
        }
    }
    // float accumulation of integers below 2^24 is exact; / (2.0 * 64) is a power-of-two scaling
    const float dXX = __fmul_rn((float)sxx, 0.0078125f), dYY = __fmul_rn((float)syy, 0.0078125f), dXY = __fmul_rn((float)sxy, 0.0078125f);
    const float s = __fadd_rn(dXX, dYY);
    const float disc = __fsub_rn(__fmul_rn(s, s), __fmul_rn(4.f, __fsub_rn(__fmul_rn(dXX, dYY), __fmul_rn(dXY, dXY))));
    return __fmul_rn(0.5f, __fsub_rn(s, __fsqrt_rn(disc)));
}

struct DsoParams {
    const uint8_t* img;     // level 0 of the frame
    const uint8_t* occ;     // w x h occupancy (255 = an existing key sits there), pitch w
    int w, h, pitch;
    int g, n_rows, n_cols;
};

// better(a, b): a precedes b in the reference's order (score desc, raster index asc)
__device__ __forceinline__ bool dso_better(float sa, int ia, float sb, int ib) { return sa > sb || (sa == sb && ia < ib); }

// one warp, one grid cell k of the frame described by P
__device__ __forceinline__ void dso_cell(const DsoParams& P, int k, uint32_t* __restrict__ cell_keys, uint8_t* __restrict__ cell_cnt,
                                         uint8_t* dso_smem, int warp_smem) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nn = k / P.n_cols, col = k - nn * P.n_cols;
    if (nn == 0 || nn == P.n_rows - 1 || col == 0 || col == P.n_cols - 1) { if (lane == 0) cell_cnt[k] = 0; return; }   // :1319-1321
    const int g = P.g, x_start = col * g, y_start = nn * g;
    const int tp = dso_tile_pitch(g);
    uint8_t* tile = dso_smem + (size_t)warp * warp_smem;          // rows y_start-3 .. y_start+g+2, cols x_start-3 ..
    for (int i = lane; i < (g + 6) * (g + 6); i += 32) {
        const int r = i / (g + 6), c = i - r * (g + 6);
        tile[r * tp + c] = P.img[(size_t)(y_start - 3 + r) * P.pitch + x_start - 3 + c];   // interior cells: always inside the image
    }
    __syncwarp();
    const int lo = g < 22 ? 0 : 3, hi = g < 22 ? g : g - 3;       // detection region of fast_corner_detect_10_sse2 on a g x g window
    const int rw = hi - lo;
    float s0 = -1.f, s1 = -1.f, s2 = -1.f;                        // per-lane best three (scores are >= 0)
    int i0 = 0x7fffffff, i1 = 0x7fffffff, i2 = 0x7fffffff;
    int n_mine = 0;
    for (int pass = 0; pass < 2; pass++) {
        const int t = pass ? 5 : 20;                              // :1331-1340
        bool any_corner = false;
        for (int i = lane; i < rw * rw + 31; i += 32) {
            bool corner = false;
            int x = 0, y = 0;
            if (i < rw * rw) {
                y = lo + i / rw; x = lo + (i - (i / rw) * rw);
                corner = fast10_corner(tile + (y + 3) * tp + x + 3, tp, t);
            }
            any_corner |= corner;
            if (corner) {
                const int X = x + x_start, Y = y + y_start;
                if (!(X < 20 || Y < 20 || X >= P.w - 20 || Y >= P.h - 20) && P.occ[(size_t)Y * P.w + X] != 255) {   // :1348-1353
                    const float sc = shi_tomasi(P.img, P.w, P.h, P.pitch, X, Y);
                    const int idx = y * g + x;                                       // raster order inside the cell
                    n_mine++;
                    if (dso_better(sc, idx, s0, i0)) { s2 = s1; i2 = i1; s1 = s0; i1 = i0; s0 = sc; i0 = idx; }
                    else if (dso_better(sc, idx, s1, i1)) { s2 = s1; i2 = i1; s1 = sc; i1 = idx; }
                    else if (dso_better(sc, idx, s2, i2)) { s2 = sc; i2 = idx; }
                }
            }
        }
        if (__any_sync(0xffffffffu, any_corner)) break;           // fast_corners.empty() decides the fallback, before the filters
    }
    // ---- merge the lanes' sorted triples: three rounds of warp arg-best
    int total = n_mine;
#pragma unroll
    for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    const int n_out = min(total, 3);
    for (int r = 0; r < n_out; r++) {
        float bs = s0; int bi = i0, bl = lane;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o), ol = __shfl_xor_sync(0xffffffffu, bl, o);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || dso_better(os, oi, bs, bi))) { bs = os; bi = oi; bl = ol; }
        }
        if (lane == 0) {
            const int y = bi / g, x = bi - y * g;
            cell_keys[(size_t)k * 3 + r] = (uint32_t)(x + x_start) | ((uint32_t)(y + y_start) << 12);
        }
        if (lane == bl) { s0 = s1; i0 = i1; s1 = s2; i1 = i2; s2 = -1.f; i2 = 0x7fffffff; }   // winner pops its head
    }
    if (lane == 0) cell_cnt[k] = (uint8_t)n_out;
}

__global__ void __launch_bounds__(DSO_WARPS * 32) k_dso_cells(DsoParams P, uint32_t* __restrict__ cell_keys,
                                                               uint8_t* __restrict__ cell_cnt, int warp_smem) {
    extern __shared__ __align__(16) uint8_t dso_smem[];
    const int k = blockIdx.x * DSO_WARPS + (threadIdx.x >> 5);
    if (k >= P.n_rows * P.n_cols) return;
    dso_cell(P, k, cell_keys, cell_cnt, dso_smem, warp_smem);
}

// ---- batched form: B independent streams, each with its own adaptive cell size.  The `while (cnt < n)` loop of :1300-1371 runs as a
// fixed number of (cells, compact) launch pairs; a frame whose loop has ended is skipped on the device, so the host synchronises once
// per group of passes instead of once per pass per frame.
struct DsoFrameState { int g, cnt, done, n_existing; };

__global__ void __launch_bounds__(DSO_WARPS * 32) k_dso_cells_batch(DsoParams P0, size_t img_stride, int max_cells,
                                                                     const DsoFrameState* __restrict__ state,
                                                                     uint32_t* __restrict__ cell_keys, uint8_t* __restrict__ cell_cnt) {
    extern __shared__ __align__(16) uint8_t dso_smem[];
    const int f = blockIdx.y;
    const DsoFrameState st = state[f];
    if (st.done) return;
    DsoParams P = P0;
    P.img += (size_t)f * img_stride;
    P.occ += (size_t)f * P.w * P.h;
    P.g = st.g; P.n_rows = P.h / st.g; P.n_cols = P.w / st.g;
    const int k = blockIdx.x * DSO_WARPS + (threadIdx.x >> 5);
    if (k >= P.n_rows * P.n_cols) return;
    dso_cell(P, k, cell_keys + (size_t)f * max_cells * 3, cell_cnt + (size_t)f * max_cells, dso_smem, dso_warp_smem(st.g));
}

// cell-order compaction of the per-cell winners (one CTA; cells <= (h/7) * (w/7)); returns the total to every thread
__device__ __forceinline__ int dso_compact(const uint32_t* __restrict__ cell_keys, const uint8_t* __restrict__ cell_cnt, int n_cells,
                                           uint32_t* __restrict__ keys_out) {
    __shared__ int wsum[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n_cells; c0 += 1024) {
        const int c = c0 + threadIdx.x;
        const int n = c < n_cells ? cell_cnt[c] : 0;
        int v = n;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        if (lane == 31) wsum[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int s = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
            wsum[lane] = s;
        }
        __syncthreads();
        const int off = base + (warp ? wsum[warp - 1] : 0) + v - n;
        for (int r = 0; r < n; r++) keys_out[off + r] = cell_keys[(size_t)c * 3 + r];
        __syncthreads();
        if (threadIdx.x == 0) base += wsum[31];
        __syncthreads();
    }
    return base;
}

__global__ void __launch_bounds__(1024) k_dso_compact(const uint32_t* __restrict__ cell_keys, const uint8_t* __restrict__ cell_cnt,
                                                      int n_cells, uint32_t* __restrict__ keys_out, int* __restrict__ total_out) {
    const int total = dso_compact(cell_keys, cell_cnt, n_cells, keys_out);
    if (threadIdx.x == 0) *total_out = total;
}

// one CTA per frame: compaction, then the loop control of :1300-1375 for that frame
//   while (cnt < n) { if (cnt > 0) { g -= 5; if (g < 7) { g = 7; break; } }  <pass>  }   (a pass that finds nothing ends the loop, as the oracle)
//   if (cnt > n) g += 5;
__global__ void __launch_bounds__(1024) k_dso_compact_batch(const uint32_t* __restrict__ cell_keys, const uint8_t* __restrict__ cell_cnt,
                                                            int w, int h, int max_cells, int n_wanted, uint32_t* __restrict__ keys_out,
                                                            DsoFrameState* __restrict__ state) {
    const int f = blockIdx.x;
    DsoFrameState st = state[f];
    if (st.done) return;
    const int n_cells = (h / st.g) * (w / st.g);
    const int total = dso_compact(cell_keys + (size_t)f * max_cells * 3, cell_cnt + (size_t)f * max_cells, n_cells,
                                  keys_out + (size_t)f * max_cells * 3);
    if (threadIdx.x == 0) {
        st.cnt = total;
        if (total == 0 || total >= n_wanted) {
            st.done = 1;
            if (total > n_wanted) st.g += 5;
        } else {
            st.g -= 5;
            if (st.g < 7) { st.g = 7; st.done = 1; }
        }
        state[f] = st;
    }
}

// mOccupancy.ptr(cvRound(kp.pt.y))[cvRound(kp.pt.x)] = 255, :1283-1285
__global__ void k_dso_mark_occupancy(const float2* __restrict__ pts, int n, uint8_t* __restrict__ occ, int w, int h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = __float2int_rn(pts[i].x), y = __float2int_rn(pts[i].y);
    if (x >= 0 && y >= 0 && x < w && y < h) occ[(size_t)y * w + x] = 255;
}

// the same for a batch: pts[i] belongs to frame fid[i]
__global__ void k_dso_mark_occupancy_batch(const float2* __restrict__ pts, const int* __restrict__ fid, int n, uint8_t* __restrict__ occ,
                                           int w, int h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = __float2int_rn(pts[i].x), y = __float2int_rn(pts[i].y);
    if (x >= 0 && y >= 0 && x < w && y < h) occ[(size_t)fid[i] * w * h + (size_t)y * w + x] = 255;
}

}  // namespace ygz_b200
