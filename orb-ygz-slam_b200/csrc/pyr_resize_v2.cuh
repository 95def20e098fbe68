// k_pyr_resize_v2 — ComputePyramid's cv::resize(INTER_LINEAR) on CV_8UC1 (/root/reference/src/ORBextractor.cc:1139),
// 11-bit fixed point exactly as OpenCV computes it; same bytes as k_pyr_resize (orb_kernels.cuh), 3 x fewer instructions.
//
// Round-1 profile of k_pyr_resize (level 1, B = 256): 40 thread-instructions per output pixel, ALU pipe 63 %, issue 65 %:
// 12 per pixel in the horizontal taps (two LDS.U8, an IMUL and an IMAD per value, addresses rebuilt per tap), 7 in the
// vertical blend, 6 in the staging loop (a division per 16 bytes) and the row bookkeeping.  Here
//   * the source window of a 256 x 16 output tile is staged in PAIR format: for every source pixel x the 16-bit word
//     S[x] | S[x+1] << 8.  A horizontal tap is then ONE aligned LDS.U16 and ONE IDP.2A (dp2a: a0 * S[x] + a1 * S[x+1],
//     with {a0, a1} being the 32-bit word the xofs/ialpha table already stores);
//   * h >> 4 is taken once per tap instead of once per use, and (b * (h >> 4)) >> 16 is one IMAD.HI with b pre-shifted
//     by 16 (the +2 rounding constant rides in the accumulator of the first one);
//   * 128 threads: two groups of 8 output rows share one staged window (staging loop walks (row, chunk) incrementally).
#pragma once
#include "orb_kernels.cuh"

namespace ygz_b200 {

constexpr int PY2_THREADS = 128;
constexpr int PY2_ROWS = 8;                        // output rows per thread
constexpr int PY2_TILE_ROWS = 2 * PY2_ROWS;        // output rows per CTA
constexpr int PY2_COLS = 256;                      // output columns per CTA (4 per thread)

// stage_pitch: bytes per staged source row (32 per 16 source pixels); the host sizes the window from the level's tables
__global__ void __launch_bounds__(PY2_THREADS) k_pyr_resize_v2(const uint8_t* __restrict__ src, int spitch, size_t sframe, int sw, int sh,
                                                               uint8_t* __restrict__ dst, int dpitch, size_t dframe, int dw, int dh,
                                                               const ResizeTabX* __restrict__ tx, const ResizeTabY* __restrict__ ty,
                                                               int stage_pitch) {
    extern __shared__ __align__(16) uint8_t py2_stage[];
    const int tid = threadIdx.x;
    const int xg0 = blockIdx.x * PY2_COLS;
    const int x4 = xg0 + (tid & 63) * 4;
    const int dyT = blockIdx.y * PY2_TILE_ROWS;
    const int dy0 = dyT + (tid >> 6) * PY2_ROWS;
    const uint8_t* S = src + (size_t)blockIdx.z * sframe;
    uint8_t* D = dst + (size_t)blockIdx.z * dframe;
    // source window of the tile
    const int xlast = min(xg0 + PY2_COLS, dw) - 1;
    const int sx_lo = __ldg(reinterpret_cast<const int2*>(tx) + xg0).x & ~15;
    const int sx_hi = min(__ldg(reinterpret_cast<const int2*>(tx) + xlast).x + 1, sw - 1);
    const int sy_lo = (short)(__ldg(reinterpret_cast<const int2*>(ty) + dyT).x & 0xFFFF);
    const int sy_hi = (short)(__ldg(reinterpret_cast<const int2*>(ty) + min(dyT + PY2_TILE_ROWS, dh) - 1).x >> 16);
    const int nrows = sy_hi - sy_lo + 1;
    const int nvec = (sx_hi - sx_lo) / 16 + 1;                 // 16-byte source vectors per staged row
    {
        const int q = PY2_THREADS / nvec, m = PY2_THREADS - q * nvec;
        int r = tid / nvec, c = tid - r * nvec;
        while (r < nrows) {
            const int col = sx_lo + c * 16;
            const uint8_t* g = S + (size_t)(sy_lo + r) * spitch + col;
            const uint4 v = ld_nc_u4(g);
            const unsigned nb = col + 16 < spitch ? (unsigned)__ldg(g + 16) : 0u;        // S[x + 1] of the vector's last pixel
            uint4 lo, hi;
            lo.x = __byte_perm(v.x, v.x, 0x2110); lo.y = __byte_perm(v.x, v.y, 0x4332);
            lo.z = __byte_perm(v.y, v.y, 0x2110); lo.w = __byte_perm(v.y, v.z, 0x4332);
            hi.x = __byte_perm(v.z, v.z, 0x2110); hi.y = __byte_perm(v.z, v.w, 0x4332);
            hi.z = __byte_perm(v.w, v.w, 0x2110); hi.w = __byte_perm(v.w, nb, 0x4332);
            uint4* o = reinterpret_cast<uint4*>(py2_stage + (size_t)r * stage_pitch + c * 32);
            o[0] = lo; o[1] = hi;
            r += q; c += m;
            if (c >= nvec) { c -= nvec; r++; }
        }
    }
    // row tables of this thread's 8 rows, fetched up front (independent loads)
    ResizeTabY yt[PY2_ROWS];
#pragma unroll
    for (int r = 0; r < PY2_ROWS; r++) {
        const int2 v = __ldg(reinterpret_cast<const int2*>(ty) + min(dy0 + r, dh - 1));
        yt[r].y0 = (short)(v.x & 0xFFFF); yt[r].y1 = (short)(v.x >> 16);
        yt[r].b0 = (short)(v.y & 0xFFFF); yt[r].b1 = (short)(v.y >> 16);
    }
    __syncthreads();
    if (x4 >= dw || dy0 >= dh) return;
    const int dy1 = min(dy0 + PY2_ROWS, dh) - 1;
    const uint8_t* pcol[4];
    unsigned a01[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int2 tv = __ldg(reinterpret_cast<const int2*>(tx) + min(x4 + i, dw - 1));
        pcol[i] = py2_stage + 2 * (tv.x - sx_lo);          // at the right border a1 == 0, so the pair's second byte is never used
        a01[i] = (unsigned)tv.y;                           // a0 | a1 << 16
    }
    unsigned hA[4], hB[4];               // (horizontally interpolated source row) >> 4: rows yA (= y0 of the current row) and yB (= y1)
    int yA = -1, yB = -1;
    auto hrow = [&](int y, unsigned* h) {
        const int ro = (y - sy_lo) * stage_pitch;
#pragma unroll
        for (int i = 0; i < 4; i++)
            h[i] = __dp2a_lo(a01[i], (unsigned)*reinterpret_cast<const uint16_t*>(pcol[i] + ro), 0u) >> 4;
    };
#pragma unroll
    for (int r = 0; r < PY2_ROWS; r++) {
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
        const unsigned b0s = (unsigned)(int)yy.b0 << 16, b1s = (unsigned)(int)yy.b1 << 16;
        unsigned t[4];
#pragma unroll
        for (int i = 0; i < 4; i++) t[i] = (__umulhi(b1s, hB[i]) + (__umulhi(b0s, hA[i]) + 2u)) >> 2;   // <= 255
        const uint32_t outw = t[0] | (t[1] << 8) | (t[2] << 16) | (t[3] << 24);
        *reinterpret_cast<uint32_t*>(D + (size_t)dy * dpitch + x4) = outw;   // pitch is a multiple of 16; pad bytes are scratch
    }
}

}  // namespace ygz_b200
