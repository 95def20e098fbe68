// Dense 256-bit Hamming brute force, best / second best per query (first index wins ties, as the `dist < bestDist` scans of
// /root/reference/src/ORBmatcher.cc:207-213): the warp-level body shared by k_hamming_bf_top2 (matcher.cu, one query set)
// and k_hamming_bf_top2_batch (stereo.cu, one query / target set per stereo pair, device-resident extractor outputs).
#pragma once
#include "common.cuh"
#include <climits>

namespace ygz_b200 {

// merge two (best, idx, second) triples; earlier index wins ties, `second` = second smallest value of the multiset
__device__ __forceinline__ void merge_top2(int& b, int& bi, int& s, int ob, int obi, int os) {
    if (ob < b || (ob == b && obi < bi)) { s = min(b, os); b = ob; bi = obi; }
    else s = min(s, ob);
}

// The same over the 32 lanes of a warp in two REDUX instructions (result on every lane).  Positions must be below 65536 (INT_MAX, "no
// candidate", folds to 0xFFFF and only ever meets b = 256).  K = min of (b << 16 | position): best distance, earliest position on ties;
// the second smallest value of the multiset is the minimum of the owner's own second and every other lane's best.
__device__ __forceinline__ void warp_top2(int& b, int& bi, int& s) {
    const unsigned key = ((unsigned)b << 16) | ((unsigned)bi & 0xFFFFu);
    const unsigned K = __reduce_min_sync(0xffffffffu, key);
    s = (int)__reduce_min_sync(0xffffffffu, (unsigned)(key == K ? s : b));
    b = (int)(K >> 16);
    bi = (int)(K & 0xFFFFu);
}

constexpr int BF_WARPS = 8;
constexpr int BF_QPW = 4;           // queries per warp
constexpr int BF_TILE = 256;        // targets per shared-memory tile

// One warp owns BF_QPW queries (registers); the CTA streams target tiles through shared memory stored as
// 8 word-planes so that lane j reading target j is bank-conflict free.
__device__ __forceinline__ void bf_top2_body(const uint32_t* __restrict__ q, int nq, const uint32_t* __restrict__ t, int nt,
                                             int* __restrict__ best_idx, int* __restrict__ best_d, int* __restrict__ second_d,
                                             int qblock) {
    __shared__ uint32_t tile[8][BF_TILE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = (qblock * BF_WARPS + warp) * BF_QPW;
    uint32_t qr[BF_QPW][8];
#pragma unroll
    for (int k = 0; k < BF_QPW; k++)
#pragma unroll
        for (int w = 0; w < 8; w++) qr[k][w] = (q0 + k < nq) ? q[(size_t)(q0 + k) * 8 + w] : 0u;
    int b[BF_QPW], bi[BF_QPW], s[BF_QPW];
#pragma unroll
    for (int k = 0; k < BF_QPW; k++) { b[k] = 256; bi[k] = INT_MAX; s[k] = 256; }
    for (int t0 = 0; t0 < nt; t0 += BF_TILE) {
        const int m = min(BF_TILE, nt - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < m * 8; i += BF_WARPS * 32) {      // coalesced global read, transposed smem write
            const int j = i >> 3, w = i & 7;
            tile[w][j] = t[(size_t)(t0 + j) * 8 + w];
        }
        __syncthreads();
        for (int j = lane; j < m; j += 32) {
            uint32_t tw[8];
#pragma unroll
            for (int w = 0; w < 8; w++) tw[w] = tile[w][j];
#pragma unroll
            for (int k = 0; k < BF_QPW; k++) {
                int d = 0;
#pragma unroll
                for (int w = 0; w < 8; w++) d += __popc(qr[k][w] ^ tw[w]);
                // lanes see targets in increasing index order, so `<` keeps the first index on ties
                if (d < b[k]) { s[k] = b[k]; b[k] = d; bi[k] = t0 + j; }
                else if (d < s[k]) s[k] = d;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < BF_QPW; k++) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, b[k], o);
            const int obi = __shfl_xor_sync(0xffffffffu, bi[k], o);
            const int os = __shfl_xor_sync(0xffffffffu, s[k], o);
            merge_top2(b[k], bi[k], s[k], ob, obi, os);
        }
        if (lane == 0 && q0 + k < nq) {
            best_idx[q0 + k] = (b[k] < 256) ? bi[k] : -1;
            best_d[q0 + k] = b[k];
            second_d[q0 + k] = s[k];
        }
    }
}

}  // namespace ygz_b200
