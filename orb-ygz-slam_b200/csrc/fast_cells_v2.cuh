// k_fast_cells_v2 — the FAST-9/16 cell detector of ComputeKeyPointsOctTree (/root/reference/src/ORBextractor.cc:725-781:
// cv::FAST(cell, iniTh, true) and, for cells it leaves empty, cv::FAST(cell, minTh, true)) for cells whose detection
// window is at most 32 columns wide (every cell of the C1 / C5 configurations except C1's coarsest level; the rest go
// through k_fast_cells, the generic kernel in orb_kernels.cuh).  Same results as k_fast_cells, bit for bit.
//
// What changed against k_fast_cells (round-1 profile: 106 thread-instructions per pixel, 80 % issue-bound):
//   * persistent warps: each warp walks a strided list of (frame, cell) items; the cell image is fetched by ONE
//     cp.async.bulk.tensor (TMA) box of 64 x (ch+1) bytes per cell straight into the warp's tile — no load loop, and
//     the next cell's box is requested as soon as the current tile has been scored, so it lands while the warp does
//     NMS / emission;
//   * constant tile pitch (64): every ring offset is an immediate of the LDS;
//   * pre-test: the packed vertical pairs of the ring rows are REUSED down the column — the pair (dy=+3) of step k is
//     the pair (dy=-3) of step k+3, the pairs (dy=+2, dx=+-2) of step k are the pairs (dy=-2, dx=+-2) of step k+2 — so
//     a step loads 6 pairs instead of 9;
//   * the FAST-9 score pass appends the true corners (12 % of the pixels on the bench stream, against 19 % pre-test
//     survivors) to a second list, in place, and NMS + ordered emission run once over that list in one fused loop;
//   * the score map is cleared sparsely (only the entries a cell wrote) instead of a dense memset per cell.
#pragma once
#include "orb_kernels.cuh"

namespace ygz_b200 {

constexpr int F2_WARPS = 8;
constexpr int F2_THREADS = F2_WARPS * 32;
constexpr int F2_PITCH = 64;                    // TMA box width = tile pitch: the box must start on a 16-byte boundary in global
                                                // memory, so it begins at x0 & ~15 and spans 15 + 38 (cell image) columns
constexpr int F2_BOXH_MAX = 44;                 // box rows = ch + 1 (the packed pre-test reads one row past the cell)
constexpr int F2_MAX_DW = 32;                   // detection window columns (one lane each)
constexpr int F2_MAX_CH = F2_BOXH_MAX - 1;      // 43 -> dh <= 37
constexpr int F2_TILE_BYTES = F2_PITCH * F2_BOXH_MAX;   // 2816 = 22 * 128 (TMA destination alignment)
constexpr int F2_SCP = 36;                      // score-map pitch (dw + 2 <= 34)
constexpr int F2_SC_BYTES = F2_SCP * 40;        // (dh + 2 <= 39) rows
constexpr int F2_ITEMS_BYTES = 2 * F2_MAX_DW * 38;      // u16 per pre-test survivor, worst case every pixel of dh (rounded up to even) rows
constexpr int F2_WARP_SMEM = 6784;              // 2816 + 1440 + 2432 + 8 (mbarrier), rounded to 128; 8 warps x 4 CTAs fit one SM

struct FastMaps {                               // per level: raw u8 [frames][h][pitch], box F2_PITCH x boxh[l] x 1
    CUtensorMap level[MAX_LEVELS];
    unsigned char boxh[MAX_LEVELS];
};

__device__ __forceinline__ unsigned f2_pack(unsigned lo, unsigned hi) {
    unsigned r;
    asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(r) : "r"(hi), "r"(lo));     // FMA pipe: the ALU pipe is the busier one here
    return r;
}

__global__ void __launch_bounds__(F2_THREADS, 4) k_fast_cells_v2(const __grid_constant__ OrbParams P,
                                                                const __grid_constant__ FastMaps maps,
                                                                const CellDesc* __restrict__ cells,
                                                                const int* __restrict__ ids, int n_ids, int n_frames,
                                                                int n_static, int* __restrict__ queue,
                                                                uint32_t* __restrict__ pool,
                                                                uint16_t* __restrict__ cell_count) {
    extern __shared__ __align__(128) uint8_t f2_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* const tile = f2_smem + (size_t)warp * F2_WARP_SMEM;
    uint8_t* const sc = tile + F2_TILE_BYTES;
    uint16_t* const items = reinterpret_cast<uint16_t*>(sc + F2_SC_BYTES);
    unsigned long long* const bar = reinterpret_cast<unsigned long long*>(sc + F2_SC_BYTES + F2_ITEMS_BYTES);
    constexpr int TP = F2_PITCH;
    const unsigned sc_s = smem_u32(sc), items_s = smem_u32(items);

    // Work items (frame, cell), frame-major.  The first n_static items are dealt to the warps round-robin; the rest are
    // drawn from a global counter (one atomic per item) so that no warp is left with a long tail: cells differ in cost
    // by the texture they cover, and a static deal left a quarter of the warp slots idle at the end of the launch.
    const int n_items = n_ids * n_frames;
    const int stride = gridDim.x * F2_WARPS;
    int item = blockIdx.x * F2_WARPS + warp;
    auto next_item = [&](int cur) {
        int nx = cur + stride;
        if (nx >= n_static) {
            if (lane == 0) nx = n_static + atomicAdd(queue, 1);
            nx = __shfl_sync(0xffffffffu, nx, 0);
        }
        return nx;
    };
    if (item >= n_static) item = next_item(n_static);       // fewer static items than warps: this one starts on the queue
    if (item >= n_items) return;

    if (lane == 0) mbar_init(bar, 1);
    for (int i = lane; i < F2_SC_BYTES / 4; i += 32) reinterpret_cast<uint32_t*>(sc)[i] = 0u;
    __syncwarp();

    // one lane: request the box of work item i (frame-major order: neighbouring warps read neighbouring cells)
    auto issue = [&](int i) {
        if (lane == 0) {
            const int f = i / n_ids;
            const CellDesc c = cells[ids[i - f * n_ids]];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // the tile was read through the generic proxy
            mbar_expect_tx(bar, TP * (int)maps.boxh[c.level]);
            tma_load_3d(tile, &maps.level[c.level], c.x0 & ~15, c.y0, f, bar);
        }
    };
    int phase = 0;
    issue(item);

    // ring offsets (cv::FAST offsets16) with the constant pitch
    constexpr int o0 = 3 * TP, o1 = 3 * TP + 1, o2 = 2 * TP + 2, o3 = TP + 3, o4 = 3, o5 = -TP + 3, o6 = -2 * TP + 2,
                  o7 = -3 * TP + 1, o8 = -3 * TP, o9 = -3 * TP - 1, o10 = -2 * TP - 2, o11 = -TP - 3, o12 = -3,
                  o13 = TP - 3, o14 = 2 * TP - 2, o15 = 3 * TP - 1;

    while (item < n_items) {
        const int nxt = next_item(item);
        const int frame = item / n_ids;
        const int cell = ids[item - frame * n_ids];
        const CellDesc cd = cells[cell];
        const int dw = cd.cw - 6, dh = cd.ch - 6;
        const uint8_t* const t00 = tile + 3 * TP + 3 + (cd.x0 & 15);       // pixel (0,0) of the detection window
        const bool has_next = nxt < n_items;
        mbar_wait(bar, phase);
        phase ^= 1;
        bool prefetched = false;
        uint32_t* out = pool + (size_t)frame * P.pool_size + cd.out_off;
        int nkeep = 0, ncorn = 0;

        for (int pass = 0; pass < 2; pass++) {
            if (pass == 1 && prefetched) {
                // rare: corners at iniTh, but NMS kept none (equal-score neighbours).  The tile already belongs to the next
                // cell: let that box land, fetch this cell's box again, and request the next one again after scoring.
                mbar_wait(bar, phase); phase ^= 1;
                __syncwarp();
                issue(item);
                mbar_wait(bar, phase); phase ^= 1;
                prefetched = false;
            }
            const int tmin = pass ? P.min_th : P.ini_th;
            // ---- phase A: opposite-pair pre-test, two detection rows per step packed in the halves of a register
            //      (see k_fast_cells for the arithmetic); ordered ballot compaction keeps raster order
            int total = 0;
            {
                const unsigned K3 = (0x7FFFu - (unsigned)tmin) * 0x00010001u;
                const unsigned lane_mask = lane < dw ? 0x80008000u : 0u;
                const uint8_t* p = t00 + lane;
#define F2_LD2(o) f2_pack(p[(o)], p[(o) + TP])
                unsigned q8a = F2_LD2(-3 * TP), q8b = F2_LD2(-TP), q8c = F2_LD2(TP);            // rows (-3,-2) (-1,0) (1,2), dx 0
                unsigned q6a = F2_LD2(-2 * TP + 2), q6b = F2_LD2(2);                             // rows (-2,-1) (0,1), dx +2
                unsigned q10a = F2_LD2(-2 * TP - 2), q10b = F2_LD2(-2);                          // rows (-2,-1) (0,1), dx -2
                unsigned itv = (unsigned)lane;                                                   // x | y << 6
                int n1 = 0;
#pragma unroll 6
                for (int y = 0; y < dh; y += 2, p += 2 * TP, itv += 128u) {
                    const unsigned c2 = F2_LD2(0), A4 = F2_LD2(o4), A12 = F2_LD2(o12);
                    const unsigned A0 = F2_LD2(o0), A2 = F2_LD2(o2), A14 = F2_LD2(o14);
                    const unsigned A8 = q8a, A6 = q6a, A10 = q10a;
                    q8a = q8b; q8b = q8c; q8c = A0;
                    q6a = q6b; q6b = A2;
                    q10a = q10b; q10b = A14;
                    const unsigned bmin = __vminu2(__vminu2(__vmaxu2(A0, A8), __vmaxu2(A4, A12)), __vminu2(__vmaxu2(A2, A10), __vmaxu2(A6, A14)));
                    const unsigned dmax = __vmaxu2(__vmaxu2(__vminu2(A0, A8), __vminu2(A4, A12)), __vmaxu2(__vminu2(A2, A10), __vminu2(A6, A14)));
                    const unsigned z = ((bmin + K3 - c2) | (c2 + K3 - dmax)) & lane_mask;
                    const bool p0 = (z & 0x8000u) != 0u, p1 = (int)z < 0;
                    unsigned bal0, bal1;                        // the warp is converged here: two plain VOTEs, one predicate each
                    asm volatile("{\n\t.reg .pred q0, q1;\n\t.reg .b32 t;\n\t"
                                 "and.b32 t, %2, 0x8000;\n\tsetp.ne.u32 q0, t, 0;\n\tsetp.lt.s32 q1, %2, 0;\n\t"
                                 "vote.sync.ballot.b32 %0, q0, 0xffffffff;\n\tvote.sync.ballot.b32 %1, q1, 0xffffffff;\n\t}"
                                 : "=r"(bal0), "=r"(bal1) : "r"(z));
                    const int n0 = __popc(bal0);
                    n1 = __popc(bal1);
                    if (p0) items[total + __popc(bal0 & lanemask_lt())] = (uint16_t)itv;
                    if (p1) items[total + n0 + __popc(bal1 & lanemask_lt())] = (uint16_t)(itv + 64u);
                    total += n0 + n1;
                }
#undef F2_LD2
                // odd dh: the last step also tested row dh, one past the window; its survivors are the tail of the list
                if (dh & 1) total -= n1;
            }
            __syncwarp();

            // ---- phase B: FAST-9 score of the survivors; true corners are written to the score map and appended, in
            //      order and in place (slot index <= item index), to the corner list
            ncorn = 0;
            for (int g0 = 0; g0 < total; g0 += 32) {
                const int g = g0 + lane;
                const bool valid = g < total;
                const int it = valid ? items[g] : 0;
                __syncwarp();                                   // every lane holds its item before any slot is rewritten
                const int x = it & 63, y = it >> 6;
                const uint8_t* q = t00 + y * TP + x;
                const int r[16] = {q[o0], q[o1], q[o2], q[o3], q[o4], q[o5], q[o6], q[o7],
                                   q[o8], q[o9], q[o10], q[o11], q[o12], q[o13], q[o14], q[o15]};
                const int best = fast9_best_packed(q[0], r);
                const bool corner = valid && best > tmin;       // response = best - 1
                // predicated stores on 32-bit shared addresses (the generic-pointer form cost 12 instructions + a branch per group)
                const unsigned bal = __ballot_sync(0xffffffffu, corner);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %0, 0;\n\t@p st.shared.u8 [%1], %2;\n\t@p st.shared.u16 [%3], %4;\n\t}"
                             :: "r"((int)corner), "r"(sc_s + (unsigned)(y * F2_SCP + x + F2_SCP + 1)), "r"(best - 1),
                                "r"(items_s + 2u * (unsigned)(ncorn + __popc(bal & lanemask_lt()))), "r"(it) : "memory");
                ncorn += __popc(bal);
            }
            __syncwarp();
            // the tile is dead unless the minTh pass may still be needed (only when iniTh found no corner at all)
            if (!prefetched && has_next && (pass == 1 || ncorn > 0)) { issue(nxt); prefetched = true; }

            // ---- phase C+D: 3x3 non-max suppression inside the detection window (strict >, zeros outside) and ordered
            //      emission, coordinates relative to (16,16)
            nkeep = 0;
            for (int g0 = 0; g0 < ncorn; g0 += 32) {
                const int g = g0 + lane;
                const bool valid = g < ncorn;
                const int it = valid ? items[g] : 0;
                const int x = it & 63, y = it >> 6;
                const uint8_t* q = sc + (y + 1) * F2_SCP + x + 1;
                const int s = q[0];
                const bool keep = valid && s > 0 &&
                                  ((s > q[-1]) & (s > q[1]) & (s > q[-F2_SCP - 1]) & (s > q[-F2_SCP]) & (s > q[-F2_SCP + 1]) &
                                   (s > q[F2_SCP - 1]) & (s > q[F2_SCP]) & (s > q[F2_SCP + 1]));
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int pos = nkeep + __popc(bal & lanemask_lt());
                    if (pos < cd.cap) out[pos] = pack_cand(x + 3 + cd.offx, y + 3 + cd.offy, s);
                }
                nkeep += __popc(bal);
            }
            __syncwarp();       // every lane's neighbour reads of the score map are done before any entry is cleared below
            // vKeysCell.empty() after the iniTh call decides the fallback to minTh (src/ORBextractor.cc:765-770)
            if (nkeep > 0 || pass == 1) break;
            // the minTh pass recomputes every score it needs: drop this pass's entries (minTh > iniTh would leave stale ones)
            for (int g = lane; g < ncorn; g += 32) { const int it = items[g]; sc[((it >> 6) + 1) * F2_SCP + (it & 63) + 1] = 0; }
            __syncwarp();
        }
        if (!prefetched && has_next) { __syncwarp(); issue(nxt); }
        if (lane == 0) cell_count[(size_t)frame * P.n_cells + cell] = (uint16_t)min(nkeep, (int)cd.cap);
        // leave the score map clear for the next cell
        for (int g = lane; g < ncorn; g += 32) { const int it = items[g]; sc[((it >> 6) + 1) * F2_SCP + (it & 63) + 1] = 0; }
        __syncwarp();
        item = nxt;
    }
}

}  // namespace ygz_b200
