// ORBmatcher Hamming-distance search on the device + the host-side greedy/ratio/rotation passes that
// reproduce the reference's sequential semantics (/root/reference/src/ORBmatcher.cc).
//
//   k_hamming_bf_top2   warp-per-query 256-bit popcount brute force, best / second best       (:1507-1523 inner loop)
//   k_hamming_pairs     candidate-list distances (BoW bucket, grid window)                     (:197-214, :86-113, :1289-1311)
//   k_bow_match_batch   whole SearchByBoW(KF,F) on the device for a batch of frame pairs       (:155-263)
//
// The Search* entry points walk flat views of the pointer graph on the host (candidate enumeration through the
// 64x48 frame grid of src/Frame.cc:314-330,424-493), compute every candidate distance in one launch, then run the
// order-dependent accept logic on the host exactly in the reference's order.
#include "common.cuh"
#include "hamming_bf.cuh"
#include <algorithm>
#include <climits>
#include <cmath>
#include <vector>

using namespace ygz_b200;

namespace {

constexpr int TH_HIGH = 100;        // src/ORBmatcher.cc:36
constexpr int TH_LOW = 50;          // :37
constexpr int HISTO_LENGTH = 30;    // :38
constexpr int GRID_COLS = 64;       // include/Frame.h:38
constexpr int GRID_ROWS = 48;       // include/Frame.h:39

// ---------------------------------------------------------------------------------------------- device
__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__global__ void __launch_bounds__(BF_WARPS * 32) k_hamming_bf_top2(const uint32_t* __restrict__ q, int nq,
                                                                   const uint32_t* __restrict__ t, int nt,
                                                                   int* __restrict__ best_idx, int* __restrict__ best_d,
                                                                   int* __restrict__ second_d) {
    bf_top2_body(q, nq, t, nt, best_idx, best_d, second_d, blockIdx.x);
}

__global__ void __launch_bounds__(256) k_hamming_pairs(const uint4* __restrict__ q, const uint4* __restrict__ t,
                                                       const int* __restrict__ pq, const int* __restrict__ pt, int n,
                                                       int* __restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* a = q + (size_t)pq[i] * 2;
    const uint4* b = t + (size_t)pt[i] * 2;
    dist[i] = hamming256(__ldg(a), __ldg(a + 1), __ldg(b), __ldg(b + 1));
}

// ---- batched device-resident SearchByBoW (src/ORBmatcher.cc:155-263)
// One CTA per (KF, F) pair.  Features of both frames are sorted by (vocabulary node, feature index) — the
// iteration order of two DBoW2::FeatureVector maps whose per-node vectors hold ascending feature indices — then
// one warp per common node runs the reference's sequential loop over that node's KF features with the F
// candidates spread over the lanes (a frame feature belongs to exactly one node, so nodes are independent).
constexpr int BOW_THREADS = 512;     // measured (C2, 256 pairs): 256 threads 0.238 ms, 512 0.152 ms, 1024 0.187 ms
constexpr int BOW_MAX_FEATS = 2048;

__device__ inline void bitonic_sort_u64(unsigned long long* a, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = a[i], y = a[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
}

struct BowBatchArgs {
    const ygz_keypoint* kps;     // [frames][cap]
    const uint4* desc;           // [frames][cap][2]
    const int* counts;           // [frames]
    const uint32_t* node;        // [frames][cap]
    int cap;
    const int* kf_idx;           // [pairs]
    const int* f_idx;
    float nnratio;
    int check_ori;
    int* matches;                // [pairs][cap]
    int* nmatches;               // [pairs]
};

// dynamic shared memory of k_bow_match_batch for a per-frame capacity `cap` (features), cap2 = next power of two
__host__ __device__ inline size_t bow_smem_bytes(int cap) {
    int p2 = 1; while (p2 < cap) p2 <<= 1;
    return (size_t)p2 * 8 * 2        // sorted (node, index) keys of both frames
         + (size_t)cap * 32 * 2      // descriptors of both frames
         + (size_t)cap * 2 * 2       // match_of_f (short) + seg_start (short)
         + (size_t)cap               // bin_of_f
         + 64;
}

__global__ void __launch_bounds__(BOW_THREADS) k_bow_match_batch(BowBatchArgs A) {
    extern __shared__ __align__(16) uint8_t bow_smem[];
    __shared__ int hist[HISTO_LENGTH];
    __shared__ int s_keep[3];
    __shared__ int s_n;
    __shared__ int s_wcnt[BOW_THREADS / 32];
    __shared__ int s_nseg;
    __shared__ int s_next;
    const int cap = A.cap;
    int p2 = 1; while (p2 < cap) p2 <<= 1;
    unsigned long long* skf = reinterpret_cast<unsigned long long*>(bow_smem);   // (node << 32) | feature index, sorted
    unsigned long long* sf = skf + p2;
    uint4* kds = reinterpret_cast<uint4*>(sf + p2);                                // KF descriptors
    uint4* fds = kds + (size_t)cap * 2;                                            // F descriptors
    short* match_of_f = reinterpret_cast<short*>(fds + (size_t)cap * 2);           // KF feature matched to F feature (or -1)
    short* seg_start = match_of_f + cap;                                           // first sorted position of every KF node
    signed char* bin_of_f = reinterpret_cast<signed char*>(seg_start + cap);
    const int pair = blockIdx.x;
    const int kf = A.kf_idx[pair], fr = A.f_idx[pair];
    const int nk = min(A.counts[kf], cap), nf = min(A.counts[fr], cap);
    const ygz_keypoint* kkp = A.kps + (size_t)kf * A.cap;
    const ygz_keypoint* fkp = A.kps + (size_t)fr * A.cap;
    const uint4* kd = A.desc + (size_t)kf * A.cap * 2;
    const uint4* fd = A.desc + (size_t)fr * A.cap * 2;
    int pk = 1; while (pk < nk) pk <<= 1;
    int pf = 1; while (pf < nf) pf <<= 1;
    // stage both descriptor sets (coalesced 16-byte loads); every later access is a shared-memory read
    for (int i = threadIdx.x; i < nk * 2; i += BOW_THREADS) kds[i] = __ldg(kd + i);
    for (int i = threadIdx.x; i < nf * 2; i += BOW_THREADS) fds[i] = __ldg(fd + i);
    __syncthreads();
    // A.node == nullptr: synthetic vocabulary (bench stand-in for DBoW2 transform): node = top 6 bits of byte 0
    auto node_of = [&](int frame, const uint4* d, int i) -> unsigned long long {
        return A.node ? (unsigned long long)A.node[(size_t)frame * A.cap + i]
                      : (unsigned long long)((reinterpret_cast<const uint8_t*>(d + (size_t)i * 2)[0]) >> 2);
    };
    for (int i = threadIdx.x; i < pk; i += BOW_THREADS)
        skf[i] = i < nk ? ((node_of(kf, kds, i) << 32) | (unsigned)i) : ~0ull;
    for (int i = threadIdx.x; i < pf; i += BOW_THREADS)
        sf[i] = i < nf ? ((node_of(fr, fds, i) << 32) | (unsigned)i) : ~0ull;
    for (int i = threadIdx.x; i < cap; i += BOW_THREADS) { match_of_f[i] = -1; bin_of_f[i] = -1; }
    if (threadIdx.x < HISTO_LENGTH) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_n = 0; s_nseg = 0; s_next = 0; }
    __syncthreads();
    bitonic_sort_u64(skf, pk);
    bitonic_sort_u64(sf, pf);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = BOW_THREADS / 32;
    // ---- KF node segments: ordered compaction of the positions where the node id changes
    for (int base = 0; base < nk; base += BOW_THREADS) {
        const int i = base + threadIdx.x;
        const bool st = i < nk && (uint32_t)(skf[i] >> 32) != 0xFFFFFFFFu && (i == 0 || (uint32_t)(skf[i] >> 32) != (uint32_t)(skf[i - 1] >> 32));
        const unsigned bal = __ballot_sync(0xffffffffu, st);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int off = s_nseg;
        for (int w = 0; w < warp; w++) off += s_wcnt[w];
        if (st) seg_start[off + __popc(bal & lanemask_lt())] = (short)i;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < nwarps; w++) t += s_wcnt[w]; s_nseg += t; }
        __syncthreads();
    }
    const int nseg = s_nseg;
    // ---- one warp per common node: the reference's sequential loop over that node's KF features (:180-234)
    // nodes are drawn from a counter: bucket sizes are skewed, a static deal left the warps waiting at the barrier below (22 % of the
    // kernel's stall samples); nodes are independent of each other, so the order does not matter
    for (;;) {
        int sgm = 0;
        if (lane == 0) sgm = atomicAdd(&s_next, 1);
        sgm = __shfl_sync(0xffffffffu, sgm, 0);
        if (sgm >= nseg) break;
        const int i0 = seg_start[sgm];
        const uint32_t node = (uint32_t)(skf[i0] >> 32);
        int i1 = (sgm + 1 < nseg) ? (int)seg_start[sgm + 1] : nk;
        while (i1 > i0 && (uint32_t)(skf[i1 - 1] >> 32) != node) i1--;      // trailing 0xFFFFFFFF (no node) entries
        int lo = 0, hi = nf;                                                 // F segment with the same node (lower_bound)
        while (lo < hi) { const int mid = (lo + hi) >> 1; if ((uint32_t)(sf[mid] >> 32) < node) lo = mid + 1; else hi = mid; }
        int j1 = lo;
        while (j1 < nf && (uint32_t)(sf[j1] >> 32) == node) j1++;
        const int j0 = lo;
        for (int i = i0; i < i1 && j1 > j0; i++) {
            const int realIdxKF = (int)(skf[i] & 0xFFFFFFFFu);
            const uint4 a0 = kds[realIdxKF * 2], a1 = kds[realIdxKF * 2 + 1];
            int b = 256, bi = INT_MAX, s = 256;
            for (int j = j0 + lane; j < j1; j += 32) {
                const int realIdxF = (int)(sf[j] & 0xFFFFFFFFu);
                if (match_of_f[realIdxF] >= 0) continue;                      // :200-201
                const int d = hamming256(a0, a1, fds[realIdxF * 2], fds[realIdxF * 2 + 1]);
                // order inside the bucket = position j (ascending); keep position for the tie-break
                if (d < b) { s = b; b = d; bi = j; }
                else if (d < s) s = d;
            }
            warp_top2(b, bi, s);                                                  // two REDUX (the 15-shuffle merge was the loop's largest part)
            if (lane == 0 && b <= TH_LOW && (float)b < A.nnratio * (float)s) {    // :216-217
                const int bestIdxF = (int)(sf[bi] & 0xFFFFFFFFu);
                match_of_f[bestIdxF] = (short)realIdxKF;
                if (A.check_ori) {
                    float rot = kkp[realIdxKF].angle - fkp[bestIdxF].angle;        // :221-226
                    if (rot < 0.0f) rot += 360.0f;
                    int bin = (int)roundf(rot * (1.0f / HISTO_LENGTH));
                    if (bin == HISTO_LENGTH) bin = 0;
                    bin_of_f[bestIdxF] = (signed char)bin;
                    atomicAdd(&hist[bin], 1);
                }
                atomicAdd(&s_n, 1);
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int i1 = -1, i2 = -1, i3 = -1;
        if (A.check_ori) {                      // ComputeThreeMaxima :1471-1502
            int max1 = 0, max2 = 0, max3 = 0;
            for (int i = 0; i < HISTO_LENGTH; i++) {
                const int sz = hist[i];
                if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; i3 = i2; i2 = i1; i1 = i; }
                else if (sz > max2) { max3 = max2; max2 = sz; i3 = i2; i2 = i; }
                else if (sz > max3) { max3 = sz; i3 = i; }
            }
            if (max2 < 0.1f * (float)max1) { i2 = -1; i3 = -1; }
            else if (max3 < 0.1f * (float)max1) { i3 = -1; }
        }
        s_keep[0] = i1; s_keep[1] = i2; s_keep[2] = i3;
    }
    __syncthreads();
    int removed = 0;
    int* out = A.matches + (size_t)pair * A.cap;
    for (int i = threadIdx.x; i < A.cap; i += BOW_THREADS) {
        int m = i < nf ? match_of_f[i] : -1;
        if (m >= 0 && A.check_ori) {
            const int bin = bin_of_f[i];
            if (bin != s_keep[0] && bin != s_keep[1] && bin != s_keep[2]) { m = -1; removed++; }
        }
        out[i] = m;
    }
    if (removed) atomicSub(&s_n, removed);
    __syncthreads();
    if (threadIdx.x == 0) A.nmatches[pair] = s_n;
}

// ---------------------------------------------------------------------------------------------- host helpers
struct Grid {     // Frame::AssignFeaturesToGrid src/Frame.cc:314-330, PosInGrid :483-493
    std::vector<int> cell[GRID_COLS][GRID_ROWS];
    float invW, invH;
    explicit Grid(const ygz_frame_view* f) {
        invW = static_cast<float>(GRID_COLS) / (f->max_x - f->min_x);
        invH = static_cast<float>(GRID_ROWS) / (f->max_y - f->min_y);
        for (int i = 0; i < f->n; i++) {
            const int px = (int)std::round((f->keys[i].x - f->min_x) * invW);
            const int py = (int)std::round((f->keys[i].y - f->min_y) * invH);
            if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
            cell[px][py].push_back(i);
        }
    }
};

// Frame::GetFeaturesInArea src/Frame.cc:424-481 — appends to `out`
void features_in_area(const ygz_frame_view* f, const Grid& g, float x, float y, float r, int minLevel, int maxLevel,
                      std::vector<int>& out) {
    const int nMinCellX = std::max(0, (int)std::floor((x - f->min_x - r) * g.invW));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - f->min_x + r) * g.invW));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)std::floor((y - f->min_y - r) * g.invH));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - f->min_y + r) * g.invH));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
            for (int idx : g.cell[ix][iy]) {
                const ygz_keypoint& kp = f->keys[idx];
                if (bCheckLevels) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                if (std::fabs(kp.x - x) < r && std::fabs(kp.y - y) < r) out.push_back(idx);
            }
}

void three_maxima(const int* sz, int L, int& ind1, int& ind2, int& ind3) {   // src/ORBmatcher.cc:1471-1502
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

inline int rot_bin(float a1, float a2) {      // :221-226
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)std::round(rot * (1.0f / HISTO_LENGTH));
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

struct RotHist {
    std::vector<int> bins[HISTO_LENGTH];
    void add(int bin, int v) { bins[bin].push_back(v); }
    // calls drop(v) for every entry outside the three fullest bins
    template <class F> void prune(F drop) {
        int sz[HISTO_LENGTH], i1 = -1, i2 = -1, i3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) sz[i] = (int)bins[i].size();
        three_maxima(sz, HISTO_LENGTH, i1, i2, i3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int v : bins[i]) drop(v);
        }
    }
};

inline bool mp_bad(const ygz_map_points* m, int i) { return m->bad && m->bad[i]; }
inline int mp_obs(const ygz_map_points* m, int i) { return m->observations ? m->observations[i] : 1; }
inline float uright(const ygz_frame_view* f, int i) { return f->u_right ? f->u_right[i] : -1.f; }

}  // namespace

#include "projection_batch.cuh"

struct ygz_matcher {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int max_q = 0, max_t = 0, max_pairs = 0;
    uint8_t *d_q = nullptr, *d_t = nullptr;
    int *d_pq = nullptr, *d_pt = nullptr, *d_dist = nullptr;
    int *d_bi = nullptr, *d_bd = nullptr, *d_sd = nullptr;
    // batch BoW
    int *d_kf_idx = nullptr, *d_f_idx = nullptr, *d_matches = nullptr, *d_nmatches = nullptr;
    uint32_t* d_node = nullptr;
    size_t node_elems = 0;              // capacity of d_node in elements (frames x cap)
    ygz_vocab* vocab = nullptr;         // attached vocabulary: node ids come from k_vocab_transform on the device
    int levelsup = 4;
    int bow_pairs_cap = 0, bow_cap = 0, bow_frames = 0;
    BowBatchArgs last_args{};
    int last_pairs = 0;
    // batched projection searches: device scratch slots (grown on demand) and the last launch (for the timing entry)
    void* pj_buf[16] = {};
    size_t pj_bytes[16] = {};
    ProjArgs pj_last{};
    int pj_last_pairs = 0, pj_last_mode = -1, pj_last_nq = 0;
};

static int matcher_free(ygz_matcher* h) {
    if (!h) return YGZ_OK;
    cudaSetDevice(h->device);
    cudaFree(h->d_q); cudaFree(h->d_t); cudaFree(h->d_pq); cudaFree(h->d_pt); cudaFree(h->d_dist);
    cudaFree(h->d_bi); cudaFree(h->d_bd); cudaFree(h->d_sd);
    cudaFree(h->d_kf_idx); cudaFree(h->d_f_idx); cudaFree(h->d_matches); cudaFree(h->d_nmatches); cudaFree(h->d_node);
    for (void* b : h->pj_buf) cudaFree(b);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return YGZ_OK;
}

// distances of all (query, target) pairs in one launch; host buffers in, host dist out
static int pair_distances(ygz_matcher* h, const uint8_t* q, int nq, const uint8_t* t, int nt, const std::vector<int>& pq,
                          const std::vector<int>& pt, std::vector<int>& dist) {
    const int n = (int)pq.size();
    dist.resize(n);
    if (n == 0) return YGZ_OK;
    return ygz_hamming_pairs(h, q, nq, t, nt, pq.data(), pt.data(), n, dist.data());
}

extern "C" {

int ygz_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t x, y;
        std::memcpy(&x, a + 4 * i, 4);
        std::memcpy(&y, b + 4 * i, 4);
        d += __builtin_popcount(x ^ y);
    }
    return d;
}

int ygz_matcher_create(int device, int max_queries, int max_targets, int max_pairs, ygz_matcher** out) {
    YGZ_ARG_CHECK(out && max_queries >= 1 && max_targets >= 1 && max_pairs >= 1);
    if (ygz_device_count() <= device) { set_last_error("no such CUDA device"); return YGZ_ERR_NO_DEVICE; }
    YGZ_CUDA_TRY(cudaSetDevice(device));
    ygz_matcher* h = new ygz_matcher();
    h->device = device; h->max_q = max_queries; h->max_t = max_targets; h->max_pairs = max_pairs;
#define TRY_OR_FREE(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); matcher_free(h); return YGZ_ERR_CUDA; } } while (0)
    TRY_OR_FREE(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    TRY_OR_FREE(cudaEventCreate(&h->ev0));
    TRY_OR_FREE(cudaEventCreate(&h->ev1));
    TRY_OR_FREE(cudaMalloc(&h->d_q, (size_t)max_queries * 32));
    TRY_OR_FREE(cudaMalloc(&h->d_t, (size_t)max_targets * 32));
    TRY_OR_FREE(cudaMalloc(&h->d_pq, (size_t)max_pairs * 4));
    TRY_OR_FREE(cudaMalloc(&h->d_pt, (size_t)max_pairs * 4));
    TRY_OR_FREE(cudaMalloc(&h->d_dist, (size_t)max_pairs * 4));
    TRY_OR_FREE(cudaMalloc(&h->d_bi, (size_t)max_queries * 4));
    TRY_OR_FREE(cudaMalloc(&h->d_bd, (size_t)max_queries * 4));
    TRY_OR_FREE(cudaMalloc(&h->d_sd, (size_t)max_queries * 4));
#undef TRY_OR_FREE
    *out = h;
    return YGZ_OK;
}

int ygz_matcher_destroy(ygz_matcher* h) { return matcher_free(h); }

int ygz_hamming_bf_top2(ygz_matcher* h, const uint8_t* q, int nq, const uint8_t* t, int nt, int* best_idx, int* best_dist,
                        int* second_dist) {
    YGZ_ARG_CHECK(h && nq >= 0 && nt >= 0 && nq <= h->max_q && nt <= h->max_t);
    if (nq == 0) return YGZ_OK;
    YGZ_ARG_CHECK(q && best_idx && best_dist && second_dist && (nt == 0 || t));
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_q, q, (size_t)nq * 32, cudaMemcpyHostToDevice, h->stream));
    if (nt) YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_t, t, (size_t)nt * 32, cudaMemcpyHostToDevice, h->stream));
    const int qpb = BF_WARPS * BF_QPW;
    k_hamming_bf_top2<<<ceil_div(nq, qpb), BF_WARPS * 32, 0, h->stream>>>((const uint32_t*)h->d_q, nq, (const uint32_t*)h->d_t, nt,
                                                                         h->d_bi, h->d_bd, h->d_sd);
    YGZ_CUDA_TRY(cudaGetLastError());
    YGZ_CUDA_TRY(cudaMemcpyAsync(best_idx, h->d_bi, (size_t)nq * 4, cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(best_dist, h->d_bd, (size_t)nq * 4, cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(second_dist, h->d_sd, (size_t)nq * 4, cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return YGZ_OK;
}

int ygz_hamming_pairs(ygz_matcher* h, const uint8_t* q, int nq, const uint8_t* t, int nt, const int* pair_q, const int* pair_t,
                      int n_pairs, int* dist) {
    YGZ_ARG_CHECK(h && nq >= 0 && nt >= 0 && n_pairs >= 0 && nq <= h->max_q && nt <= h->max_t);
    if (n_pairs == 0) return YGZ_OK;
    YGZ_ARG_CHECK(q && t && pair_q && pair_t && dist);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_q, q, (size_t)nq * 32, cudaMemcpyHostToDevice, h->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_t, t, (size_t)nt * 32, cudaMemcpyHostToDevice, h->stream));
    for (int off = 0; off < n_pairs; off += h->max_pairs) {          // chunk through the pair buffers
        const int m = std::min(h->max_pairs, n_pairs - off);
        YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_pq, pair_q + off, (size_t)m * 4, cudaMemcpyHostToDevice, h->stream));
        YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_pt, pair_t + off, (size_t)m * 4, cudaMemcpyHostToDevice, h->stream));
        { NvtxRange nvtx_range("ygz/Hamming pairs");
        k_hamming_pairs<<<ceil_div(m, 256), 256, 0, h->stream>>>((const uint4*)h->d_q, (const uint4*)h->d_t, h->d_pq, h->d_pt, m, h->d_dist);
        }
        YGZ_CUDA_TRY(cudaGetLastError());
        YGZ_CUDA_TRY(cudaMemcpyAsync(dist + off, h->d_dist, (size_t)m * 4, cudaMemcpyDeviceToHost, h->stream));
        YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return YGZ_OK;
}

// ---- src/ORBmatcher.cc:155-263
int ygz_search_by_bow(ygz_matcher* h, float nnratio, int check_ori, const ygz_frame_view* kf, const ygz_map_points* mps,
                      const ygz_frame_view* f, int* out) {
    YGZ_ARG_CHECK(h && kf && mps && f && out);
    for (int i = 0; i < f->n; i++) out[i] = -1;
    // 1) enumerate every (KF feature, F feature) pair that shares a vocabulary node, in the reference's loop order
    struct Job { int kfIdx, first, count; };
    std::vector<Job> jobs;
    std::vector<int> pq, pt;
    int ki = 0, fi = 0;
    while (ki < kf->fv_nodes && fi < f->fv_nodes) {
        const uint32_t kn = kf->fv_node_id[ki], fn = f->fv_node_id[fi];
        if (kn == fn) {
            for (int a = kf->fv_offset[ki]; a < kf->fv_offset[ki + 1]; a++) {
                const int realIdxKF = (int)kf->fv_index[a];
                const int pMP = kf->map_point[realIdxKF];
                if (pMP < 0 || mp_bad(mps, pMP)) continue;
                Job j{realIdxKF, (int)pq.size(), f->fv_offset[fi + 1] - f->fv_offset[fi]};
                for (int b = f->fv_offset[fi]; b < f->fv_offset[fi + 1]; b++) { pq.push_back(realIdxKF); pt.push_back((int)f->fv_index[b]); }
                jobs.push_back(j);
            }
            ki++; fi++;
        } else if (kn < fn) {
            while (ki < kf->fv_nodes && kf->fv_node_id[ki] < fn) ki++;
        } else {
            while (fi < f->fv_nodes && f->fv_node_id[fi] < kn) fi++;
        }
    }
    // 2) all distances in one launch
    std::vector<int> dist;
    int rc = pair_distances(h, kf->desc, kf->n, f->desc, f->n, pq, pt, dist);
    if (rc) return rc;
    // 3) greedy accept in the reference's order
    int nmatches = 0;
    RotHist rh;
    for (const Job& j : jobs) {
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        for (int p = j.first; p < j.first + j.count; p++) {
            const int realIdxF = pt[p];
            if (out[realIdxF] >= 0) continue;
            const int d = dist[p];
            if (d < bestDist1) { bestDist2 = bestDist1; bestDist1 = d; bestIdxF = realIdxF; }
            else if (d < bestDist2) bestDist2 = d;
        }
        if (bestDist1 <= TH_LOW && static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
            out[bestIdxF] = kf->map_point[j.kfIdx];
            if (check_ori) rh.add(rot_bin(kf->keys[j.kfIdx].angle, f->keys[bestIdxF].angle), bestIdxF);
            nmatches++;
        }
    }
    if (check_ori) rh.prune([&](int idx) { out[idx] = -1; nmatches--; });
    return nmatches;
}

// ---- src/ORBmatcher.cc:480-595  SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)
int ygz_search_by_bow_kf(ygz_matcher* h, float nnratio, int check_ori, const ygz_frame_view* kf1, const ygz_frame_view* kf2,
                         const ygz_map_points* mps, int* matches12) {
    YGZ_ARG_CHECK(h && kf1 && kf2 && mps && matches12);
    for (int i = 0; i < kf1->n; i++) matches12[i] = -1;
    struct Job { int idx1, first, count; };
    std::vector<Job> jobs;
    std::vector<int> pq, pt;
    int i1n = 0, i2n = 0;
    while (i1n < kf1->fv_nodes && i2n < kf2->fv_nodes) {
        const uint32_t n1 = kf1->fv_node_id[i1n], n2 = kf2->fv_node_id[i2n];
        if (n1 == n2) {
            for (int a = kf1->fv_offset[i1n]; a < kf1->fv_offset[i1n + 1]; a++) {
                const int idx1 = (int)kf1->fv_index[a];
                const int pMP1 = kf1->map_point[idx1];
                if (pMP1 < 0 || mp_bad(mps, pMP1)) continue;
                Job j{idx1, (int)pq.size(), 0};
                for (int b = kf2->fv_offset[i2n]; b < kf2->fv_offset[i2n + 1]; b++) {
                    const int idx2 = (int)kf2->fv_index[b];
                    const int pMP2 = kf2->map_point[idx2];
                    if (pMP2 < 0 || mp_bad(mps, pMP2)) continue;            // state-independent part of :528-532
                    pq.push_back(idx1); pt.push_back(idx2); j.count++;
                }
                jobs.push_back(j);
            }
            i1n++; i2n++;
        } else if (n1 < n2) {
            while (i1n < kf1->fv_nodes && kf1->fv_node_id[i1n] < n2) i1n++;
        } else {
            while (i2n < kf2->fv_nodes && kf2->fv_node_id[i2n] < n1) i2n++;
        }
    }
    std::vector<int> dist;
    int rc = pair_distances(h, kf1->desc, kf1->n, kf2->desc, kf2->n, pq, pt, dist);
    if (rc) return rc;
    std::vector<char> vbMatched2(kf2->n, 0);
    int nmatches = 0;
    RotHist rh;
    for (const Job& j : jobs) {
        int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
        for (int p = j.first; p < j.first + j.count; p++) {
            const int idx2 = pt[p];
            if (vbMatched2[idx2]) continue;
            const int d = dist[p];
            if (d < bestDist1) { bestDist2 = bestDist1; bestDist1 = d; bestIdx2 = idx2; }
            else if (d < bestDist2) bestDist2 = d;
        }
        if (bestDist1 < TH_LOW && static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
            matches12[j.idx1] = kf2->map_point[bestIdx2];
            vbMatched2[bestIdx2] = 1;
            if (check_ori) rh.add(rot_bin(kf1->keys[j.idx1].angle, kf2->keys[bestIdx2].angle), j.idx1);
            nmatches++;
        }
    }
    if (check_ori) rh.prune([&](int idx) { matches12[idx] = -1; nmatches--; });
    return nmatches;
}

// ---- src/ORBmatcher.cc:597-746  SearchForTriangulation (+ CheckDistEpipolarLine :136-153)
int ygz_search_for_triangulation(ygz_matcher* h, int check_ori, const ygz_frame_view* kf1, const ygz_frame_view* kf2,
                                 const float* F12, float ex, float ey, const float* level_sigma2, int only_stereo, int* pairs) {
    YGZ_ARG_CHECK(h && kf1 && kf2 && F12 && level_sigma2 && pairs);
    struct Job { int idx1, first, count; };
    std::vector<Job> jobs;
    std::vector<int> pq, pt;
    int i1n = 0, i2n = 0;
    while (i1n < kf1->fv_nodes && i2n < kf2->fv_nodes) {
        const uint32_t n1 = kf1->fv_node_id[i1n], n2 = kf2->fv_node_id[i2n];
        if (n1 == n2) {
            for (int a = kf1->fv_offset[i1n]; a < kf1->fv_offset[i1n + 1]; a++) {
                const int idx1 = (int)kf1->fv_index[a];
                if (kf1->map_point[idx1] >= 0) continue;
                if (only_stereo && !(uright(kf1, idx1) >= 0)) continue;
                Job j{idx1, (int)pq.size(), 0};
                for (int b = kf2->fv_offset[i2n]; b < kf2->fv_offset[i2n + 1]; b++) {
                    const int idx2 = (int)kf2->fv_index[b];
                    if (kf2->map_point[idx2] >= 0) continue;
                    if (only_stereo && !(uright(kf2, idx2) >= 0)) continue;
                    pq.push_back(idx1); pt.push_back(idx2); j.count++;
                }
                jobs.push_back(j);
            }
            i1n++; i2n++;
        } else if (n1 < n2) {
            while (i1n < kf1->fv_nodes && kf1->fv_node_id[i1n] < n2) i1n++;
        } else {
            while (i2n < kf2->fv_nodes && kf2->fv_node_id[i2n] < n1) i2n++;
        }
    }
    std::vector<int> dist;
    int rc = pair_distances(h, kf1->desc, kf1->n, kf2->desc, kf2->n, pq, pt, dist);
    if (rc) return rc;
    std::vector<char> vbMatched2(kf2->n, 0);
    std::vector<int> vMatches12(kf1->n, -1);
    int nmatches = 0;
    RotHist rh;
    for (const Job& j : jobs) {
        const ygz_keypoint& kp1 = kf1->keys[j.idx1];
        const bool bStereo1 = uright(kf1, j.idx1) >= 0;
        int bestDist = TH_LOW, bestIdx2 = -1;
        for (int p = j.first; p < j.first + j.count; p++) {
            const int idx2 = pt[p];
            if (vbMatched2[idx2]) continue;
            const int d = dist[p];
            if (d > TH_LOW || d > bestDist) continue;
            const ygz_keypoint& kp2 = kf2->keys[idx2];
            const bool bStereo2 = uright(kf2, idx2) >= 0;
            if (!bStereo1 && !bStereo2) {
                const float distex = ex - kp2.x, distey = ey - kp2.y;
                if (distex * distex + distey * distey < 100 * kf2->scale_factors[kp2.octave]) continue;
            }
            // CheckDistEpipolarLine: l = x1' F12, distance of kp2 to l against the chi-square bound of its level
            const float a = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
            const float b = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
            const float c = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
            const float num = a * kp2.x + b * kp2.y + c;
            const float den = a * a + b * b;
            if (den == 0) continue;
            const float dsqr = num * num / den;
            if (dsqr < 3.84 * level_sigma2[kp2.octave]) { bestIdx2 = idx2; bestDist = d; }
        }
        if (bestIdx2 >= 0) {
            vMatches12[j.idx1] = bestIdx2;
            // (the reference never sets vbMatched2 in this function: a KF2 feature can be matched by several KF1 features)
            nmatches++;
            if (check_ori) rh.add(rot_bin(kp1.angle, kf2->keys[bestIdx2].angle), j.idx1);
        }
    }
    if (check_ori) rh.prune([&](int idx) { vMatches12[idx] = -1; nmatches--; });
    int np = 0;
    for (int i = 0; i < kf1->n; i++) {
        if (vMatches12[i] < 0) continue;
        pairs[2 * np] = i; pairs[2 * np + 1] = vMatches12[i];
        np++;
    }
    return nmatches;
}

// ---- src/ORBmatcher.cc:43-126
int ygz_search_by_projection_points(ygz_matcher* h, float nnratio, ygz_frame_view* F, const ygz_map_points* mps,
                                    const int* query_ids, int nq, float th, int check_level) {
    YGZ_ARG_CHECK(h && F && mps && (nq == 0 || query_ids));
    const bool bFactor = th != 1.0;
    Grid g(F);
    struct Job { int pMP, first, count, level; float r; };
    std::vector<Job> jobs;
    std::vector<int> pq, pt, cand;
    for (int q = 0; q < nq; q++) {
        const int pMP = query_ids[q];
        if (!mps->track_in_view[pMP] || mp_bad(mps, pMP)) continue;
        const int lvl = mps->track_scale_level[pMP];
        float r = mps->track_view_cos[pMP] > 0.998 ? 2.5f : 4.0f;     // RadiusByViewingCos :128-133
        if (bFactor) r *= th;
        cand.clear();
        features_in_area(F, g, mps->track_proj_x[pMP], mps->track_proj_y[pMP], r * F->scale_factors[lvl],
                         check_level ? lvl - 1 : -1, check_level ? lvl : -1, cand);
        if (cand.empty()) continue;
        Job j{pMP, (int)pq.size(), (int)cand.size(), lvl, r};
        for (int idx : cand) { pq.push_back(pMP); pt.push_back(idx); }
        jobs.push_back(j);
    }
    std::vector<int> dist;
    int rc = pair_distances(h, mps->desc, mps->m, F->desc, F->n, pq, pt, dist);
    if (rc) return rc;
    int nmatches = 0;
    for (const Job& j : jobs) {
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int p = j.first; p < j.first + j.count; p++) {
            const int idx = pt[p];
            if (F->map_point[idx] >= 0 && mp_obs(mps, F->map_point[idx]) > 0) continue;       // :89-91
            if (uright(F, idx) > 0) {                                                          // :93-97
                const float er = std::fabs(mps->track_proj_xr[j.pMP] - uright(F, idx));
                if (er > j.r * F->scale_factors[j.level]) continue;
            }
            const int d = dist[p];
            if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestLevel2 = bestLevel; bestLevel = F->keys[idx].octave; bestIdx = idx; }
            else if (d < bestDist2) { bestLevel2 = F->keys[idx].octave; bestDist2 = d; }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;           // :117-118
            F->map_point[bestIdx] = j.pMP;
            nmatches++;
        }
    }
    return nmatches;
}

static inline void project(const ygz_frame_view* C, const float* Pw, float& xc, float& yc, float& zc) {
    const float* R = C->Rcw; const float* t = C->tcw;
    xc = R[0] * Pw[0] + R[1] * Pw[1] + R[2] * Pw[2] + t[0];
    yc = R[3] * Pw[0] + R[4] * Pw[1] + R[5] * Pw[2] + t[1];
    zc = R[6] * Pw[0] + R[7] * Pw[1] + R[8] * Pw[2] + t[2];
}

// ---- src/ORBmatcher.cc:1218-1350
int ygz_search_by_projection_last(ygz_matcher* h, int check_ori, ygz_frame_view* C, const ygz_frame_view* L,
                                  const ygz_map_points* mps, float th, int mono, int check_level) {
    YGZ_ARG_CHECK(h && C && L && mps);
    const float* R = C->Rcw; const float* t = C->tcw;
    float twc[3], tlc[3];
    for (int i = 0; i < 3; i++) twc[i] = -1 * (R[0 * 3 + i] * t[0] + R[1 * 3 + i] * t[1] + R[2 * 3 + i] * t[2]);
    for (int i = 0; i < 3; i++) tlc[i] = L->Rcw[i * 3 + 0] * twc[0] + L->Rcw[i * 3 + 1] * twc[1] + L->Rcw[i * 3 + 2] * twc[2] + L->tcw[i];
    const bool bForward = tlc[2] > C->mb && !mono;
    const bool bBackward = -tlc[2] > C->mb && !mono;
    Grid g(C);
    struct Job { int i, pMP, first, count; float u, invzc, radius; };
    std::vector<Job> jobs;
    std::vector<int> pq, pt, cand;
    for (int i = 0; i < L->n; i++) {
        const int pMP = L->map_point[i];
        if (pMP < 0) continue;
        if (L->outlier && L->outlier[i]) continue;
        float xc, yc, zc;
        project(C, mps->world_pos + 3 * (size_t)pMP, xc, yc, zc);
        const float invzc = (float)(1.0 / zc);
        if (invzc < 0) continue;
        const float u = C->fx * xc * invzc + C->cx;
        const float v = C->fy * yc * invzc + C->cy;
        if (u < C->min_x || u > C->max_x) continue;
        if (v < C->min_y || v > C->max_y) continue;
        const int nLastOctave = L->keys[i].octave;
        const float radius = th * C->scale_factors[nLastOctave];
        cand.clear();
        if (!check_level) features_in_area(C, g, u, v, radius, -1, -1, cand);
        else if (bForward) features_in_area(C, g, u, v, radius, nLastOctave, -1, cand);
        else if (bBackward) features_in_area(C, g, u, v, radius, 0, nLastOctave, cand);
        else features_in_area(C, g, u, v, radius, nLastOctave - 1, nLastOctave + 1, cand);
        if (cand.empty()) continue;
        Job j{i, pMP, (int)pq.size(), (int)cand.size(), u, invzc, radius};
        for (int idx : cand) { pq.push_back(pMP); pt.push_back(idx); }
        jobs.push_back(j);
    }
    std::vector<int> dist;
    int rc = pair_distances(h, mps->desc, mps->m, C->desc, C->n, pq, pt, dist);
    if (rc) return rc;
    int nmatches = 0;
    RotHist rh;
    for (const Job& j : jobs) {
        int bestDist = 256, bestIdx2 = -1;
        for (int p = j.first; p < j.first + j.count; p++) {
            const int i2 = pt[p];
            if (C->map_point[i2] >= 0 && mp_obs(mps, C->map_point[i2]) > 0) continue;
            if (uright(C, i2) > 0) {
                const float ur = j.u - C->mbf * j.invzc;
                if (std::fabs(ur - uright(C, i2)) > j.radius) continue;
            }
            if (dist[p] < bestDist) { bestDist = dist[p]; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            C->map_point[bestIdx2] = j.pMP;
            nmatches++;
            if (check_ori) rh.add(rot_bin(L->keys[j.i].angle, C->keys[bestIdx2].angle), bestIdx2);
        }
    }
    if (check_ori) rh.prune([&](int idx) { C->map_point[idx] = -1; nmatches--; });
    return nmatches;
}

// ---- src/ORBmatcher.cc:1352-1469
int ygz_search_by_projection_kf(ygz_matcher* h, int check_ori, ygz_frame_view* C, const ygz_frame_view* KF,
                                const ygz_map_points* mps, const uint8_t* found, float th, int ORBdist) {
    YGZ_ARG_CHECK(h && C && KF && mps);
    const float* R = C->Rcw; const float* t = C->tcw;
    float Ow[3];
    for (int i = 0; i < 3; i++) Ow[i] = -1 * (R[0 * 3 + i] * t[0] + R[1 * 3 + i] * t[1] + R[2 * 3 + i] * t[2]);
    Grid g(C);
    struct Job { int i, pMP, first, count; };
    std::vector<Job> jobs;
    std::vector<int> pq, pt, cand;
    for (int i = 0; i < KF->n; i++) {
        const int pMP = KF->map_point[i];
        if (pMP < 0) continue;
        if (mp_bad(mps, pMP) || (found && found[pMP])) continue;
        const float* Pw = mps->world_pos + 3 * (size_t)pMP;
        float xc, yc, zc;
        project(C, Pw, xc, yc, zc);
        const float invzc = (float)(1.0 / zc);
        const float u = C->fx * xc * invzc + C->cx;
        const float v = C->fy * yc * invzc + C->cy;
        if (u < C->min_x || u > C->max_x) continue;
        if (v < C->min_y || v > C->max_y) continue;
        const float PO[3] = {Pw[0] - Ow[0], Pw[1] - Ow[1], Pw[2] - Ow[2]};
        const float dist3D = std::sqrt(PO[0] * PO[0] + PO[1] * PO[1] + PO[2] * PO[2]);
        const float maxDistance = 1.2f * mps->max_dist[pMP], minDistance = 0.8f * mps->min_dist[pMP];   // src/MapPoint.cc:333-341
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const float ratio = mps->max_dist[pMP] / dist3D;                                                // PredictScale src/MapPoint.cc:359-373
        int lvl = (int)std::ceil(std::log(ratio) / C->log_scale_factor);
        if (lvl < 0) lvl = 0; else if (lvl >= C->n_levels) lvl = C->n_levels - 1;
        const float radius = th * C->scale_factors[lvl];
        cand.clear();
        features_in_area(C, g, u, v, radius, lvl - 1, lvl + 1, cand);
        if (cand.empty()) continue;
        Job j{i, pMP, (int)pq.size(), (int)cand.size()};
        for (int idx : cand) { pq.push_back(pMP); pt.push_back(idx); }
        jobs.push_back(j);
    }
    std::vector<int> dist;
    int rc = pair_distances(h, mps->desc, mps->m, C->desc, C->n, pq, pt, dist);
    if (rc) return rc;
    int nmatches = 0;
    RotHist rh;
    for (const Job& j : jobs) {
        int bestDist = 256, bestIdx2 = -1;
        for (int p = j.first; p < j.first + j.count; p++) {
            const int i2 = pt[p];
            if (C->map_point[i2] >= 0) continue;
            if (dist[p] < bestDist) { bestDist = dist[p]; bestIdx2 = i2; }
        }
        if (bestDist <= ORBdist) {
            C->map_point[bestIdx2] = j.pMP;
            nmatches++;
            if (check_ori) rh.add(rot_bin(KF->keys[j.i].angle, C->keys[bestIdx2].angle), bestIdx2);
        }
    }
    if (check_ori) rh.prune([&](int idx) { C->map_point[idx] = -1; nmatches--; });
    return nmatches;
}

// ---- src/ORBmatcher.cc:375-478
int ygz_search_for_initialization(ygz_matcher* h, float nnratio, int check_ori, const ygz_frame_view* F1,
                                  const ygz_frame_view* F2, float* prev, int* m12, int windowSize) {
    YGZ_ARG_CHECK(h && F1 && F2 && prev && m12);
    for (int i = 0; i < F1->n; i++) m12[i] = -1;
    Grid g(F2);
    struct Job { int i1, first, count; };
    std::vector<Job> jobs;
    std::vector<int> pq, pt, cand;
    for (int i1 = 0; i1 < F1->n; i1++) {
        const int level1 = F1->keys[i1].octave;
        if (level1 > 0) continue;
        cand.clear();
        features_in_area(F2, g, prev[2 * i1], prev[2 * i1 + 1], (float)windowSize, level1, level1, cand);
        if (cand.empty()) continue;
        Job j{i1, (int)pq.size(), (int)cand.size()};
        for (int idx : cand) { pq.push_back(i1); pt.push_back(idx); }
        jobs.push_back(j);
    }
    std::vector<int> dist;
    int rc = pair_distances(h, F1->desc, F1->n, F2->desc, F2->n, pq, pt, dist);
    if (rc) return rc;
    int nmatches = 0;
    RotHist rh;
    std::vector<int> vMatchedDistance(F2->n, INT_MAX), vnMatches21(F2->n, -1);
    for (const Job& j : jobs) {
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int p = j.first; p < j.first + j.count; p++) {
            const int i2 = pt[p], d = dist[p];
            if (vMatchedDistance[i2] <= d) continue;
            if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestIdx2 = i2; }
            else if (d < bestDist2) bestDist2 = d;
        }
        if (bestDist <= TH_LOW && bestDist < (float)bestDist2 * nnratio) {
            if (vnMatches21[bestIdx2] >= 0) { m12[vnMatches21[bestIdx2]] = -1; nmatches--; }
            m12[j.i1] = bestIdx2;
            vnMatches21[bestIdx2] = j.i1;
            vMatchedDistance[bestIdx2] = bestDist;
            nmatches++;
            if (check_ori) rh.add(rot_bin(F1->keys[j.i1].angle, F2->keys[bestIdx2].angle), j.i1);
        }
    }
    if (check_ori) rh.prune([&](int idx1) { if (m12[idx1] >= 0) { m12[idx1] = -1; nmatches--; } });
    for (int i1 = 0; i1 < F1->n; i1++)
        if (m12[i1] >= 0) { prev[2 * i1] = F2->keys[m12[i1]].x; prev[2 * i1 + 1] = F2->keys[m12[i1]].y; }
    return nmatches;
}

// ---- batched device-resident SearchByBoW
static int bow_batch_prepare(ygz_matcher* h, ygz_orb* orb, int n_pairs, int* cap_out, int* frames_out) {
    void *d_kps, *d_desc, *d_counts;
    int cap;
    int rc = ygz_orb_device_outputs(orb, &d_kps, &d_desc, &d_counts, &cap);
    if (rc) return rc;
    if (cap > BOW_MAX_FEATS) { set_last_error("batched SearchByBoW supports up to 2048 features per frame"); return YGZ_ERR_CAPACITY; }
    if (n_pairs > h->bow_pairs_cap || cap != h->bow_cap) {
        cudaFree(h->d_kf_idx); cudaFree(h->d_f_idx); cudaFree(h->d_matches); cudaFree(h->d_nmatches);
        h->d_kf_idx = h->d_f_idx = h->d_matches = h->d_nmatches = nullptr;
        YGZ_CUDA_TRY(cudaMalloc(&h->d_kf_idx, (size_t)n_pairs * 4));
        YGZ_CUDA_TRY(cudaMalloc(&h->d_f_idx, (size_t)n_pairs * 4));
        YGZ_CUDA_TRY(cudaMalloc(&h->d_matches, (size_t)n_pairs * cap * 4));
        YGZ_CUDA_TRY(cudaMalloc(&h->d_nmatches, (size_t)n_pairs * 4));
        h->bow_pairs_cap = n_pairs;
        h->bow_cap = cap;
        YGZ_CUDA_TRY(cudaFuncSetAttribute(k_bow_match_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bow_smem_bytes(cap)));
    }
    *cap_out = cap;
    (void)frames_out;
    return YGZ_OK;
}

static int node_capacity(ygz_matcher* h, size_t elems) {
    if (elems <= h->node_elems) return YGZ_OK;
    cudaFree(h->d_node); h->d_node = nullptr; h->node_elems = 0;
    YGZ_CUDA_TRY(cudaMalloc(&h->d_node, elems * 4));
    h->node_elems = elems;
    return YGZ_OK;
}

static int frames_of_pairs(const int* kf_idx, const int* f_idx, int n_pairs) {
    int max_frame = 0;
    for (int i = 0; i < n_pairs; i++) max_frame = std::max(max_frame, std::max(kf_idx[i], f_idx[i]));
    return max_frame + 1;
}

// node ids for the batched search: Frame::ComputeBoW (src/Frame.cc:495-501) of every frame the pairs reference, on stream st
static int launch_node_ids(ygz_matcher* h, ygz_orb* orb, int frames, int cap, cudaStream_t st) {
    void *d_kps, *d_desc, *d_counts;
    int rc = ygz_orb_device_outputs(orb, &d_kps, &d_desc, &d_counts, nullptr);
    if (rc) return rc;
    return vocab_node_ids_async(h->vocab, d_desc, (const int*)d_counts, cap, frames, h->levelsup, h->d_node, st);
}

int ygz_matcher_set_vocabulary(ygz_matcher* h, ygz_vocab* vocab, int levelsup) {
    YGZ_ARG_CHECK(h && levelsup >= 0);
    if (vocab) YGZ_ARG_CHECK(vocab_device(vocab) == h->device);
    h->vocab = vocab;
    h->levelsup = levelsup;
    h->last_pairs = 0;                 // cached launch arguments referred to the previous node source
    return YGZ_OK;
}

int ygz_search_by_bow_batch(ygz_matcher* h, ygz_orb* orb, float nnratio, int check_ori, const int* kf_idx, const int* f_idx,
                            int n_pairs, const uint32_t* node_of_feature, int* out_matches, int* out_nmatches) {
    YGZ_ARG_CHECK(h && orb && kf_idx && f_idx && n_pairs >= 1 && out_matches && out_nmatches);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    int cap = 0, rc;
    if ((rc = bow_batch_prepare(h, orb, n_pairs, &cap, nullptr))) return rc;
    const int frames = frames_of_pairs(kf_idx, f_idx, n_pairs);
    YGZ_ARG_CHECK(frames <= ygz_orb_max_batch(orb));
    for (int i = 0; i < n_pairs; i++) YGZ_ARG_CHECK(kf_idx[i] >= 0 && f_idx[i] >= 0);
    const bool use_nodes = node_of_feature || h->vocab;
    if (use_nodes && (rc = node_capacity(h, (size_t)frames * cap))) return rc;
    void *d_kps, *d_desc, *d_counts;
    ygz_orb_device_outputs(orb, &d_kps, &d_desc, &d_counts, nullptr);
    // the extractor's stream produced the inputs: order our stream after it
    YGZ_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)ygz_orb_stream(orb)));
    if (node_of_feature)
        YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_node, node_of_feature, (size_t)frames * cap * 4, cudaMemcpyHostToDevice, h->stream));
    else if (h->vocab && (rc = launch_node_ids(h, orb, frames, cap, h->stream))) return rc;
    h->bow_frames = frames;
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_kf_idx, kf_idx, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, h->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_f_idx, f_idx, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, h->stream));
    BowBatchArgs A;
    A.kps = (const ygz_keypoint*)d_kps; A.desc = (const uint4*)d_desc; A.counts = (const int*)d_counts;
    A.node = use_nodes ? h->d_node : nullptr; A.cap = cap; A.kf_idx = h->d_kf_idx; A.f_idx = h->d_f_idx; A.nnratio = nnratio;
    A.check_ori = check_ori; A.matches = h->d_matches; A.nmatches = h->d_nmatches;
    { NvtxRange nvtx_range("ygz/SearchByBoW batch");
    k_bow_match_batch<<<n_pairs, BOW_THREADS, bow_smem_bytes(A.cap), h->stream>>>(A);
    }
    YGZ_CUDA_TRY(cudaGetLastError());
    h->last_args = A;
    h->last_pairs = n_pairs;
    YGZ_CUDA_TRY(cudaMemcpyAsync(out_matches, h->d_matches, (size_t)n_pairs * cap * 4, cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(out_nmatches, h->d_nmatches, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return YGZ_OK;
}

// Asynchronous variant on the EXTRACTOR's stream (stream order gives the dependency on the descriptors): synthetic
// vocabulary only (node_of_feature == NULL path); results land in the (pinned) host arrays after ygz_orb_wait(orb).
int ygz_search_by_bow_batch_async(ygz_matcher* h, ygz_orb* orb, float nnratio, int check_ori, const int* kf_idx, const int* f_idx,
                                  int n_pairs, int* out_matches, int* out_nmatches) {
    YGZ_ARG_CHECK(h && orb && kf_idx && f_idx && n_pairs >= 1 && out_matches && out_nmatches);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    int cap = 0, rc;
    if ((rc = bow_batch_prepare(h, orb, n_pairs, &cap, nullptr))) return rc;
    cudaStream_t st = (cudaStream_t)ygz_orb_stream(orb);
    void *d_kps, *d_desc, *d_counts;
    ygz_orb_device_outputs(orb, &d_kps, &d_desc, &d_counts, nullptr);
    const int frames = frames_of_pairs(kf_idx, f_idx, n_pairs);
    YGZ_ARG_CHECK(frames <= ygz_orb_max_batch(orb));
    for (int i = 0; i < n_pairs; i++) YGZ_ARG_CHECK(kf_idx[i] >= 0 && f_idx[i] >= 0);
    if (h->vocab) {                     // Frame::ComputeBoW of the batch, on the extractor's stream right after the descriptors
        if ((rc = node_capacity(h, (size_t)frames * cap))) return rc;
        if ((rc = launch_node_ids(h, orb, frames, cap, st))) return rc;
    }
    h->bow_frames = frames;
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_kf_idx, kf_idx, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st));
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_f_idx, f_idx, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st));
    BowBatchArgs A;
    A.kps = (const ygz_keypoint*)d_kps; A.desc = (const uint4*)d_desc; A.counts = (const int*)d_counts;
    A.node = h->vocab ? h->d_node : nullptr; A.cap = cap; A.kf_idx = h->d_kf_idx; A.f_idx = h->d_f_idx; A.nnratio = nnratio;
    A.check_ori = check_ori; A.matches = h->d_matches; A.nmatches = h->d_nmatches;
    { NvtxRange nvtx_range("ygz/SearchByBoW batch");
    k_bow_match_batch<<<n_pairs, BOW_THREADS, bow_smem_bytes(A.cap), st>>>(A);
    }
    YGZ_CUDA_TRY(cudaGetLastError());
    h->last_args = A;
    h->last_pairs = n_pairs;
    YGZ_CUDA_TRY(cudaMemcpyAsync(out_matches, h->d_matches, (size_t)n_pairs * cap * 4, cudaMemcpyDeviceToHost, st));
    YGZ_CUDA_TRY(cudaMemcpyAsync(out_nmatches, h->d_nmatches, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, st));
    return YGZ_OK;
}

// the cached launch arguments of the last batched search, re-pointed at the extractor's CURRENT output buffers
// (ygz_orb_extract_frame / _dso may have re-allocated them since); a changed capacity invalidates the cache
static int refreshed_args(ygz_matcher* h, ygz_orb* orb, float nnratio, int check_ori, BowBatchArgs* out) {
    void *d_kps, *d_desc, *d_counts;
    int cap = 0;
    int rc = ygz_orb_device_outputs(orb, &d_kps, &d_desc, &d_counts, &cap);
    if (rc) return rc;
    if (cap != h->last_args.cap) {
        h->last_pairs = 0;
        set_last_error("the extractor's output capacity changed since the last batched SearchByBoW: call it again first");
        return YGZ_ERR_ARG;
    }
    BowBatchArgs A = h->last_args;
    A.kps = (const ygz_keypoint*)d_kps; A.desc = (const uint4*)d_desc; A.counts = (const int*)d_counts;
    A.nnratio = nnratio; A.check_ori = check_ori;
    *out = A;
    return YGZ_OK;
}

int ygz_search_by_bow_batch_time(ygz_matcher* h, ygz_orb* orb, float nnratio, int check_ori, int n_pairs, int iters, float* ms) {
    YGZ_ARG_CHECK(h && orb && ms && iters >= 1 && h->last_pairs >= 1 && n_pairs == h->last_pairs);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    BowBatchArgs A;
    int rc;
    if ((rc = refreshed_args(h, orb, nnratio, check_ori, &A))) return rc;
    YGZ_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)ygz_orb_stream(orb)));
    YGZ_CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    for (int i = 0; i < iters; i++) {
        if (h->vocab && (rc = launch_node_ids(h, orb, h->bow_frames, A.cap, h->stream))) return rc;
        { NvtxRange nvtx_range("ygz/SearchByBoW batch");
        k_bow_match_batch<<<n_pairs, BOW_THREADS, bow_smem_bytes(A.cap), h->stream>>>(A);
        }
    }
    YGZ_CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    YGZ_CUDA_TRY(cudaEventSynchronize(h->ev1));
    YGZ_CUDA_TRY(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return YGZ_OK;
}


// whole front-end step on ONE stream (the extractor's): pyramid + FAST + octree + describe + batched SearchByBoW.
// Needs a prior ygz_search_by_bow_batch call (pair lists / node ids resident).  CUDA-event timed.
int ygz_frontend_time(ygz_orb* orb, ygz_matcher* h, int n_frames, int iters, float nnratio, int check_ori, float* ms) {
    YGZ_ARG_CHECK(orb && h && ms && iters >= 1 && h->last_pairs >= 1);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)ygz_orb_stream(orb);
    BowBatchArgs A;
    int rc;
    if ((rc = refreshed_args(h, orb, nnratio, check_ori, &A))) return rc;
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    YGZ_CUDA_TRY(cudaEventRecord(h->ev0, st));
    for (int i = 0; i < iters; i++) {
        if ((rc = ygz_orb_run(orb, n_frames))) return rc;
        if (h->vocab && (rc = launch_node_ids(h, orb, h->bow_frames, A.cap, st))) return rc;
        { NvtxRange nvtx_range("ygz/SearchByBoW batch");
        k_bow_match_batch<<<h->last_pairs, BOW_THREADS, bow_smem_bytes(A.cap), st>>>(A);
        }
    }
    YGZ_CUDA_TRY(cudaGetLastError());
    YGZ_CUDA_TRY(cudaEventRecord(h->ev1, st));
    YGZ_CUDA_TRY(cudaEventSynchronize(h->ev1));
    YGZ_CUDA_TRY(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return YGZ_OK;
}


// Same as ygz_frontend_time with TWO independent batches in flight on two streams (handle pairs A and B, both with
// resident inputs and prepared pair lists): steps alternate A, B, A, ... so the latency-bound kernels of one batch
// (octree, BoW, small pyramid levels) fill the issue slots the other batch's FAST / describe kernels leave idle.
// Timed with CUDA events: start on A's stream (B waits on it), end after both streams have drained.
int ygz_frontend_time2(ygz_orb* orbA, ygz_matcher* hA, ygz_orb* orbB, ygz_matcher* hB, int n_frames, int iters,
                       float nnratio, int check_ori, float* ms) {
    YGZ_ARG_CHECK(orbA && hA && orbB && hB && ms && iters >= 1 && hA->last_pairs >= 1 && hB->last_pairs >= 1);
    YGZ_CUDA_TRY(cudaSetDevice(hA->device));
    cudaStream_t sA = (cudaStream_t)ygz_orb_stream(orbA), sB = (cudaStream_t)ygz_orb_stream(orbB);
    BowBatchArgs A, B;
    int rc;
    if ((rc = refreshed_args(hA, orbA, nnratio, check_ori, &A))) return rc;
    if ((rc = refreshed_args(hB, orbB, nnratio, check_ori, &B))) return rc;
    YGZ_CUDA_TRY(cudaStreamSynchronize(hA->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(hB->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(sA));
    YGZ_CUDA_TRY(cudaStreamSynchronize(sB));
    YGZ_CUDA_TRY(cudaEventRecord(hA->ev0, sA));
    YGZ_CUDA_TRY(cudaStreamWaitEvent(sB, hA->ev0, 0));
    for (int i = 0; i < iters; i++) {
        ygz_orb* orb = (i & 1) ? orbB : orbA;
        ygz_matcher* h = (i & 1) ? hB : hA;
        cudaStream_t st = (i & 1) ? sB : sA;
        if ((rc = ygz_orb_run(orb, n_frames))) return rc;
        if (h->vocab && (rc = launch_node_ids(h, orb, h->bow_frames, (i & 1) ? B.cap : A.cap, st))) return rc;
        { NvtxRange nvtx_range("ygz/SearchByBoW batch");
        k_bow_match_batch<<<h->last_pairs, BOW_THREADS, bow_smem_bytes((i & 1) ? B.cap : A.cap), st>>>((i & 1) ? B : A);
        }
    }
    YGZ_CUDA_TRY(cudaGetLastError());
    YGZ_CUDA_TRY(cudaEventRecord(hB->ev1, sB));
    YGZ_CUDA_TRY(cudaStreamWaitEvent(sA, hB->ev1, 0));
    YGZ_CUDA_TRY(cudaEventRecord(hA->ev1, sA));
    YGZ_CUDA_TRY(cudaEventSynchronize(hA->ev1));
    YGZ_CUDA_TRY(cudaEventElapsedTime(ms, hA->ev0, hA->ev1));
    return YGZ_OK;
}

}  // extern "C"

// The same with k handle pairs in flight (k streams): step i runs on pair i % k.  k = 2 is ygz_frontend_time2.
extern "C" int ygz_frontend_time_n(ygz_orb* const* orbs, ygz_matcher* const* hs, int k, int n_frames, int iters, float nnratio, int check_ori,
                                   float* ms) {
    YGZ_ARG_CHECK(orbs && hs && k >= 1 && k <= 8 && ms && iters >= 1);
    for (int i = 0; i < k; i++) YGZ_ARG_CHECK(orbs[i] && hs[i] && hs[i]->last_pairs >= 1);
    YGZ_CUDA_TRY(cudaSetDevice(hs[0]->device));
    BowBatchArgs A[8];
    cudaStream_t st[8];
    int rc;
    for (int i = 0; i < k; i++) {
        if ((rc = refreshed_args(hs[i], orbs[i], nnratio, check_ori, &A[i]))) return rc;
        st[i] = (cudaStream_t)ygz_orb_stream(orbs[i]);
        YGZ_CUDA_TRY(cudaStreamSynchronize(hs[i]->stream));
        YGZ_CUDA_TRY(cudaStreamSynchronize(st[i]));
    }
    YGZ_CUDA_TRY(cudaEventRecord(hs[0]->ev0, st[0]));
    for (int i = 1; i < k; i++) YGZ_CUDA_TRY(cudaStreamWaitEvent(st[i], hs[0]->ev0, 0));
    for (int it = 0; it < iters; it++) {
        const int i = it % k;
        if ((rc = ygz_orb_run(orbs[i], n_frames))) return rc;
        if (hs[i]->vocab && (rc = launch_node_ids(hs[i], orbs[i], hs[i]->bow_frames, A[i].cap, st[i]))) return rc;
        { NvtxRange nvtx_range("ygz/SearchByBoW batch");
        k_bow_match_batch<<<hs[i]->last_pairs, BOW_THREADS, bow_smem_bytes(A[i].cap), st[i]>>>(A[i]);
        }
    }
    YGZ_CUDA_TRY(cudaGetLastError());
    for (int i = 1; i < k; i++) {
        YGZ_CUDA_TRY(cudaEventRecord(hs[i]->ev1, st[i]));
        YGZ_CUDA_TRY(cudaStreamWaitEvent(st[0], hs[i]->ev1, 0));
    }
    YGZ_CUDA_TRY(cudaEventRecord(hs[0]->ev1, st[0]));
    YGZ_CUDA_TRY(cudaEventSynchronize(hs[0]->ev1));
    YGZ_CUDA_TRY(cudaEventElapsedTime(ms, hs[0]->ev0, hs[0]->ev1));
    return YGZ_OK;
}

// ---------------------------------------------------------------------------------- batched projection searches (host side)
namespace {
enum { PJ_QF, PJ_TF, PJ_WP, PJ_VALID, PJ_POSE, PJ_PX, PJ_PY, PJ_PXR, PJ_VC, PJ_LVL, PJ_USABLE, PJ_QDESC, PJ_OCC, PJ_UR, PJ_OUT, PJ_CNT };

// device copy of a host array in scratch slot `slot` (NULL stays NULL)
template <class T>
int pj_upload(ygz_matcher* h, int slot, const T* host, size_t count, const T** dev) {
    *dev = nullptr;
    if (!host || count == 0) return YGZ_OK;
    const size_t bytes = count * sizeof(T);
    if (bytes > h->pj_bytes[slot]) {
        cudaFree(h->pj_buf[slot]); h->pj_buf[slot] = nullptr; h->pj_bytes[slot] = 0;
        YGZ_CUDA_TRY(cudaMalloc(&h->pj_buf[slot], bytes + bytes / 4));
        h->pj_bytes[slot] = bytes + bytes / 4;
    }
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->pj_buf[slot], host, bytes, cudaMemcpyHostToDevice, h->stream));
    *dev = static_cast<const T*>(h->pj_buf[slot]);
    return YGZ_OK;
}
int pj_alloc(ygz_matcher* h, int slot, size_t bytes, void** dev) {
    if (bytes > h->pj_bytes[slot]) {
        cudaFree(h->pj_buf[slot]); h->pj_buf[slot] = nullptr; h->pj_bytes[slot] = 0;
        YGZ_CUDA_TRY(cudaMalloc(&h->pj_buf[slot], bytes));
        h->pj_bytes[slot] = bytes;
    }
    *dev = h->pj_buf[slot];
    return YGZ_OK;
}

int pj_common(ygz_matcher* h, ygz_orb* orb, const ygz_projection_camera* cam, int n_pairs, const int* q_frame, const int* t_frame,
              const float* u_right, const uint8_t* occupied, ProjArgs* A) {
    void *d_kps, *d_desc, *d_counts;
    int cap = 0, rc;
    if ((rc = ygz_orb_device_outputs(orb, &d_kps, &d_desc, &d_counts, &cap))) return rc;
    if (cap > PJ_MAX_CAP) { set_last_error("batched projection search supports up to 4096 key slots per frame"); return YGZ_ERR_CAPACITY; }
    const int B = ygz_orb_max_batch(orb);
    for (int i = 0; i < n_pairs; i++) YGZ_ARG_CHECK(q_frame[i] >= 0 && t_frame[i] >= 0 && t_frame[i] < B);
    std::memset(A, 0, sizeof(*A));
    A->kps = (const ygz_keypoint*)d_kps; A->desc = (const uint4*)d_desc; A->counts = (const int*)d_counts; A->cap = cap;
    float scale[16], inv[16], s2[16], is2[16];
    int fpl[16];
    if ((rc = ygz_orb_tables(orb, scale, inv, s2, is2, fpl))) return rc;
    for (int l = 0; l < 16; l++) A->scale[l] = scale[l];
    A->fx = cam->fx; A->fy = cam->fy; A->cx = cam->cx; A->cy = cam->cy; A->mb = cam->mb; A->mbf = cam->mbf;
    A->min_x = cam->min_x; A->min_y = cam->min_y; A->max_x = cam->max_x; A->max_y = cam->max_y;
    YGZ_ARG_CHECK(cam->max_x > cam->min_x && cam->max_y > cam->min_y);
    if ((rc = pj_upload(h, PJ_QF, q_frame, n_pairs, &A->q_frame)) || (rc = pj_upload(h, PJ_TF, t_frame, n_pairs, &A->t_frame))) return rc;
    if ((rc = pj_upload(h, PJ_UR, u_right, (size_t)B * cap, &A->u_right))) return rc;
    if ((rc = pj_upload(h, PJ_OCC, occupied, (size_t)n_pairs * cap, &A->occupied))) return rc;
    void* out = nullptr;
    if ((rc = pj_alloc(h, PJ_OUT, (size_t)n_pairs * cap * 4, &out))) return rc;
    A->matches = (int*)out;
    if ((rc = pj_alloc(h, PJ_CNT, (size_t)n_pairs * 8, &out))) return rc;
    A->nmatches = (int*)out; A->rounds = (int*)out + n_pairs;
    return YGZ_OK;
}

int pj_launch(ygz_matcher* h, const ProjArgs& A, int n_pairs, int mode, int nq, cudaStream_t st) {
    const size_t smem = proj_smem_bytes(A.cap, nq);
    if (smem > 220 * 1024) { set_last_error("batched projection search: shared-memory budget exceeded (too many key slots / query points)"); return YGZ_ERR_CAPACITY; }
    NvtxRange nvtx_range(mode == 0 ? "ygz/SearchByProjection(Cur,Last) batch" : "ygz/SearchByProjection(F,points) batch");
    if (mode == 0) {
        YGZ_CUDA_TRY(cudaFuncSetAttribute(k_projection_batch<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_projection_batch<0><<<n_pairs, PJ_THREADS, smem, st>>>(A);
    } else {
        YGZ_CUDA_TRY(cudaFuncSetAttribute(k_projection_batch<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_projection_batch<1><<<n_pairs, PJ_THREADS, smem, st>>>(A);
    }
    YGZ_CUDA_TRY(cudaGetLastError());
    h->pj_last = A; h->pj_last_pairs = n_pairs; h->pj_last_mode = mode; h->pj_last_nq = nq;
    return YGZ_OK;
}

int pj_finish(ygz_matcher* h, const ProjArgs& A, int n_pairs, int* out_matches, int* out_nmatches, int* out_rounds) {
    if (out_matches) YGZ_CUDA_TRY(cudaMemcpyAsync(out_matches, A.matches, (size_t)n_pairs * A.cap * 4, cudaMemcpyDeviceToHost, h->stream));
    if (out_nmatches) YGZ_CUDA_TRY(cudaMemcpyAsync(out_nmatches, A.nmatches, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, h->stream));
    if (out_rounds) YGZ_CUDA_TRY(cudaMemcpyAsync(out_rounds, A.rounds, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return YGZ_OK;
}
}  // namespace

extern "C" {

int ygz_search_by_projection_last_batch(ygz_matcher* h, ygz_orb* orb, const ygz_projection_camera* cam, int n_pairs, const int* last_frame,
                                        const int* cur_frame, const float* world_pos, const uint8_t* valid, const float* pose,
                                        const uint8_t* mp_desc, const float* u_right, float th, int mono, int check_level, int check_orientation,
                                        int* out_matches, int* out_nmatches, int* out_rounds) {
    YGZ_ARG_CHECK(h && orb && cam && n_pairs >= 1 && last_frame && cur_frame && world_pos && valid && pose);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    ProjArgs A;
    int rc;
    YGZ_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)ygz_orb_stream(orb)));      // the extractor's stream produced the inputs
    if ((rc = pj_common(h, orb, cam, n_pairs, last_frame, cur_frame, u_right, nullptr, &A))) return rc;
    const int B = ygz_orb_max_batch(orb);
    for (int i = 0; i < n_pairs; i++) YGZ_ARG_CHECK(last_frame[i] < B);
    if ((rc = pj_upload(h, PJ_WP, world_pos, (size_t)B * A.cap * 3, &A.world_pos)) || (rc = pj_upload(h, PJ_VALID, valid, (size_t)B * A.cap, &A.valid)) ||
        (rc = pj_upload(h, PJ_POSE, pose, (size_t)B * 12, &A.pose))) return rc;
    const uint4* qd = nullptr;
    if ((rc = pj_upload(h, PJ_QDESC, reinterpret_cast<const uint4*>(mp_desc), mp_desc ? (size_t)B * A.cap * 2 : 0, &qd))) return rc;
    A.q_desc = qd;
    A.th = th; A.mono = mono; A.check_level = check_level; A.check_ori = check_orientation;
    if ((rc = pj_launch(h, A, n_pairs, 0, A.cap, h->stream))) return rc;
    return pj_finish(h, A, n_pairs, out_matches, out_nmatches, out_rounds);
}

int ygz_search_by_projection_points_batch(ygz_matcher* h, ygz_orb* orb, const ygz_projection_camera* cam, int n_pairs, const int* frame,
                                          int M, const float* proj_x, const float* proj_y, const float* proj_xr, const float* view_cos,
                                          const int* level, const uint8_t* usable, const uint8_t* mp_desc, const uint8_t* occupied,
                                          const float* u_right, float nnratio, float th, int check_level, int* out_matches,
                                          int* out_nmatches, int* out_rounds) {
    YGZ_ARG_CHECK(h && orb && cam && n_pairs >= 1 && frame && M >= 1 && proj_x && proj_y && proj_xr && view_cos && level && usable && mp_desc);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    ProjArgs A;
    int rc;
    YGZ_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)ygz_orb_stream(orb)));
    if ((rc = pj_common(h, orb, cam, n_pairs, frame, frame, u_right, occupied, &A))) return rc;
    const size_t n = (size_t)n_pairs * M;
    for (size_t i = 0; i < n; i++) YGZ_ARG_CHECK(!usable[i] || (level[i] >= 0 && level[i] < 16));
    A.M = M;
    const uint4* qd = nullptr;
    if ((rc = pj_upload(h, PJ_PX, proj_x, n, &A.proj_x)) || (rc = pj_upload(h, PJ_PY, proj_y, n, &A.proj_y)) ||
        (rc = pj_upload(h, PJ_PXR, proj_xr, n, &A.proj_xr)) || (rc = pj_upload(h, PJ_VC, view_cos, n, &A.view_cos)) ||
        (rc = pj_upload(h, PJ_LVL, level, n, &A.level)) || (rc = pj_upload(h, PJ_USABLE, usable, n, &A.usable)) ||
        (rc = pj_upload(h, PJ_QDESC, reinterpret_cast<const uint4*>(mp_desc), n * 2, &qd))) return rc;
    A.q_desc = qd;
    A.th = th; A.nnratio = nnratio; A.check_level = check_level;
    if ((rc = pj_launch(h, A, n_pairs, 1, M, h->stream))) return rc;
    return pj_finish(h, A, n_pairs, out_matches, out_nmatches, out_rounds);
}

/* device-timed repeats of the last batched projection search (inputs resident) on the extractor's stream, for bench.py */
int ygz_search_by_projection_batch_time(ygz_matcher* h, ygz_orb* orb, int iters, float* ms) {
    YGZ_ARG_CHECK(h && orb && ms && iters >= 1 && h->pj_last_pairs >= 1);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    void *d_kps, *d_desc, *d_counts;
    int cap = 0, rc;
    if ((rc = ygz_orb_device_outputs(orb, &d_kps, &d_desc, &d_counts, &cap))) return rc;
    if (cap != h->pj_last.cap) { set_last_error("the extractor's output capacity changed since the last batched projection search"); return YGZ_ERR_ARG; }
    ProjArgs A = h->pj_last;
    A.kps = (const ygz_keypoint*)d_kps; A.desc = (const uint4*)d_desc; A.counts = (const int*)d_counts;
    cudaStream_t st = (cudaStream_t)ygz_orb_stream(orb);
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    YGZ_CUDA_TRY(cudaEventRecord(h->ev0, st));
    for (int i = 0; i < iters; i++)
        if ((rc = pj_launch(h, A, h->pj_last_pairs, h->pj_last_mode, h->pj_last_nq, st))) return rc;
    YGZ_CUDA_TRY(cudaEventRecord(h->ev1, st));
    YGZ_CUDA_TRY(cudaEventSynchronize(h->ev1));
    YGZ_CUDA_TRY(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return YGZ_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// The Sim3 family of the back end (loop closing), src/ORBmatcher.cc:265-373 (SearchByProjection(KF, Scw, ...)), :888-1004
// (Fuse(KF, Scw, ...)), :1006-1216 (SearchBySim3).  Same pattern as the front-end searches: the host projects the points (CV_32F
// cv::Mat algebra upstream: a product accumulates in double and rounds once) and enumerates the window candidates, ONE launch of
// k_hamming_pairs computes every candidate distance, the order-dependent decisions follow on the host in the reference's order.
namespace {
struct SimCamH { float R[9], t[3], Ow[3]; };
inline void mat33_vec_h(const float* R, const float* p, float* o) {
    for (int i = 0; i < 3; i++) { double a = 0; for (int k = 0; k < 3; k++) a += (double)R[3 * i + k] * p[k]; o[i] = (float)a; }
}
inline float norm3_h(const float* p) { double a = 0; for (int k = 0; k < 3; k++) a += (double)p[k] * p[k]; return (float)std::sqrt(a); }
inline double dot3_h(const float* a, const float* b) { double r = 0; for (int k = 0; k < 3; k++) r += (double)a[k] * b[k]; return r; }
void decompose_scw_h(const float* S, SimCamH& C) {          // :274-278
    double d = 0;
    for (int k = 0; k < 3; k++) d += (double)S[k] * S[k];
    const float scw = (float)std::sqrt(d);
    const double inv = 1.0 / scw;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C.R[3 * i + j] = (float)(S[4 * i + j] * inv);
    for (int i = 0; i < 3; i++) C.t[i] = (float)(S[4 * i + 3] * inv);
    float nRt[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) nRt[3 * i + j] = (float)(C.R[3 * j + i] * -1.0);
    mat33_vec_h(nRt, C.t, C.Ow);
}
inline int predict_scale_h(float max_dist_raw, float dist, const ygz_frame_view* kf) {     // MapPoint::PredictScale, src/MapPoint.cc:343-357
    const float ratio = max_dist_raw / dist;
    int l = (int)std::ceil(std::log(ratio) / kf->log_scale_factor);
    if (l < 0) l = 0; else if (l >= kf->n_levels) l = kf->n_levels - 1;
    return l;
}
inline bool in_image_h(const ygz_frame_view* kf, float x, float y) {                        // KeyFrame::IsInImage (integer bounds)
    return x >= (int)kf->min_x && x < (int)kf->max_x && y >= (int)kf->min_y && y < (int)kf->max_y;
}
struct SimJob { int q, pMP, level, first, count; };
// project map point pMP through (R, t) [+ the depth / angle gates when Ow is given], predict its level, list the window candidates
// of `kf` (level band [l-1, l]) into (pq, pt).  invz_double: `1.0 / z` (Fuse, SearchBySim3) vs `1 / z` (SearchByProjection).
bool sim_candidates(const ygz_frame_view* kf, const Grid& g, const ygz_map_points* mps, int pMP, const float* pc, float fx, float fy,
                    float cx, float cy, const float* Ow, bool invz_double, float th, std::vector<int>& cand, int* level_out) {
    const float invz = invz_double ? (float)(1.0 / pc[2]) : 1 / pc[2];
    const float x = pc[0] * invz, y = pc[1] * invz;
    const float u = fx * x + cx, v = fy * y + cy;
    if (!in_image_h(kf, u, v)) return false;
    const float maxDistance = 1.2f * mps->max_dist[pMP], minDistance = 0.8f * mps->min_dist[pMP];
    float dist;
    if (Ow) {
        const float* Pw = mps->world_pos + 3 * (size_t)pMP;
        float PO[3];
        for (int k = 0; k < 3; k++) PO[k] = Pw[k] - Ow[k];
        dist = norm3_h(PO);
        if (dist < minDistance || dist > maxDistance) return false;
        if (dot3_h(PO, mps->normal + 3 * (size_t)pMP) < 0.5 * dist) return false;       // viewing angle below 60 degrees
    } else {
        dist = norm3_h(pc);
        if (dist < minDistance || dist > maxDistance) return false;
    }
    const int lvl = predict_scale_h(mps->max_dist[pMP], dist, kf);
    const float radius = th * kf->scale_factors[lvl];
    std::vector<int> all;
    features_in_area(kf, g, u, v, radius, -1, -1, all);
    cand.clear();
    for (int idx : all) {
        const int o = kf->keys[idx].octave;
        if (o < lvl - 1 || o > lvl) continue;
        cand.push_back(idx);
    }
    *level_out = lvl;
    return !cand.empty();
}
}  // namespace

extern "C" {

int ygz_search_by_projection_sim3(ygz_matcher* h, const ygz_frame_view* kf, const float* Scw, const ygz_map_points* mps, const int* point_ids,
                                  int n_points, int* matched, int th) {
    YGZ_ARG_CHECK(h && kf && Scw && mps && mps->normal && matched && (n_points == 0 || point_ids));
    SimCamH C;
    decompose_scw_h(Scw, C);
    std::vector<char> already(mps->m, 0);
    for (int i = 0; i < kf->n; i++) if (matched[i] >= 0) { YGZ_ARG_CHECK(matched[i] < mps->m); already[matched[i]] = 1; }
    Grid g(kf);
    std::vector<SimJob> jobs;
    std::vector<int> pq, pt, cand;
    for (int i = 0; i < n_points; i++) {
        const int pMP = point_ids[i];
        YGZ_ARG_CHECK(pMP >= 0 && pMP < mps->m);
        if (mp_bad(mps, pMP) || already[pMP]) continue;
        float pc[3];
        mat33_vec_h(C.R, mps->world_pos + 3 * (size_t)pMP, pc);
        for (int k = 0; k < 3; k++) pc[k] = pc[k] + C.t[k];
        if (pc[2] < 0.0) continue;
        int lvl;
        if (!sim_candidates(kf, g, mps, pMP, pc, kf->fx, kf->fy, kf->cx, kf->cy, C.Ow, false, (float)th, cand, &lvl)) continue;
        jobs.push_back(SimJob{i, pMP, lvl, (int)pq.size(), (int)cand.size()});
        for (int idx : cand) { pq.push_back(pMP); pt.push_back(idx); }
    }
    std::vector<int> dist;
    int rc = pair_distances(h, mps->desc, mps->m, kf->desc, kf->n, pq, pt, dist);
    if (rc) return rc;
    int nmatches = 0;
    for (const SimJob& j : jobs) {
        int bestDist = 256, bestIdx = -1;
        for (int p = j.first; p < j.first + j.count; p++) {
            if (matched[pt[p]] >= 0) continue;                  // vpMatched[idx] set by an earlier point of this very loop, too
            if (dist[p] < bestDist) { bestDist = dist[p]; bestIdx = pt[p]; }
        }
        if (bestDist <= TH_LOW) { matched[bestIdx] = j.pMP; nmatches++; }
    }
    return nmatches;
}

int ygz_fuse_sim3(ygz_matcher* h, ygz_frame_view* kf, const float* Scw, const ygz_map_points* mps, const int* point_ids, int n_points,
                  float th, int* replace_point) {
    YGZ_ARG_CHECK(h && kf && Scw && mps && mps->normal && kf->map_point && replace_point && (n_points == 0 || point_ids));
    SimCamH C;
    decompose_scw_h(Scw, C);
    std::vector<char> already(mps->m, 0);                       // pKF->GetMapPoints() at entry
    for (int i = 0; i < kf->n; i++) if (kf->map_point[i] >= 0 && !mp_bad(mps, kf->map_point[i])) already[kf->map_point[i]] = 1;
    Grid g(kf);
    std::vector<SimJob> jobs;
    std::vector<int> pq, pt, cand;
    for (int i = 0; i < n_points; i++) {
        const int pMP = point_ids[i];
        YGZ_ARG_CHECK(pMP >= 0 && pMP < mps->m);
        if (mp_bad(mps, pMP) || already[pMP]) continue;
        float pc[3];
        mat33_vec_h(C.R, mps->world_pos + 3 * (size_t)pMP, pc);
        for (int k = 0; k < 3; k++) pc[k] = pc[k] + C.t[k];
        if (pc[2] < 0.0f) continue;
        int lvl;
        if (!sim_candidates(kf, g, mps, pMP, pc, kf->fx, kf->fy, kf->cx, kf->cy, C.Ow, true, th, cand, &lvl)) continue;
        jobs.push_back(SimJob{i, pMP, lvl, (int)pq.size(), (int)cand.size()});
        for (int idx : cand) { pq.push_back(pMP); pt.push_back(idx); }
    }
    std::vector<int> dist;
    int rc = pair_distances(h, mps->desc, mps->m, kf->desc, kf->n, pq, pt, dist);
    if (rc) return rc;
    int nFused = 0;
    for (const SimJob& j : jobs) {
        int bestDist = INT_MAX, bestIdx = -1;
        for (int p = j.first; p < j.first + j.count; p++)
            if (dist[p] < bestDist) { bestDist = dist[p]; bestIdx = pt[p]; }
        if (bestDist <= TH_LOW) {
            const int pMPinKF = kf->map_point[bestIdx];
            if (pMPinKF >= 0) {
                if (!mp_bad(mps, pMPinKF)) replace_point[j.q] = pMPinKF;
            } else {
                kf->map_point[bestIdx] = j.pMP;                 // AddObservation + AddMapPoint: the slot is taken for the points that follow
            }
            nFused++;
        }
    }
    return nFused;
}

int ygz_fuse_candidates(ygz_matcher* h, const ygz_frame_view* kf, const float* Ow, const float* inv_level_sigma2, const ygz_map_points* mps,
                        const int* point_ids, int n_points, float th, int* best_idx, int* best_dist) {
    YGZ_ARG_CHECK(h && kf && Ow && inv_level_sigma2 && mps && mps->normal && mps->min_dist && mps->max_dist && best_idx && best_dist &&
                  (n_points == 0 || point_ids));
    Grid g(kf);
    std::vector<SimJob> jobs;
    std::vector<int> pq, pt, all;
    const float* R = kf->Rcw;
    const float* t = kf->tcw;
    for (int i = 0; i < n_points; i++) {
        best_idx[i] = -1;
        best_dist[i] = 256;
        const int pMP = point_ids[i];
        if (pMP < 0) continue;                                              // NULL entry of vpMapPoints
        YGZ_ARG_CHECK(pMP < mps->m);
        const float* Pw = mps->world_pos + 3 * (size_t)pMP;
        float pc[3];
        for (int r = 0; r < 3; r++) {                                       // Eigen Matrix3f * Vector3f + Vector3f: float, left to right
            float a = R[3 * r] * Pw[0];
            a += R[3 * r + 1] * Pw[1];
            a += R[3 * r + 2] * Pw[2];
            pc[r] = a + t[r];
        }
        if (pc[2] < 0.0f) continue;
        const float invz = 1 / pc[2];
        const float u = kf->fx * (pc[0] * invz) + kf->cx, v = kf->fy * (pc[1] * invz) + kf->cy;
        if (!in_image_h(kf, u, v)) continue;
        const float ur = u - kf->mbf * invz;
        float PO[3];
        for (int k = 0; k < 3; k++) PO[k] = Pw[k] - Ow[k];
        float sq = PO[0] * PO[0]; sq += PO[1] * PO[1]; sq += PO[2] * PO[2];
        const float dist3D = std::sqrt(sq);
        if (dist3D < 0.8f * mps->min_dist[pMP] || dist3D > 1.2f * mps->max_dist[pMP]) continue;
        const float* Pn = mps->normal + 3 * (size_t)pMP;
        float dt = PO[0] * Pn[0]; dt += PO[1] * Pn[1]; dt += PO[2] * Pn[2];
        if (dt < 0.5 * dist3D) continue;
        const int lvl = predict_scale_h(mps->max_dist[pMP], dist3D, kf);
        all.clear();                                                        // features_in_area appends
        features_in_area(kf, g, u, v, th * kf->scale_factors[lvl], -1, -1, all);
        const int first = (int)pq.size();
        for (int idx : all) {
            const ygz_keypoint& kp = kf->keys[idx];
            if (kp.octave < lvl - 1 || kp.octave > lvl) continue;
            const float ex = u - kp.x, ey = v - kp.y;
            const float kr = uright(kf, idx);
            if (kr >= 0) {                                                  // stereo key: three-dimensional reprojection error
                const float er = ur - kr;
                const float e2 = ex * ex + ey * ey + er * er;
                if (e2 * inv_level_sigma2[kp.octave] > 7.8) continue;
            } else {
                const float e2 = ex * ex + ey * ey;
                if (e2 * inv_level_sigma2[kp.octave] > 5.99) continue;
            }
            pq.push_back(pMP); pt.push_back(idx);
        }
        if ((int)pq.size() > first) jobs.push_back(SimJob{i, pMP, lvl, first, (int)pq.size() - first});
    }
    std::vector<int> dist;
    int rc = pair_distances(h, mps->desc, mps->m, kf->desc, kf->n, pq, pt, dist);
    if (rc) return rc;
    for (const SimJob& j : jobs) {
        int bestDist = 256, bestIdx = -1;
        for (int p = j.first; p < j.first + j.count; p++)
            if (dist[p] < bestDist) { bestDist = dist[p]; bestIdx = pt[p]; }
        best_idx[j.q] = bestIdx;
        best_dist[j.q] = bestDist;
    }
    return YGZ_OK;
}

int ygz_search_by_sim3(ygz_matcher* h, const ygz_frame_view* kf1, const ygz_frame_view* kf2, const ygz_map_points* mps, int* matches12,
                       const int* idx_in_kf2, float s12, const float* R12, const float* t12, float th) {
    YGZ_ARG_CHECK(h && kf1 && kf2 && mps && matches12 && idx_in_kf2 && R12 && t12 && kf1->map_point && kf2->map_point && s12 > 0.f);
    const float fx = kf1->fx, fy = kf1->fy, cx = kf1->cx, cy = kf1->cy;          // pKF1's intrinsics in both directions, as upstream
    float sR12[9], sR21[9], nsR21[9], t21[3];
    const double inv_s = 1.0 / s12;
    for (int i = 0; i < 9; i++) sR12[i] = (float)(R12[i] * (double)s12);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) sR21[3 * i + j] = (float)(R12[3 * j + i] * inv_s);
    for (int i = 0; i < 9; i++) nsR21[i] = (float)(sR21[i] * -1.0);
    mat33_vec_h(nsR21, t12, t21);
    const int N1 = kf1->n, N2 = kf2->n;
    std::vector<char> am1(N1, 0), am2(N2, 0);
    for (int i = 0; i < N1; i++) {
        const int pMP = matches12[i];
        if (pMP >= 0) {
            YGZ_ARG_CHECK(pMP < mps->m);
            am1[i] = 1;
            const int i2 = idx_in_kf2[pMP];
            if (i2 >= 0 && i2 < N2) am2[i2] = 1;
        }
    }
    Grid g1(kf1), g2(kf2);
    // both directions are independent per point: enumerate all candidates, one distance launch per direction
    auto one_way = [&](const ygz_frame_view* from, const ygz_frame_view* to, const Grid& gto, const std::vector<char>& am, const float* sR,
                       const float* tt, std::vector<int>& out) -> int {
        std::vector<SimJob> jobs;
        std::vector<int> pq, pt, cand;
        for (int i = 0; i < from->n; i++) {
            const int pMP = from->map_point[i];
            if (pMP < 0 || am[i]) continue;
            if (pMP >= mps->m) { set_last_error("map point index out of range"); return YGZ_ERR_ARG; }
            if (mp_bad(mps, pMP)) continue;
            float pa[3], pb[3];
            mat33_vec_h(from->Rcw, mps->world_pos + 3 * (size_t)pMP, pa);
            for (int k = 0; k < 3; k++) pa[k] = pa[k] + from->tcw[k];
            mat33_vec_h(sR, pa, pb);
            for (int k = 0; k < 3; k++) pb[k] = pb[k] + tt[k];
            if (pb[2] < 0.0) continue;
            int lvl;
            if (!sim_candidates(to, gto, mps, pMP, pb, fx, fy, cx, cy, nullptr, true, th, cand, &lvl)) continue;
            jobs.push_back(SimJob{i, pMP, lvl, (int)pq.size(), (int)cand.size()});
            for (int idx : cand) { pq.push_back(pMP); pt.push_back(idx); }
        }
        std::vector<int> dist;
        int rc = pair_distances(h, mps->desc, mps->m, to->desc, to->n, pq, pt, dist);
        if (rc) return rc;
        for (const SimJob& j : jobs) {
            int bestDist = INT_MAX, bestIdx = -1;
            for (int p = j.first; p < j.first + j.count; p++)
                if (dist[p] < bestDist) { bestDist = dist[p]; bestIdx = pt[p]; }
            if (bestDist <= TH_HIGH) out[j.q] = bestIdx;
        }
        return YGZ_OK;
    };
    std::vector<int> vnMatch1(N1, -1), vnMatch2(N2, -1);
    int rc;
    if ((rc = one_way(kf1, kf2, g2, am1, sR21, t21, vnMatch1))) return rc;
    if ((rc = one_way(kf2, kf1, g1, am2, sR12, t12, vnMatch2))) return rc;
    int nFound = 0;
    for (int i1 = 0; i1 < N1; i1++) {
        const int i2 = vnMatch1[i1];
        if (i2 >= 0 && vnMatch2[i2] == i1) { matches12[i1] = kf2->map_point[i2]; nFound++; }
    }
    return nFound;
}

}  // extern "C"
