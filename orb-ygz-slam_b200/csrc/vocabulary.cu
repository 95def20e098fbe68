// DBoW2 vocabulary on the device — SURVEY.md §8(f) rank 4:
//   TemplatedVocabulary::loadFromBinaryFile   /root/reference/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1478-1522
//   TemplatedVocabulary::transform(feature, word_id, weight, &nid, levelsup)   :1249-1282
//   FORB::distance                            Thirdparty/DBoW2/DBoW2/FORB.cpp:82-102
// `Frame::ComputeBoW` (src/Frame.cc:495-501, levelsup = 4) turns every descriptor into a word (leaf) and the id of its
// ancestor at level L - levelsup; the latter is the FeatureVector SearchByBoW walks.  The tree (1 082 073 nodes x 32 B
// for ORBvoc.bin = 35 MB) stays resident in HBM/L2; one WARP descends the tree per descriptor: lanes = the k children
// of the current node, 256-bit Hamming distance each, first-minimum by child order (the reference's strict `<`).
#include "common.cuh"
#include <algorithm>
#include <climits>
#include <cstdio>
#include <vector>

namespace {
using namespace ygz_b200;

// Device layout: nodes re-indexed breadth first so that the children of a node are CONTIGUOUS and in the reference's
// insertion order (ascending original id): a level of the descent is two dependent loads — the node's 16-byte record
// {first child, child count, word id, original id} and the children's descriptors (k x 32 contiguous bytes) — instead of
// the five of a CSR (count, begin, child id, descriptor, leaf flag).
struct VocabDev {
    const uint4* desc;          // 2 x uint4 per node, BFS order
    const int4* meta;           // {first_child, child_count, word_id (-1: inner node), original node id}
    const float* weight;
    int L;
};

// Lanes per descriptor (GROUP): 10 for trees with k <= 10 (ORBvoc.bin: k = 10) — three descents per warp, every lane of a group busy at
// every level — and 16 for wider trees (two descents per warp).  Measured on the ORBvoc-shaped tree, 256 frames: 16 lanes 0.146 ms,
// 10 lanes 0.135 ms; persistent CTAs with the top three levels of the tree (1 111 nodes, 53 KB) in shared memory: 0.150 ms (the top of
// the tree already hits L1; the shared-memory copy only costs occupancy) — not kept.
template <int GROUP> struct VocGroup { static constexpr int per_warp = 32 / GROUP; };

// mask_stopped: node id 0xFFFFFFFF for words of weight 0 — TemplatedVocabulary::transform adds a feature to the FeatureVector
// only `if (w > 0)` (TemplatedVocabulary.h:1186-1190), so SearchByBoW never sees the features of stopped words.
template <int GROUP>
__global__ void __launch_bounds__(256) k_vocab_transform(VocabDev V, const uint4* __restrict__ desc, size_t desc_stride_u4,
                                                         const int* __restrict__ counts, int cap, int n_total, int levelsup,
                                                         int mask_stopped, int* __restrict__ word_id, float* __restrict__ weight,
                                                         int* __restrict__ node_id) {
    constexpr int PER_WARP = 32 / GROUP;
    // feature f = (frame, slot): frame = f / cap, slot = f % cap (cap = n_total and counts = NULL for a flat list)
    const int lane = threadIdx.x & 31;
    const int grp = lane / GROUP;                       // group inside the warp; lanes past PER_WARP * GROUP idle
    const int gl = lane - grp * GROUP;
    const int f = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * PER_WARP + grp;
    bool live = grp < PER_WARP && f < n_total;
    if (live && counts) {
        const int frame = f / cap, slot = f - frame * cap;
        live = slot < __ldg(counts + frame);
    }
    const unsigned glanes = (GROUP == 32 ? 0xffffffffu : ((1u << GROUP) - 1u)) << (grp * GROUP);
    const unsigned gmask = __ballot_sync(0xffffffffu, live) & glanes;     // this group's lanes, if live (a group is live or dead as a whole)
    if (!live) return;
    const uint4* d = desc + (size_t)f * desc_stride_u4;
    const uint4 a0 = __ldg(d), a1 = __ldg(d + 1);
    const int nid_level = V.L - levelsup;
    int cur = 0, level = 0, nid_node = 0;
    int4 m = __ldg(V.meta);
    while (m.y > 0) {
        ++level;
        // first minimum in child order (the reference's strict <) = minimum of (distance << 20 | child): one REDUX over the group's lanes
        unsigned key = 0xFFFFFFFFu;
        for (int c = gl; c < m.y; c += GROUP) {
            const uint4* b = V.desc + 2 * (size_t)(m.x + c);
            const uint4 b0 = __ldg(b), b1 = __ldg(b + 1);
            const int dist = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                             __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
            key = min(key, ((unsigned)dist << 20) | (unsigned)c);
        }
        const int bestc = (int)(__reduce_min_sync(gmask, key) & 0xFFFFFu);
        cur = m.x + bestc;
        m = __ldg(V.meta + cur);
        if (level == nid_level) nid_node = m.w;
        if (m.z >= 0) break;                                   // isLeaf()
    }
    if (gl == 0) {
        const float w = __ldg(V.weight + cur);
        if (word_id) word_id[f] = m.z;
        if (weight) weight[f] = w;
        node_id[f] = (mask_stopped && !(w > 0.f)) ? -1 : nid_node;
    }
}

// one launch over n descriptors with the group width that fits the tree
static cudaError_t launch_vocab_transform(int k, cudaStream_t st, VocabDev V, const uint4* desc, size_t stride_u4, const int* counts, int cap, int n,
                                          int levelsup, int mask_stopped, int* word_id, float* weight, int* node_id) {
    if (k <= 10) {
        const int per_cta = (256 / 32) * (32 / 10);
        k_vocab_transform<10><<<ceil_div(n, per_cta), 256, 0, st>>>(V, desc, stride_u4, counts, cap, n, levelsup, mask_stopped, word_id, weight, node_id);
    } else {
        const int per_cta = (256 / 32) * (32 / 16);
        k_vocab_transform<16><<<ceil_div(n, per_cta), 256, 0, st>>>(V, desc, stride_u4, counts, cap, n, levelsup, mask_stopped, word_id, weight, node_id);
    }
    return cudaGetLastError();
}

}  // namespace

struct ygz_vocab {
    int device = 0;
    cudaStream_t stream = nullptr;
    int k = 0, L = 0, scoring = 0, weighting = 0, n_nodes = 0, n_words = 0;
    uint4* d_desc = nullptr;
    int4* d_meta = nullptr;
    float* d_weight = nullptr;
    // scratch for transforms
    int cap = 0;
    uint8_t* d_in = nullptr;
    int *d_wid = nullptr, *d_nid = nullptr;
    float* d_w = nullptr;
    VocabDev dev() const { return VocabDev{d_desc, d_meta, d_weight, L}; }
};

static int vocab_free(ygz_vocab* v) {
    if (!v) return YGZ_OK;
    cudaSetDevice(v->device);
    cudaFree(v->d_desc); cudaFree(v->d_meta); cudaFree(v->d_weight); cudaFree(v->d_in); cudaFree(v->d_wid); cudaFree(v->d_nid); cudaFree(v->d_w);
    if (v->stream) cudaStreamDestroy(v->stream);
    delete v;
    return YGZ_OK;
}

static int vocab_scratch(ygz_vocab* v, int n) {
    if (n <= v->cap) return YGZ_OK;
    cudaFree(v->d_in); cudaFree(v->d_wid); cudaFree(v->d_nid); cudaFree(v->d_w);
    v->d_in = nullptr; v->d_wid = v->d_nid = nullptr; v->d_w = nullptr; v->cap = 0;
    YGZ_CUDA_TRY(cudaMalloc(&v->d_in, (size_t)n * 32));
    YGZ_CUDA_TRY(cudaMalloc(&v->d_wid, (size_t)n * 4));
    YGZ_CUDA_TRY(cudaMalloc(&v->d_nid, (size_t)n * 4));
    YGZ_CUDA_TRY(cudaMalloc(&v->d_w, (size_t)n * 4));
    v->cap = n;
    return YGZ_OK;
}

namespace ygz_b200 {
// Frame::ComputeBoW's FeatureVector node ids (levelsup levels above the leaves) of the descriptors an extractor holds on the
// device, written to d_node_out[n_frames][cap] on the CALLER's stream (no host round trip): what the batched SearchByBoW
// consumes when a vocabulary is attached to the matcher (ygz_matcher_set_vocabulary).  Slots past a frame's count are not written.
int vocab_node_ids_async(ygz_vocab* v, const void* d_desc, const int* d_counts, int cap, int n_frames, int levelsup,
                         uint32_t* d_node_out, cudaStream_t st) {
    const int n = n_frames * cap;
    { NvtxRange nvtx_range("ygz/ComputeBoW");
    if (launch_vocab_transform(v->k, st, v->dev(), reinterpret_cast<const uint4*>(d_desc), 2, d_counts, cap, n, levelsup, 1, nullptr, nullptr,
                               reinterpret_cast<int*>(d_node_out)) != cudaSuccess) return YGZ_ERR_CUDA;
    }
    return YGZ_OK;
}
int vocab_device(const ygz_vocab* v) { return v->device; }
}  // namespace ygz_b200

extern "C" {

int ygz_vocab_from_memory(const uint8_t* buf, size_t bytes, int device, ygz_vocab** out) {
    YGZ_ARG_CHECK(buf && out && bytes >= 24);
    if (ygz_device_count() <= device) { set_last_error("no such CUDA device"); return YGZ_ERR_NO_DEVICE; }
    unsigned nb_nodes, size_node;
    int k, L, scoring, weighting;
    std::memcpy(&nb_nodes, buf, 4); std::memcpy(&size_node, buf + 4, 4); std::memcpy(&k, buf + 8, 4); std::memcpy(&L, buf + 12, 4);
    std::memcpy(&scoring, buf + 16, 4); std::memcpy(&weighting, buf + 20, 4);
    if (size_node != 41) { set_last_error("vocabulary: size_node != 41 (not an ORB binary vocabulary)"); return YGZ_ERR_ARG; }
    const size_t nrec = (bytes - 24) / size_node, n = nrec + 1;
    // each record once (the reference's eof loop re-reads the last one; the duplicate child can never be selected)
    std::vector<int> parent(n, 0), word(n, -1), cbegin(n, 0), children(nrec > 0 ? nrec : 1);
    std::vector<uint8_t> ccount(n, 0), desc(n * 32, 0);
    std::vector<float> weight(n, 0.f);
    std::vector<int> cnt(n, 0);
    int n_words = 0;
    for (size_t r = 0; r < nrec; r++) {
        const uint8_t* p = buf + 24 + r * size_node;
        const int nid = (int)r + 1;
        int par; std::memcpy(&par, p, 4);
        if (par < 0 || par >= nid) { set_last_error("vocabulary: parent id out of order"); return YGZ_ERR_ARG; }
        parent[nid] = par;
        if (++cnt[par] > 255) { set_last_error("vocabulary: more than 255 children"); return YGZ_ERR_ARG; }
        std::memcpy(&desc[(size_t)nid * 32], p + 4, 32);
        std::memcpy(&weight[nid], p + 36, 4);
        if (p[40]) word[nid] = n_words++;
    }
    int off = 0;
    for (size_t i = 0; i < n; i++) { cbegin[i] = off; off += cnt[i]; ccount[i] = (uint8_t)cnt[i]; }
    std::vector<int> fill(cbegin);
    for (size_t nid = 1; nid < n; nid++) children[fill[parent[nid]]++] = (int)nid;      // ascending node id = insertion order
    // breadth-first re-indexing: children of a node contiguous, in insertion order (nodes unreachable from the root are dropped)
    std::vector<int> order;
    order.reserve(n);
    order.push_back(0);
    std::vector<int4> meta;
    meta.reserve(n);
    for (size_t q = 0; q < order.size(); q++) {
        const int id = order[q];
        const int first = (int)order.size();
        for (int c = 0; c < cnt[id]; c++) order.push_back(children[cbegin[id] + c]);
        meta.push_back(make_int4(first, cnt[id], word[id], id));
    }
    const size_t nn = order.size();
    std::vector<uint8_t> desc2(nn * 32);
    std::vector<float> weight2(nn);
    for (size_t q = 0; q < nn; q++) {
        std::memcpy(&desc2[q * 32], &desc[(size_t)order[q] * 32], 32);
        weight2[q] = weight[order[q]];
    }
    if (k >= (1 << 20)) { set_last_error("branching factor above 2^20"); return YGZ_ERR_ARG; }     // k_vocab_transform packs (distance, child) in 32 bits
    YGZ_CUDA_TRY(cudaSetDevice(device));
    ygz_vocab* v = new ygz_vocab();
    v->device = device; v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting; v->n_nodes = (int)n; v->n_words = n_words;
#define TRY_OR_FREE(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); vocab_free(v); return YGZ_ERR_CUDA; } } while (0)
    TRY_OR_FREE(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
    TRY_OR_FREE(cudaMalloc(&v->d_desc, nn * 32));
    TRY_OR_FREE(cudaMalloc(&v->d_meta, nn * sizeof(int4)));
    TRY_OR_FREE(cudaMalloc(&v->d_weight, nn * 4));
    TRY_OR_FREE(cudaMemcpy(v->d_desc, desc2.data(), nn * 32, cudaMemcpyHostToDevice));
    TRY_OR_FREE(cudaMemcpy(v->d_meta, meta.data(), nn * sizeof(int4), cudaMemcpyHostToDevice));
    TRY_OR_FREE(cudaMemcpy(v->d_weight, weight2.data(), nn * 4, cudaMemcpyHostToDevice));
#undef TRY_OR_FREE
    (void)nb_nodes;
    *out = v;
    return YGZ_OK;
}

int ygz_vocab_load(const char* path, int device, ygz_vocab** out) {
    YGZ_ARG_CHECK(path && out);
    FILE* f = std::fopen(path, "rb");
    if (!f) { set_last_error(std::string("cannot open vocabulary file ") + path); return YGZ_ERR_ARG; }
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> buf((size_t)std::max(sz, 0L));
    const size_t got = std::fread(buf.data(), 1, buf.size(), f);
    std::fclose(f);
    if (got != buf.size()) { set_last_error("short read on vocabulary file"); return YGZ_ERR_ARG; }
    return ygz_vocab_from_memory(buf.data(), buf.size(), device, out);
}

int ygz_vocab_destroy(ygz_vocab* v) { return vocab_free(v); }

int ygz_vocab_info(const ygz_vocab* v, int* k, int* L, int* scoring, int* weighting, int* n_nodes, int* n_words) {
    YGZ_ARG_CHECK(v);
    if (k) *k = v->k;
    if (L) *L = v->L;
    if (scoring) *scoring = v->scoring;
    if (weighting) *weighting = v->weighting;
    if (n_nodes) *n_nodes = v->n_nodes;
    if (n_words) *n_words = v->n_words;
    return YGZ_OK;
}

int ygz_vocab_transform(ygz_vocab* v, const uint8_t* desc, int n, int levelsup, int* word_id, float* weight, int* node_id) {
    YGZ_ARG_CHECK(v && n >= 0 && (n == 0 || (desc && word_id && weight && node_id)));
    if (n == 0) return YGZ_OK;
    YGZ_CUDA_TRY(cudaSetDevice(v->device));
    int rc = vocab_scratch(v, n);
    if (rc) return rc;
    YGZ_CUDA_TRY(cudaMemcpyAsync(v->d_in, desc, (size_t)n * 32, cudaMemcpyHostToDevice, v->stream));
    { NvtxRange nvtx_range("ygz/ComputeBoW");
    YGZ_CUDA_TRY(launch_vocab_transform(v->k, v->stream, v->dev(), reinterpret_cast<const uint4*>(v->d_in), 2, nullptr, n, n, levelsup, 0, v->d_wid,
                                        v->d_w, v->d_nid));
    }
    YGZ_CUDA_TRY(cudaMemcpyAsync(word_id, v->d_wid, (size_t)n * 4, cudaMemcpyDeviceToHost, v->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(weight, v->d_w, (size_t)n * 4, cudaMemcpyDeviceToHost, v->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(node_id, v->d_nid, (size_t)n * 4, cudaMemcpyDeviceToHost, v->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(v->stream));
    return YGZ_OK;
}

/* transform of the descriptors an extractor holds on the device (its last batch of n_frames frames, zero copy):
 * outputs are [n_frames][cap] arrays with cap = ygz_orb_max_keypoints(orb) rows per frame (entries past a frame's count:
 * word -1, node -1, weight 0).  The node ids are what ygz_search_by_bow_batch takes as `node_of_feature`. */
int ygz_vocab_transform_orb(ygz_vocab* v, ygz_orb* orb, int n_frames, int levelsup, int* word_id, float* weight, int* node_id,
                            float* ms) {
    YGZ_ARG_CHECK(v && orb && n_frames >= 1 && n_frames <= ygz_orb_max_batch(orb));
    YGZ_CUDA_TRY(cudaSetDevice(v->device));
    void *d_kps = nullptr, *d_desc = nullptr, *d_counts = nullptr;
    int cap = 0;
    int rc = ygz_orb_device_outputs(orb, &d_kps, &d_desc, &d_counts, &cap);
    if (rc) return rc;
    {   // the vocabulary and the extractor must live on the same device (the kernel dereferences both)
        cudaPointerAttributes at;
        YGZ_CUDA_TRY(cudaPointerGetAttributes(&at, d_desc));
        if (at.device != v->device) { set_last_error("vocabulary and extractor are on different devices"); return YGZ_ERR_ARG; }
    }
    const int n = n_frames * cap;
    if ((rc = vocab_scratch(v, n))) return rc;
    YGZ_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)ygz_orb_stream(orb)));
    YGZ_CUDA_TRY(cudaMemsetAsync(v->d_wid, 0xFF, (size_t)n * 4, v->stream));      // slots past a frame's count: word -1, node -1, weight 0
    YGZ_CUDA_TRY(cudaMemsetAsync(v->d_nid, 0xFF, (size_t)n * 4, v->stream));
    YGZ_CUDA_TRY(cudaMemsetAsync(v->d_w, 0, (size_t)n * 4, v->stream));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ms) { YGZ_CUDA_TRY(cudaEventCreate(&e0)); YGZ_CUDA_TRY(cudaEventCreate(&e1)); YGZ_CUDA_TRY(cudaEventRecord(e0, v->stream)); }
    { NvtxRange nvtx_range("ygz/ComputeBoW");
    YGZ_CUDA_TRY(launch_vocab_transform(v->k, v->stream, v->dev(), reinterpret_cast<const uint4*>(d_desc), 2, reinterpret_cast<const int*>(d_counts), cap,
                                        n, levelsup, 0, v->d_wid, v->d_w, v->d_nid));
    }
    if (ms) { YGZ_CUDA_TRY(cudaEventRecord(e1, v->stream)); }
    if (word_id) YGZ_CUDA_TRY(cudaMemcpyAsync(word_id, v->d_wid, (size_t)n * 4, cudaMemcpyDeviceToHost, v->stream));
    if (weight) YGZ_CUDA_TRY(cudaMemcpyAsync(weight, v->d_w, (size_t)n * 4, cudaMemcpyDeviceToHost, v->stream));
    if (node_id) YGZ_CUDA_TRY(cudaMemcpyAsync(node_id, v->d_nid, (size_t)n * 4, cudaMemcpyDeviceToHost, v->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(v->stream));
    if (ms) { YGZ_CUDA_TRY(cudaEventElapsedTime(ms, e0, e1)); cudaEventDestroy(e0); cudaEventDestroy(e1); }
    return YGZ_OK;
}

}  // extern "C"
