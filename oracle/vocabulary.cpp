// ORACLE — TEST INFRASTRUCTURE ONLY (see ygz_oracle.h).  CPU restatement of the DBoW2 pieces the front end uses,
// SURVEY.md §8(f) rank 4:
//   TemplatedVocabulary::loadFromBinaryFile   /root/reference/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1478-1522
//   TemplatedVocabulary::transform (feature)  :1249-1292      transform (features -> BowVector, FeatureVector) :1151-1227
//   FORB::distance                            Thirdparty/DBoW2/DBoW2/FORB.cpp:82-102
//   BowVector::addWeight / normalize          BowVector.cpp:34-46, 62-84      FeatureVector::addFeature  FeatureVector.cpp:31-45
// File layout: u32 nb_nodes, u32 size_node, i32 k, i32 L, i32 scoring, i32 weighting, then records of size_node bytes
// {i32 parent, u8 descriptor[32], f32 weight, u8 is_leaf}; node ids are record order starting at 1, word ids leaf order.
// The reference's `while (!f.eof())` loop processes the last record twice (one extra child with an identical
// descriptor, never selected because the comparison is strict, and one extra word id at the very end); the oracle reads
// each record once.  ORB-SLAM scoring is L1 with TF-IDF weighting: weights are summed per word in feature order (double)
// and the vector is L1-normalised.
// Pinned on the reference's own Vocabulary/ORBvoc.bin where present (tests/test_oracle_vocab.py): header values, node
// and word counts, tree depth.
#include "ygz_oracle.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>

struct orc_vocab {
    int k = 0, L = 0, scoring = 0, weighting = 0;
    std::vector<int> parent;
    std::vector<std::vector<int>> children;
    std::vector<uint8_t> desc;            // 32 B per node
    std::vector<float> weight;
    std::vector<int> word_id;             // -1 for inner nodes
    int n_words = 0;
};

static int hamming256(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; i++) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

extern "C" {

orc_vocab* orc_vocab_from_memory(const uint8_t* buf, size_t bytes) {
    if (bytes < 24) return nullptr;
    unsigned nb_nodes, size_node;
    std::memcpy(&nb_nodes, buf, 4); std::memcpy(&size_node, buf + 4, 4);
    orc_vocab* v = new orc_vocab();
    std::memcpy(&v->k, buf + 8, 4); std::memcpy(&v->L, buf + 12, 4);
    std::memcpy(&v->scoring, buf + 16, 4); std::memcpy(&v->weighting, buf + 20, 4);
    if (size_node != 41) { delete v; return nullptr; }
    const size_t nrec = (bytes - 24) / size_node;
    const size_t n = nrec + 1;
    v->parent.assign(n, 0); v->children.assign(n, {}); v->desc.assign(n * 32, 0); v->weight.assign(n, 0.f); v->word_id.assign(n, -1);
    for (size_t r = 0; r < nrec; r++) {
        const uint8_t* p = buf + 24 + r * size_node;
        const int nid = (int)r + 1;
        int par; std::memcpy(&par, p, 4);
        if (par < 0 || par >= nid) { delete v; return nullptr; }
        v->parent[nid] = par;
        v->children[par].push_back(nid);
        std::memcpy(&v->desc[(size_t)nid * 32], p + 4, 32);
        std::memcpy(&v->weight[nid], p + 36, 4);
        if (p[40]) v->word_id[nid] = v->n_words++;
    }
    (void)nb_nodes;
    return v;
}

orc_vocab* orc_vocab_load(const char* path) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return nullptr;
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> buf((size_t)sz);
    const size_t got = std::fread(buf.data(), 1, (size_t)sz, f);
    std::fclose(f);
    return got == (size_t)sz ? orc_vocab_from_memory(buf.data(), buf.size()) : nullptr;
}

void orc_vocab_destroy(orc_vocab* v) { delete v; }

void orc_vocab_info(const orc_vocab* v, int* k, int* L, int* scoring, int* weighting, int* n_nodes, int* n_words) {
    *k = v->k; *L = v->L; *scoring = v->scoring; *weighting = v->weighting; *n_nodes = (int)v->parent.size(); *n_words = v->n_words;
}

/* transform(feature, word_id, weight, &nid, levelsup) :1249-1282 for n descriptors */
void orc_vocab_transform(const orc_vocab* v, const uint8_t* desc, int n, int levelsup, int* word_id, float* weight, int* node_id) {
    const int nid_level = v->L - levelsup;
    for (int i = 0; i < n; i++) {
        const uint8_t* f = desc + (size_t)i * 32;
        int nid = 0, final_id = 0, current_level = 0;
        if (nid_level <= 0) nid = 0;
        do {
            ++current_level;
            const std::vector<int>& nodes = v->children[final_id];
            final_id = nodes[0];
            double best_d = hamming256(f, &v->desc[(size_t)final_id * 32]);
            for (size_t c = 1; c < nodes.size(); c++) {
                const double d = hamming256(f, &v->desc[(size_t)nodes[c] * 32]);
                if (d < best_d) { best_d = d; final_id = nodes[c]; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (v->word_id[final_id] < 0 && !v->children[final_id].empty());
        word_id[i] = v->word_id[final_id];
        weight[i] = v->weight[final_id];
        node_id[i] = nid;
    }
}

/* transform(features, BowVector&, FeatureVector&, levelsup) :1151-1227 with TF-IDF weighting and L1 scoring.
 * bow_ids / bow_vals: ascending word id, L1-normalised (capacity n); fv as CSR (ascending node id; fv_index holds n entries).
 * Returns the BowVector size; *fv_nodes = FeatureVector size. */
int orc_vocab_transform_bow(const orc_vocab* v, const uint8_t* desc, int n, int levelsup, int* bow_ids, double* bow_vals,
                            int* fv_nodes, int* fv_node_id, int* fv_offset, int* fv_index) {
    std::vector<int> wid(n), nid(n);
    std::vector<float> w(n);
    orc_vocab_transform(v, desc, n, levelsup, wid.data(), w.data(), nid.data());
    std::map<int, double> bow;
    std::map<int, std::vector<int>> fv;
    for (int i = 0; i < n; i++) {
        const double wv = w[i];
        if (wv > 0) { bow[wid[i]] += wv; fv[nid[i]].push_back(i); }
    }
    double norm = 0.0;
    for (auto& kv : bow) norm += std::fabs(kv.second);
    if (norm > 0.0) for (auto& kv : bow) kv.second /= norm;
    int nb = 0;
    for (auto& kv : bow) { bow_ids[nb] = kv.first; bow_vals[nb] = kv.second; nb++; }
    int nn = 0, off = 0;
    for (auto& kv : fv) {
        fv_node_id[nn] = kv.first; fv_offset[nn] = off;
        for (int i : kv.second) fv_index[off++] = i;
        nn++;
    }
    fv_offset[nn] = off;
    *fv_nodes = nn;
    return nb;
}

}  // extern "C"
