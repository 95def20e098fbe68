// ORACLE — TEST INFRASTRUCTURE ONLY.  Never part of the shipped product path.
//
// C shim over the REFERENCE's own ygz::ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>
// (include/ORBVocabulary.h:32-33; Thirdparty/DBoW2/DBoW2/{TemplatedVocabulary.h, FORB.cpp, BowVector.cpp, FeatureVector.cpp,
// ScoringObject.cpp} compiled unmodified, from where they lie, against oracle/ref_mock's cv::Mat; oracle/Makefile target
// _ref/libvocab_ref.so).  Entry points mirror the oracle's orc_vocab_* (ygz_oracle.h) so tests/test_ref_vocab.py runs
// loadFromBinaryFile (:1478-1522) and transform (:1151-1292) of both on the same files and descriptors: word ids, weights,
// FeatureVector node ids, the normalised BowVector.
#include <opencv2/core/core.hpp>  // (and the standard headers it pulls) before the access macro below
#include <algorithm>
#include <cassert>
#include <fstream>
#include <limits>
#include <numeric>
#define protected public          // transform(feature, id, weight, nid, levelsup) and m_nodes are protected upstream
#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"
#undef protected
#include <cstdint>
#include <vector>

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;

extern "C" {

void *refv_load(const char *path) {
    ORBVocabulary *v = new ORBVocabulary();
    if (!v->loadFromBinaryFile(path)) { delete v; return nullptr; }
    return v;
}
void refv_destroy(void *h) { delete (ORBVocabulary *) h; }

void refv_info(void *h, int *k, int *L, int *scoring, int *weighting, int *n_nodes, int *n_words) {
    ORBVocabulary *v = (ORBVocabulary *) h;
    *k = v->getBranchingFactor(); *L = v->getDepthLevels();
    *scoring = (int) v->getScoringType(); *weighting = (int) v->getWeightingType();
    *n_nodes = (int) v->m_nodes.size(); *n_words = (int) v->size();
}

/* per feature: transform(feature, word_id, weight, &nid, levelsup) :1258-1292 */
void refv_transform(void *h, const uint8_t *desc, int n, int levelsup, int *word_id, float *weight, int *node_id) {
    ORBVocabulary *v = (ORBVocabulary *) h;
    for (int i = 0; i < n; i++) {
        cv::Mat f(1, 32, CV_8U, (void *) (desc + 32 * (size_t) i));
        DBoW2::WordId id; DBoW2::WordValue w; DBoW2::NodeId nid = (DBoW2::NodeId) -1;      // see tests/test_ref_vocab.py: upstream leaves it unwritten for early leaves
        v->transform(f, id, w, &nid, levelsup);
        word_id[i] = (int) id; weight[i] = (float) w; node_id[i] = (int) nid;
    }
}

/* transform(features, BowVector&, FeatureVector&, levelsup) :1200-1256, flattened like orc_vocab_transform_bow */
int refv_transform_bow(void *h, const uint8_t *desc, int n, int levelsup, int *bow_ids, double *bow_vals, int *fv_nodes,
                       int *fv_node_id, int *fv_offset, int *fv_index) {
    ORBVocabulary *v = (ORBVocabulary *) h;
    std::vector<cv::Mat> feats(n);
    for (int i = 0; i < n; i++) feats[i] = cv::Mat(1, 32, CV_8U, (void *) (desc + 32 * (size_t) i));
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    v->transform(feats, bv, fv, levelsup);
    int nb = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++nb) { bow_ids[nb] = (int) it->first; bow_vals[nb] = it->second; }
    int nn = 0, o = 0;
    fv_offset[0] = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++nn) {
        fv_node_id[nn] = (int) it->first;
        for (size_t j = 0; j < it->second.size(); j++) fv_index[o++] = (int) it->second[j];
        fv_offset[nn + 1] = o;
    }
    *fv_nodes = nn;
    return nb;
}

}  // extern "C"
