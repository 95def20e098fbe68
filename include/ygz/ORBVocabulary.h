// Host-side mirror of ygz::ORBVocabulary (= DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>,
// /root/reference/include/ORBVocabulary.h, Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h) over the C ABI of libygz_b200.so,
// for the two calls the front end makes: loadFromBinaryFile (:1478-1522) and
// transform(features, BowVector&, FeatureVector&, levelsup) (:1151-1227, called by Frame::ComputeBoW src/Frame.cc:495-501).
// The tree descent runs on the device (one warp per descriptor); the std::map assembly of the two vectors, whose insertion
// order defines the floating-point sums, is the reference's own code shape and stays on the host.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>
#include "../ygz_b200.h"

// The classes live in namespace ygz (the reference's) unless ygz/adapters.h moved them to ygz::flat to free the reference's
// class names for the literal-signature adapters.
#ifndef YGZ_MIRROR_BEGIN
#define YGZ_MIRROR_BEGIN namespace ygz {
#define YGZ_MIRROR_END }
#define YGZ_MIRROR_NS ygz
#endif
YGZ_MIRROR_BEGIN

typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
typedef std::map<WordId, WordValue> BowVector;                       // DBoW2/BowVector.h
typedef std::map<NodeId, std::vector<unsigned int> > FeatureVector;  // DBoW2/FeatureVector.h

class ORBVocabulary {
public:
    explicit ORBVocabulary(int device = 0) : device_(device) {}
    ~ORBVocabulary() { if (v_) ygz_vocab_destroy(v_); }
    ORBVocabulary(const ORBVocabulary&) = delete;
    ORBVocabulary& operator=(const ORBVocabulary&) = delete;

    bool loadFromBinaryFile(const std::string& filename) {
        if (v_) { ygz_vocab_destroy(v_); v_ = nullptr; }
        if (ygz_vocab_load(filename.c_str(), device_, &v_) != YGZ_OK) {
            std::fprintf(stderr, "ygz::ORBVocabulary: %s\n", ygz_last_error());
            v_ = nullptr;
            return false;
        }
        return true;
    }
    bool empty() const { return v_ == nullptr; }
    unsigned int size() const { int n = 0; if (v_) ygz_vocab_info(v_, nullptr, nullptr, nullptr, nullptr, nullptr, &n); return (unsigned)n; }

    // features: n x 32 bytes (rows of mDescriptors).  TF-IDF weighting + L1 scoring (what ORBvoc.bin declares).
    void transform(const uint8_t* features, int n, BowVector& v, FeatureVector& fv, int levelsup) const {
        v.clear();
        fv.clear();
        if (empty() || n <= 0) return;
        std::vector<int> wid(n), nid(n);
        std::vector<float> w(n);
        if (ygz_vocab_transform(v_, features, n, levelsup, wid.data(), w.data(), nid.data()) != YGZ_OK) {
            std::fprintf(stderr, "ygz::ORBVocabulary: %s\n", ygz_last_error());
            return;
        }
        for (int i = 0; i < n; i++) {
            if (w[i] > 0) {                                              // not stopped (:1181)
                v[(WordId)wid[i]] += (WordValue)w[i];                    // BowVector::addWeight
                fv[(NodeId)nid[i]].push_back((unsigned)i);               // FeatureVector::addFeature
            }
        }
        double norm = 0.0;                                               // BowVector::normalize(L1)
        for (BowVector::iterator it = v.begin(); it != v.end(); ++it) norm += std::fabs(it->second);
        if (norm > 0.0)
            for (BowVector::iterator it = v.begin(); it != v.end(); ++it) it->second /= norm;
    }
    ygz_vocab* handle() { return v_; }

private:
    int device_;
    ygz_vocab* v_ = nullptr;
};

YGZ_MIRROR_END  // namespace
