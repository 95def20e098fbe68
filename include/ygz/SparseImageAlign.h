// ygz::SparseImgAlign — header-only host mirror of the reference class (include/SparseImageAlign.h:14-113 and its
// base NLLSSolver, include/NLSSolver.h:31-164) over the C ABI: same constructor, run() returning n_meas/patch_area,
// getFisherInformation(), public knobs n_iter_ / eps_.
#pragma once
#include <cstdio>
#include <cstring>
#include "../ygz_b200.h"

// The classes live in namespace ygz (the reference's) unless ygz/adapters.h moved them to ygz::flat to free the reference's
// class names for the literal-signature adapters.
#ifndef YGZ_MIRROR_BEGIN
#define YGZ_MIRROR_BEGIN namespace ygz {
#define YGZ_MIRROR_END }
#define YGZ_MIRROR_NS ygz
#endif
YGZ_MIRROR_BEGIN

struct SE3fPOD { float q[4]; float t[3]; };     // Sophus::SE3f: unit quaternion (x,y,z,w) + translation

class SparseImgAlign {
public:
    enum Method { GaussNewton, LevenbergMarquardt };
    SparseImgAlign(int max_level, int min_level, int n_iter = 10, Method method = GaussNewton, bool display = false,
                   bool verbose = false, int device = 0)
        : n_iter_(n_iter), eps_(0.000001f), verbose_(verbose), max_level_(max_level), min_level_(min_level) {
        (void)method; (void)display;            // GaussNewton is the only method Tracking constructs (src/Tracking.cc:207)
        std::memset(H_, 0, sizeof(H_));
        if (ygz_sia_create(device, 1, 4096, 2 * 1920 * 1200, &h_) != YGZ_OK) {
            std::fprintf(stderr, "ygz::SparseImgAlign: %s\n", ygz_last_error());
            h_ = nullptr;
        }
    }
    ~SparseImgAlign() { if (h_) ygz_sia_destroy(h_); }
    SparseImgAlign(const SparseImgAlign&) = delete;
    SparseImgAlign& operator=(const SparseImgAlign&) = delete;

    // size_t run(Frame* ref_frame, Frame* cur_frame, SE3f& TCR), src/SparseImageAlign.cc:20-49.
    // `problem` carries the members of the two frames run() reads (INTEGRATION.md builds it from Frame*).
    size_t run(const ygz_sia_problem& problem, SE3fPOD& TCR) {
        if (!h_ || problem.n <= 0) return 0;                           // "no features to track" (:24-27)
        SE3fPOD out;
        int n = 0;
        // run() upstream overwrites n_iter_ with its own per-level table {10, 10, ...} (src/SparseImageAlign.cc:40-45): the constructor's
        // n_iter never takes effect (checked against the compiled reference, tests/test_ref_sia.py); the C ABI keeps it explicit
        n_iter_ = 10;
        int rc = ygz_sia_run(h_, &problem, 1, max_level_, min_level_, n_iter_, eps_, out.q, out.t, H_, &n);
        if (rc != YGZ_OK) { std::fprintf(stderr, "ygz::SparseImgAlign::run: %d %s\n", rc, ygz_last_error()); return 0; }
        TCR = out;
        return (size_t)n;
    }
    // Matrix<float,6,6> getFisherInformation(), src/SparseImageAlign.cc:51-55 (row-major 36 floats)
    void getFisherInformation(float I[36]) const {
        const float sigma_i_sq = 5e-4f * 255 * 255;
        for (int i = 0; i < 36; i++) I[i] = H_[i] / sigma_i_sq;
    }
    const float* getInformationMatrix() const { return H_; }

    int n_iter_;
    float eps_;
    bool verbose_;

protected:
    int max_level_, min_level_;
    float H_[36];
    ygz_sia* h_ = nullptr;
};

YGZ_MIRROR_END  // namespace
