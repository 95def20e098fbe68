// ygz::ORBextractor — header-only host mirror of the reference class (include/ORBextractor.h:45-172) over the C ABI
// of libygz_b200.so.  Same constructor arguments, getters, public mvImagePyramid and operator() overloads over flat views; the
// OpenCV / Frame* typed signatures Tracking.cc / Frame.cc use are in ygz/adapters.h.
#pragma once
#include <cstdint>
#include <cstdio>
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

typedef ygz_keypoint KeyPoint;      // == cv::KeyPoint: pt.x, pt.y, size, angle, response, octave, class_id

struct ImageView {                  // CV_8UC1 image: data, cols, rows, step (cv::Mat fields)
    const uint8_t* data; int cols, rows, step;
    ImageView() : data(nullptr), cols(0), rows(0), step(0) {}
    ImageView(const uint8_t* d, int c, int r, int s) : data(d), cols(c), rows(r), step(s) {}
    bool empty() const { return !data || cols <= 0 || rows <= 0; }
};

struct PyramidLevel { int cols = 0, rows = 0; std::vector<uint8_t> data; };   // continuous, step == cols (src/Frame.cc:810-813)

// the members of ygz::Frame the extractor reads/writes (src/ORBextractor.cc:1037-1099)
struct ExtractorFrame {
    ImageView mImGray, mImRight;
    std::vector<KeyPoint> mvKeys;            // pre-existing (direct-tracked) keys, N = mvKeys.size()
};

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };
    typedef enum { ORBSLAM_KEYPOINT, FAST_KEYPOINT, DSO_KEYPOINT } KeyPointMethod;

    ORBextractor(int nfeatures_, float scaleFactor_, int nlevels_, int iniThFAST_, int minThFAST_, int device = 0)
        : nfeatures(nfeatures_), scaleFactor(scaleFactor_), nlevels(nlevels_), iniThFAST(iniThFAST_), minThFAST(minThFAST_),
          device_(device) {
        mvImagePyramid.resize(nlevels);
        // tables do not depend on the image size: build them through a throw-away geometry once a handle exists
    }
    ~ORBextractor() { if (h_) ygz_orb_destroy(h_); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // operator()(InputArray image, InputArray mask, vector<KeyPoint>&, OutputArray descriptors), src/ORBextractor.cc:970-1028.
    // Mask is ignored (include/ORBextractor.h:66).  descriptors: N x 32 bytes, row-major.
    void operator()(const ImageView& image, const ImageView& /*mask*/, std::vector<KeyPoint>& keypoints,
                    std::vector<uint8_t>& descriptors) {
        if (image.empty()) return;                                  // :972-973
        if (!ensure(image.cols, image.rows)) { keypoints.clear(); descriptors.clear(); return; }
        const int cap = ygz_orb_max_keypoints(h_);
        keypoints.resize(cap);
        descriptors.resize((size_t)cap * 32);
        int n = 0;
        int rc = ygz_orb_extract(h_, image.data, 1, (size_t)image.step * image.rows, image.step, keypoints.data(),
                                 descriptors.data(), &n, cap);
        if (rc != YGZ_OK) { report("operator()", rc); n = 0; }      // reference failure convention: 0 features
        keypoints.resize(n);
        descriptors.resize((size_t)n * 32);
        pyramid_stale_ = true;
    }

    // operator()(Frame*, vector<KeyPoint>&, OutputArray, KeyPointMethod, bool leftEye), src/ORBextractor.cc:1031-1127.
    // ORBSLAM_KEYPOINT (octree) and DSO_KEYPOINT (ComputeKeyPointsDSOSingleLevel, :1277-1381; mnGridSize is kept in the
    // handle like the reference's member) are device paths; FAST_KEYPOINT (ComputeKeyPointsFast, which upstream marks as
    // broken, :1190) is reported, not silently substituted.
    void operator()(ExtractorFrame* frame, std::vector<KeyPoint>& keypoints, std::vector<uint8_t>& descriptors,
                    KeyPointMethod method, bool leftEye = true) {
        if (method == FAST_KEYPOINT) {
            std::fprintf(stderr, "ygz::ORBextractor: KeyPointMethod FAST_KEYPOINT is not on the B200 path\n");
            return;
        }
        if (method == DSO_KEYPOINT) {
            const ImageView& im = leftEye ? frame->mImGray : frame->mImRight;
            if (im.empty() || !ensure(im.cols, im.rows)) return;
            const int n_existing = (int)keypoints.size();             // ComputeKeyPointsDSOSingleLevel(kps, frame->mvKeys)
            const int cap = 3 * (im.rows / 7) * (im.cols / 7) + n_existing;
            keypoints.resize(cap);
            descriptors.resize((size_t)cap * 32);
            int n = 0;
            int rc = ygz_orb_extract_frame_dso(h_, im.data, im.step, n_existing, keypoints.data(), descriptors.data(), &n, cap, nullptr);
            if (rc != YGZ_OK) { report("operator()(Frame*, DSO_KEYPOINT)", rc); n = n_existing; }
            keypoints.resize(n);
            descriptors.resize((size_t)n * 32);
            pyramid_stale_ = true;
            return;
        }
        const ImageView& im = leftEye ? frame->mImGray : frame->mImRight;
        if (im.empty() || !ensure(im.cols, im.rows)) return;
        const int n_existing = leftEye ? (int)keypoints.size() : 0;
        if (!leftEye) keypoints.clear();
        const int cap = ygz_orb_max_keypoints(h_) + n_existing;
        keypoints.resize(cap);
        descriptors.resize((size_t)cap * 32);
        int n = 0;
        int rc = ygz_orb_extract_frame(h_, im.data, im.step, n_existing, keypoints.data(), descriptors.data(), &n, cap);
        if (rc != YGZ_OK) { report("operator()(Frame*)", rc); n = n_existing; }
        keypoints.resize(n);
        descriptors.resize((size_t)n * 32);
        pyramid_stale_ = true;
    }

    // ComputePyramid(cv::Mat), src/ORBextractor.cc:1129-1150: fills mvImagePyramid (host copies, step == cols)
    void ComputePyramid(const ImageView& image) {
        if (image.empty() || !ensure(image.cols, image.rows)) return;
        int rc = ygz_orb_compute_pyramid(h_, image.data, 1, (size_t)image.step * image.rows, image.step);
        if (rc != YGZ_OK) { report("ComputePyramid", rc); return; }
        pyramid_stale_ = true;
        SyncPyramid();
    }
    // lazily mirrors the device pyramid into mvImagePyramid (callers index it: src/Frame.cc:812, :515,604,623)
    void SyncPyramid() {
        if (!h_ || !pyramid_stale_) return;
        for (int l = 0; l < nlevels; l++) {
            PyramidLevel& L = mvImagePyramid[l];
            ygz_orb_level_size(h_, l, &L.cols, &L.rows);
            L.data.resize((size_t)L.cols * L.rows);
            ygz_orb_get_level(h_, 0, l, 0, L.data.data());
        }
        pyramid_stale_ = false;
    }

    int GetLevels() { return nlevels; }
    float GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> GetScaleFactors() { return table(0); }
    std::vector<float> GetInverseScaleFactors() { return table(1); }
    std::vector<float> GetScaleSigmaSquares() { return table(2); }
    std::vector<float> GetInverseScaleSigmaSquares() { return table(3); }

    std::vector<PyramidLevel> mvImagePyramid;
    ygz_orb* handle() { return h_; }

protected:
    bool ensure(int w, int h) {
        if (h_ && w == w_ && h == hgt_) return true;
        if (h_) { ygz_orb_destroy(h_); h_ = nullptr; }
        ygz_orb_config c = {nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST, w, h, 1, device_};
        int rc = ygz_orb_create(&c, &h_);
        if (rc != YGZ_OK) { report("create", rc); h_ = nullptr; return false; }
        w_ = w; hgt_ = h;
        return true;
    }
    std::vector<float> table(int which) {
        std::vector<float> t[4];
        for (auto& v : t) v.resize(nlevels);
        if (!h_ && !ensure(752, 480)) return t[which];      // tables are size independent
        ygz_orb_tables(h_, t[0].data(), t[1].data(), t[2].data(), t[3].data(), nullptr);
        return t[which];
    }
    static void report(const char* where, int rc) { std::fprintf(stderr, "ygz::ORBextractor::%s: %d %s\n", where, rc, ygz_last_error()); }

    int nfeatures = 0;
    double scaleFactor = 0;
    int nlevels = 0, iniThFAST = 0, minThFAST = 0;
    int device_ = 0, w_ = 0, hgt_ = 0;
    ygz_orb* h_ = nullptr;
    bool pyramid_stale_ = false;

};

YGZ_MIRROR_END  // namespace
