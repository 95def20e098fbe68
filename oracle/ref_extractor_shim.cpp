// ORACLE — TEST INFRASTRUCTURE ONLY.  Never part of the shipped product path.
//
// C shim over the REFERENCE's own ygz::ORBextractor: /root/reference/src/ORBextractor.cc is compiled unmodified, from
// where it lies, against the mock OpenCV / Eigen / glog / Frame headers of oracle/ref_mock (oracle/Makefile target
// _ref/libextractor_ref.so; no reference source is copied into this repo).  tests/test_ref_extractor.py compares the
// oracle restatement (oracle/orb_extractor.cpp, dso_detect.cpp) with this library: that is what pins the octree list
// order, IC_Angle, computeOrbDescriptor, the cell loop and both operator() overloads to the reference's own code.
//
// Pointer-order tie-break: DistributeOctTree sorts pair<int, ExtractorNode*>, so equal node sizes are
// ordered by ADDRESS (src/ORBextractor.cc:656).  With ref_arena(1) every allocation made inside a call is served from a
// bump arena with strictly increasing addresses, i.e. address order == creation order — the definition the oracle and
// the CUDA kernel use.  With ref_arena(0) the process allocator decides (what an upstream build does).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <list>
#include <new>
#include <stdexcept>
#include <vector>

#include <opencv2/core/core.hpp>
#include "ORBextractor.h"
#include "Frame.h"

namespace {
// Only list nodes of std::list<ExtractorNode> are served from the arena: their addresses are the only ones the
// reference ever compares, and they all die inside DistributeOctTree, so nothing in the arena outlives a call
// (checked: `live` must be back to zero when the call returns).
const size_t kNodeBytes = sizeof(std::_List_node<ygz::ExtractorNode>);
struct Arena {
    char* base = nullptr;
    size_t cap = 0, off = 0;
    long live = 0;
    bool on = false, active = false;
};
thread_local Arena g_arena;
}  // namespace

void* operator new(size_t n) {
    Arena& a = g_arena;
    if (a.active && n == kNodeBytes) {
        size_t o = (a.off + 15) & ~(size_t)15;
        if (o + n > a.cap) std::abort();                    // order guarantee would be lost
        a.off = o + n; a.live++;
        return a.base + o;
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept {
    Arena& a = g_arena;
    if (a.base && (char*)p >= a.base && (char*)p < a.base + a.cap) { a.live--; return; }    // bump arena: no reuse
    std::free(p);
}
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }

namespace {
struct ArenaScope {
    ArenaScope() {
        Arena& a = g_arena;
        if (a.on) {
            if (!a.base) { a.cap = (size_t)256 << 20; a.base = (char*)std::malloc(a.cap); }
            a.off = 0; a.live = 0; a.active = true;
        }
    }
    ~ArenaScope() {
        if (g_arena.active && g_arena.live != 0) std::abort();
        g_arena.active = false;
    }
};

struct Ref {
    ygz::ORBextractor ex;
    ygz::Frame frame;
    Ref(int n, float sf, int nl, int ini, int mn) : ex(n, sf, nl, ini, mn) {}
};

int flatten(const std::vector<cv::KeyPoint>& kps, const cv::Mat& desc, float* kp_out, uint8_t* desc_out, int cap) {
    int n = (int)kps.size();
    for (int i = 0; i < n && i < cap; i++) {
        std::memcpy(kp_out + 7 * i, &kps[i], 28);
        std::memcpy(desc_out + 32 * i, desc.ptr(i), 32);
    }
    return n;
}
}  // namespace

static_assert(sizeof(cv::KeyPoint) == 28, "KeyPoint layout");

extern "C" {

void ref_arena(int on) { g_arena.on = on != 0; }

void* ref_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    return new Ref(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void ref_orb_destroy(void* h) { delete (Ref*)h; }

// getters of include/ORBextractor.h:82-104; which: 0 scale, 1 inverse scale, 2 sigma2, 3 inverse sigma2
void ref_orb_tables(void* h, int which, float* out) {
    Ref* r = (Ref*)h;
    std::vector<float> v = which == 0 ? r->ex.GetScaleFactors() : which == 1 ? r->ex.GetInverseScaleFactors()
                         : which == 2 ? r->ex.GetScaleSigmaSquares() : r->ex.GetInverseScaleSigmaSquares();
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
}

// operator()(InputArray image, InputArray mask, vector<KeyPoint>&, OutputArray)   src/ORBextractor.cc:970-1028
// kp_out: 7 x 4 bytes per key (cv::KeyPoint layout); returns the number of keys (may exceed cap: then truncated).
int ref_orb_extract(void* h, const uint8_t* img, int w, int hgt, int stride, float* kp_out, uint8_t* desc_out, int cap) {
    Ref* r = (Ref*)h;
    int n;
    {
        ArenaScope scope;
        cv::Mat image(hgt, w, CV_8UC1, (void*)img, (size_t)stride), desc;
        std::vector<cv::KeyPoint> kps;
        try {
            r->ex(image, cv::Mat(), kps, desc);
        } catch (const std::exception&) {       // e.g. std::length_error from vpIniNodes.resize(nIni < 0), :545
            return -2;
        }
        n = flatten(kps, desc, kp_out, desc_out, cap);
    }
    return n;
}

// ComputePyramid (src/ORBextractor.cc:1129-1150) followed by the copy of Frame::ComputeImagePyramid
// (src/Frame.cc:807-813: mvImagePyramid[l] = extractor level l, cloned).  Kept in the handle's Frame for the next call.
void ref_orb_frame_pyramid(void* h, const uint8_t* img, int w, int hgt, int stride) {
    Ref* r = (Ref*)h;
    cv::Mat image(hgt, w, CV_8UC1, (void*)img, (size_t)stride);
    r->ex.ComputePyramid(image);
    r->frame.mvImagePyramid.resize(r->ex.mvImagePyramid.size());
    for (size_t l = 0; l < r->ex.mvImagePyramid.size(); l++) r->frame.mvImagePyramid[l] = r->ex.mvImagePyramid[l].clone();
}

// level l of the extractor's pyramid with `border` (<= 19) pixels of its REFLECT_101 frame, into out (contiguous)
int ref_orb_level(void* h, int level, int border, uint8_t* out, int* w, int* hgt) {
    Ref* r = (Ref*)h;
    const cv::Mat& m = r->ex.mvImagePyramid[level];
    if (m.empty()) return -1;
    *w = m.cols + 2 * border; *hgt = m.rows + 2 * border;
    if (out)
        for (int y = -border; y < m.rows + border; y++)
            std::memcpy(out + (size_t)(y + border) * *w, m.data + (ptrdiff_t)y * (ptrdiff_t)m.step.buf[0] - border, *w);
    return 0;
}

// operator()(Frame*, vector<KeyPoint>&, OutputArray, KeyPointMethod, bool leftEye)   src/ORBextractor.cc:1031-1127
// on the pyramid stored by ref_orb_frame_pyramid.  exist: n_exist keys (28 B each) already in frame->mvKeys;
// as Frame::ExtractORB (src/Frame.cc:332-348) the same vector is passed as `_keypoints` and N = its size.
// method: 0 ORBSLAM_KEYPOINT, 2 DSO_KEYPOINT.  Returns the total number of keys; exist_out receives the refreshed
// existing keys (DSO path rewrites their angles, :1383-1385).
int ref_orb_extract_frame(void* h, int method, const float* exist, int n_exist, float* kp_out, uint8_t* desc_out, int cap) {
    Ref* r = (Ref*)h;
    int n;
    {
        ArenaScope scope;
        r->frame.mvKeys.resize(n_exist);
        for (int i = 0; i < n_exist; i++) std::memcpy(&r->frame.mvKeys[i], exist + 7 * i, 28);
        r->frame.N = n_exist;
        cv::Mat desc;
        try {
            r->ex(&r->frame, r->frame.mvKeys, desc, (ygz::ORBextractor::KeyPointMethod)method, true);
        } catch (const std::exception&) {
            return -2;
        }
        n = flatten(r->frame.mvKeys, desc, kp_out, desc_out, cap);
    }
    return n;
}

}  // extern "C"
