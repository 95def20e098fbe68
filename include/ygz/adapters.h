// ygz/adapters.h — the reference's LITERAL class surface over libygz_b200.so.
//
// include/ygz/{ORBextractor,ORBmatcher,SparseImageAlign}.h mirror the reference classes over flat views; this header adds
// the pointer-graph gather that SURVEY.md §8(b) assigns to the host shim, so that the call sites of Tracking.cc /
// Frame.cc / LocalMapping.cc compile unchanged:
//
//   ORBextractor::operator()(cv::InputArray, cv::InputArray, std::vector<cv::KeyPoint>&, cv::OutputArray)   include/ORBextractor.h:69
//   ORBextractor::operator()(Frame*, std::vector<cv::KeyPoint>&, cv::OutputArray, KeyPointMethod, bool)     :72
//   ORBextractor::ComputePyramid(cv::Mat) + public std::vector<cv::Mat> mvImagePyramid                     :76, :106
//   ORBmatcher::SearchByProjection(Frame&, const std::vector<MapPoint*>&, float, bool)                      include/ORBmatcher.h:62
//   ORBmatcher::SearchByProjection(Frame&, const Frame&, float, bool, bool)                                 :79
//   ORBmatcher::SearchByProjection(Frame&, KeyFrame*, const std::set<MapPoint*>&, float, int)               :84
//   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, std::vector<MapPoint*>&)                                     :98
//   ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, std::vector<MapPoint*>&)                                  :100
//   ORBmatcher::SearchForInitialization(Frame&, Frame&, std::vector<cv::Point2f>&, std::vector<int>&, int)  :103
//   ORBmatcher::SearchForTriangulation(KeyFrame*, KeyFrame*, Matrix3f&, vector<pair<size_t,size_t>>&, bool) :108
//   ORBmatcher::DescriptorDistance(const cv::Mat&, const cv::Mat&)                                          :47
//   ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, ...) / Fuse(KeyFrame*, cv::Mat Scw, ...) / SearchBySim3     :66, :123, :113
//   SparseImgAlign::run(Frame*, Frame*, SE3f&)                                                              include/SparseImageAlign.h:33
//
// The classes are templates over the SLAM types (duck-typed: only the members the reference methods themselves touch are
// used, each cited below), so the header compiles against the reference's real Frame.h / KeyFrame.h / MapPoint.h and
// against the small stand-ins of oracle/ref_mock — which is how this repo builds and runs it (oracle/Makefile target
// _ref/libadapters_e2e.so: the SAME caller code, oracle/ref_matcher_shim.cpp, is compiled once against the reference's
// ORBmatcher.cc and once against this header; tests/test_adapters_gpu.py compares the two on a B200).
//
// Usage in the reference tree: let include/ORBextractor.h, ORBmatcher.h and SparseImageAlign.h each be
//     #define YGZ_B200_REFERENCE_NAMES
//     #include <ygz/adapters.h>
// which declares ygz::ORBmatcher / ORBextractor / SparseImgAlign as the specialisations for ygz::Frame / KeyFrame / MapPoint (those
// may still be incomplete at that point; they must be complete where a method is called).  INTEGRATION.md §2.
#pragma once
#include <cmath>
#include <cstring>
#include <map>
#include <set>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

// the flat-view mirrors move to ygz::flat: ygz::ORBmatcher etc. are the reference's names and belong to the adapters
#ifdef YGZ_MIRROR_BEGIN
#error "include <ygz/adapters.h> before (or instead of) the flat-view headers <ygz/ORBmatcher.h> ..."
#endif
#define YGZ_MIRROR_BEGIN namespace ygz { namespace flat {
#define YGZ_MIRROR_END } }
#define YGZ_MIRROR_NS ygz::flat
#include "ORBextractor.h"
#include "ORBmatcher.h"
#include "SparseImageAlign.h"

namespace ygz {
namespace b200 {
template <class Frame, class KeyFrame, class MapPoint> class ORBmatcherT;
template <class Frame> class ORBextractorT;
template <class Frame> class SparseImgAlignT;
}  // namespace b200
#ifdef YGZ_B200_REFERENCE_NAMES
// The reference's own class names, for an unchanged Tracking.cc / Frame.cc.  Naming a specialisation does not instantiate it, so
// this works with Frame / KeyFrame / MapPoint still incomplete: the header can take the place of include/ORBextractor.h,
// ORBmatcher.h and SparseImageAlign.h, which Frame.h itself includes (Frame holds ORBextractor* members, include/Frame.h:225).
class Frame;
class KeyFrame;
class MapPoint;
typedef b200::ORBmatcherT<Frame, KeyFrame, MapPoint> ORBmatcher;
typedef b200::ORBextractorT<Frame> ORBextractor;
typedef b200::SparseImgAlignT<Frame> SparseImgAlign;
#endif
namespace b200 {
using flat::ExtractorFrame;
using flat::ImageView;
using flat::KeyPoint;
using flat::PyramidLevel;
using flat::SE3fPOD;

static_assert(sizeof(ygz_keypoint) == 28, "ygz_keypoint must have cv::KeyPoint's layout");

// ---------------------------------------------------------------------------------------------------------- helpers
// MapPoint::mfMinDistance / mfMaxDistance are private upstream (include/MapPoint.h:164-165); the searches need the raw
// values (PredictScale, src/MapPoint.cc:343-373).  If the member is accessible (friend declaration, or the mock) it is
// read directly; otherwise it is recovered from the public getter (0.8f * min, 1.2f * max, :333-341), at most one ulp off.
template <class MP> auto raw_max_distance(MP* p, int) -> decltype(p->mfMaxDistance) { return p->mfMaxDistance; }
template <class MP> float raw_max_distance(MP* p, long) { return p->GetMaxDistanceInvariance() / 1.2f; }
template <class MP> auto raw_min_distance(MP* p, int) -> decltype(p->mfMinDistance) { return p->mfMinDistance; }
template <class MP> float raw_min_distance(MP* p, long) { return p->GetMinDistanceInvariance() / 0.8f; }

// Sophus::SE3f (or any type with rotationMatrix() / translation()) -> row-major R, t
template <class SE3> inline void pose_to_Rt(const SE3& T, float R[9], float t[3]) {
    const auto Rm = T.rotationMatrix();
    const auto tv = T.translation();
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) R[3 * i + j] = Rm(i, j);
        t[i] = tv(i);
    }
}
// ... -> unit quaternion (x, y, z, w) + t.  Goes through the rotation matrix (Shepperd's method) so that it works for any
// SE3 type; the round trip costs ~1e-7 on the quaternion, far inside the 1e-4 pose tolerance of the aligner.
template <class SE3> inline void pose_to_qt(const SE3& T, float q[4], float t[3]) {
    float R[9];
    pose_to_Rt(T, R, t);
    const float tr = R[0] + R[4] + R[8];
    float x, y, z, w;
    if (tr > 0.f) {
        const float s = std::sqrt(tr + 1.f) * 2.f;
        w = 0.25f * s; x = (R[7] - R[5]) / s; y = (R[2] - R[6]) / s; z = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        const float s = std::sqrt(1.f + R[0] - R[4] - R[8]) * 2.f;
        w = (R[7] - R[5]) / s; x = 0.25f * s; y = (R[1] + R[3]) / s; z = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
        const float s = std::sqrt(1.f + R[4] - R[0] - R[8]) * 2.f;
        w = (R[2] - R[6]) / s; x = (R[1] + R[3]) / s; y = 0.25f * s; z = (R[5] + R[7]) / s;
    } else {
        const float s = std::sqrt(1.f + R[8] - R[0] - R[4]) * 2.f;
        w = (R[3] - R[1]) / s; x = (R[2] + R[6]) / s; y = (R[5] + R[7]) / s; z = 0.25f * s;
    }
    const float n = std::sqrt(x * x + y * y + z * z + w * w);
    q[0] = x / n; q[1] = y / n; q[2] = z / n; q[3] = w / n;
}
// unit quaternion (x, y, z, w) + t -> SE3(Matrix3f, Vector3f) (both Sophus and the mock have that constructor)
template <class SE3> inline SE3 qt_to_pose(const float q[4], const float t[3]) {
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    auto Rm = SE3().rotationMatrix();
    auto tv = SE3().translation();
    Rm(0, 0) = 1 - 2 * (y * y + z * z); Rm(0, 1) = 2 * (x * y - z * w); Rm(0, 2) = 2 * (x * z + y * w);
    Rm(1, 0) = 2 * (x * y + z * w); Rm(1, 1) = 1 - 2 * (x * x + z * z); Rm(1, 2) = 2 * (y * z - x * w);
    Rm(2, 0) = 2 * (x * z - y * w); Rm(2, 1) = 2 * (y * z + x * w); Rm(2, 2) = 1 - 2 * (x * x + y * y);
    tv(0) = t[0]; tv(1) = t[1]; tv(2) = t[2];
    return SE3(Rm, tv);
}

// N x 32 CV_8U descriptor matrix -> contiguous rows (zero copy when the matrix already is)
struct DescRows {
    std::vector<uint8_t> buf;
    const uint8_t* ptr = nullptr;
    template <class Mat> void fill(const Mat& m, int n) {
        if (n <= 0 || m.rows <= 0) { ptr = nullptr; return; }
        if ((size_t)m.step == 32 || n == 1) { ptr = m.data; return; }
        buf.resize((size_t)n * 32);
        for (int i = 0; i < n; i++) std::memcpy(&buf[(size_t)i * 32], m.data + (size_t)i * (size_t)m.step, 32);
        ptr = buf.data();
    }
};

// The MapPoint objects a call can touch, as the table ygz_map_points describes.  Accessors: src/MapPoint.cc:63-71
// (GetWorldPos), :284-287 (GetDescriptor), :186-190 (isBad), :132-135 (Observations); tracking fields include/MapPoint.h:114-119.
template <class MapPoint> struct PointTable {
    std::vector<MapPoint*> pts;
    std::unordered_map<MapPoint*, int> index;
    std::vector<float> wp, mind, maxd, px, py, pxr, vcos, nrm;
    std::vector<uint8_t> desc, bad, inview;
    std::vector<int> obs, lvl;

    int add(MapPoint* p) {
        if (!p) return -1;
        auto it = index.find(p);
        if (it != index.end()) return it->second;
        const int i = (int)pts.size();
        index[p] = i;
        pts.push_back(p);
        const auto w = p->GetWorldPos();
        wp.push_back(w(0)); wp.push_back(w(1)); wp.push_back(w(2));
        const auto nv = p->GetNormal();
        nrm.push_back(nv(0)); nrm.push_back(nv(1)); nrm.push_back(nv(2));
        const auto d = p->GetDescriptor();
        desc.resize(desc.size() + 32, 0);
        if (d.data) std::memcpy(&desc[(size_t)i * 32], d.data, 32);
        bad.push_back(p->isBad() ? 1 : 0);
        obs.push_back(p->Observations());
        mind.push_back(raw_min_distance(p, 0));
        maxd.push_back(raw_max_distance(p, 0));
        inview.push_back(p->mbTrackInView ? 1 : 0);
        px.push_back(p->mTrackProjX); py.push_back(p->mTrackProjY); pxr.push_back(p->mTrackProjXR);
        lvl.push_back(p->mnTrackScaleLevel); vcos.push_back(p->mTrackViewCos);
        return i;
    }
    MapPoint* at(int i) const { return i < 0 ? static_cast<MapPoint*>(nullptr) : pts[i]; }
    ygz_map_points view() const {
        ygz_map_points v = {};
        v.m = (int)pts.size();
        v.world_pos = wp.data(); v.desc = desc.data(); v.bad = bad.data(); v.observations = obs.data();
        v.min_dist = mind.data(); v.max_dist = maxd.data(); v.track_in_view = inview.data();
        v.track_proj_x = px.data(); v.track_proj_y = py.data(); v.track_proj_xr = pxr.data();
        v.track_scale_level = lvl.data(); v.track_view_cos = vcos.data(); v.normal = nrm.data();
        return v;
    }
};

// Frame / KeyFrame -> ygz_frame_view (+ the storage the view points into).  Members read: N, mvKeys, mDescriptors, mvuRight,
// mFeatVec, fx fy cx cy mb mbf, mnMinX.., mnScaleLevels, mvScaleFactors, mvInvScaleFactors, mfLogScaleFactor
// (include/Frame.h:231-328, include/KeyFrame.h:262-324).
template <class MapPoint> struct FrameGather {
    ygz_frame_view v;
    std::vector<int> mp;
    std::vector<uint8_t> outl;
    std::vector<uint32_t> node, idx;
    std::vector<int> off;
    DescRows rows;
    std::vector<float> ur;

    template <class F> void common(const F& f, int n) {
        std::memset(&v, 0, sizeof(v));
        v.n = n;
        static_assert(sizeof(f.mvKeys[0]) == sizeof(ygz_keypoint), "cv::KeyPoint layout");
        v.keys = n > 0 ? reinterpret_cast<const ygz_keypoint*>(f.mvKeys.data()) : nullptr;
        rows.fill(f.mDescriptors, n);
        v.desc = rows.ptr;
        if ((int)f.mvuRight.size() >= n && n > 0) v.u_right = f.mvuRight.data();
        v.fx = f.fx; v.fy = f.fy; v.cx = f.cx; v.cy = f.cy; v.mb = f.mb; v.mbf = f.mbf;
        v.min_x = (float)f.mnMinX; v.min_y = (float)f.mnMinY; v.max_x = (float)f.mnMaxX; v.max_y = (float)f.mnMaxY;
        v.n_levels = f.mnScaleLevels;
        v.scale_factors = f.mvScaleFactors.data();
        v.inv_scale_factors = f.mvInvScaleFactors.data();
        v.log_scale_factor = f.mfLogScaleFactor;
        node.clear(); idx.clear(); off.clear();
        for (auto it = f.mFeatVec.begin(); it != f.mFeatVec.end(); ++it) {          // std::map: ascending node id
            node.push_back((uint32_t)it->first);
            off.push_back((int)idx.size());
            for (size_t j = 0; j < it->second.size(); j++) idx.push_back((uint32_t)it->second[j]);
        }
        off.push_back((int)idx.size());
        v.fv_nodes = (int)node.size();
        v.fv_node_id = node.data(); v.fv_offset = off.data(); v.fv_index = idx.data();
    }
    template <class Frame> void frame(const Frame& F, PointTable<MapPoint>& T) {
        common(F, F.N);
        pose_to_Rt(F.mTcw, v.Rcw, v.tcw);
        mp.assign(F.N > 0 ? F.N : 1, -1);
        for (int i = 0; i < F.N && i < (int)F.mvpMapPoints.size(); i++) mp[i] = T.add(F.mvpMapPoints[i]);
        v.map_point = mp.data();
        if ((int)F.mvbOutlier.size() >= F.N && F.N > 0) {
            outl.resize(F.N);
            for (int i = 0; i < F.N; i++) outl[i] = F.mvbOutlier[i] ? 1 : 0;
            v.outlier = outl.data();
        }
    }
    template <class KeyFrame> void keyframe(KeyFrame* K, PointTable<MapPoint>& T) {
        common(*K, K->N);
        pose_to_Rt(K->GetPose(), v.Rcw, v.tcw);
        const std::vector<MapPoint*> m = K->GetMapPointMatches();
        mp.assign(K->N > 0 ? K->N : 1, -1);
        for (int i = 0; i < K->N && i < (int)m.size(); i++) mp[i] = T.add(m[i]);
        v.map_point = mp.data();
    }
    // mvpMapPoints <- index array the search wrote
    template <class Frame> void store(Frame& F, const PointTable<MapPoint>& T) const {
        for (int i = 0; i < F.N; i++) F.mvpMapPoints[i] = T.at(mp[i]);
    }
};

// ------------------------------------------------------------------------------------------------------- ORBmatcher
template <class Frame, class KeyFrame, class MapPoint>
class ORBmatcherT : public ::ygz::flat::ORBmatcher {
    typedef ::ygz::flat::ORBmatcher Base;
    typedef PointTable<MapPoint> Table;
    typedef FrameGather<MapPoint> Gather;

public:
    ORBmatcherT(float nnratio = 0.6, bool checkOri = true, int device = 0) : Base(nnratio, checkOri, device) {}

    // src/ORBmatcher.cc:1507-1523
    template <class Mat> static int DescriptorDistance(const Mat& a, const Mat& b) { return ygz_descriptor_distance(a.data, b.data); }

    // :43-126
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3, bool checkLevel = true) {
        Table T; Gather g;
        std::vector<int> q;
        q.reserve(vpMapPoints.size());
        for (MapPoint* p : vpMapPoints) q.push_back(T.add(p));
        g.frame(F, T);
        const ygz_map_points mps = T.view();
        const int n = ok(ygz_search_by_projection_points(h_, mfNNratio, &g.v, &mps, q.data(), (int)q.size(), th, checkLevel));
        g.store(F, T);
        return n;
    }
    // :1218-1350
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono, bool checkLevel = true) {
        Table T; Gather gc, gl;
        gl.frame(LastFrame, T);
        gc.frame(CurrentFrame, T);
        const ygz_map_points mps = T.view();
        const int n = ok(ygz_search_by_projection_last(h_, mbCheckOrientation, &gc.v, &gl.v, &mps, th, bMono, checkLevel));
        gc.store(CurrentFrame, T);
        return n;
    }
    // :1352-1469
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist) {
        Table T; Gather gc, gk;
        gk.keyframe(pKF, T);
        gc.frame(CurrentFrame, T);
        std::vector<uint8_t> found(T.pts.size() ? T.pts.size() : 1, 0);
        for (size_t i = 0; i < T.pts.size(); i++) found[i] = sAlreadyFound.count(T.pts[i]) ? 1 : 0;
        const ygz_map_points mps = T.view();
        const int n = ok(ygz_search_by_projection_kf(h_, mbCheckOrientation, &gc.v, &gk.v, &mps, found.data(), th, ORBdist));
        gc.store(CurrentFrame, T);
        return n;
    }
    // :155-263
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
        Table T; Gather gk, gf;
        gk.keyframe(pKF, T);
        gf.frame(F, T);
        const ygz_map_points mps = T.view();
        std::vector<int> out(F.N > 0 ? F.N : 1, -1);
        const int n = ok(ygz_search_by_bow(h_, mfNNratio, mbCheckOrientation, &gk.v, &mps, &gf.v, out.data()));
        vpMapPointMatches.assign(F.N, static_cast<MapPoint*>(nullptr));
        if (n > 0) for (int i = 0; i < F.N; i++) vpMapPointMatches[i] = T.at(out[i]);
        return n;
    }
    // :480-595
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
        Table T; Gather g1, g2;
        g1.keyframe(pKF1, T);
        g2.keyframe(pKF2, T);
        const ygz_map_points mps = T.view();
        std::vector<int> out(pKF1->N > 0 ? pKF1->N : 1, -1);
        const int n = ok(ygz_search_by_bow_kf(h_, mfNNratio, mbCheckOrientation, &g1.v, &g2.v, &mps, out.data()));
        vpMatches12.assign(pKF1->N, static_cast<MapPoint*>(nullptr));
        if (n > 0) for (int i = 0; i < pKF1->N; i++) vpMatches12[i] = T.at(out[i]);
        return n;
    }
    // :375-478.  cv::Point2f is two packed floats.
    template <class Point2f>
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10) {
        static_assert(sizeof(Point2f) == 2 * sizeof(float), "cv::Point2f layout");
        Table T; Gather g1, g2;
        g1.frame(F1, T);
        g2.frame(F2, T);
        vnMatches12.assign(F1.mvKeys.size() ? F1.mvKeys.size() : 1, -1);
        const int n = ok(ygz_search_for_initialization(h_, mfNNratio, mbCheckOrientation, &g1.v, &g2.v,
                                                       reinterpret_cast<float*>(vbPrevMatched.data()), vnMatches12.data(), windowSize));
        vnMatches12.resize(F1.mvKeys.size());
        return n;
    }
    // :597-746
    template <class Matrix3>
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, Matrix3& F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs,
                               const bool bOnlyStereo) {
        Table T; Gather g1, g2;
        g1.keyframe(pKF1, T);
        g2.keyframe(pKF2, T);
        // epipole in the second image, exactly as :603-609
        const auto Cw = pKF1->GetCameraCenter();
        const auto R2w = pKF2->GetRotation();
        const auto t2w = pKF2->GetTranslation();
        const auto C2 = R2w * Cw + t2w;
        const float invz = 1.0f / C2[2];
        const float ex = pKF2->fx * C2[0] * invz + pKF2->cx;
        const float ey = pKF2->fy * C2[1] * invz + pKF2->cy;
        float f12[9];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) f12[3 * i + j] = F12(i, j);
        std::vector<int> pairs((size_t)2 * (pKF1->N > 0 ? pKF1->N : 1));
        const int n = ok(ygz_search_for_triangulation(h_, mbCheckOrientation, &g1.v, &g2.v, f12, ex, ey, pKF2->mvLevelSigma2.data(),
                                                      bOnlyStereo, pairs.data()));
        vMatchedPairs.clear();
        vMatchedPairs.reserve(n);
        for (int i = 0; i < n; i++) vMatchedPairs.push_back(std::make_pair((size_t)pairs[2 * i], (size_t)pairs[2 * i + 1]));
        return n;
    }

    // ---- the Sim3 family of the back end (LoopClosing.cc).  Scw / R12 / t12 are CV_32F cv::Mat upstream: any type with rows, cols and at<float>().
    template <class Mat> static void mat_to_floats(const Mat& m, int r, int c, float* out) {
        for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) out[i * c + j] = m.template at<float>(i, j);
    }
    // SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th)  :265-373
    template <class Mat>
    int SearchByProjection(KeyFrame* pKF, Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th) {
        Table T; Gather gk;
        gk.keyframe(pKF, T);
        std::vector<int> ids, matched(pKF->N > 0 ? pKF->N : 1, -1);
        ids.reserve(vpPoints.size());
        for (MapPoint* p : vpPoints) ids.push_back(T.add(p));
        for (int i = 0; i < pKF->N && i < (int)vpMatched.size(); i++) matched[i] = T.add(vpMatched[i]);
        float S[16];
        mat_to_floats(Scw, 4, 4, S);
        const ygz_map_points mps = T.view();
        const int n = ok(ygz_search_by_projection_sim3(h_, &gk.v, S, &mps, ids.data(), (int)ids.size(), matched.data(), th));
        for (int i = 0; i < pKF->N && i < (int)vpMatched.size(); i++) vpMatched[i] = T.at(matched[i]);
        return n;
    }
    // Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint)  :888-1004
    template <class Mat>
    int Fuse(KeyFrame* pKF, Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint) {
        Table T; Gather gk;
        gk.keyframe(pKF, T);
        std::vector<int> ids, rep(vpPoints.size() ? vpPoints.size() : 1, -1);
        ids.reserve(vpPoints.size());
        for (MapPoint* p : vpPoints) ids.push_back(T.add(p));
        for (size_t i = 0; i < vpPoints.size() && i < vpReplacePoint.size(); i++) rep[i] = T.add(vpReplacePoint[i]);
        const std::vector<int> before = gk.mp;
        float S[16];
        mat_to_floats(Scw, 4, 4, S);
        const ygz_map_points mps = T.view();
        const int n = ok(ygz_fuse_sim3(h_, &gk.v, S, &mps, ids.data(), (int)ids.size(), th, rep.data()));
        for (size_t i = 0; i < vpPoints.size() && i < vpReplacePoint.size(); i++) vpReplacePoint[i] = T.at(rep[i]);
        for (int i = 0; i < pKF->N; i++)                       // slots the search filled: the two calls of :996-997, in slot order
            if (gk.mp[i] != before[i] && gk.mp[i] >= 0) { T.at(gk.mp[i])->AddObservation(pKF, (size_t)i); pKF->AddMapPoint(T.at(gk.mp[i]), (size_t)i); }
        return n;
    }
    // Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th = 3.0)  :748-886 (LocalMapping::SearchInNeighbors).
    // The per-point search runs once on the device (ygz_fuse_candidates); the loop below is upstream's, with the search replaced by its
    // result: the skip tests are evaluated when the point's turn comes and the graph calls (Replace / AddObservation / AddMapPoint, :866-880)
    // are made on the caller's objects in upstream's order.  A point that was bad or already in pKF at entry stays so, hence it is not searched.
    int Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th = 3.0) {
        Table T; Gather gk;
        gk.common(*pKF, pKF->N);
        const auto R = pKF->GetRotation();
        const auto t = pKF->GetTranslation();
        const auto ow = pKF->GetCameraCenter();
        for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) gk.v.Rcw[3 * i + j] = R(i, j); gk.v.tcw[i] = t(i); }
        const float Ow[3] = {ow(0), ow(1), ow(2)};
        std::vector<int> ids;
        ids.reserve(vpMapPoints.size());
        for (MapPoint* p : vpMapPoints) ids.push_back(p && !p->isBad() && !p->IsInKeyFrame(pKF) ? T.add(p) : -1);
        const ygz_map_points mps = T.view();
        std::vector<int> bestIdx, bestDist;
        if (!Base::FuseCandidates(gk.v, Ow, pKF->mvInvLevelSigma2.data(), mps, ids, th, bestIdx, bestDist)) return 0;
        int nFused = 0;
        for (size_t i = 0; i < vpMapPoints.size(); i++) {
            MapPoint* pMP = vpMapPoints[i];
            if (!pMP) continue;
            if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
            if (bestDist[i] <= Base::TH_LOW) {
                const size_t idx = (size_t)bestIdx[i];
                MapPoint* pMPinKF = pKF->GetMapPoint(idx);
                if (pMPinKF) {
                    if (!pMPinKF->isBad()) {
                        if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                        else pMPinKF->Replace(pMP);
                    }
                } else {
                    pMP->AddObservation(pKF, idx);
                    pKF->AddMapPoint(pMP, idx);
                }
                nFused++;
            }
        }
        return nFused;
    }
    // SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12, const cv::Mat& t12, float th)  :1006-1216
    template <class Mat>
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const Mat& R12, const Mat& t12, const float th) {
        Table T; Gather g1, g2;
        g1.keyframe(pKF1, T);
        g2.keyframe(pKF2, T);
        std::vector<int> m12(pKF1->N > 0 ? pKF1->N : 1, -1);
        for (int i = 0; i < pKF1->N && i < (int)vpMatches12.size(); i++) m12[i] = T.add(vpMatches12[i]);
        std::vector<int> idx2(T.pts.size() ? T.pts.size() : 1, -1);
        for (size_t i = 0; i < T.pts.size(); i++) idx2[i] = T.pts[i]->GetIndexInKeyFrame(pKF2);
        float R[9], t[3];
        mat_to_floats(R12, 3, 3, R);
        mat_to_floats(t12, 3, 1, t);
        const ygz_map_points mps = T.view();
        const int n = ok(ygz_search_by_sim3(h_, &g1.v, &g2.v, &mps, m12.data(), idx2.data(), s12, R, t, th));
        for (int i = 0; i < pKF1->N && i < (int)vpMatches12.size(); i++) vpMatches12[i] = T.at(m12[i]);
        return n;
    }

protected:
    using Base::ok;
};

// cv::InputArray / cv::OutputArray proxies have getMat() / create() / release(); a cv::Mat passed directly is its own Mat
template <class A> auto arr_mat(A& a, int) -> decltype(a.getMat()) { return a.getMat(); }
template <class A> A& arr_mat(A& a, long) { return a; }

// ----------------------------------------------------------------------------------------------------- ORBextractor
// Mat / KeyPoint types are taken from the arguments so that the header itself does not name OpenCV.
template <class Frame>
class ORBextractorT : public ::ygz::flat::ORBextractor {
    typedef ::ygz::flat::ORBextractor Base;

public:
    ORBextractorT(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device = 0)
        : Base(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device) {}

    // operator()(InputArray image, InputArray mask, vector<KeyPoint>&, OutputArray descriptors), src/ORBextractor.cc:970-1028
    template <class InArr, class MaskArr, class CvKeyPoint, class OutArr>
    void operator()(const InArr& image, const MaskArr& /*mask*/, std::vector<CvKeyPoint>& keypoints, OutArr&& descriptors) {
        static_assert(sizeof(CvKeyPoint) == sizeof(KeyPoint), "cv::KeyPoint layout");
        if (image.empty()) return;                                          // :972-973
        const auto im = arr_mat(image, 0);
        std::vector<KeyPoint> k;
        std::vector<uint8_t> d;
        Base::operator()(ImageView(im.data, im.cols, im.rows, (int)im.step), ImageView(), k, d);
        put(k, d, keypoints, descriptors);
    }
    // operator()(Frame*, vector<KeyPoint>&, OutputArray, KeyPointMethod, bool leftEye), :1031-1127.  The left pyramid is
    // frame->mvImagePyramid (:1037): its level 0 is the image it was built from, and the device rebuilds the same levels.
    template <class CvKeyPoint, class OutArr>
    void operator()(Frame* frame, std::vector<CvKeyPoint>& keypoints, OutArr&& descriptors, KeyPointMethod method, bool leftEye = true) {
        static_assert(sizeof(CvKeyPoint) == sizeof(KeyPoint), "cv::KeyPoint layout");
        ExtractorFrame ef;
        if (leftEye && !frame->mvImagePyramid.empty()) {
            const auto& l0 = frame->mvImagePyramid[0];
            ef.mImGray = ImageView(l0.data, l0.cols, l0.rows, (int)l0.step);
        }
        if (!leftEye) ef.mImRight = ImageView(frame->mImRight.data, frame->mImRight.cols, frame->mImRight.rows, (int)frame->mImRight.step);
        std::vector<KeyPoint> k(keypoints.size());
        if (!k.empty()) std::memcpy(static_cast<void*>(k.data()), keypoints.data(), k.size() * sizeof(KeyPoint));
        std::vector<uint8_t> d;
        Base::operator()(&ef, k, d, method, leftEye);
        put(k, d, keypoints, descriptors);
    }
    // ComputePyramid(cv::Mat image), :1129-1150: fills the public mvImagePyramid with headers over host copies
    template <class Mat> void ComputePyramid(const Mat& image) {
        Base::ComputePyramid(ImageView(image.data, image.cols, image.rows, (int)image.step));
        mvImagePyramid.resize(Base::mvImagePyramid.size());
        for (size_t l = 0; l < mvImagePyramid.size(); l++) {
            PyramidLevel& L = Base::mvImagePyramid[l];
            mvImagePyramid[l] = Mat(L.rows, L.cols, 0 /* CV_8U */, static_cast<void*>(L.data.data()));
        }
    }
    // the reference's public member (include/ORBextractor.h:106); element type = the Frame's own pyramid element (cv::Mat)
    typedef typename std::remove_reference<decltype(static_cast<Frame*>(nullptr)->mvImagePyramid[0])>::type MatT;
    std::vector<MatT> mvImagePyramid;

private:
    template <class CvKeyPoint, class OutArr>
    static void put(const std::vector<KeyPoint>& k, const std::vector<uint8_t>& d, std::vector<CvKeyPoint>& keypoints, OutArr& descriptors) {
        keypoints.resize(k.size());
        if (!k.empty()) std::memcpy(static_cast<void*>(keypoints.data()), k.data(), k.size() * sizeof(KeyPoint));
        if (k.empty()) { descriptors.release(); return; }                    // :1000-1001
        descriptors.create((int)k.size(), 32, 0 /* CV_8U */);
        auto m = arr_mat(descriptors, 0);
        for (size_t i = 0; i < k.size(); i++) std::memcpy(m.data + i * (size_t)m.step, &d[i * 32], 32);
    }
};

// --------------------------------------------------------------------------------------------------- SparseImgAlign
template <class Frame>
class SparseImgAlignT : public ::ygz::flat::SparseImgAlign {
    typedef ::ygz::flat::SparseImgAlign Base;

public:
    SparseImgAlignT(int max_level, int min_level, int n_iter = 10, Method method = GaussNewton, bool display = false, bool verbose = false,
                    int device = 0)
        : Base(max_level, min_level, n_iter, method, display, verbose, device) {}

    // size_t run(Frame* ref_frame, Frame* cur_frame, SE3f& TCR), src/SparseImageAlign.cc:20-49.  Reads what run() /
    // precomputeReferencePatches / computeResiduals read: both pyramids, ref mvKeys / mvpMapPoints / mvbOutlier / mTcw,
    // cur mTcw, Frame::fx.. and ref mvInvScaleFactors (:57-128, :130-231).
    template <class SE3> size_t run(Frame* ref, Frame* cur, SE3& TCR) {
        if (ref->mvKeys.empty()) return 0;                                  // "no features to track" (:24-27)
        const int nl = (int)ref->mvImagePyramid.size();
        std::vector<const uint8_t*> rl(nl), cl(nl);
        std::vector<int> w(nl), h(nl);
        std::vector<std::vector<uint8_t> > keep;
        auto level_ptr = [&keep](const decltype(ref->mvImagePyramid[0])& m) -> const uint8_t* {
            if ((int)m.step == m.cols) return m.data;                       // continuous (Frame clones its levels, src/Frame.cc:810-813)
            keep.emplace_back((size_t)m.cols * m.rows);
            for (int y = 0; y < m.rows; y++) std::memcpy(&keep.back()[(size_t)y * m.cols], m.data + (size_t)y * (size_t)m.step, m.cols);
            return keep.back().data();
        };
        for (int l = 0; l < nl; l++) {
            rl[l] = level_ptr(ref->mvImagePyramid[l]);
            cl[l] = level_ptr(cur->mvImagePyramid[l]);
            w[l] = ref->mvImagePyramid[l].cols; h[l] = ref->mvImagePyramid[l].rows;
        }
        const int n = (int)ref->mvKeys.size();
        std::vector<float> px(2 * (size_t)n), xyz(3 * (size_t)n, 0.f);
        std::vector<uint8_t> valid(n, 0);
        for (int i = 0; i < n; i++) {
            px[2 * i] = ref->mvKeys[i].pt.x; px[2 * i + 1] = ref->mvKeys[i].pt.y;
            auto* mp = i < (int)ref->mvpMapPoints.size() ? ref->mvpMapPoints[i] : nullptr;
            if (!mp || mp->isBad() || (i < (int)ref->mvbOutlier.size() && ref->mvbOutlier[i])) continue;   // :71-73
            const auto p = mp->GetWorldPos();
            xyz[3 * i] = p(0); xyz[3 * i + 1] = p(1); xyz[3 * i + 2] = p(2);
            valid[i] = 1;
        }
        ygz_sia_problem P = {};
        P.n_levels = nl; P.ref_levels = rl.data(); P.cur_levels = cl.data(); P.level_w = w.data(); P.level_h = h.data();
        P.inv_scale_factors = ref->mvInvScaleFactors.data();
        P.fx = ref->fx; P.fy = ref->fy; P.cx = ref->cx; P.cy = ref->cy;
        P.n = n; P.ref_px = px.data(); P.world_xyz = xyz.data(); P.valid = valid.data();
        pose_to_qt(ref->mTcw, P.ref_q, P.ref_t);
        pose_to_qt(cur->mTcw, P.cur_q, P.cur_t);
        SE3fPOD T;
        const size_t tracked = Base::run(P, T);
        if (tracked > 0) TCR = qt_to_pose<SE3>(T.q, T.t);
        return tracked;
    }
};

}  // namespace b200


}  // namespace ygz
