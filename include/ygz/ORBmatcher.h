// ygz::ORBmatcher — header-only host mirror of the reference class (include/ORBmatcher.h:38-220) over the C ABI.
// The Search* functions take flat views (ygz_frame_view / ygz_map_points) of the Frame/KeyFrame/MapPoint graph;
// INTEGRATION.md shows the few lines that gather them from the reference's objects.  Return values and the
// mutation of mvpMapPoints (as index arrays) follow the reference.
#pragma once
#include <cstdio>
#include <utility>
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

class ORBmatcher {
public:
    static const int TH_LOW = 50;          // src/ORBmatcher.cc:37
    static const int TH_HIGH = 100;        // :36
    static const int HISTO_LENGTH = 30;    // :38

    ORBmatcher(float nnratio = 0.6, bool checkOri = true, int device = 0) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {
        if (ygz_matcher_create(device, 1 << 14, 1 << 14, 1 << 20, &h_) != YGZ_OK) {
            std::fprintf(stderr, "ygz::ORBmatcher: %s\n", ygz_last_error());
            h_ = nullptr;
        }
    }
    ~ORBmatcher() { if (h_) ygz_matcher_destroy(h_); if (d_) ygz_direct_destroy(d_); }
    ORBmatcher(const ORBmatcher&) = delete;
    ORBmatcher& operator=(const ORBmatcher&) = delete;

    // src/ORBmatcher.cc:1507-1523
    static int DescriptorDistance(const uint8_t* a, const uint8_t* b) { return ygz_descriptor_distance(a, b); }

    // SearchByProjection(Frame&, const vector<MapPoint*>&, th, checkLevel)  :43-126
    int SearchByProjection(ygz_frame_view& F, const ygz_map_points& mps, const std::vector<int>& vpMapPoints, float th = 3, bool checkLevel = true) {
        return ok(ygz_search_by_projection_points(h_, mfNNratio, &F, &mps, vpMapPoints.data(), (int)vpMapPoints.size(), th, checkLevel));
    }
    // SearchByProjection(Frame& Current, const Frame& Last, th, bMono, checkLevel)  :1218-1350
    int SearchByProjection(ygz_frame_view& CurrentFrame, const ygz_frame_view& LastFrame, const ygz_map_points& mps, float th, bool bMono, bool checkLevel = true) {
        return ok(ygz_search_by_projection_last(h_, mbCheckOrientation, &CurrentFrame, &LastFrame, &mps, th, bMono, checkLevel));
    }
    // SearchByProjection(Frame& Current, KeyFrame*, const set<MapPoint*>& sAlreadyFound, th, ORBdist)  :1352-1469
    int SearchByProjection(ygz_frame_view& CurrentFrame, const ygz_frame_view& KF, const ygz_map_points& mps, const std::vector<uint8_t>& sAlreadyFound, float th, int ORBdist) {
        return ok(ygz_search_by_projection_kf(h_, mbCheckOrientation, &CurrentFrame, &KF, &mps, sAlreadyFound.empty() ? nullptr : sAlreadyFound.data(), th, ORBdist));
    }
    // SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>& vpMapPointMatches)  :155-263
    int SearchByBoW(const ygz_frame_view& KF, const ygz_map_points& mps, const ygz_frame_view& F, std::vector<int>& vpMapPointMatches) {
        vpMapPointMatches.assign(F.n > 0 ? F.n : 0, -1);
        std::vector<int> tmp((size_t)(F.n > 0 ? F.n : 1));
        int n = ok(ygz_search_by_bow(h_, mfNNratio, mbCheckOrientation, &KF, &mps, &F, tmp.data()));
        for (int i = 0; i < F.n; i++) vpMapPointMatches[i] = tmp[i];
        return n;
    }
    // SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)  :480-595
    int SearchByBoW(const ygz_frame_view& KF1, const ygz_frame_view& KF2, const ygz_map_points& mps, std::vector<int>& vpMatches12) {
        vpMatches12.assign(KF1.n > 0 ? KF1.n : 1, -1);
        int n = ok(ygz_search_by_bow_kf(h_, mfNNratio, mbCheckOrientation, &KF1, &KF2, &mps, vpMatches12.data()));
        vpMatches12.resize(KF1.n > 0 ? KF1.n : 0);
        return n;
    }
    // SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)  :597-746; epipole = KF1's centre seen from KF2
    int SearchForTriangulation(const ygz_frame_view& KF1, const ygz_frame_view& KF2, const float F12[9], float ex, float ey,
                               const std::vector<float>& mvLevelSigma2, std::vector<std::pair<size_t, size_t> >& vMatchedPairs,
                               bool bOnlyStereo) {
        std::vector<int> p((size_t)2 * (KF1.n > 0 ? KF1.n : 1));
        int n = ok(ygz_search_for_triangulation(h_, mbCheckOrientation, &KF1, &KF2, F12, ex, ey, mvLevelSigma2.data(), bOnlyStereo, p.data()));
        vMatchedPairs.clear();
        for (int i = 0; i < n; i++) vMatchedPairs.push_back(std::make_pair((size_t)p[2 * i], (size_t)p[2 * i + 1]));
        return n;
    }
    // SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)  :375-478
    int SearchForInitialization(const ygz_frame_view& F1, const ygz_frame_view& F2, std::vector<float>& vbPrevMatchedXY, std::vector<int>& vnMatches12, int windowSize = 10) {
        vnMatches12.assign(F1.n > 0 ? F1.n : 1, -1);
        int n = ok(ygz_search_for_initialization(h_, mfNNratio, mbCheckOrientation, &F1, &F2, vbPrevMatchedXY.data(), vnMatches12.data(), windowSize));
        vnMatches12.resize(F1.n > 0 ? F1.n : 0);
        return n;
    }
    // The Sim3 family of the back end (loop closing); the map-point table must carry `normal`.
    // SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th)  :265-373   (Scw 4x4 row-major)
    int SearchByProjection(const ygz_frame_view& KF, const float Scw[16], const ygz_map_points& mps, const std::vector<int>& vpPoints,
                           std::vector<int>& vpMatched, int th) {
        vpMatched.resize(KF.n > 0 ? KF.n : 1, -1);
        const int n = ok(ygz_search_by_projection_sim3(h_, &KF, Scw, &mps, vpPoints.data(), (int)vpPoints.size(), vpMatched.data(), th));
        vpMatched.resize(KF.n > 0 ? KF.n : 0);
        return n;
    }
    // Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint)  :888-1004   (KF.map_point is the key frame's slot table, updated)
    int Fuse(ygz_frame_view& KF, const float Scw[16], const ygz_map_points& mps, const std::vector<int>& vpPoints, float th,
             std::vector<int>& vpReplacePoint) {
        vpReplacePoint.resize(vpPoints.size() ? vpPoints.size() : 1, -1);
        const int n = ok(ygz_fuse_sim3(h_, &KF, Scw, &mps, vpPoints.data(), (int)vpPoints.size(), th, vpReplacePoint.data()));
        vpReplacePoint.resize(vpPoints.size());
        return n;
    }
    // The search of Fuse(KeyFrame*, const vector<MapPoint*>&, th)  :748-886: best key point / distance per candidate map point (-1 / 256
    // = none); the caller replays the observation-graph mutation (ygz::b200::ORBmatcherT::Fuse does).  false: the call failed.
    bool FuseCandidates(const ygz_frame_view& KF, const float Ow[3], const float* invLevelSigma2, const ygz_map_points& mps,
                        const std::vector<int>& vpMapPoints, float th, std::vector<int>& bestIdx, std::vector<int>& bestDist) {
        const size_t n = vpMapPoints.size();
        bestIdx.assign(n ? n : 1, -1);
        bestDist.assign(n ? n : 1, 256);
        const int rc = ygz_fuse_candidates(h_, &KF, Ow, invLevelSigma2, &mps, vpMapPoints.data(), (int)n, th, bestIdx.data(), bestDist.data());
        if (rc < 0) { ok(rc); bestIdx.assign(n, -1); bestDist.assign(n, 256); return false; }    // "nothing to fuse" + a log line
        bestIdx.resize(n); bestDist.resize(n);
        return true;
    }
    // SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th)  :1006-1216
    int SearchBySim3(const ygz_frame_view& KF1, const ygz_frame_view& KF2, const ygz_map_points& mps, std::vector<int>& vpMatches12,
                     const std::vector<int>& idxInKF2, float s12, const float R12[9], const float t12[3], float th) {
        vpMatches12.resize(KF1.n > 0 ? KF1.n : 1, -1);
        const int n = ok(ygz_search_by_sim3(h_, &KF1, &KF2, &mps, vpMatches12.data(), idxInKF2.data(), s12, R12, t12, th));
        vpMatches12.resize(KF1.n > 0 ? KF1.n : 0);
        return n;
    }
    ygz_matcher* handle() { return h_; }

    // ---- direct (patch) projection search, SURVEY.md §8(f) rank 1 -------------------------------------------------
    // FindDirectProjection(KeyFrame* ref, Frame* curr, MapPoint* mp, Vector2f& px_curr, int& search_level)  :1575-1602,
    // batched over the (map point, observing keyframe) pairs Tracking::SearchLocalPointsDirect tries for one frame
    // (src/Tracking.cc:2174-2326).  Pyramids live in device slots: BindDirect once, UploadDirectFrame per keyframe /
    // per current frame, then one FindDirectProjection call per frame.
    bool BindDirect(int max_slots, int max_items, size_t max_pyramid_bytes, int device = 0) {
        if (d_) { ygz_direct_destroy(d_); d_ = nullptr; }
        return ygz_direct_create(device, max_slots, max_items, max_pyramid_bytes, &d_) == YGZ_OK || (ok(-1), false);
    }
    bool UploadDirectFrame(int slot, const ygz_direct_frame& f, int nlevels) {
        return d_ && (ygz_direct_upload_frame(d_, slot, &f, nlevels) == YGZ_OK || (ok(-1), false));
    }
    bool SetDirectPose(int slot, const float q_xyzw[4], const float t[3]) {
        return d_ && (ygz_direct_set_pose(d_, slot, q_xyzw, t) == YGZ_OK || (ok(-1), false));
    }
    // px_curr (in: mp->mTrackProjX/Y, out: refined), search_level and the returned bool per item; returns #successes
    int FindDirectProjection(int cur_slot, const float K4[4], const std::vector<float>& mvScaleFactors,
                             const std::vector<float>& mvInvScaleFactors, float invLevelSigma2_1,
                             const std::vector<ygz_direct_item>& items, std::vector<float>& px_curr,
                             std::vector<int>& search_level, std::vector<uint8_t>& success) {
        const int n = (int)items.size();
        px_curr.assign((size_t)2 * n, 0.f); search_level.assign(n, 0); success.assign(n, 0);
        if (!d_ || n == 0) return 0;
        if (ygz_find_direct_projection(d_, cur_slot, K4, mvScaleFactors.data(), mvInvScaleFactors.data(), invLevelSigma2_1,
                                       (int)mvScaleFactors.size(), items.data(), n, px_curr.data(), search_level.data(),
                                       success.data()) != YGZ_OK)
            return ok(-1);
        int cnt = 0;
        for (uint8_t s : success) cnt += s;
        return cnt;
    }

protected:
    static int ok(int rc) {        // reference convention: failures surface as "0 matches" + a log line
        if (rc < 0) { std::fprintf(stderr, "ygz::ORBmatcher: %d %s\n", rc, ygz_last_error()); return 0; }
        return rc;
    }
    float mfNNratio;
    bool mbCheckOrientation;
    ygz_matcher* h_ = nullptr;
    ygz_direct* d_ = nullptr;
};

YGZ_MIRROR_END  // namespace
