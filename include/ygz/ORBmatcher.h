// ygz::ORBmatcher — header-only host mirror of the reference class (include/ORBmatcher.h:38-220) over the C ABI.
// The Search* functions take flat views (ygz_frame_view / ygz_map_points) of the Frame/KeyFrame/MapPoint graph;
// INTEGRATION.md shows the few lines that gather them from the reference's objects.  Return values and the
// mutation of mvpMapPoints (as index arrays) follow the reference.
#pragma once
#include <cstdio>
#include <vector>
#include "../ygz_b200.h"

namespace ygz {

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
    ~ORBmatcher() { if (h_) ygz_matcher_destroy(h_); }
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
    // SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)  :375-478
    int SearchForInitialization(const ygz_frame_view& F1, const ygz_frame_view& F2, std::vector<float>& vbPrevMatchedXY, std::vector<int>& vnMatches12, int windowSize = 10) {
        vnMatches12.assign(F1.n > 0 ? F1.n : 1, -1);
        int n = ok(ygz_search_for_initialization(h_, mfNNratio, mbCheckOrientation, &F1, &F2, vbPrevMatchedXY.data(), vnMatches12.data(), windowSize));
        vnMatches12.resize(F1.n > 0 ? F1.n : 0);
        return n;
    }
    ygz_matcher* handle() { return h_; }

protected:
    static int ok(int rc) {        // reference convention: failures surface as "0 matches" + a log line
        if (rc < 0) { std::fprintf(stderr, "ygz::ORBmatcher: %d %s\n", rc, ygz_last_error()); return 0; }
        return rc;
    }
    float mfNNratio;
    bool mbCheckOrientation;
    ygz_matcher* h_ = nullptr;
};

}  // namespace ygz
