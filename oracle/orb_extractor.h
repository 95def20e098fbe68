// ORACLE — TEST INFRASTRUCTURE ONLY (see cvprims.h).  CPU restatement of ygz::ORBextractor,
// following /root/reference/src/ORBextractor.cc line by line (citations on each method).
//
// Determinism rules the oracle fixes where the reference is non-deterministic:
//   * DistributeOctTree sorts pair<int, ExtractorNode*>; equal counts are ordered by POINTER VALUE
//     in the reference (src/ORBextractor.cc:656).  Oracle: pointer order := node creation order.
//   * x*b + y*a in computeOrbDescriptor is evaluated in float WITHOUT fma contraction
//     (this file is compiled with -ffp-contract=off); cos/sin are glibc cosf/sinf.
//   * Frame::N (uninitialised in the reference) := number of pre-existing keys passed in.
#pragma once
#include "cvprims.h"
#include <utility>
#include <vector>

namespace orc {

struct Image {
    int w = 0, h = 0;
    std::vector<uint8_t> px;       // continuous, stride == w
    const uint8_t* row(int y) const { return px.data() + (size_t)y * w; }
};

class Extractor {
public:
    // src/ORBextractor.cc:412-470
    Extractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);

    // src/ORBextractor.cc:1129-1150 (levels stored WITHOUT the 19-px border; the border is never
    // read on this path; border_reflect101 exists in cvprims for completeness)
    void ComputePyramid(const uint8_t* img, int w, int h, int stride);

    // src/ORBextractor.cc:970-1028  operator()(InputArray image, mask, keypoints, descriptors)
    void extract(const uint8_t* img, int w, int h, int stride,
                 std::vector<KeyPoint>& keypoints, std::vector<uint8_t>& descriptors);

    // src/ORBextractor.cc:1031-1127  operator()(Frame*, keypoints, descriptors, ORBSLAM_KEYPOINT, leftEye)
    // `keypoints` holds the frame's pre-existing (direct-tracked) keys on entry (level-0 coords,
    // octave/angle set); their descriptors are recomputed first, new keys are appended.
    // The pyramid must already have been built with ComputePyramid (Frame ctor does that).
    void extract_frame(std::vector<KeyPoint>& keypoints, std::vector<uint8_t>& descriptors);

    // The in-tree part of operator() only — DistributeOctTree, IC_Angle, computeOrbDescriptor, scaling (:782-803, 1012-1026) — on
    // a pyramid, blurred pyramid and per-level FAST candidates (relative to (16,16)) produced elsewhere: bench.py's OpenCV-path
    // CPU baseline feeds it with cv2's own resize / FAST / GaussianBlur outputs.  `pyramid` / `blurred` must be filled first.
    void extract_tail(const std::vector<std::vector<KeyPoint>>& cands, std::vector<KeyPoint>& keypoints,
                      std::vector<uint8_t>& descriptors);

    // stage outputs kept for parity tests
    std::vector<Image> pyramid;                        // raw levels
    std::vector<Image> blurred;                        // 7x7 sigma-2 blurred levels
    std::vector<std::vector<KeyPoint>> candidates;     // per level, pre-octree, coords rel. (16,16)
    std::vector<std::vector<KeyPoint>> level_keys;     // per level, after octree+orientation (level coords)

    int nfeatures, nlevels, iniThFAST, minThFAST;
    double scaleFactor;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<int> umax;

    // src/ORBextractor.cc:725-804
    void ComputeKeyPointsOctTree(std::vector<std::vector<KeyPoint>>& all);
    // src/ORBextractor.cc:533-723
    std::vector<KeyPoint> DistributeOctTree(const std::vector<KeyPoint>& keys, int minX, int maxX,
                                            int minY, int maxY, int N);
    // src/ORBextractor.cc:77-101
    float IC_Angle(const Image& im, float px, float py) const;
    // src/ORBextractor.cc:105-149
    void computeOrbDescriptor(const KeyPoint& kp, const Image& im, uint8_t* desc) const;

    // ---- YGZ "DSO" key points, oracle/dso_detect.cpp
    int mnGridSize = -1;                                // adaptive cell size, carried from frame to frame (:1295-1381)
    // src/ORBextractor.cc:1152-1186
    float ShiTomasiScore(const Image& img, int u, int v) const;
    // src/ORBextractor.cc:1277-1381 (level-0 key points, size 7; also refreshes the angles of exist_kps)
    void ComputeKeyPointsDSOSingleLevel(std::vector<KeyPoint>& allKeypoints, std::vector<KeyPoint>& exist_kps);
    // src/ORBextractor.cc:1031-1127 with DSO_KEYPOINT
    void extract_frame_dso(std::vector<KeyPoint>& keypoints, std::vector<uint8_t>& descriptors);
    void blur_all_public() { blur_all(); }

private:
    void blur_all();
};

// FAST-10 of Thirdparty/fast as the reference calls it on a w x h window (plain below 22 px width, SSE2 region otherwise)
bool fast10_is_corner(const uint8_t* p, int stride, int barrier);
void fast10_detect_cell(const uint8_t* img, int w, int h, int stride, int barrier, std::vector<std::pair<int, int>>& out);

extern const int ORB_PATTERN[256 * 4];

}  // namespace orc
