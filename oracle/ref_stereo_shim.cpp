// ORACLE — TEST INFRASTRUCTURE ONLY.  Never part of the shipped product path.
//
// C shim over the REFERENCE's own Frame::ComputeStereoMatches.  The function lives in /root/reference/src/Frame.cc (lines 509-682), a
// translation unit that needs the whole SLAM system to compile, so it cannot be compiled as a file.  oracle/Makefile (target
// _ref/libstereo_ref.so) therefore cuts exactly that function — the reference's own text, located by its signature line and its
// closing brace, verified at build time — out of src/Frame.cc into oracle/_ref/gen/Frame_ComputeStereoMatches.inc (an intermediate of
// the build, deleted again after the compile; nothing of it is committed or shipped) and this file includes it, which compiles it as the member of the stand-in ygz::Frame of
// oracle/ref_mock/Frame.h.  ORBmatcher::DescriptorDistance / TH_HIGH / TH_LOW come from the reference's src/ORBmatcher.cc compiled
// alongside.  tests/test_ref_stereo.py compares the oracle restatement (oracle/stereo_matches.cpp) with it: mvuRight / mvDepth bit
// for bit, match counts.
#include <opencv2/core/core.hpp>
#include "Common.h"
#include "MapPoint.h"
#include "KeyFrame.h"
#include "Frame.h"
#include "ORBmatcher.h"
#include <climits>
#include <cstring>
#include <vector>

#include "ygz_oracle.h"

namespace ygz {
// the one member of the extractor the function reads (include/ORBextractor.h:106)
class ORBextractor {
public:
    std::vector<cv::Mat> mvImagePyramid;
};

float Frame::fx, Frame::fy, Frame::cx, Frame::cy;
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
std::vector<MapPoint::LogEntry> MapPoint::mLog;
int MapPoint::PredictScale(const float &currentDist, Frame *pF) {             // src/MapPoint.cc:359-373 (needed by ORBmatcher.cc)
    float ratio = mfMaxDistance / currentDist;
    int nScale = ceil(log(ratio) / pF->mfLogScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
    return nScale;
}

#include "Frame_ComputeStereoMatches.inc"        // void Frame::ComputeStereoMatches() { ... }  — the reference's text
}  // namespace ygz

extern "C" int refst_compute_stereo_matches(const orc_direct_frame *pyrL, const orc_direct_frame *pyrR, int nlevels, const orc_keypoint *keysL,
                                            const uint8_t *descL, int N, const orc_keypoint *keysR, const uint8_t *descR, int Nr,
                                            const float *scale, const float *inv_scale, float mb, float mbf, float *mvuRight, float *mvDepth) {
    using namespace ygz;
    ORBextractor L, R;
    L.mvImagePyramid.resize(nlevels); R.mvImagePyramid.resize(nlevels);
    for (int l = 0; l < nlevels; l++) {
        L.mvImagePyramid[l] = cv::Mat(pyrL->h[l], pyrL->w[l], CV_8U, (void *) pyrL->level[l], (size_t) pyrL->pitch[l]);
        R.mvImagePyramid[l] = cv::Mat(pyrR->h[l], pyrR->w[l], CV_8U, (void *) pyrR->level[l], (size_t) pyrR->pitch[l]);
    }
    Frame F;
    F.N = N;
    F.mvKeys.resize(N); F.mvKeysRight.resize(Nr);
    for (int i = 0; i < N; i++) F.mvKeys[i] = cv::KeyPoint(keysL[i].x, keysL[i].y, keysL[i].size, keysL[i].angle, keysL[i].response, keysL[i].octave, keysL[i].class_id);
    for (int i = 0; i < Nr; i++) F.mvKeysRight[i] = cv::KeyPoint(keysR[i].x, keysR[i].y, keysR[i].size, keysR[i].angle, keysR[i].response, keysR[i].octave, keysR[i].class_id);
    F.mDescriptors = cv::Mat(N > 0 ? N : 1, 32, CV_8U); F.mDescriptors.rows = N;
    F.mDescriptorsRight = cv::Mat(Nr > 0 ? Nr : 1, 32, CV_8U); F.mDescriptorsRight.rows = Nr;
    if (N) std::memcpy(F.mDescriptors.data, descL, (size_t) N * 32);
    if (Nr) std::memcpy(F.mDescriptorsRight.data, descR, (size_t) Nr * 32);
    F.mvScaleFactors.assign(scale, scale + nlevels);
    F.mvInvScaleFactors.assign(inv_scale, inv_scale + nlevels);
    F.mb = mb; F.mbf = mbf;
    F.mpORBextractorLeft = &L; F.mpORBextractorRight = &R;
    F.ComputeStereoMatches();
    int n = 0;
    for (int i = 0; i < N; i++) { mvuRight[i] = F.mvuRight[i]; mvDepth[i] = F.mvDepth[i]; n += F.mvuRight[i] >= 0; }
    return n;
}
