// ORACLE — TEST INFRASTRUCTURE ONLY.  Never part of the shipped product path.
//
// C shim over the REFERENCE's own ygz::SparseImgAlign: /root/reference/src/SparseImageAlign.cc (with include/NLSSolver.h,
// NLSSolver_impl.hpp, RobustCost.h, SparseImageAlign.h) compiled unmodified, from where it lies, against oracle/ref_mock
// (oracle/Makefile target _ref/libsia_ref.so, -DMOCK_SE3_QUATERNION).  What this pins: everything IN the reference tree —
// run()'s level loop, precomputeReferencePatches (visibility, bilinear weights, Jacobian cache layout and scaling),
// computeResiduals (projection, border tests, residual / H / Jres accumulation order), the Gauss-Newton driver of
// NLSSolver_impl.hpp (rollback on a 1.2x chi2 increase, eps stop, sticky stop_).  What it cannot pin: Eigen's LDLT and Sophus'
// SE3 arithmetic are absent third-party code; the stand-ins forward to the oracle's restatement of their published algorithms
// (ref_mock/mock_eigen.hpp), so those two pieces are common to both sides of the comparison.
#include <opencv2/core/core.hpp>
#include "SparseImageAlign.h"
#include "MapPoint.h"
#include "KeyFrame.h"
#include <cstring>
#include <vector>

#include "ygz_oracle.h"

namespace ygz {
float Frame::fx, Frame::fy, Frame::cx, Frame::cy;
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
std::vector<MapPoint::LogEntry> MapPoint::mLog;
}  // namespace ygz

extern "C" int refs_sia_run(const orc_sia_problem *p, int max_level, int min_level, int n_iter, float *out_q, float *out_t, float *H_out,
                            float *chi2_out) {
    using namespace ygz;
    Frame::fx = p->fx; Frame::fy = p->fy; Frame::cx = p->cx; Frame::cy = p->cy;
    Frame R, Cf;
    std::vector<MapPoint> pts(p->n > 0 ? p->n : 0);
    R.N = p->n;
    R.mvKeys.resize(p->n > 0 ? p->n : 0);
    R.mvpMapPoints.assign(R.mvKeys.size(), nullptr);
    R.mvbOutlier.assign(R.mvKeys.size(), false);
    for (int i = 0; i < p->n; i++) {
        R.mvKeys[i] = cv::KeyPoint(p->ref_px[2 * i], p->ref_px[2 * i + 1], 31.f);
        pts[i].mWorldPos = Vector3f(p->world_xyz[3 * i], p->world_xyz[3 * i + 1], p->world_xyz[3 * i + 2]);
        if (p->valid[i]) R.mvpMapPoints[i] = &pts[i];
    }
    R.mvInvScaleFactors.assign(p->inv_scale_factors, p->inv_scale_factors + p->n_levels);
    R.mvImagePyramid.resize(p->n_levels);
    Cf.mvImagePyramid.resize(p->n_levels);
    for (int l = 0; l < p->n_levels; l++) {
        R.mvImagePyramid[l] = cv::Mat(p->level_h[l], p->level_w[l], CV_8U, (void *) p->ref_levels[l]);
        Cf.mvImagePyramid[l] = cv::Mat(p->level_h[l], p->level_w[l], CV_8U, (void *) p->cur_levels[l]);
    }
    R.mTcw = SE3f::fromQT(p->ref_q, p->ref_t);
    Cf.mTcw = SE3f::fromQT(p->cur_q, p->cur_t);
    SparseImgAlign align(max_level, min_level, n_iter);          // SparseImgAlign(nLevels - 1, 1), src/Tracking.cc:207
    SE3f TCR = Cf.mTcw * R.mTcw.inverse();                       // callers read TCR only when run() succeeded
    const size_t n = align.run(&R, &Cf, TCR);
    std::memcpy(out_q, TCR.q, 16);
    std::memcpy(out_t, TCR.t, 12);
    if (H_out) {
        const Eigen::Matrix<float, 6, 6> &H = align.getInformationMatrix();
        for (int i = 0; i < 36; i++) H_out[i] = H.d[i];
    }
    if (chi2_out) *chi2_out = align.getChi2();
    return (int) n;
}
