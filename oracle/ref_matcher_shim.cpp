// ORACLE — TEST INFRASTRUCTURE ONLY.  Never part of the shipped product path.
//
// C shim over the REFERENCE's own ygz::ORBmatcher: /root/reference/src/ORBmatcher.cc (and src/Align.cc) are compiled
// unmodified, from where they lie, against the mock OpenCV / Eigen / Sophus / Frame / KeyFrame / MapPoint headers of
// oracle/ref_mock (oracle/Makefile target _ref/libmatcher_ref.so; no reference source is copied into this repo).
// Every entry point has the signature of the oracle's orc_* function of the same name (ygz_oracle.h) and rebuilds the
// pointer graph (Frame, KeyFrame, MapPoint objects) the reference method walks from the same flat views, so
// tests/test_ref_matcher.py can run one scene through both and compare index arrays and counts: that is what pins the
// greedy order, the ratio / threshold tests, the rotation histogram and the level bands of every Search* variant to
// the reference's own code instead of to a restatement.
//
// What stays mocked (stated, not hidden): Frame::GetFeaturesInArea / AssignFeaturesToGrid / PosInGrid live in
// src/Frame.cc, which needs the whole SLAM system to compile — ref_mock/Frame.h restates them (48 lines); Eigen and
// Sophus are absent third-party dependencies — ref_mock/mock_eigen.hpp evaluates fixed-size products left to right.
//
// Second use (-DSHIM_USE_ADAPTERS, oracle/Makefile target _ref/libadapters_e2e.so): the SAME caller code below is compiled
// against include/ygz/adapters.h instead of the reference's ORBmatcher.h — ygz::ORBmatcher is then the product's
// literal-signature adapter over libygz_b200.so and the entry points are named adp_* instead of refm_*.  That the calls
// `m.SearchByBoW(K, F, vp)` ... compile unchanged against both is the drop-in claim of INTEGRATION.md;
// tests/test_adapters_gpu.py runs both libraries on the same scenes on a B200.
#ifdef SHIM_USE_ADAPTERS
#include <opencv2/core/core.hpp>
#include "Common.h"
#define YGZ_B200_REFERENCE_NAMES   // adapters first, as the replaced include/ORBextractor.h would be (Frame.h includes it upstream)
#include <ygz/adapters.h>
#include "MapPoint.h"
#include "KeyFrame.h"
#include "Frame.h"
#define REFM(name) adp_##name
#else
#include <opencv2/core/core.hpp>  // (and the standard headers it pulls) before the access macros below
#include "Common.h"
#define protected public          // ComputeThreeMaxima, RadiusByViewingCos and CheckDistEpipolarLine are protected upstream
#define private public
#include "ORBmatcher.h"
#undef protected
#undef private
#define REFM(name) refm_##name
#endif
#include <cstring>
#include <map>
#include <set>
#include <vector>

#include "ygz_oracle.h"           // the flat view structs only (orc_frame_view, orc_map_points, orc_direct_*)

namespace ygz {
// statics of the mocked classes (include/Frame.h:231-236, 292-293, 325-328 upstream)
float Frame::fx, Frame::fy, Frame::cx, Frame::cy;
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
std::vector<MapPoint::LogEntry> MapPoint::mLog;

// src/MapPoint.cc:359-373
int MapPoint::PredictScale(const float &currentDist, Frame *pF) {
    float ratio = mfMaxDistance / currentDist;
    int nScale = ceil(log(ratio) / pF->mfLogScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
    return nScale;
}
}  // namespace ygz

namespace {
using namespace ygz;

cv::KeyPoint to_cv(const orc_keypoint &k) {
    cv::KeyPoint kp(k.x, k.y, k.size, k.angle, k.response, k.octave, k.class_id);
    return kp;
}

SE3f pose_of(const orc_frame_view *v) {
    Matrix3f R;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R(i, j) = v->Rcw[3 * i + j];
    return SE3f(R, Vector3f(v->tcw[0], v->tcw[1], v->tcw[2]));
}

void fill_featvec(DBoW2::FeatureVector &fv, const orc_frame_view *v) {
    for (int n = 0; n < v->fv_nodes; n++) {
        std::vector<unsigned int> &dst = fv[v->fv_node_id[n]];
        for (int j = v->fv_offset[n]; j < v->fv_offset[n + 1]; j++) dst.push_back(v->fv_index[j]);
    }
}

cv::Mat desc_mat(const uint8_t *d, int n) {
    cv::Mat m(n > 0 ? n : 1, 32, CV_8U);
    if (n > 0) std::memcpy(m.data, d, (size_t) n * 32);
    m.rows = n;
    return m;
}

// the map-point table: MapPoint objects with the accessors' backing fields filled from the flat arrays
struct Points {
    std::vector<MapPoint> mp;
    explicit Points(const orc_map_points *p) : mp(p ? p->m : 0) {
        for (size_t i = 0; i < mp.size(); i++) {
            MapPoint &m = mp[i];
            if (p->world_pos) m.mWorldPos = Vector3f(p->world_pos[3 * i], p->world_pos[3 * i + 1], p->world_pos[3 * i + 2]);
            if (p->normal) m.mNormalVector = Vector3f(p->normal[3 * i], p->normal[3 * i + 1], p->normal[3 * i + 2]);
            m.mDescriptor = cv::Mat(1, 32, CV_8U);
            if (p->desc) std::memcpy(m.mDescriptor.data, p->desc + 32 * i, 32);
            m.mbBad = p->bad ? p->bad[i] != 0 : false;
            m.nObs = p->observations ? p->observations[i] : 1;
            m.mfMinDistance = p->min_dist ? p->min_dist[i] : 0.f;
            m.mfMaxDistance = p->max_dist ? p->max_dist[i] : 0.f;
            m.mbTrackInView = p->track_in_view ? p->track_in_view[i] != 0 : false;
            m.mTrackProjX = p->track_proj_x ? p->track_proj_x[i] : 0.f;
            m.mTrackProjY = p->track_proj_y ? p->track_proj_y[i] : 0.f;
            m.mTrackProjXR = p->track_proj_xr ? p->track_proj_xr[i] : 0.f;
            m.mnTrackScaleLevel = p->track_scale_level ? p->track_scale_level[i] : 0;
            m.mTrackViewCos = p->track_view_cos ? p->track_view_cos[i] : 0.f;
        }
    }
    explicit Points(int m) : mp(m) {}             // anonymous points: only "is there a map point" matters to the caller
    MapPoint *at(int i) { return (i >= 0 && i < (int) mp.size()) ? &mp[i] : nullptr; }
    int index(const MapPoint *p) const { return p ? (int) (p - mp.data()) : -1; }
};

// Frame from a view.  The statics (intrinsics, image bounds, grid pitch) are process-wide upstream too: the caller sets
// them from the CURRENT frame's view last, as Tracking does (all frames of a run share them).
void set_statics(const orc_frame_view *v) {
    Frame::fx = v->fx; Frame::fy = v->fy; Frame::cx = v->cx; Frame::cy = v->cy;
    Frame::mnMinX = v->min_x; Frame::mnMinY = v->min_y; Frame::mnMaxX = v->max_x; Frame::mnMaxY = v->max_y;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);      // src/Frame.cc:151-152
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);
}

void fill_frame(Frame &F, const orc_frame_view *v, Points &P) {
    set_statics(v);
    F.N = v->n;
    F.mbf = v->mbf; F.mb = v->mb;
    F.mvKeys.resize(v->n);
    for (int i = 0; i < v->n; i++) F.mvKeys[i] = to_cv(v->keys[i]);
    F.mDescriptors = desc_mat(v->desc, v->n);
    F.mvuRight.assign(v->n, -1.f);
    if (v->u_right) F.mvuRight.assign(v->u_right, v->u_right + v->n);
    F.mvbOutlier.assign(v->n, false);
    if (v->outlier) for (int i = 0; i < v->n; i++) F.mvbOutlier[i] = v->outlier[i] != 0;
    F.mvpMapPoints.assign(v->n, nullptr);
    if (v->map_point) for (int i = 0; i < v->n; i++) F.mvpMapPoints[i] = P.at(v->map_point[i]);
    F.mnScaleLevels = v->n_levels;
    F.mfLogScaleFactor = v->log_scale_factor;
    if (v->scale_factors) F.mvScaleFactors.assign(v->scale_factors, v->scale_factors + v->n_levels);
    if (v->inv_scale_factors) F.mvInvScaleFactors.assign(v->inv_scale_factors, v->inv_scale_factors + v->n_levels);
    F.mTcw = pose_of(v);
    fill_featvec(F.mFeatVec, v);
    F.AssignFeaturesToGrid();
}

KeyFrame *make_keyframe(const orc_frame_view *v, Points &P) {
    KeyFrame *K = new KeyFrame(v->n, v->fx, v->fy, v->cx, v->cy, v->mbf, v->mb, (int) v->min_x, (int) v->min_y, (int) v->max_x,
                               (int) v->max_y, v->n_levels, v->log_scale_factor);
    K->mvKeys.resize(v->n);
    for (int i = 0; i < v->n; i++) K->mvKeys[i] = to_cv(v->keys[i]);
    K->mDescriptors = desc_mat(v->desc, v->n);
    K->mvuRight.assign(v->n, -1.f);
    if (v->u_right) K->mvuRight.assign(v->u_right, v->u_right + v->n);
    K->mvpMapPoints.assign(v->n, nullptr);
    if (v->map_point) for (int i = 0; i < v->n; i++) K->mvpMapPoints[i] = P.at(v->map_point[i]);
    if (v->scale_factors) K->mvScaleFactors.assign(v->scale_factors, v->scale_factors + v->n_levels);
    if (v->inv_scale_factors) K->mvInvScaleFactors.assign(v->inv_scale_factors, v->inv_scale_factors + v->n_levels);
    K->Tcw = pose_of(v);
    {   // src/KeyFrame.cc:100-119 SetPose: Ow = -Rcw^T tcw
        Matrix3f Rwc = K->Tcw.rotationMatrix().transpose();
        K->Ow = -(Rwc * K->Tcw.translation());
    }
    fill_featvec(K->mFeatVec, v);
    // the KeyFrame copies the Frame's grid (src/KeyFrame.cc:51-56)
    K->mGrid.assign(64, std::vector<std::vector<size_t> >(48));
    for (int i = 0; i < v->n; i++) {
        int px = round((K->mvKeys[i].pt.x - K->mnMinX) * K->mfGridElementWidthInv);
        int py = round((K->mvKeys[i].pt.y - K->mnMinY) * K->mfGridElementHeightInv);
        if (px < 0 || px >= 64 || py < 0 || py >= 48) continue;
        K->mGrid[px][py].push_back(i);
    }
    return K;
}

void store_map_points(const std::vector<MapPoint *> &v, const Points &P, int *out) {
    for (size_t i = 0; i < v.size(); i++) out[i] = P.index(v[i]);
}
}  // namespace

extern "C" {

int REFM(hamming256)(const uint8_t *a, const uint8_t *b) {
    cv::Mat ma(1, 32, CV_8U, (void *) a), mb(1, 32, CV_8U, (void *) b);
    return ORBmatcher::DescriptorDistance(ma, mb);                              // src/ORBmatcher.cc:1507
}

#ifndef SHIM_USE_ADAPTERS
void REFM(three_maxima)(const int *hist_sizes, int L, int *ind1, int *ind2, int *ind3) {
    std::vector<std::vector<int> > histo(L);
    for (int i = 0; i < L; i++) histo[i].assign(hist_sizes[i], 0);
    ORBmatcher m;
    m.ComputeThreeMaxima(histo.data(), L, *ind1, *ind2, *ind3);                 // :1471
}
#endif

int REFM(search_by_bow)(float nnratio, int check_ori, const orc_frame_view *kf, const orc_map_points *mps,
                       const orc_frame_view *f, int *out_matches) {
    Points P(mps);
    KeyFrame *K = make_keyframe(kf, P);
    Frame F;
    fill_frame(F, f, P);
    ORBmatcher m(nnratio, check_ori != 0);
    std::vector<MapPoint *> vp;
    int n = m.SearchByBoW(K, F, vp);                                            // :155
    store_map_points(vp, P, out_matches);
    delete K;
    return n;
}

int REFM(search_by_projection_points)(float nnratio, orc_frame_view *f, const orc_map_points *mps, const int *query_ids, int nq,
                                     float th, int check_level) {
    Points P(mps);
    Frame F;
    fill_frame(F, f, P);
    std::vector<MapPoint *> q(nq);
    for (int i = 0; i < nq; i++) q[i] = P.at(query_ids[i]);
    ORBmatcher m(nnratio);
    int n = m.SearchByProjection(F, q, th, check_level != 0);                   // :43
    store_map_points(F.mvpMapPoints, P, f->map_point);
    return n;
}

int REFM(search_by_projection_last)(int check_ori, orc_frame_view *cur, const orc_frame_view *last, const orc_map_points *mps,
                                   float th, int mono, int check_level) {
    Points P(mps);
    Frame L, Cf;
    fill_frame(L, last, P);
    fill_frame(Cf, cur, P);                                                     // statics from the current frame
    ORBmatcher m(0.9f, check_ori != 0);
    int n = m.SearchByProjection(Cf, L, th, mono != 0, check_level != 0);       // :1218
    store_map_points(Cf.mvpMapPoints, P, cur->map_point);
    return n;
}

int REFM(search_by_projection_kf)(int check_ori, orc_frame_view *cur, const orc_frame_view *kf, const orc_map_points *mps,
                                 const uint8_t *already_found, float th, int orb_dist) {
    Points P(mps);
    KeyFrame *K = make_keyframe(kf, P);
    Frame Cf;
    fill_frame(Cf, cur, P);
    std::set<MapPoint *> found;
    if (already_found)
        for (size_t i = 0; i < P.mp.size(); i++) if (already_found[i]) found.insert(&P.mp[i]);
    ORBmatcher m(0.9f, check_ori != 0);
    int n = m.SearchByProjection(Cf, K, found, th, orb_dist);                   // :1352
    store_map_points(Cf.mvpMapPoints, P, cur->map_point);
    delete K;
    return n;
}

int REFM(search_for_initialization)(float nnratio, int check_ori, const orc_frame_view *f1, const orc_frame_view *f2,
                                   float *prev_matched_xy, int *matches12, int window) {
    Points P(nullptr);
    Frame F1, F2;
    fill_frame(F1, f1, P);
    fill_frame(F2, f2, P);
    std::vector<cv::Point2f> prev(f1->n);
    for (int i = 0; i < f1->n; i++) prev[i] = cv::Point2f(prev_matched_xy[2 * i], prev_matched_xy[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher m(nnratio, check_ori != 0);
    int n = m.SearchForInitialization(F1, F2, prev, m12, window);               // :375
    for (int i = 0; i < f1->n; i++) { matches12[i] = m12[i]; prev_matched_xy[2 * i] = prev[i].x; prev_matched_xy[2 * i + 1] = prev[i].y; }
    return n;
}

int REFM(search_by_bow_kf)(float nnratio, int check_ori, const orc_frame_view *kf1, const orc_frame_view *kf2,
                          const orc_map_points *mps, int *matches12) {
    Points P(mps);
    KeyFrame *K1 = make_keyframe(kf1, P), *K2 = make_keyframe(kf2, P);
    ORBmatcher m(nnratio, check_ori != 0);
    std::vector<MapPoint *> vp;
    int n = m.SearchByBoW(K1, K2, vp);                                          // :480
    store_map_points(vp, P, matches12);
    delete K1; delete K2;
    return n;
}

/* (ex, ey) of the oracle's signature are OUTPUTS here: the reference computes the epipole itself from the two poses
 * (:603-609); the test feeds them to the oracle. */
int REFM(search_for_triangulation)(int check_ori, const orc_frame_view *kf1, const orc_frame_view *kf2, const float *F12,
                                  float *ex, float *ey, const float *level_sigma2, int only_stereo, int *pairs) {
    int npts = 0;
    for (int i = 0; i < kf1->n; i++) if (kf1->map_point && kf1->map_point[i] >= npts) npts = kf1->map_point[i] + 1;
    for (int i = 0; i < kf2->n; i++) if (kf2->map_point && kf2->map_point[i] >= npts) npts = kf2->map_point[i] + 1;
    Points P(npts);
    KeyFrame *K1 = make_keyframe(kf1, P), *K2 = make_keyframe(kf2, P);
    K2->mvLevelSigma2.assign(level_sigma2, level_sigma2 + kf2->n_levels);
    Matrix3f F;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) F(i, j) = F12[3 * i + j];
    {   // the epipole exactly as :603-609 evaluates it (reported for the oracle's input)
        Vector3f C2 = K2->GetRotation() * K1->GetCameraCenter() + K2->GetTranslation();
        const float invz = 1.0f / C2[2];
        *ex = K2->fx * C2[0] * invz + K2->cx;
        *ey = K2->fy * C2[1] * invz + K2->cy;
    }
    ORBmatcher m(0.6f, check_ori != 0);
    std::vector<std::pair<size_t, size_t> > vp;
    int n = m.SearchForTriangulation(K1, K2, F, vp, only_stereo != 0);          // :597
    for (size_t i = 0; i < vp.size(); i++) { pairs[2 * i] = (int) vp[i].first; pairs[2 * i + 1] = (int) vp[i].second; }
    delete K1; delete K2;
    return n;
}

}  // extern "C"

#ifndef SHIM_USE_ADAPTERS      // the batched device form of FindDirectProjection has its own tests (tests/test_direct_gpu.py)
namespace {
// Unit quaternion (x, y, z, w) -> rotation matrix with Eigen's QuaternionBase::toRotationMatrix arithmetic
// (tx = 2x, twx = tx*w, ...; m(0,0) = 1 - (tyy + tzz), ...), which is what Sophus' SE3::rotationMatrix() returns.
SE3f pose_q(const float *q, const float *t) {
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const float twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    Matrix3f R;
    R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
    R(1, 0) = txy + twz; R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = 1 - (txx + tyy);
    return SE3f(R, Vector3f(t[0], t[1], t[2]));
}
void fill_pyramid(std::vector<cv::Mat> &pyr, const orc_direct_frame *f, int nlevels) {
    pyr.resize(nlevels);
    for (int l = 0; l < nlevels; l++) pyr[l] = cv::Mat(f->h[l], f->w[l], CV_8U, (void *) f->level[l], (size_t) f->pitch[l]);
}
}  // namespace

/* ORBmatcher::FindDirectProjection :1573-1602 (GetWarpAffineMatrix :1525, WarpAffine :1549, GetBestSearchLevel
 * include/ORBmatcher.h:185-197, ygz::Align2D src/Align.cc:8-103) per item; same arguments as orc_find_direct_projection.
 * The mock SE3 holds a rotation MATRIX where Sophus holds a quaternion, so poses are converted once and the float results can
 * differ from a Sophus build in the last bits: the test compares levels / success flags exactly and pixels to 1e-3. */
extern "C" int REFM(find_direct_projection)(const orc_direct_frame *refs, int n_refs, const orc_direct_frame *cur, const float *K4,
                                           const float *scale, const float *inv_scale, float inv_level_sigma2_1, int nlevels,
                                           const orc_direct_item *items, int n_items, float *out_px, int *out_level, uint8_t *out_ok) {
    Frame::fx = K4[0]; Frame::fy = K4[1]; Frame::cx = K4[2]; Frame::cy = K4[3];
    Frame C;
    C.mTcw = pose_q(cur->q, cur->t);
    fill_pyramid(C.mvImagePyramid, cur, nlevels);
    C.mvScaleFactors.assign(scale, scale + nlevels);
    C.mvInvScaleFactors.assign(inv_scale, inv_scale + nlevels);
    std::vector<KeyFrame *> kfs(n_refs);
    for (int r = 0; r < n_refs; r++) {
        KeyFrame *Kf = new KeyFrame(1, K4[0], K4[1], K4[2], K4[3], 0.f, 0.f, 0, 0, refs[r].w[0], refs[r].h[0], nlevels, logf(scale[1]));
        Kf->Tcw = pose_q(refs[r].q, refs[r].t);
        fill_pyramid(Kf->mvImagePyramid, &refs[r], nlevels);
        Kf->mvScaleFactors.assign(scale, scale + nlevels);
        Kf->mvInvScaleFactors.assign(inv_scale, inv_scale + nlevels);
        Kf->mvInvLevelSigma2.assign(nlevels, 1.f);
        if (nlevels > 1) Kf->mvInvLevelSigma2[1] = inv_level_sigma2_1;
        Kf->mvKeys.resize(1);
        kfs[r] = Kf;
    }
    ORBmatcher m;
    for (int i = 0; i < n_items; i++) {
        const orc_direct_item &it = items[i];
        KeyFrame *Kf = kfs[it.ref];
        Kf->mvKeys[0] = cv::KeyPoint(it.px_ref[0], it.px_ref[1], 31.f, -1.f, 0.f, it.octave, -1);
        MapPoint mp;
        mp.mWorldPos = Vector3f(it.pw[0], it.pw[1], it.pw[2]);
        mp.mObservations[Kf] = 0;
        Vector2f px(it.px_cur[0], it.px_cur[1]);
        int level = -1;
        bool ok = m.FindDirectProjection(Kf, &C, &mp, px, level);
        out_px[2 * i] = px[0]; out_px[2 * i + 1] = px[1];
        out_level[i] = level; out_ok[i] = ok ? 1 : 0;
    }
    for (KeyFrame *k : kfs) delete k;
    return 0;
}
#endif

#ifdef SHIM_USE_ADAPTERS
// ---- the extractor and aligner adapters, driven the way Frame.cc / Tracking.cc drive the reference classes ---------------
namespace {
SE3f pose_from_q(const float *q, const float *t) { return ygz::b200::qt_to_pose<SE3f>(q, t); }
}

extern "C" {

/* ORBextractor::operator()(InputArray image, InputArray mask, vector<KeyPoint>&, OutputArray), called as Frame's mono constructor does */
int adp_extract_image(int nfeatures, float scale, int nlevels, int ini_th, int min_th, const uint8_t *img, int w, int h, int stride,
                      orc_keypoint *kps, uint8_t *desc, int cap) {
    ygz::ORBextractor ext(nfeatures, scale, nlevels, ini_th, min_th);
    cv::Mat im(h, w, CV_8U, (void *) img, (size_t) stride), descriptors;
    std::vector<cv::KeyPoint> keys;
    ext(im, cv::Mat(), keys, descriptors);
    const int n = (int) keys.size();
    if (n > cap) return -n;
    for (int i = 0; i < n; i++) {
        std::memcpy(&kps[i], &keys[i], sizeof(orc_keypoint));
        std::memcpy(desc + 32 * (size_t) i, descriptors.data + (size_t) i * descriptors.step, 32);
    }
    return n;
}

/* Frame::ExtractORB (src/Frame.cc:332-348): ComputePyramid into frame->mvImagePyramid, then operator()(Frame*, mvKeys, mDescriptors, method);
 * method 0 = ORBSLAM_KEYPOINT, 2 = DSO_KEYPOINT.  kps holds the n_existing direct-tracked keys on entry.  Also returns level
 * `probe_level` of the extractor's public mvImagePyramid (probe may be NULL). */
int adp_extract_frame(int nfeatures, float scale, int nlevels, int ini_th, int min_th, const uint8_t *img, int w, int h, int stride,
                      int method, int n_existing, orc_keypoint *kps, uint8_t *desc, int cap, int probe_level, uint8_t *probe) {
    ygz::ORBextractor ext(nfeatures, scale, nlevels, ini_th, min_th);
    cv::Mat im(h, w, CV_8U, (void *) img, (size_t) stride);
    ext.ComputePyramid(im);                                                   // src/Frame.cc:807-813
    Frame F;
    F.mvImagePyramid.resize(nlevels);
    for (int l = 0; l < nlevels; l++) F.mvImagePyramid[l] = ext.mvImagePyramid[l].clone();
    if (probe) {
        const cv::Mat &L = ext.mvImagePyramid[probe_level];
        for (int y = 0; y < L.rows; y++) std::memcpy(probe + (size_t) y * L.cols, L.data + (size_t) y * L.step, L.cols);
    }
    F.mvKeys.resize(n_existing);
    for (int i = 0; i < n_existing; i++) F.mvKeys[i] = to_cv(kps[i]);
    F.N = n_existing;
    ext(&F, F.mvKeys, F.mDescriptors, (ygz::ORBextractor::KeyPointMethod) method, true);
    const int n = (int) F.mvKeys.size();
    if (n > cap) return -n;
    for (int i = 0; i < n; i++) {
        std::memcpy(&kps[i], &F.mvKeys[i], sizeof(orc_keypoint));
        std::memcpy(desc + 32 * (size_t) i, F.mDescriptors.data + (size_t) i * F.mDescriptors.step, 32);
    }
    return n;
}

/* SparseImgAlign::run(Frame* ref, Frame* cur, SE3f& TCR) as Tracking::TrackWithSparseAlignment calls it (src/Tracking.cc:2087) */
int adp_sia_run(const orc_sia_problem *p, int max_level, int min_level, int n_iter, float *out_q, float *out_t) {
    Frame::fx = p->fx; Frame::fy = p->fy; Frame::cx = p->cx; Frame::cy = p->cy;
    Frame R, Cf;
    std::vector<MapPoint> pts(p->n);
    R.N = p->n;
    R.mvKeys.resize(p->n);
    R.mvpMapPoints.assign(p->n, nullptr);
    R.mvbOutlier.assign(p->n, false);
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
    R.mTcw = pose_from_q(p->ref_q, p->ref_t);
    Cf.mTcw = pose_from_q(p->cur_q, p->cur_t);
    ygz::SparseImgAlign align(max_level, min_level, n_iter);
    SE3f TCR;
    const size_t n = align.run(&R, &Cf, TCR);
    ygz::b200::pose_to_qt(TCR, out_q, out_t);
    return (int) n;
}

}  // extern "C"
#endif

// ---- the Sim3 family of the back end (cv::Mat geometry): SearchByProjection(KF, Scw, ...) :265, Fuse(KF, Scw, ...) :888, SearchBySim3 :1006
namespace {
cv::Mat mat_from(const float *v, int r, int c) {
    cv::Mat m(r, c, CV_32F);
    for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) m.at<float>(i, j) = v[i * c + j];
    return m;
}
}  // namespace

extern "C" {

int REFM(search_by_projection_sim3)(const orc_frame_view *kf, const float *Scw, const orc_map_points *mps, const int *point_ids, int n_points,
                                   int *matched, int th) {
    Points P(mps);
    KeyFrame *K = make_keyframe(kf, P);
    std::vector<MapPoint *> pts(n_points), vpMatched(kf->n, nullptr);
    for (int i = 0; i < n_points; i++) pts[i] = P.at(point_ids[i]);
    for (int i = 0; i < kf->n; i++) vpMatched[i] = P.at(matched[i]);
    ORBmatcher m(0.75f, true);
    const int n = m.SearchByProjection(K, mat_from(Scw, 4, 4), pts, vpMatched, th);
    store_map_points(vpMatched, P, matched);
    delete K;
    return n;
}

int REFM(fuse_sim3)(orc_frame_view *kf, const float *Scw, const orc_map_points *mps, const int *point_ids, int n_points, float th,
                   int *replace_point) {
    Points P(mps);
    KeyFrame *K = make_keyframe(kf, P);
    std::vector<MapPoint *> pts(n_points), rep(n_points, nullptr);
    for (int i = 0; i < n_points; i++) { pts[i] = P.at(point_ids[i]); rep[i] = P.at(replace_point[i]); }
    ORBmatcher m(0.8f);
    const int n = m.Fuse(K, mat_from(Scw, 4, 4), pts, th, rep);
    store_map_points(rep, P, replace_point);
    store_map_points(K->mvpMapPoints, P, kf->map_point);
    delete K;
    return n;
}

int REFM(search_by_sim3)(const orc_frame_view *kf1, const orc_frame_view *kf2, const orc_map_points *mps, int *matches12, const int *idx_in_kf2,
                        float s12, const float *R12, const float *t12, float th) {
    Points P(mps);
    KeyFrame *K1 = make_keyframe(kf1, P), *K2 = make_keyframe(kf2, P);
    for (size_t i = 0; i < P.mp.size(); i++)
        if (idx_in_kf2[i] >= 0) P.mp[i].mObservations[K2] = (size_t) idx_in_kf2[i];
    std::vector<MapPoint *> vp(kf1->n, nullptr);
    for (int i = 0; i < kf1->n; i++) vp[i] = P.at(matches12[i]);
    ORBmatcher m(0.75f, true);
    const int n = m.SearchBySim3(K1, K2, vp, s12, mat_from(R12, 3, 3), mat_from(t12, 3, 1), th);
    store_map_points(vp, P, matches12);
    delete K1; delete K2;
    return n;
}

/* ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) :748-886 on a two-key-frame map: `kf` (the target) and a second, mono key frame
 * of n2 slots that only holds extra observations (obs2_idx[j] = slot of point j there, -1 = none), so that Observations() differs
 * between points and MapPoint::Replace has something to move.  Observations are registered with AddObservation (stereo keys of kf
 * count twice, src/MapPoint.cc:78-91).  point_ids[i] = -1 is a NULL entry of vpMapPoints.
 * out: kf->map_point (slots after the call), bad[m], n_obs[m], replaced_by[m] (mpReplaced, -1 = none), kf2_slots[n2]. */
int REFM(fuse_kf)(orc_frame_view *kf, const float *inv_level_sigma2, const orc_map_points *mps, const int *obs2_idx, int n2,
                 const int *point_ids, int n_points, float th, uint8_t *bad_out, int *nobs_out, int *replaced_out, int *kf2_slots_out) {
    Points P(mps);
    KeyFrame *K = make_keyframe(kf, P);
    K->mvInvLevelSigma2.assign(inv_level_sigma2, inv_level_sigma2 + kf->n_levels);
    KeyFrame *K2 = new KeyFrame(n2, kf->fx, kf->fy, kf->cx, kf->cy, kf->mbf, kf->mb, (int) kf->min_x, (int) kf->min_y, (int) kf->max_x,
                                (int) kf->max_y, kf->n_levels, kf->log_scale_factor);
    K2->mvuRight.assign(n2, -1.f);
    K2->mvpMapPoints.assign(n2, nullptr);
    for (size_t j = 0; j < P.mp.size(); j++) P.mp[j].nObs = 0;
    for (int i = 0; i < kf->n; i++) if (K->mvpMapPoints[i]) K->mvpMapPoints[i]->AddObservation(K, i);
    for (size_t j = 0; j < P.mp.size(); j++)
        if (obs2_idx[j] >= 0) { K2->mvpMapPoints[obs2_idx[j]] = &P.mp[j]; P.mp[j].AddObservation(K2, obs2_idx[j]); }
    MapPoint::mLog.clear();
    std::vector<MapPoint *> pts(n_points);
    for (int i = 0; i < n_points; i++) pts[i] = P.at(point_ids[i]);
    ORBmatcher m(0.6f);
    const int n = m.Fuse(K, pts, th);
    store_map_points(K->mvpMapPoints, P, kf->map_point);
    store_map_points(K2->mvpMapPoints, P, kf2_slots_out);
    for (size_t j = 0; j < P.mp.size(); j++) {
        bad_out[j] = P.mp[j].mbBad;
        nobs_out[j] = P.mp[j].nObs;
        replaced_out[j] = P.index(P.mp[j].mpReplaced);
    }
    delete K; delete K2;
    return n;
}

#ifndef SHIM_USE_ADAPTERS
/* The index Fuse picks for ONE point at a time against a key frame with no map points: best_idx[i] = bestIdx if bestDist <= TH_LOW else -1
 * (read back from the observation Fuse adds).  Pins the per-point search of the restatement (orc_fuse_candidates). */
void REFM(fuse_kf_each)(const orc_frame_view *kf, const float *inv_level_sigma2, const orc_map_points *mps, const int *point_ids, int n_points,
                       float th, int *best_idx) {
    Points P(mps);
    KeyFrame *K = make_keyframe(kf, P);
    K->mvInvLevelSigma2.assign(inv_level_sigma2, inv_level_sigma2 + kf->n_levels);
    K->mvpMapPoints.assign(kf->n, nullptr);
    ORBmatcher m(0.6f);
    for (int i = 0; i < n_points; i++) {
        best_idx[i] = -1;
        MapPoint *p = P.at(point_ids[i]);
        if (!p) continue;
        p->mbBad = false;
        p->mObservations.clear();
        std::vector<MapPoint *> one(1, p);
        if (m.Fuse(K, one, th) == 1) {
            best_idx[i] = (int) p->mObservations[K];
            K->mvpMapPoints[best_idx[i]] = nullptr;
            p->mObservations.clear();
        }
    }
    delete K;
}
#endif

}  // extern "C"
