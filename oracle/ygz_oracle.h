/* ORACLE — TEST INFRASTRUCTURE ONLY.  C API of the CPU restatement (ctypes-friendly).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (libygz_b200.so) never links or calls it. */
#pragma once
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y, size, angle, response; int octave, class_id; } orc_keypoint;  /* == cv::KeyPoint */

/* ---- primitives (cv2-pinned) */
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
void orc_border_reflect101(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int border, int dstride);
void orc_gauss7_sigma2_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
int  orc_fast9_16(const uint8_t* img, int w, int h, int stride, int threshold, int nms, orc_keypoint* out, int cap);
void orc_fast_atan2_array(const float* y, const float* x, float* out, int n);
void orc_cosf_sinf_array(const float* deg, float* cos_out, float* sin_out, int n);  /* a,b of computeOrbDescriptor */
int  orc_hamming256(const uint8_t* a, const uint8_t* b);

/* ---- ORBextractor */
typedef struct orc_extractor orc_extractor;
orc_extractor* orc_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
void orc_extractor_destroy(orc_extractor*);
void orc_extractor_tables(orc_extractor*, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                          int* features_per_level, int* umax16);
/* operator()(image, mask, keypoints, descriptors); returns n (<= cap) or -needed */
int  orc_extract(orc_extractor*, const uint8_t* img, int w, int h, int stride, orc_keypoint* kps, uint8_t* desc, int cap);
/* ComputePyramid + operator()(Frame*, ..., ORBSLAM_KEYPOINT, leftEye=true): kps[0..n_existing) are pre-existing keys */
int  orc_extract_frame(orc_extractor*, const uint8_t* img, int w, int h, int stride, int n_existing,
                       orc_keypoint* kps, uint8_t* desc, int cap);
/* stage outputs of the last extract call */
/* operator()(Frame*, ..., DSO_KEYPOINT) :1031-1127 = ComputeKeyPointsDSOSingleLevel :1277-1381 + descriptors.  kps holds the
 * n_existing direct-tracked keys on entry (their angles are refreshed, :1377-1380) and all keys on return;
 * *grid_size is mnGridSize before / after the call (-1 = first frame). */
int  orc_extract_frame_dso(orc_extractor*, const uint8_t* img, int w, int h, int stride, int n_existing,
                           orc_keypoint* kps, uint8_t* desc, int cap, int* grid_size);
/* fast::fast_corner_detect_10_sse2 semantics on a w x h window inside a larger image (Thirdparty/fast) */
int  orc_fast10_detect_cell(const uint8_t* img, int w, int h, int stride, int barrier, short* xy, int cap);
/* ShiTomasiScore(mvImagePyramid[0], u, v) :1152-1186 (after a ComputePyramid / extract call) */
float orc_shi_tomasi(orc_extractor*, int u, int v);
int  orc_level_size(orc_extractor*, int level, int* w, int* h);
int  orc_get_level(orc_extractor*, int level, int blurred, uint8_t* out);
int  orc_get_candidates(orc_extractor*, int level, orc_keypoint* out, int cap);
int  orc_get_level_keys(orc_extractor*, int level, orc_keypoint* out, int cap);
/* DistributeOctTree alone (keys rel. to (minX,minY)) */
int  orc_distribute_octtree(orc_extractor*, const orc_keypoint* keys, int n, int minX, int maxX, int minY, int maxY,
                            int N, orc_keypoint* out, int cap);
/* batch helper for CPU-baseline timing: frames processed with `threads` std::threads; returns total keypoints */
long orc_extract_batch(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh,
                       const uint8_t* imgs, int n_frames, int w, int h, int threads);


/* ---- flat views of the SLAM objects the matcher / aligner read (pointer graph -> arrays).
 * Same layout as include/ygz_b200.h's ygz_frame_view / ygz_map_points so tests can share buffers. */
typedef struct {
    int n;                          /* Frame::N */
    const orc_keypoint* keys;       /* mvKeys */
    const uint8_t* desc;            /* mDescriptors, n x 32 */
    const float* u_right;           /* mvuRight (NULL: all -1) */
    const uint8_t* outlier;         /* mvbOutlier (NULL: none) */
    int* map_point;                 /* mvpMapPoints as indices into the map-point table, -1 = NULL */
    float fx, fy, cx, cy, mb, mbf;  /* Frame statics */
    float min_x, min_y, max_x, max_y;         /* mnMinX.. (image bounds) */
    int n_levels;
    const float* scale_factors;     /* mvScaleFactors */
    const float* inv_scale_factors; /* mvInvScaleFactors */
    float log_scale_factor;         /* mfLogScaleFactor */
    float Rcw[9];                   /* mTcw rotation, row-major */
    float tcw[3];                   /* mTcw translation */
    int fv_nodes;                   /* DBoW2::FeatureVector as CSR, node ids ascending */
    const uint32_t* fv_node_id;
    const int* fv_offset;           /* fv_nodes + 1 */
    const uint32_t* fv_index;
} orc_frame_view;

typedef struct {
    int m;
    const float* world_pos;         /* m x 3, GetWorldPos */
    const uint8_t* desc;            /* m x 32, GetDescriptor */
    const uint8_t* bad;             /* isBad (NULL: none) */
    const int* observations;        /* Observations (NULL: all 1) */
    const float* min_dist;          /* mfMinDistance */
    const float* max_dist;          /* mfMaxDistance */
    const uint8_t* track_in_view;   /* fields written by Frame::isInFrustum */
    const float* track_proj_x;
    const float* track_proj_y;
    const float* track_proj_xr;
    const int* track_scale_level;
    const float* track_view_cos;
    const float* normal;            /* m x 3, GetNormal (the Sim3 searches; NULL elsewhere) */
} orc_map_points;

/* ---- ORBmatcher (src/ORBmatcher.cc).  All return nmatches like the reference. */
/* SearchByBoW(KF, F, vpMapPointMatches) :155-263. out_matches[F.n] = KF map-point index or -1 */
int orc_search_by_bow(float nnratio, int check_ori, const orc_frame_view* kf, const orc_map_points* mps,
                      const orc_frame_view* f, int* out_matches);
/* SearchByProjection(F, vpMapPoints, th, checkLevel) :43-126; query_ids[nq] index mps; mutates f->map_point */
int orc_search_by_projection_points(float nnratio, orc_frame_view* f, const orc_map_points* mps,
                                    const int* query_ids, int nq, float th, int check_level);
/* SearchByProjection(Cur, Last, th, bMono, checkLevel) :1218-1350; mutates cur->map_point */
int orc_search_by_projection_last(int check_ori, orc_frame_view* cur, const orc_frame_view* last,
                                  const orc_map_points* mps, float th, int mono, int check_level);
/* SearchByProjection(Cur, KF, sAlreadyFound, th, ORBdist) :1352-1469; already_found[m] flags */
int orc_search_by_projection_kf(int check_ori, orc_frame_view* cur, const orc_frame_view* kf,
                                const orc_map_points* mps, const uint8_t* already_found, float th, int orb_dist);
/* SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) :375-478 */
int orc_search_for_initialization(float nnratio, int check_ori, const orc_frame_view* f1, const orc_frame_view* f2,
                                  float* prev_matched_xy, int* matches12, int window);
/* SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) :480-595: matches12[i] = map point (of KF2) matched to feature i of KF1 */
int orc_search_by_bow_kf(float nnratio, int check_ori, const orc_frame_view* kf1, const orc_frame_view* kf2,
                         const orc_map_points* mps, int* matches12);
/* SearchForTriangulation :597-746 (+ CheckDistEpipolarLine :136-153).  F12 row-major; (ex, ey) = epipole of KF1 in KF2
 * (:603-609); level_sigma2 = pKF2->mvLevelSigma2; pairs = (idx1, idx2) in ascending idx1 (capacity kf1->n). */
int orc_search_for_triangulation(int check_ori, const orc_frame_view* kf1, const orc_frame_view* kf2, const float* F12,
                                 float ex, float ey, const float* level_sigma2, int only_stereo, int* pairs);
/* The Sim3 family of the back end (cv::Mat geometry; point_ids index the map-point table, which must carry `normal`):
 * SearchByProjection(KF, Scw, vpPoints, vpMatched, th) :265-373 — Scw 4x4 row-major, matched[kf->n] in / out */
int orc_search_by_projection_sim3(const orc_frame_view* kf, const float* Scw, const orc_map_points* mps, const int* point_ids, int n_points,
                                  int* matched, int th);
/* per-point search of Fuse(pKF, vpMapPoints, th) :748-886 (see orb_matcher.cpp); point_ids[i] = -1 is a NULL entry */
void orc_fuse_candidates(const orc_frame_view* kf, const float* Ow, const float* inv_level_sigma2, const orc_map_points* mps,
                         const int* point_ids, int n_points, float th, int* best_idx, int* best_dist);
/* Fuse(KF, Scw, vpPoints, th, vpReplacePoint) :888-1004 — kf->map_point in / out (AddMapPoint), replace_point[n_points] in / out */
int orc_fuse_sim3(orc_frame_view* kf, const float* Scw, const orc_map_points* mps, const int* point_ids, int n_points, float th,
                  int* replace_point);
/* SearchBySim3(KF1, KF2, vpMatches12, s12, R12, t12, th) :1006-1216 — idx_in_kf2[m] = GetIndexInKeyFrame(pKF2) */
int orc_search_by_sim3(const orc_frame_view* kf1, const orc_frame_view* kf2, const orc_map_points* mps, int* matches12,
                       const int* idx_in_kf2, float s12, const float* R12, const float* t12, float th);
/* ComputeThreeMaxima :1471-1502 */
void orc_three_maxima(const int* hist_sizes, int L, int* ind1, int* ind2, int* ind3);
/* Frame::GetFeaturesInArea src/Frame.cc:424-481 (grid built as AssignFeaturesToGrid :314-330) */
int orc_features_in_area(const orc_frame_view* f, float x, float y, float r, int min_level, int max_level,
                         int* out, int cap);
/* dense brute force best / second best (plain restatement used to check the GPU top-2 kernel) */
void orc_bf_top2(const uint8_t* q, int nq, const uint8_t* t, int nt, int* best_idx, int* best_d, int* second_d);

/* ---- SparseImgAlign (src/SparseImageAlign.cc, include/NLSSolver_impl.hpp) */
typedef struct {
    int n_levels;                   /* pyramid levels available */
    const uint8_t* const* ref_levels;   /* continuous, stride == w */
    const uint8_t* const* cur_levels;
    const int* level_w;
    const int* level_h;
    const float* inv_scale_factors; /* ref_frame->mvInvScaleFactors */
    float fx, fy, cx, cy;
    int n;                          /* ref_frame->N */
    const float* ref_px;            /* n x 2: mvKeys[i].pt */
    const float* world_xyz;         /* n x 3: mvpMapPoints[i]->GetWorldPos() */
    const uint8_t* valid;           /* n: map point non-null, !isBad, !mvbOutlier */
    float ref_q[4], ref_t[3];       /* ref_frame->mTcw: unit quaternion (x,y,z,w) + translation */
    float cur_q[4], cur_t[3];       /* cur_frame->mTcw (initial guess) */
} orc_sia_problem;
/* run(ref, cur, TCR) :20-49.  Returns n_meas_/patch_area_.  out_q/out_t = TCR; H_out 36 floats (last H_);
 * trace (optional, 8 floats per evaluated iteration: level, iter, chi2, n_meas, max|x|, accepted,0,0) */
int orc_sia_run(const orc_sia_problem* p, int max_level, int min_level, int n_iter, float eps,
                float* out_q, float* out_t, float* H_out, float* trace, int trace_cap, int* trace_n);
/* test switch (per calling thread): 1 = accumulate H, Jres and chi2 in double and round once — exact-arithmetic reference for
 * the float sums of computeResiduals (how far the reference's own float accumulation is from the exact value) */
void orc_sia_wide_sums(int on);
/* ---- direct (patch) projection search: ORBmatcher::FindDirectProjection src/ORBmatcher.cc:1575-1602 with
 * GetWarpAffineMatrix :1525-1547, WarpAffine :1549-1573 and ygz::Align2D src/Align.cc:8-103.   */
typedef struct {
    const uint8_t* level[16];       /* mvImagePyramid[l] data (row stride pitch[l]) */
    int w[16], h[16], pitch[16];
    float q[4], t[3];               /* mTcw / GetPose(): unit quaternion (x,y,z,w) + translation */
} orc_direct_frame;
typedef struct {
    int ref;                        /* index of the observing keyframe in `refs` */
    float px_ref[2];                /* ref->mvKeys[idx].pt */
    int octave;                     /* ref->mvKeys[idx].octave */
    float pw[3];                    /* mp->GetWorldPos() */
    float px_cur[2];                /* initial guess (mp->mTrackProjX/Y) */
} orc_direct_item;
/* per item: out_px = refined px_curr (level-0 pixels), out_level = search_level, out_ok = return value */
int orc_find_direct_projection(const orc_direct_frame* refs, int n_refs, const orc_direct_frame* cur,
                               const float* K4, const float* scale, const float* inv_scale,
                               float inv_level_sigma2_1, int nlevels, const orc_direct_item* items,
                               int n_items, float* out_px, int* out_level, uint8_t* out_ok);
/* Frame::ComputeStereoMatches src/Frame.cc:509-682.  pyrL / pyrR: the two extractors'
 * mvImagePyramid (pose fields unused); keys in level-0 pixels.  Outputs mvuRight / mvDepth (-1 = no match); returns the
 * number of matched left key points. */
int orc_compute_stereo_matches(const orc_direct_frame* pyrL, const orc_direct_frame* pyrR, int nlevels,
                               const orc_keypoint* keysL, const uint8_t* descL, int N,
                               const orc_keypoint* keysR, const uint8_t* descR, int Nr,
                               const float* mvScaleFactors, const float* mvInvScaleFactors, float mb, float mbf,
                               float* mvuRight, float* mvDepth);
/* ---- DBoW2 vocabulary: loadFromBinaryFile TemplatedVocabulary.h:1478-1522, transform :1151-1292 */
typedef struct orc_vocab orc_vocab;
orc_vocab* orc_vocab_from_memory(const uint8_t* buf, size_t bytes);
orc_vocab* orc_vocab_load(const char* path);
void orc_vocab_destroy(orc_vocab*);
void orc_vocab_info(const orc_vocab*, int* k, int* L, int* scoring, int* weighting, int* n_nodes, int* n_words);
void orc_vocab_transform(const orc_vocab*, const uint8_t* desc, int n, int levelsup, int* word_id, float* weight, int* node_id);
int orc_vocab_transform_bow(const orc_vocab*, const uint8_t* desc, int n, int levelsup, int* bow_ids, double* bow_vals,
                            int* fv_nodes, int* fv_node_id, int* fv_offset, int* fv_index);
/* CPU front-end baseline: extractor + SearchByBoW vs previous frame, `threads` std::threads (bench.py) */
long orc_frontend_batch(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, const uint8_t* imgs,
                        int n_frames, int w, int h, int threads, float nnratio, int check_ori, long* total_keypoints,
                        int repeat);
/* pieces of bench.py's OpenCV-path CPU baseline: the in-tree tail of ORBextractor::operator() on externally produced
 * levels / blurred levels / FAST candidates (relative to (16,16)), and ComputeBoW x2 + SearchByBoW of one frame pair */
int orc_extract_tail(orc_extractor*, int nlevels, const uint8_t* const* levels, const uint8_t* const* blurred, const int* w,
                     const int* h, const orc_keypoint* const* cands, const int* ncand, orc_keypoint* kps, uint8_t* desc, int cap);
int orc_bow_match_pair(const orc_vocab* voc, int levelsup, float nnratio, int check_ori, const orc_keypoint* ka, const uint8_t* da,
                       int na, const orc_keypoint* kb, const uint8_t* db, int nb);
/* the same with Frame::ComputeBoW through a vocabulary (NULL: the byte0 >> 2 stand-in) */
long orc_frontend_batch_voc(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, const uint8_t* imgs,
                            int n_frames, int w, int h, int threads, float nnratio, int check_ori, long* total_keypoints,
                            int repeat, const orc_vocab* voc, int levelsup);
/* Sophus pieces exposed for unit tests */
void orc_se3_exp(const float* x6, float* q, float* t);
void orc_se3_mul(const float* qa, const float* ta, const float* qb, const float* tb, float* q, float* t);
void orc_se3_inv(const float* q, const float* t, float* qo, float* to);
void orc_se3_apply(const float* q, const float* t, const float* p3, float* o3);
void orc_quat_to_matrix(const float* q, float* R9);
void orc_ldlt_solve6(const float* A_rowmajor36, const float* b6, float* x6);

#ifdef __cplusplus
}
#endif
