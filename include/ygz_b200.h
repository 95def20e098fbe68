/* ygz_b200 — C ABI of the B200-native ORB-YGZ-SLAM image front end (libygz_b200.so).
 *
 * The reference has no FFI layer: the boundary is the C++ class surface in namespace ygz, compiled
 * into libygz-SLAM.so and used directly by Tracking.cc / Frame.cc.  The host-side
 * mirror of those classes lives in include/ygz/{ORBextractor,ORBmatcher,SparseImageAlign}.h and is a
 * thin C++ layer over the entry points below; each entry point cites the reference interface it
 * replaces (file:line relative to /root/reference).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a negative
 * ygz_status on error, never throws (reference failure convention: "0 features / 0 matches",
 * SURVEY.md §8b "Errors").  Handles own their device buffers and one CUDA stream; a handle may be
 * used from one host thread at a time, different handles concurrently (include/ORBextractor.h is
 * stateful and not re-entrant either; L/R extractors run on two threads, src/Frame.cc:728-731).
 */
#ifndef YGZ_B200_H
#define YGZ_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    YGZ_OK = 0,
    YGZ_ERR_ARG = -1,        /* bad argument / unsupported size */
    YGZ_ERR_CUDA = -2,       /* CUDA runtime error (ygz_last_error() has the text) */
    YGZ_ERR_CAPACITY = -3,   /* caller buffer or internal pool too small */
    YGZ_ERR_NO_DEVICE = -4
} ygz_status;

const char* ygz_last_error(void);
int ygz_device_count(void);
/* name (<=255 chars), SM count, L2 bytes of device `dev` */
int ygz_device_info(int dev, char* name, int* sm_count, size_t* l2_bytes);

/* == cv::KeyPoint (28 bytes, same field order): the element type of mvKeys (include/Frame.h) */
typedef struct { float x, y, size, angle, response; int octave, class_id; } ygz_keypoint;

/* ------------------------------------------------------------------ ORBextractor
 * replaces ygz::ORBextractor (include/ORBextractor.h:45-172, src/ORBextractor.cc). */
typedef struct {
    int nfeatures;        /* ORBextractor ctor args, include/ORBextractor.h:59-60 */
    float scale_factor;
    int nlevels;
    int ini_th_fast;      /* 1..255: threshold of the first cv::FAST call per cell */
    int min_th_fast;      /* 1..255: threshold of the second call, made only for cells the first left empty */
    int width, height;    /* frame geometry the handle is sized for */
    int max_batch;        /* frames per launch (1 for the per-frame drop-in path) */
    int device;           /* CUDA device ordinal */
} ygz_orb_config;

typedef struct ygz_orb ygz_orb;

int ygz_orb_create(const ygz_orb_config* cfg, ygz_orb** out);
int ygz_orb_destroy(ygz_orb* h);
/* Getters of include/ORBextractor.h:82-104 (arrays of nlevels floats) and mnFeaturesPerLevel */
int ygz_orb_tables(ygz_orb* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                   int* features_per_level);
int ygz_orb_level_size(ygz_orb* h, int level, int* w, int* hgt);
int ygz_orb_max_keypoints(ygz_orb* h);    /* upper bound of keypoints per frame (caller buffer size) */
int ygz_orb_max_batch(ygz_orb* h);        /* frames per call the handle was created for (ygz_orb_config.max_batch) */

/* ORBextractor::operator()(InputArray image, mask, keypoints, descriptors), src/ORBextractor.cc:970-1028,
 * for a batch of n independent frames (n = 1 is the reference call).  HOST buffers:
 *   images   n frames, frame f at images + f*frame_stride, rows row_stride bytes apart, CV_8UC1
 *   kps      n * cap entries; desc n * cap * 32 bytes; counts n ints (keypoints of frame f)
 * Host->device and device->host copies are part of the call. */
int ygz_orb_extract(ygz_orb* h, const uint8_t* images, int n, size_t frame_stride, int row_stride,
                    ygz_keypoint* kps, uint8_t* desc, int* counts, int cap);

/* Asynchronous form of ygz_orb_extract: enqueues H2D + kernels + D2H on the handle's stream and returns; the host
 * buffers must be pinned and stay valid until ygz_orb_wait(h).  Alternating two handles overlaps copies and kernels. */
int ygz_orb_extract_async(ygz_orb* h, const uint8_t* images, int n, size_t frame_stride, int row_stride,
                          ygz_keypoint* kps, uint8_t* desc, int* counts, int cap);
int ygz_orb_wait(ygz_orb* h);

/* ORBextractor::operator()(Frame*, keypoints, descriptors, ORBSLAM_KEYPOINT, leftEye),
 * src/ORBextractor.cc:1031-1127, single frame: kps[0..n_existing) are the frame's pre-existing
 * (direct-tracked) keys whose descriptors are recomputed first (:1088-1099); new keys are appended.
 * The pyramid is (re)built from `image` (Frame ctor -> ComputePyramid, src/Frame.cc:311,807). */
int ygz_orb_extract_frame(ygz_orb* h, const uint8_t* image, int row_stride, int n_existing,
                          ygz_keypoint* kps, uint8_t* desc, int* count, int cap);

/* ORBextractor::ComputePyramid (src/ORBextractor.cc:1129-1150) only: builds levels on the device. */
/* ORBextractor::operator()(Frame*, keypoints, descriptors, DSO_KEYPOINT), src/ORBextractor.cc:1031-1127 — what
 * Frame::ExtractORB runs for frames that already carry direct-tracked points (src/Frame.cc:335-337):
 * ComputeKeyPointsDSOSingleLevel :1277-1381 (FAST-10 of Thirdparty/fast per adaptive grid cell, thresholds 20 -> 5,
 * ShiTomasiScore :1152-1186, three best corners per cell, cell size shrunk by 5 until nfeatures are found) on level 0,
 * then descriptors of the n_existing keys (angles refreshed by IC_Angle, :1377-1380) and of the new keys (size 7,
 * octave 0).  *grid_size carries mnGridSize from frame to frame (-1 before the first frame; NULL: kept in the handle).
 * Like ygz_orb_extract_frame it may grow the handle's output capacity (ygz_orb_max_keypoints) to hold existing + new keys.
 * SURVEY.md §8(f) rank 2. */
int ygz_orb_extract_frame_dso(ygz_orb* h, const uint8_t* image, int row_stride, int n_existing, ygz_keypoint* kps,
                              uint8_t* desc, int* count, int cap, int* grid_size);
/* The same for n_frames frames of n_frames independent streams in one call (batched over cameras / sequences; frames of ONE stream are
 * sequential through mnGridSize).  images as in ygz_orb_upload; kps / desc are [n_frames][cap] rows whose first n_existing[f] entries are
 * frame f's direct-tracked keys; counts[f] = keys of frame f after the call; grid_sizes[f] = that stream's mnGridSize, in / out (-1: first
 * frame).  Results are identical to n_frames calls of ygz_orb_extract_frame_dso; the adaptive cell-size loop runs on the device. */
int ygz_orb_extract_frame_dso_batch(ygz_orb* h, const uint8_t* images, int n_frames, size_t frame_stride, int row_stride,
                                    const int* n_existing, ygz_keypoint* kps, uint8_t* desc, int* counts, int cap, int* grid_sizes);
int ygz_orb_compute_pyramid(ygz_orb* h, const uint8_t* images, int n, size_t frame_stride, int row_stride);
/* mvImagePyramid[level] of frame f (include/ORBextractor.h:106), continuous w*h bytes (step == cols,
 * as Frame clones it, src/Frame.cc:810-813).  border > 0 additionally reproduces the EDGE_THRESHOLD
 * REFLECT_101 frame of src/ORBextractor.cc:1141-1146 into a (w+2b)*(h+2b) buffer. */
int ygz_orb_get_level(ygz_orb* h, int frame, int level, int border, uint8_t* out);

/* GaussianBlur(mvImagePyramid[level], 7x7, sigma 2, REFLECT_101) as the last extraction left it on the device
 * (src/ORBextractor.cc:1009-1010, :1079-1084): continuous w*h bytes.  Parity instrumentation (stage tests). */
int ygz_orb_get_blurred_level(ygz_orb* h, int frame, int level, uint8_t* out);

/* Split phases for device-resident timing (bench.py `value`) and stage-level parity tests. */
int ygz_orb_upload(ygz_orb* h, const uint8_t* images, int n, size_t frame_stride, int row_stride);
int ygz_orb_run(ygz_orb* h, int n);                       /* all kernels, inputs already in HBM */
int ygz_orb_download(ygz_orb* h, int n, ygz_keypoint* kps, uint8_t* desc, int* counts, int cap);
int ygz_orb_sync(ygz_orb* h);
/* FAST candidates of (frame, level) before DistributeOctTree, in the reference's order (cell
 * row-major, raster inside a cell), coordinates relative to (16,16) (src/ORBextractor.cc:773-777). */
int ygz_orb_get_candidates(ygz_orb* h, int frame, int level, ygz_keypoint* out, int cap);
/* device-timed repeats of ygz_orb_run on the handle's stream (CUDA events): total ms and per-stage ms
 * stage_ms[6] = {pyramid, fast_cells, octree, blur, orient+describe, total}; launches = kernels per run */
int ygz_orb_time(ygz_orb* h, int n, int iters, float* stage_ms, int* launches);
/* device pointers for zero-copy chaining into the matcher / aligner of the same process */
int ygz_orb_device_outputs(ygz_orb* h, void** d_kps, void** d_desc, void** d_counts, int* cap);
int ygz_orb_device_level(ygz_orb* h, int level, void** d_ptr, int* pitch, size_t* frame_stride);
void* ygz_orb_stream(ygz_orb* h);

/* Parity instrumentation: (cos, sin) of computeOrbDescriptor (src/ORBextractor.cc:108-109) as the device evaluates
 * them, for the n consecutive float bit patterns starting at first_bits taken as kpt.angle in degrees. */
int ygz_debug_cos_sin(int device, unsigned first_bits, unsigned n, float* cos_out, float* sin_out);

/* ------------------------------------------------------------------ ORBmatcher
 * replaces the Hamming search core of ygz::ORBmatcher (include/ORBmatcher.h:38-220, src/ORBmatcher.cc). */

/* ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:1507-1523 (host, pure) */
int ygz_descriptor_distance(const uint8_t* a, const uint8_t* b);

typedef struct ygz_matcher ygz_matcher;
typedef struct ygz_vocab ygz_vocab;      /* DBoW2 vocabulary on the device, see below */
int ygz_matcher_create(int device, int max_queries, int max_targets, int max_pairs, ygz_matcher** out);
int ygz_matcher_destroy(ygz_matcher* h);

/* Dense brute force: for each of nq query descriptors the best / second-best of nt targets
 * (first index wins ties, as the `dist < bestDist` scans of src/ORBmatcher.cc:207-213). HOST buffers. */
int ygz_hamming_bf_top2(ygz_matcher* h, const uint8_t* q, int nq, const uint8_t* t, int nt,
                        int* best_idx, int* best_dist, int* second_dist);
/* Candidate-list distances: dist[p] = Hamming(q[pair_q[p]], t[pair_t[p]]) for n_pairs pairs — the
 * inner loop of every Search* variant (BoW bucket, grid window, row band). */
int ygz_hamming_pairs(ygz_matcher* h, const uint8_t* q, int nq, const uint8_t* t, int nt,
                      const int* pair_q, const int* pair_t, int n_pairs, int* dist);

/* Flat views of the pointer graph the Search* functions walk (Frame / KeyFrame / MapPoint). */
typedef struct {
    int n;                          /* Frame::N */
    const ygz_keypoint* keys;       /* mvKeys */
    const uint8_t* desc;            /* mDescriptors, n x 32 */
    const float* u_right;           /* mvuRight (NULL: all -1) */
    const uint8_t* outlier;         /* mvbOutlier (NULL: none) */
    int* map_point;                 /* mvpMapPoints as indices into ygz_map_points, -1 = NULL */
    float fx, fy, cx, cy, mb, mbf;  /* Frame statics, include/Frame.h:231-236 */
    float min_x, min_y, max_x, max_y;         /* mnMinX.. */
    int n_levels;
    const float* scale_factors;     /* mvScaleFactors */
    const float* inv_scale_factors; /* mvInvScaleFactors */
    float log_scale_factor;         /* mfLogScaleFactor */
    float Rcw[9];                   /* mTcw rotation, row-major */
    float tcw[3];                   /* mTcw translation */
    int fv_nodes;                   /* DBoW2::FeatureVector (map<NodeId, vector<uint>>) as CSR, ids ascending */
    const uint32_t* fv_node_id;
    const int* fv_offset;           /* fv_nodes + 1 */
    const uint32_t* fv_index;
} ygz_frame_view;

typedef struct {
    int m;
    const float* world_pos;         /* m x 3, MapPoint::GetWorldPos */
    const uint8_t* desc;            /* m x 32, MapPoint::GetDescriptor */
    const uint8_t* bad;             /* isBad (NULL: none) */
    const int* observations;        /* Observations (NULL: all 1) */
    const float* min_dist;          /* mfMinDistance */
    const float* max_dist;          /* mfMaxDistance */
    const uint8_t* track_in_view;   /* written by Frame::isInFrustum, src/Frame.cc:363-422 */
    const float* track_proj_x;
    const float* track_proj_y;
    const float* track_proj_xr;
    const int* track_scale_level;
    const float* track_view_cos;
    const float* normal;            /* m x 3, MapPoint::GetNormal (used by the Sim3 searches only; may be NULL elsewhere) */
} ygz_map_points;

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), src/ORBmatcher.cc:155-263.
 * out_matches[f->n] = map-point index or -1; returns nmatches (>= 0) or a negative ygz_status. */
int ygz_search_by_bow(ygz_matcher* h, float nnratio, int check_orientation, const ygz_frame_view* kf,
                      const ygz_map_points* mps, const ygz_frame_view* f, int* out_matches);
/* SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12), src/ORBmatcher.cc:480-595 (loop closing /
 * relocalisation candidates): matches12[i] = index of the KF2 map point matched to feature i of KF1, -1 = none. */
int ygz_search_by_bow_kf(ygz_matcher* h, float nnratio, int check_ori, const ygz_frame_view* kf1, const ygz_frame_view* kf2,
                         const ygz_map_points* mps, int* matches12);
/* SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo), src/ORBmatcher.cc:597-746 with
 * CheckDistEpipolarLine :136-153.  F12 row-major 3x3; (ex, ey) = projection of KF1's camera centre into KF2 (:603-609);
 * level_sigma2 = pKF2->mvLevelSigma2.  pairs receives (idx1, idx2) in ascending idx1 (capacity 2 * kf1->n ints);
 * returns the number of pairs. */
int ygz_search_for_triangulation(ygz_matcher* h, int check_ori, const ygz_frame_view* kf1, const ygz_frame_view* kf2,
                                 const float* F12, float ex, float ey, const float* level_sigma2, int only_stereo, int* pairs);
/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, checkLevel), src/ORBmatcher.cc:43-126 */
int ygz_search_by_projection_points(ygz_matcher* h, float nnratio, ygz_frame_view* f, const ygz_map_points* mps,
                                    const int* query_ids, int nq, float th, int check_level);
/* ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono, checkLevel), src/ORBmatcher.cc:1218-1350 */
int ygz_search_by_projection_last(ygz_matcher* h, int check_orientation, ygz_frame_view* cur,
                                  const ygz_frame_view* last, const ygz_map_points* mps, float th, int mono,
                                  int check_level);
/* ORBmatcher::SearchByProjection(Frame& Cur, KeyFrame*, sAlreadyFound, th, ORBdist), src/ORBmatcher.cc:1352-1469 */
int ygz_search_by_projection_kf(ygz_matcher* h, int check_orientation, ygz_frame_view* cur,
                                const ygz_frame_view* kf, const ygz_map_points* mps, const uint8_t* already_found,
                                float th, int orb_dist);
/* ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:375-478 */
int ygz_search_for_initialization(ygz_matcher* h, float nnratio, int check_orientation, const ygz_frame_view* f1,
                                  const ygz_frame_view* f2, float* prev_matched_xy, int* matches12, int window);

/* The Sim3 family of the back end (loop closing; the map-point table must carry `normal`).  Scw: 4x4 row-major similarity.
 * SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th), src/ORBmatcher.cc:265-373: point_ids index the table;
 * matched[kf->n] = map-point index per key-frame feature, in / out (-1 = NULL). */
int ygz_search_by_projection_sim3(ygz_matcher* h, const ygz_frame_view* kf, const float* Scw, const ygz_map_points* mps,
                                  const int* point_ids, int n_points, int* matched, int th);
/* Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint), :888-1004: kf->map_point is the key frame's slot table, in / out (a free
 * slot receives the point: AddObservation + AddMapPoint); replace_point[n_points] in / out (upstream initialises it with NULL = -1):
 * the map point already sitting in the chosen slot.  Returns nFused. */
int ygz_fuse_sim3(ygz_matcher* h, ygz_frame_view* kf, const float* Scw, const ygz_map_points* mps, const int* point_ids, int n_points,
                  float th, int* replace_point);
/* The search of ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, th), src/ORBmatcher.cc:748-886 (local mapping:
 * SearchInNeighbors): for every entry of point_ids (index into mps, -1 = NULL) the key point of pKF the loop body picks - projection with
 * pKF's pose (kf->Rcw / tcw, Eigen float), positive depth, image bounds, distance invariance, viewing angle (needs mps->normal), predicted
 * level, window th * mvScaleFactors[level], level band, stereo-aware chi2 gate (kf->u_right, kf->mbf, inv_level_sigma2 = mvInvLevelSigma2),
 * smallest descriptor distance (first minimum in GetFeaturesInArea order).  best_idx[i] = -1 / best_dist[i] = 256 when nothing survives.
 * Ow = pKF->GetCameraCenter().  The isBad / IsInKeyFrame skip and the map mutation that ends the loop body (Replace / AddObservation /
 * AddMapPoint, `bestDist <= TH_LOW`) need the observation graph: ygz::b200::ORBmatcherT::Fuse (include/ygz/adapters.h) replays them on
 * the caller's objects in upstream's order.  None of the inputs read here is changed by those mutations for a point that is still
 * eligible when its turn comes, so the one-shot search equals the interleaved one. */
int ygz_fuse_candidates(ygz_matcher* h, const ygz_frame_view* kf, const float* Ow, const float* inv_level_sigma2,
                        const ygz_map_points* mps, const int* point_ids, int n_points, float th, int* best_idx, int* best_dist);
/* SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th), :1006-1216: matches12[kf1->n] in / out (map-point index of KF2's point, -1 =
 * none); idx_in_kf2[mps->m] = pMP->GetIndexInKeyFrame(pKF2) (-1 = not observed there); R12 row-major 3x3, t12 3. */
int ygz_search_by_sim3(ygz_matcher* h, const ygz_frame_view* kf1, const ygz_frame_view* kf2, const ygz_map_points* mps, int* matches12,
                       const int* idx_in_kf2, float s12, const float* R12, const float* t12, float th);

/* Batched, device-resident projection searches over the extractor's resident outputs (one CTA per pair; the sequential occupancy
 * semantics of the reference loops are resolved exactly by ordered reservation rounds, csrc/projection_batch.cuh).  Assumption: every
 * query map point has Observations() > 0 (a feature it takes blocks later queries); the per-call entry points above handle the
 * general case. */
typedef struct {
    float fx, fy, cx, cy, mb, mbf;          /* Frame statics, include/Frame.h:231-236; mb / mbf of the current frame */
    float min_x, min_y, max_x, max_y;       /* mnMinX .. mnMaxY */
} ygz_projection_camera;
/* ORBmatcher::SearchByProjection(Frame& Current, const Frame& Last, th, bMono, checkLevel), src/ORBmatcher.cc:1218-1350, for
 * n_pairs (last, current) frame pairs of the extractor batch.  Per-frame host arrays, indexed by frame of the batch
 * (B = ygz_orb_max_batch, cap = ygz_orb_max_keypoints): world_pos [B][cap][3] and valid [B][cap] = LastFrame.mvpMapPoints[i]'s
 * GetWorldPos() and (non-null && !mvbOutlier[i]); pose [B][12] = mTcw of each frame (R row-major, t); mp_desc [B][cap][32] =
 * pMP->GetDescriptor() (NULL: the last frame's own descriptor of the feature); u_right [B][cap] = mvuRight (NULL: monocular).
 * CurrentFrame.mvpMapPoints must be empty on entry (TrackWithMotionModel fills it with NULL first, src/Tracking.cc:1079).
 * out_matches [n_pairs][cap]: index of the LAST-frame feature whose map point current feature i received, -1 = none;
 * out_nmatches [n_pairs] = the function's return value; out_rounds (optional) = resolution rounds per pair. */
int ygz_search_by_projection_last_batch(ygz_matcher* h, ygz_orb* orb, const ygz_projection_camera* cam, int n_pairs, const int* last_frame,
                                        const int* cur_frame, const float* world_pos, const uint8_t* valid, const float* pose,
                                        const uint8_t* mp_desc, const float* u_right, float th, int mono, int check_level,
                                        int check_orientation, int* out_matches, int* out_nmatches, int* out_rounds);
/* ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>&, th, checkLevel), src/ORBmatcher.cc:43-126, for n_pairs frames:
 * pair p searches frame frame[p] with its own list of M map points (arrays [n_pairs][M]: mTrackProjX / Y / XR, mTrackViewCos,
 * mnTrackScaleLevel, usable = mbTrackInView && !isBad(), mp_desc [n_pairs][M][32]); occupied [n_pairs][cap] (optional) marks
 * features that hold a map point with Observations() > 0 before the call.  out_matches [n_pairs][cap]: index into the pair's
 * point list assigned to feature i, -1 = unchanged. */
int ygz_search_by_projection_points_batch(ygz_matcher* h, ygz_orb* orb, const ygz_projection_camera* cam, int n_pairs, const int* frame,
                                          int M, const float* proj_x, const float* proj_y, const float* proj_xr, const float* view_cos,
                                          const int* level, const uint8_t* usable, const uint8_t* mp_desc, const uint8_t* occupied,
                                          const float* u_right, float nnratio, float th, int check_level, int* out_matches,
                                          int* out_nmatches, int* out_rounds);
/* device-timed repeats of the last batched projection search (inputs resident) on the extractor's stream, for bench.py */
int ygz_search_by_projection_batch_time(ygz_matcher* h, ygz_orb* orb, int iters, float* ms);

/* Batched device-resident SearchByBoW over n independent (KF, F) pairs whose keypoints/descriptors are
 * extractor outputs already in HBM (stream front end, BASELINE C2): pair p matches frame kf_idx[p]
 * against frame f_idx[p] of the extractor batch; FeatureVectors are given per frame as node ids per
 * feature (node_of_feature[frame][cap], 0xFFFFFFFF = none; NULL selects the synthetic bench vocabulary
 * node = descriptor byte 0 >> 2, a stand-in for DBoW2 transform which is a "next" row).  All KF features
 * carry a map point.
 * out_matches: n x cap ints (host), out_nmatches: n ints. */
int ygz_search_by_bow_batch(ygz_matcher* h, ygz_orb* orb, float nnratio, int check_orientation,
                            const int* kf_idx, const int* f_idx, int n_pairs, const uint32_t* node_of_feature,
                            int* out_matches, int* out_nmatches);
/* Frame::ComputeBoW on the device (src/Frame.cc:495-501): attach a vocabulary and the batched searches below take the
 * FeatureVector node ids (levelsup levels above the leaves; the reference uses 4) from a k_vocab_transform launch over the
 * extractor's device-resident descriptors, on the same stream, instead of node_of_feature / the synthetic stand-in.
 * vocab == NULL detaches.  The vocabulary handle must outlive the matcher's use of it and live on the same device. */
int ygz_matcher_set_vocabulary(ygz_matcher* h, ygz_vocab* vocab, int levelsup);
/* asynchronous form on the extractor's stream (attached vocabulary, else the synthetic stand-in node = byte0 >> 2);
 * results valid after ygz_orb_wait(orb) */
int ygz_search_by_bow_batch_async(ygz_matcher* h, ygz_orb* orb, float nnratio, int check_orientation,
                                  const int* kf_idx, const int* f_idx, int n_pairs, int* out_matches, int* out_nmatches);
/* same, results left on the device; timing variant for bench.py (ms over `iters` repeats) */
int ygz_search_by_bow_batch_time(ygz_matcher* h, ygz_orb* orb, float nnratio, int check_orientation,
                                 int n_pairs, int iters, float* ms);

/* One whole front-end step (extractor kernels + batched SearchByBoW) on the extractor's stream, repeated
 * `iters` times and timed with CUDA events; inputs resident in HBM.  Requires a prior
 * ygz_search_by_bow_batch call on the same handles (pair lists resident). */
int ygz_frontend_time(ygz_orb* orb, ygz_matcher* h, int n_frames, int iters, float nnratio, int check_orientation,
                      float* ms);

/* Two batches in flight on two streams (handle pairs A and B, inputs resident, pair lists prepared): steps alternate
 * A, B, A, ...; CUDA-event timed from the start on A's stream until both streams have drained. */
int ygz_frontend_time2(ygz_orb* orbA, ygz_matcher* hA, ygz_orb* orbB, ygz_matcher* hB, int n_frames, int iters,
                       float nnratio, int check_orientation, float* ms);

/* The same with k (1..8) handle pairs in flight on k streams: step i runs on pair i % k; timed from the start on the first stream until
 * every stream has drained. */
int ygz_frontend_time_n(ygz_orb* const* orbs, ygz_matcher* const* matchers, int k, int n_frames, int iters, float nnratio,
                        int check_orientation, float* ms);

/* ------------------------------------------------------------------ SparseImgAlign
 * replaces ygz::SparseImgAlign (include/SparseImageAlign.h:14-113, src/SparseImageAlign.cc) and its
 * Gauss-Newton driver (include/NLSSolver_impl.hpp:17-91). */
typedef struct {
    int n_levels;
    const uint8_t* const* ref_levels;   /* HOST level images of the reference frame, continuous (stride == w) */
    const uint8_t* const* cur_levels;
    const int* level_w;
    const int* level_h;
    const float* inv_scale_factors;     /* ref_frame->mvInvScaleFactors */
    float fx, fy, cx, cy;
    int n;                              /* ref_frame->N */
    const float* ref_px;                /* n x 2: mvKeys[i].pt */
    const float* world_xyz;             /* n x 3: mvpMapPoints[i]->GetWorldPos() */
    const uint8_t* valid;               /* n: map point non-null && !isBad && !mvbOutlier[i] */
    float ref_q[4], ref_t[3];           /* ref_frame->mTcw as unit quaternion (x,y,z,w) + translation */
    float cur_q[4], cur_t[3];           /* cur_frame->mTcw (initial guess) */
} ygz_sia_problem;

typedef struct ygz_sia ygz_sia;
int ygz_sia_create(int device, int max_problems, int max_features, int max_level_pixels, ygz_sia** out);
int ygz_sia_destroy(ygz_sia* h);
/* SparseImgAlign::run(Frame* ref, Frame* cur, SE3f& TCR), src/SparseImageAlign.cc:20-49, for a batch of
 * n_problems independent alignments (n_problems = 1 is the reference call).  Levels max_level..min_level
 * (SparseImgAlign(nLevels-1, 1), src/Tracking.cc:207), n_iter Gauss-Newton iterations per level, eps as
 * eps_ (src/SparseImageAlign.cc:17).  Outputs per problem: T_cur_from_ref (q[4], t[3]), H (36 floats,
 * getFisherInformation's H_), n_tracked = n_meas_/patch_area_ (the return value of run). */
int ygz_sia_run(ygz_sia* h, const ygz_sia_problem* problems, int n_problems, int max_level, int min_level,
                int n_iter, float eps, float* out_q, float* out_t, float* out_H, int* out_n_tracked);
/* The same alignments with the level images read IN PLACE from extractor handles (the tracker's case: Frame::mvImagePyramid is the
 * extractor's pyramid, src/Frame.cc ComputeImagePyramid -> ORBextractor::ComputePyramid, and both frames were just extracted on this
 * device).  Problem p aligns frame ref_frame[p] of orb_ref with frame cur_frame[p] of orb_cur (the same handle may be passed twice, e.g.
 * consecutive frames of one batch).  Of ygz_sia_problem only fx..cy, n, ref_px, world_xyz, valid and the poses are read; the level
 * fields (n_levels, *_levels, level_w/h, inv_scale_factors) are ignored - sizes and mvInvScaleFactors come from the handles.  No image
 * bytes cross PCIe; the call waits for the extractors' streams on the device, not on the host.  All three handles must live on the same
 * device (the level pointers are used as they are), and the extractor handles must not be re-used for other frames until the call returns. */
int ygz_sia_run_orb(ygz_sia* h, ygz_orb* orb_ref, ygz_orb* orb_cur, const ygz_sia_problem* problems, const int* ref_frame,
                    const int* cur_frame, int n_problems, int max_level, int min_level, int n_iter, float eps, float* out_q,
                    float* out_t, float* out_H, int* out_n_tracked);
/* per-iteration record of the last ygz_sia_run for problem `problem`: 8 floats per evaluated iteration
 * {level, iter, chi2, n_meas, max|x|, accepted, 0, 0} (the verbose_ printout of NLSSolver_impl.hpp:76-83) */
int ygz_sia_last_trace(ygz_sia* h, int problem, float* out, int cap, int* n);
/* device-timed repeats (inputs resident) for bench.py */
int ygz_sia_time(ygz_sia* h, int n_problems, int max_level, int min_level, int n_iter, float eps, int iters,
                 float* ms);

/* ------------------------------------------------------------------ bag-of-words vocabulary
 * the DBoW2 ORB vocabulary of Frame::ComputeBoW (src/Frame.cc:495-501):
 * TemplatedVocabulary::loadFromBinaryFile (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1478-1522; ORBvoc.bin: 24-byte
 * header {nb_nodes, size_node = 41, k, L, scoring, weighting} + records {i32 parent, u8 desc[32], f32 weight, u8 is_leaf})
 * and transform(feature, word_id, weight, &nid, levelsup) (:1249-1282) for batches of descriptors.  The BowVector /
 * FeatureVector assembly (std::map inserts, :1151-1227) is a per-frame host step over these arrays (see the C++ and
 * Python mirrors). */
int ygz_vocab_load(const char* path, int device, ygz_vocab** out);
int ygz_vocab_from_memory(const uint8_t* buf, size_t bytes, int device, ygz_vocab** out);
int ygz_vocab_destroy(ygz_vocab* v);
int ygz_vocab_info(const ygz_vocab* v, int* k, int* L, int* scoring, int* weighting, int* n_nodes, int* n_words);
/* per descriptor: word id (leaf index), its weight (0 = stopped word), node id at level L - levelsup */
int ygz_vocab_transform(ygz_vocab* v, const uint8_t* desc, int n, int levelsup, int* word_id, float* weight, int* node_id);
/* the same for the descriptors an extractor holds on the device (last batch, n_frames frames, zero copy); outputs are
 * [n_frames][ygz_orb_max_keypoints] (NULL to skip one); ms (optional) = device time of the kernel */
int ygz_vocab_transform_orb(ygz_vocab* v, ygz_orb* orb, int n_frames, int levelsup, int* word_id, float* weight, int* node_id,
                            float* ms);

/* ------------------------------------------------------------------ stereo association
 * Frame::ComputeStereoMatches, /root/reference/src/Frame.cc:509-682: for every left key point the
 * best right key point of the row band (Hamming, compatible octave, disparity in [0, mbf/mb]), an 11x11 SAD sliding
 * window of +-5 px on the key's pyramid level, parabola refinement, then the median-based outlier cut.  The pyramids are
 * read in place from the two extractor handles (frame `frame` of their last batch).  Outputs mvuRight / mvDepth
 * (-1 = no match); returns the number of stereo matches or a negative ygz_status. */
int ygz_compute_stereo_matches(ygz_orb* left, ygz_orb* right, int frame, const ygz_keypoint* keysL, const uint8_t* descL, int N,
                               const ygz_keypoint* keysR, const uint8_t* descR, int Nr, float mb, float mbf,
                               float* mvuRight, float* mvDepth);

/* Batched, device-resident form (BASELINE C5: many stereo pairs per call): frame f of the left handle's last batch with frame f of
 * the right handle's, both extractors' key points / descriptors / pyramids read in place on the device (zero copy): the row-band
 * association and sub-pixel refinement of ComputeStereoMatches for every left key, the median cut (:668-681) per frame on the
 * device, and — when any bf_* pointer is given — C5's dense brute-force matcher (best / second-best right descriptor of every
 * left descriptor, first index wins ties).  Host outputs, each [n_frames][capL] with capL = ygz_orb_max_keypoints(left) (NULL to
 * skip); slots past a frame's key count hold -1.  n_matches[n_frames] = the return value of ComputeStereoMatches' counting per
 * frame.  The _async form enqueues on the left extractor's stream (ordered after the right one's); outputs (pinned) are valid
 * after ygz_orb_wait(left).  At most 4096 left key slots per frame. */
int ygz_stereo_batch(ygz_orb* left, ygz_orb* right, int n_frames, float mb, float mbf, float* mvuRight, float* mvDepth,
                     int* n_matches, int* bf_best, int* bf_dist, int* bf_second);
int ygz_stereo_batch_async(ygz_orb* left, ygz_orb* right, int n_frames, float mb, float mbf, float* mvuRight, float* mvDepth,
                           int* n_matches, int* bf_best, int* bf_dist, int* bf_second);
/* device-timed repeats of one whole C5 step, inputs resident: both extractors (each on its own stream) + the batched
 * association + median cut (+ brute force).  ms = total over `iters`; stage_ms[4] (optional, last iteration) =
 * {extractors L || R, association, median cut, brute force}. */
int ygz_stereo_frontend_time(ygz_orb* left, ygz_orb* right, int n_frames, int iters, float mb, float mbf, int with_bf, float* ms,
                             float* stage_ms);

/* ------------------------------------------------------------------ direct (patch) projection search
 * ORBmatcher::FindDirectProjection(KeyFrame* ref, Frame* curr, MapPoint* mp,
 * Vector2f& px_curr, int& search_level), /root/reference/src/ORBmatcher.cc:1575-1602, with GetWarpAffineMatrix
 * :1525-1547, WarpAffine :1549-1573, GetBestSearchLevel / GetBilateralInterpUchar include/ORBmatcher.h:185-211 and
 * ygz::Align2D src/Align.cc:8-103 — batched over the (map point, observing keyframe) pairs that
 * Tracking::SearchLocalPointsDirect (src/Tracking.cc:2174-2326) tries for one current frame.  Keyframe and
 * current-frame pyramids live in device slots (a keyframe is uploaded once, the current frame once per frame). */
typedef struct {
    const uint8_t* level[16];       /* mvImagePyramid[l].data (host), row stride pitch[l] */
    int w[16], h[16], pitch[16];
    float q[4], t[3];               /* mTcw / GetPose(): unit quaternion (x,y,z,w) + translation */
} ygz_direct_frame;
typedef struct {
    int ref;                        /* slot of the observing keyframe */
    float px_ref[2];                /* ref->mvKeys[idx].pt */
    int octave;                     /* ref->mvKeys[idx].octave */
    float pw[3];                    /* mp->GetWorldPos() */
    float px_cur[2];                /* initial guess: mp->mTrackProjX / mTrackProjY */
} ygz_direct_item;
typedef struct ygz_direct ygz_direct;
int ygz_direct_create(int device, int max_slots, int max_items, size_t max_pyramid_bytes, ygz_direct** out);
int ygz_direct_destroy(ygz_direct* h);
/* copy a frame's pyramid (nlevels images) and pose into device slot `slot` */
int ygz_direct_upload_frame(ygz_direct* h, int slot, const ygz_direct_frame* f, int nlevels);
/* new pose for a resident slot (the tracker refines curr->mTcw between searches) */
int ygz_direct_set_pose(ygz_direct* h, int slot, const float* q, const float* t);
/* K4 = {fx, fy, cx, cy} (static members of Frame / KeyFrame), scale / inv_scale = mvScaleFactors / mvInvScaleFactors,
 * inv_level_sigma2_1 = ref->mvInvLevelSigma2[1].  Per item: out_px[2] = px_curr after the search (level-0 pixels),
 * out_level = search_level, out_ok = the bool FindDirectProjection returns. */
int ygz_find_direct_projection(ygz_direct* h, int cur_slot, const float* K4, const float* scale, const float* inv_scale,
                               float inv_level_sigma2_1, int nlevels, const ygz_direct_item* items, int n_items,
                               float* out_px, int* out_level, uint8_t* out_ok);
/* device-timed repeats of the last search (items resident) for bench.py */
int ygz_direct_time(ygz_direct* h, int iters, float* ms);

/* ------------------------------------------------------------------ throughput counters (multi-GPU)
 * The only cross-rank exchange of the path is one all-gather of these; the library
 * fills them, the launcher (torch.distributed / NCCL) gathers them. */
typedef struct {
    uint64_t frames, keypoints, matches, sia_iters, bytes_in, bytes_out, kernel_launches, elapsed_ns;
} ygz_counters;
int ygz_orb_counters(ygz_orb* h, ygz_counters* out, int reset);

#ifdef __cplusplus
}
#endif
#endif /* YGZ_B200_H */
