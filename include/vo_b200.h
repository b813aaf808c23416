/*
 * vo_b200.h -- C ABI of the B200-native hot path of visual_odometry_ros.
 *
 * The reference has no FFI layer: its boundary is the public C++ surface of
 *   core/visual_odometry/feature_tracker.h:44-104          (FeatureTracker)
 *   core/visual_odometry/motion_estimator.h:107-147        (MotionEstimator)
 *   core/visual_odometry/ba_solver/sparse_bundle_adjustment.h:105-130
 *   core/util/triangulate_3d.h:16-30                       (mapping::triangulateDLT)
 *   standalone/motion_estimator/motion_estimator.h:22-35
 *   standalone/depth_filter/depth_filter.h:13-21
 * The C++ shim classes in visual_odometry_ros_b200/host/ keep those signatures and
 * call ONLY the functions below.  Plain pointers and sizes, no torch / Eigen / cv types.
 *
 * Conventions
 *  - every function returns an int status: VO_OK (0) or a negative VO_ERR_* code; the
 *    shim maps codes back to the reference's std::runtime_error texts.
 *  - pointers are HOST pointers unless the parameter name ends in `_d` (device pointer)
 *    or the function name ends in `_d` (all array arguments are device pointers; the
 *    call is asynchronous on the context's stream and does not synchronise).
 *  - 2-D points are interleaved float32 (x, y) == cv::Point2f (define_type.h:16);
 *    3-D points interleaved float32 (x, y, z) == Eigen::Vector3f (define_type.h:17);
 *    poses are 4x4 float32 ROW-major here (the shim transposes Eigen's column-major
 *    PoseSE3, define_type.h:42); masks are one uint8 per element (the shim unpacks
 *    std::vector<bool>, define_type.h:36).
 *  - one vo_ctx == one GPU + one CUDA stream + a set of image slots holding
 *    device-resident pyramids.  There is NO CPU fallback: without a CUDA device
 *    vo_ctx_create fails with VO_ERR_NO_DEVICE.
 */
#ifndef VO_B200_H_
#define VO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VO_API __attribute__((visibility("default")))
#else
#define VO_API
#endif

#define VO_OK                 0
#define VO_ERR_INVALID_ARG   -1
#define VO_ERR_CUDA          -2
#define VO_ERR_SIZE_MISMATCH -3  /* reference: "X.size() != pts1.size()" etc. */
#define VO_ERR_NAN           -4  /* reference: NaN runtime_errors */
#define VO_ERR_MODE          -5  /* reference: stereo/mono mode misuse */
#define VO_ERR_NO_DEVICE     -6
#define VO_ERR_LARGE_UPDATE  -7  /* reference: LBA "large update!" (sparse_bundle_adjustment.cpp:731) */

/* cv::OPTFLOW_USE_INITIAL_FLOW (feature_tracker.cpp:71,110,119,188) */
#define VO_KLT_USE_INITIAL_FLOW 4

typedef struct vo_ctx vo_ctx;

VO_API const char *vo_status_string(int status);
/* Last CUDA / argument error text recorded on this context (never NULL). */
VO_API const char *vo_last_error(const vo_ctx *ctx);
/* Library build info: "vo_b200 <version> sm_100a nvcc <ver>". */
VO_API const char *vo_build_info(void);

/* ------------------------------------------------------------------ context */
/* device: CUDA ordinal. n_slots image slots, each sized for max_w x max_h u8 images
 * (pyramid + Scharr derivative pyramid, all levels).  max_feat: capacity of the
 * internal staging used by the host-pointer entry points.  stream: a cudaStream_t to
 * launch on (e.g. torch's current stream) or NULL to let the context own one. */
VO_API int vo_ctx_create(int device, int max_w, int max_h, int n_slots, int max_feat,
                  void *stream, vo_ctx **out);
VO_API int vo_ctx_destroy(vo_ctx *ctx);
VO_API int vo_ctx_synchronize(vo_ctx *ctx);
/* Number of kernels this context has launched since creation (bench `gpu_launches`). */
VO_API long long vo_ctx_launch_count(const vo_ctx *ctx);

/* ------------------------------------------------------------------ images / pyramids
 * Replaces the per-call cv::buildOpticalFlowPyramid inside cv::calcOpticalFlowPyrLK
 * (feature_tracker.cpp:29 ...) -- pyramids are built once per image, on device. */
/* Async H2D of a CV_8UC1 image (row pitch `step` bytes) into slot; marks pyramid stale. */
VO_API int vo_upload_image(vo_ctx *ctx, int slot, const uint8_t *data, int w, int h, size_t step);
/* Same, source already on the device. */
VO_API int vo_set_image_d(vo_ctx *ctx, int slot, const uint8_t *data_d, int w, int h, size_t step);
/* Build pyramid levels 0..n_levels-1 (+ Scharr derivative if with_deriv) for a batch of
 * slots now (otherwise built lazily by the first tracking call that needs them). */
VO_API int vo_build_pyramids(vo_ctx *ctx, const int *slots, int n_slots, int n_levels, int with_deriv);
/* Mark the derived pyramid levels / derivatives of these slots stale (level-0 pixels stay):
 * the next tracking call rebuilds them, as the reference does on every call. */
VO_API int vo_invalidate_pyramids(vo_ctx *ctx, const int *slots, int n_slots);
/* Effective maxLevel after OpenCV's clamp (next level w or h <= win -> stop). */
VO_API int vo_effective_max_level(int w, int h, int win, int max_level);
/* Debug/parity read-back of one level: img (w_l*h_l u8) and/or deriv (w_l*h_l*2 int16). */
VO_API int vo_read_pyramid_level(vo_ctx *ctx, int slot, int level, uint8_t *img, int16_t *deriv,
                          int *w_l, int *h_l);

/* ------------------------------------------------------------------ stereo rectification
 * StereoCamera::generateStereoImagesUndistortAndRectifyMaps (core/visual_odometry/camera.cpp:364-546): builds the four
 * CV_32FC1 rectification maps on the device from the pinhole-radtan parameters (K = fx, fy, cx, cy; D = k1, k2, p1, p2,
 * k3 as in camera.cpp:30-35; T_lr 4x4 row-major) and returns the rectified intrinsics (fx, fy, cx, cy) and T_lr_rect. */
VO_API int vo_rectify_init(vo_ctx *ctx, const float *K_l4, const float *D_l5, const float *K_r4, const float *D_r5,
                    const float *T_lr, int w, int h, float *K_rect4_out, float *T_lr_rect_out);
/* Debug / parity read-back of one camera's maps (w*h floats each). */
VO_API int vo_read_rectify_maps(vo_ctx *ctx, int right, float *map_u, float *map_v);
/* StereoCamera::rectifyStereoImages for one image + the convertTo(CV_8UC1) of StereoVO (camera.cpp:300-336,
 * stereo_vo.cpp:416-421): H2D of the distorted image, cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) on the device, result
 * becomes the slot's image (pyramid stale). right = 0 / 1 selects the left / right maps. */
VO_API int vo_upload_image_rectified(vo_ctx *ctx, int slot, int right, const uint8_t *data, int w, int h, size_t step);
/* Single-camera undistortion for MonoVO with flagDoUndistortion: 1 (mono_vo.cpp:509-513): the map of
 * Camera::generateImageUndistortMaps (core/visual_odometry/camera.cpp:57-87; D5 = k1 k2 p1 p2 k3) becomes map set 0, so
 * vo_upload_image_rectified(ctx, slot, 0, ...) = Camera::undistortImage (:163-183) + convertTo(CV_8UC1).  The camera
 * matrix is unchanged. */
VO_API int vo_undistort_init(vo_ctx *ctx, const float *K4, const float *D5, int w, int h);

/* ------------------------------------------------------------------ raw pyramidal LK
 * == cv::calcOpticalFlowPyrLK(img[slot0], img[slot1], pts0, pts1, status, err,
 *        Size(win,win), max_level, TermCriteria(COUNT+EPS,30,0.01), flags, 1e-4)
 * err is written 0 where status==0 (OpenCV leaves it undefined there). */
VO_API int vo_klt_track(vo_ctx *ctx, int slot0, int slot1, const float *pts0, int n, int win,
                 int max_level, int flags, float *pts1_inout, uint8_t *status, float *err);
/* Batched, device-resident: n_pairs independent image pairs, n features each.
 * slots0/slots1 are HOST arrays of slot ids; *_d arrays are [n_pairs][n] row-major.
 * counters_d (nullable): int64[2*VO_MAX_LEVELS] = per level {windows templated,
 * LK iterations executed}, accumulated (used for the algorithmic-bytes model). */
#define VO_MAX_LEVELS 8
VO_API int vo_klt_track_batch_d(vo_ctx *ctx, int n_pairs, const int *slots0, const int *slots1,
                         const float *pts0_d, int n, int win, int max_level, int flags,
                         float *pts1_inout_d, uint8_t *status_d, float *err_d,
                         long long *counters_d);

/* ------------------------------------------------------------------ FeatureTracker methods
 * (feature_tracker.cpp:13-206).  mask_inout: pre-existing entries are ANDed in, exactly as
 * `mask_valid.resize(n, true)` keeps them (Appendix B #7); pass all-1 for a fresh mask. */
VO_API int vo_ft_track(vo_ctx *ctx, int slot0, int slot1, const float *pts0, int n, int window_size,
                int max_pyr_lvl, float thres_err, float *pts_track, uint8_t *mask_inout);
VO_API int vo_ft_track_with_prior(vo_ctx *ctx, int slot0, int slot1, const float *pts0, int n,
                           int window_size, int max_pyr_lvl, float thres_err,
                           float *pts_track_inout, uint8_t *mask_inout);
VO_API int vo_ft_track_bidirection(vo_ctx *ctx, int slot0, int slot1, const float *pts0, int n,
                            int window_size, int max_pyr_lvl, float thres_err,
                            float thres_bidirection, float *pts_track, uint8_t *mask_inout);
VO_API int vo_ft_track_bidirection_with_prior(vo_ctx *ctx, int slot0, int slot1, const float *pts0,
                                       int n, int window_size, int max_pyr_lvl, float thres_err,
                                       float thres_bidirection, float *pts_track_inout,
                                       uint8_t *mask_inout);
/* Batched FeatureTracker::track over n_pairs independent image pairs with HOST buffers:
 * imgs0/imgs1 are arrays of n_pairs host image pointers (CV_8UC1, w x h, row pitch `step`;
 * pinned memory makes the copies asynchronous; a NULL entry keeps the slot's current image),
 * pts0 / pts_track / mask_inout are [n_pairs][n].  One H2D per image, batched kernels, one
 * D2H of the results, one synchronisation. with_prior != 0 selects trackWithPrior.
 * Pairs whose slots are all distinct are pipelined in chunks (image DMA on a copy stream, chunks round-robin over four
 * compute streams).  A batch that reuses a slot across pairs (e.g. a chain slots1[i] == slots0[i+1] with imgs0[i+1] == NULL)
 * is detected and processed strictly in order on the context's stream: same results, no overlap. */
VO_API int vo_ft_track_batch(vo_ctx *ctx, int n_pairs, const int *slots0, const int *slots1,
                      const uint8_t *const *imgs0, const uint8_t *const *imgs1, int w, int h,
                      size_t step, const float *pts0, int n, int window_size, int max_pyr_lvl,
                      float thres_err, int with_prior, float *pts_track_inout, uint8_t *mask_inout);
/* FeatureTracker::calcPrior (feature_tracker.cpp:208-234). Tw1: 4x4 row-major; K: fx,fy,cx,cy. */
VO_API int vo_ft_calc_prior(vo_ctx *ctx, const float *pts0, const float *Xw, int n, const float *Tw1,
                     const float *K4, float *pts1_prior);
/* FeatureTracker::trackWithScale (feature_tracker.cpp:236-504). The float image and its
 * 3x3 Sobel derivatives (stereo_vo.cpp:551-552) are computed on device from slot0.
 * Samples that leave the image: the reference keeps its per-sample buffers and masks across features and iterations
 * (feature_tracker.cpp:324-333, image_processing.cpp:88-89), so such a sample reuses the value the previous feature /
 * iteration left at that index.  vo_set_scale_mode(ctx, 1) reproduces exactly that (the affected features -- a few per
 * frame -- are recomputed in a second pass against the buffer state their predecessors leave; every trackWithScale stage
 * of the context, also inside the frame steps); mode 0 (default) masks those samples out, the intended semantics. */
VO_API int vo_set_scale_mode(vo_ctx *ctx, int faithful_borders);
VO_API int vo_get_scale_mode(const vo_ctx *ctx);
VO_API int vo_ft_track_with_scale(vo_ctx *ctx, int slot0, int slot1, const float *pts0,
                           const float *scale_est, int n, float *pts_track_inout,
                           uint8_t *mask_inout);

/* ------------------------------------------------------------------ pose-only Gauss-Newton
 * MotionEstimator::poseOnlyBundleAdjustment (core motion_estimator.cpp:665-861 ==
 * standalone motion_estimator.cpp:4-193) and _Stereo (core :863-1088 == standalone :195-411).
 * R01/t01/T01 are in-out (row-major). Returns VO_OK and *success=0 when the result is NaN
 * (reference returns false and leaves the pose untouched). iters_out nullable.
 * standalone_variant: 0 = core error accounting (weighted y row adds w*ry^2, core :799),
 * 1 = standalone (adds ry^2, standalone :135); only the stopping test can differ. */
VO_API int vo_pose_gn_mono(vo_ctx *ctx, const float *X, const float *pts1, int n, float fx, float fy,
                    float cx, float cy, int thres_reproj_outlier, int standalone_variant,
                    float *R01_inout, float *t01_inout, uint8_t *mask_inlier, int *success,
                    int *iters_out);
VO_API int vo_pose_gn_stereo(vo_ctx *ctx, const float *X, const float *pts_l1, const float *pts_r1, int n,
                      const float *K_l4, const float *K_r4, const float *T_lr,
                      float thres_reproj_outlier, float *T01_inout, uint8_t *mask_inlier,
                      int *success, int *iters_out);
/* Accumulation mode of the pose-only GN (and of the frame steps that run it).
 * VO_POSE_FAST (default): residuals, weights, Jacobian rows and inlier masks per point in the reference's FP32 operation
 *   order; the row products of JtWJ / mJtWr are formed and summed in FP64 (fused multiply-add on the FP32 factors, fixed
 *   reduction tree) and rounded once.  More accurate than the reference's sums, but the reference's stop test
 *   `delta_err < 1e-7` (motion_estimator.cpp:1044,1063) fires when ITS sequential FP32 error sum repeats bit for bit, so
 *   the stop iteration can differ by one or two and the poses then differ by the size of the last updates.
 * VO_POSE_STRICT: the 28 sums are accumulated sequentially in FP32 in point order, exactly like
 *   motion_estimator.cpp:720-810 / :925-1040 -- same iterates, same stop iteration, same pose as the reference's
 *   arithmetic; costs the latency of a 4N-link dependent add chain per iteration (about 16 us at N = 2000).
 * VO_POSE_NO_EARLY_STOP (only the _ex entry points): ignore the stop test and run max_iter iterations. */
#define VO_POSE_FAST 0
#define VO_POSE_STRICT 1
#define VO_POSE_NO_EARLY_STOP 2
VO_API int vo_set_pose_mode(vo_ctx *ctx, int flags);   /* VO_POSE_FAST or VO_POSE_STRICT; used by every non-_ex pose call */
VO_API int vo_get_pose_mode(const vo_ctx *ctx);
/* _ex: explicit flags (VO_POSE_STRICT | VO_POSE_NO_EARLY_STOP), max_iter (0 = the reference's 100) and an optional
 * per-iteration trace [max_iter][24] = {T10 after the update (16, row-major), err_curr, delta_xi (6), delta_err}
 * (rows past the executed iterations are zero). */
VO_API int vo_pose_gn_mono_ex(vo_ctx *ctx, const float *X, const float *pts1, int n, float fx, float fy,
                       float cx, float cy, int thres_reproj_outlier, int standalone_variant,
                       float *R01_inout, float *t01_inout, uint8_t *mask_inlier, int *success,
                       int *iters_out, int flags, int max_iter, float *trace);
VO_API int vo_pose_gn_stereo_ex(vo_ctx *ctx, const float *X, const float *pts_l1, const float *pts_r1, int n,
                         const float *K_l4, const float *K_r4, const float *T_lr,
                         float thres_reproj_outlier, float *T01_inout, uint8_t *mask_inlier,
                         int *success, int *iters_out, int flags, int max_iter, float *trace);
VO_API int vo_pose_gn_stereo_batch_ex_d(vo_ctx *ctx, int n_prob, const int *offsets_d, const float *X_d,
                                 const float *pts_l1_d, const float *pts_r1_d, const float *K_l4,
                                 const float *K_r4, const float *T_lr, float thres_reproj_outlier,
                                 float *T01_inout_d, uint8_t *mask_inlier_d, int *success_d,
                                 int *iters_d, int flags, int max_iter, float *trace_d);
/* Batched device-resident variant: n_prob independent problems, problem p owns points
 * [offsets[p], offsets[p+1]).  T01_inout_d: [n_prob][16]. One CTA per problem. */
VO_API int vo_pose_gn_stereo_batch_d(vo_ctx *ctx, int n_prob, const int *offsets_d, const float *X_d,
                              const float *pts_l1_d, const float *pts_r1_d, const float *K_l4,
                              const float *K_r4, const float *T_lr, float thres_reproj_outlier,
                              float *T01_inout_d, uint8_t *mask_inlier_d, int *success_d,
                              int *iters_d);

/* ------------------------------------------------------------------ stereo tracking step
 * Steady-state part of StereoVO::trackStereoImages (core/visual_odometry/stereo_vo/stereo_vo.cpp:475-670,
 * steps [2]..[7]): constant-velocity prior, trackWithPrior(l0->l1), trackWithScale, trackWithPrior(l1->r1),
 * stereo pose-only GN on the triangulated survivors, the y > sampson_y stub, stable compaction.
 * Device resident: one H2D (two images + landmark state), one D2H (pose + survivors), one sync.
 * slot_l0 must already hold the previous left image (its pyramid stays cached between frames).
 * img_l1 / img_r1: host CV_8UC1 images (NULL = the slot already holds the image).
 * Xw: landmark 3-D points in the world frame; triangulated[i] = Landmark::isTriangulated().
 * Outputs: T_wc = T_wp * dT_pc (4x4 row-major), dT_pc (the GN result, next frame's dT_pc_prev),
 * index_out[k] = original index of the k-th surviving landmark (stable order == the reference's
 * StereoLandmarkTracking compactions), its tracked pixels, counts_out[5] (nullable) = survivors after
 * l0->l1, scale refinement, l1->r1, points fed to the GN, final.
 * Returns VO_ERR_NAN with the reference's text if the GN fails ("PoseOnlyStereoBA is failed!"). */
typedef struct vo_stereo_step_params {
    int window_size, max_level;      /* feature_tracker.window_size / max_level */
    float thres_error;               /* feature_tracker.thres_error */
    float thres_poseba_error;        /* motion_estimator.thres_poseba_error */
    float K_l[4], K_r[4];            /* fx, fy, cx, cy */
    float T_lr[16];                  /* row-major */
    int do_scale_refine;             /* 1 = run trackWithScale (stereo_vo.cpp:553) */
    float sampson_y;                 /* 660 (stereo_vo.cpp:663) */
} vo_stereo_step_params;
VO_API int vo_stereo_track_step(vo_ctx *ctx, const vo_stereo_step_params *prm, int slot_l0, int slot_l1,
                         int slot_r1, const uint8_t *img_l1, const uint8_t *img_r1, int w, int h,
                         size_t step, int n, const float *pts_l0, const float *pts_r0, const float *Xw,
                         const uint8_t *triangulated, const float *T_wp, const float *dT_pc_prev,
                         float *T_wc_out, float *dT_pc_out, int *n_out, int *index_out,
                         float *pts_l1_out, float *pts_r1_out, int *counts_out);

/* ------------------------------------------------------------------ feature extraction (bucketed)
 * FeatureExtractor::updateWeightBin(pts_occupied) + extractORBwithBinning_fast(img)
 * (core/visual_odometry/feature_extractor.cpp:94-98, 211-282; WeightBin feature_extractor.h:56-134): bins that hold
 * an occupied point are skipped, every other bin returns its best-response keypoint, in bin-index order.
 * The response is NOT cv::ORB's (third-party): it is the exact-integer Harris response on the slot's resident
 * Scharr plane defined in DESIGN.md ("K-det"); edge = border without detections (ORB edgeThreshold, 31). */
VO_API int vo_detect_bucketed(vo_ctx *ctx, int slot, const float *pts_occupied, int n_occupied, int n_bins_u,
                       int n_bins_v, int edge, long long min_score, float *pts_out, int max_out, int *n_out);

/* Keypoint detector behind vo_detect_bucketed and both frame steps.
 * VO_DETECTOR_HARRIS_SCHARR (default): K-det, exact-integer Harris on the resident Scharr plane, non-occupied bins only.
 * VO_DETECTOR_ORB: the reference's extractor, cv::ORB::detect as configured at feature_extractor.cpp:26-60 (10000 features,
 * scale 1.2, 8 levels, edge 31, HARRIS_SCORE, FAST threshold = feature_extractor.thres_fastscore), restated stage by stage
 * (INTER_LINEAR_EXACT pyramid, FAST-9/16 + non-maximum suppression, retainBest on the FAST score, Harris responses,
 * retainBest per level) and pinned bit-exact against cv2.ORB 4.13 through its numpy twin. */
#define VO_DETECTOR_HARRIS_SCHARR 0
#define VO_DETECTOR_ORB 1
VO_API int vo_set_detector(vo_ctx *ctx, int kind, int fast_threshold);
/* The whole keypoint list of cv::ORB::detect on the slot's image (unordered): pts [max][2] in level-0 pixels, Harris
 * response, octave.  VO_ERR_INVALID_ARG if max_keypoints is too small. */
VO_API int vo_orb_detect(vo_ctx *ctx, int slot, int fast_threshold, int edge, int max_keypoints, float *pts, float *response,
                  int *octave, int *n_out);
/* Test hook: plane of the last K-orb run (0 pyramid level, 1 FAST score, 2 score after non-maximum suppression; the last two
 * are defined inside the evaluated border band only).  dst nullable (size query), dense w x h. */
VO_API int vo_orb_read_level(vo_ctx *ctx, int level, int plane, uint8_t *dst, int *w, int *h);

/* ------------------------------------------------------------------ stereo frame step (tracking + new features)
 * vo_stereo_track_step followed, on the same stream and before the single synchronisation, by step [10] of
 * StereoVO::trackStereoImages (stereo_vo.cpp:690-740): bucketed detection on the current left image with the
 * survivors as occupancy, trackBidirection(l1 -> r1) of the detected points (feature_tracker.cpp:39-86), DLT of
 * every match (triangulate_3d.cpp:91-130) and the `both depths > 0` gate, stable compaction.
 * n == 0 with new_depth_gate == 0 is the first frame (stereo_vo.cpp:850-883): detection + bidirectional match only
 * (slot_l0, T_wp, dT_pc_prev may then be -1 / NULL).  n_bins_u * n_bins_v == 0 disables the new-feature stage. */
typedef struct vo_stereo_frame_params {
    vo_stereo_step_params track;
    float thres_bidirection;         /* feature_tracker.thres_bidirection */
    int n_bins_u, n_bins_v;          /* feature_extractor.n_bins_u / n_bins_v */
    int det_edge;                    /* 31 */
    long long det_min_score;         /* candidates need score > det_min_score */
    int new_depth_gate;              /* 1: keep a new feature only if both DLT depths are > 0 (stereo_vo.cpp:725) */
} vo_stereo_frame_params;
typedef struct vo_stereo_frame_result {
    float *T_wc, *dT_pc;             /* [16] row-major each */
    int n_tracked;                   /* survivors of the tracking step */
    int *index;                      /* [n] original index of every survivor */
    float *pts_l1, *pts_r1;          /* [n][2] */
    int *counts;                     /* [5] nullable, as vo_stereo_track_step */
    int n_detected;                  /* points returned by the detector */
    int n_new;                       /* new features that passed the bidirectional match (and the depth gate) */
    float *new_l1, *new_r1;          /* [n_bins_u * n_bins_v][2] */
} vo_stereo_frame_result;
VO_API int vo_stereo_frame_step(vo_ctx *ctx, const vo_stereo_frame_params *prm, int slot_l0, int slot_l1, int slot_r1,
                         const uint8_t *img_l1, const uint8_t *img_r1, int w, int h, size_t step, int n,
                         const float *pts_l0, const float *pts_r0, const float *Xw, const uint8_t *triangulated,
                         const float *T_wp, const float *dT_pc_prev, vo_stereo_frame_result *res);
/* Keyframe / first-frame reconstruction (stereo_vo.cpp:767-797, 911-941): DLT of every stereo match, 1-px^2
 * reprojection gates in both images (Camera::projectToPixel, camera.cpp:208-213), both depths > 0;
 * Xw_out = T_wc * X_left for every point, ok_out = 1 where all gates pass. */
VO_API int vo_stereo_reconstruct(vo_ctx *ctx, const float *pts_l, const float *pts_r, int n, const float *K_l4,
                          const float *K_r4, const float *T_lr, const float *T_wc, float *Xw_out, uint8_t *ok_out);

/* ------------------------------------------------------------------ mono frame step
 * Steady-state branch of MonoVO::trackImage (core/visual_odometry/mono_vo/mono_vo.cpp:724-992), device resident, one
 * synchronisation: constant-velocity prior and patch scale for the bundled landmarks (:739-761),
 * trackBidirectionWithPrior(I0 -> I1) (:768), trackWithScale (:783), selection of the pose-only-BA landmarks (bundled
 * when use_bundled_only, i.e. more than 5 keyframes, else triangulated; depth in the previous frame > 0.1, :799-827),
 * mono pose-only GN from dT01_prior (:860-866), inlier scatter, dT10 = inverseSE3_f(dT01), T_wc = T_wc_prev * dT01,
 * Sampson gate with F10 = Kinv^T [t10]x R10 Kinv (motion_estimator.cpp:539-568; mono_vo.cpp:957-962), stable
 * compaction; then bucketed detection on I1 with the survivors as occupancy and trackBidirection(I1 -> I0) of the new
 * points (:981-992).  flags[i]: bit 0 = Landmark::isTriangulated(), bit 1 = Landmark::isBundled().
 * The reference's fallback to cv::findEssentialMat when fewer than 11 landmarks are selected or the GN fails
 * (:909-949) runs vo_pose_5point's device stage on the K7 survivors and scales the unit translation to the length of
 * the previous motion; with thres_5p <= 0 the fallback is disabled and the call returns VO_ERR_MODE instead. */
typedef struct vo_mono_frame_params {
    int window_size, max_level;      /* feature_tracker.window_size / max_level */
    float thres_error, thres_bidirection, thres_sampson;
    float thres_poseba_error;        /* truncated to int as the reference's `const int &` parameter does */
    float K[4];                      /* fx, fy, cx, cy */
    int use_bundled_only;            /* keyframes_->getList().size() > 5 (mono_vo.cpp:800) */
    int do_scale_refine;
    int n_bins_u, n_bins_v, det_edge;
    long long det_min_score;
    float thres_5p;                  /* motion_estimator.thres_5p_error; <= 0 disables the five-point fallback (VO_ERR_MODE) */
    int n_hypotheses;                /* five-point RANSAC hypotheses, 0 = 1024 */
    unsigned seed;                   /* five-point sampling seed */
    int init_mode;                   /* 1: the second image of a sequence (mono_vo.cpp:562-659): K1 track + five-point, |t10| = 1;
                                        Xw / flags / dT01_prior are not read */
} vo_mono_frame_params;
typedef struct vo_mono_frame_result {
    float *T_wc, *dT01, *dT10;       /* [16] row-major each */
    int n_tracked;
    int *index;                      /* [n] */
    float *pts1;                     /* [n][2] */
    int *counts;                     /* [5] nullable: after K4, after K7, GN points, after the motion gate, final */
    int n_detected, n_new;
    float *new_p1, *new_p0;          /* [n_bins_u * n_bins_v][2]: new points in I1 and their back-tracked position in I0 */
    int used_5point;                 /* out: the pose came from calcPose5PointsAlgorithm (init, or the fallback of :909-949) */
    int n_5p_ransac;                 /* out: RANSAC inliers of that model */
} vo_mono_frame_result;
VO_API int vo_mono_frame_step(vo_ctx *ctx, const vo_mono_frame_params *prm, int slot_0, int slot_1, const uint8_t *img_1,
                       int w, int h, size_t step, int n, const float *pts0, const float *Xw, const uint8_t *flags,
                       const float *T_wc_prev, const float *dT01_prior, vo_mono_frame_result *res);

/* ------------------------------------------------------------------ five-point relative pose
 * MotionEstimator::calcPose5PointsAlgorithm (core/visual_odometry/motion_estimator.cpp:21-123): essential matrix by
 * five-point RANSAC (the reference calls cv::findEssentialMat(pts0, pts1, K, RANSAC, 0.999, thres_5p), :41), its SVD
 * decomposition into four (R10, t10) candidates (:70-96) and findCorrectRT (:205-263): the candidate with most
 * DLT-triangulated points in front of both cameras.  All n_hypotheses (0 = 1024) minimal samples are solved and scored
 * in parallel; sampling is a counter-based hash of (seed, hypothesis), so a call is reproducible.
 * R10[9] row-major, t10[3] (unit norm), X0 [n][3] nullable (points of the chosen candidate in camera 0),
 * mask[n] = RANSAC inlier AND cheirality (:115), E[9] nullable (row-major, Frobenius norm 1),
 * info[3] nullable = {RANSAC inliers, cheirality inliers, ok}.  VO_ERR_MODE when no model was found
 * ("calcPose5PointsAlgorithm() is failed.", mono_vo.cpp:590). */
VO_API int vo_pose_5point(vo_ctx *ctx, const float *pts0, const float *pts1, int n, const float *K4, float thres_px,
                   int n_hypotheses, unsigned seed, float *R10, float *t10, float *X0, uint8_t *mask, float *E, int *info);
/* The minimal solver alone (Nister): q [n_sets][5][4] = normalised (x0, y0, x1, y1) with x1^T E x0 = 0;
 * E [n_sets][10][9] row-major unit-norm candidates, n_solutions [n_sets]. */
VO_API int vo_five_point_minimal(vo_ctx *ctx, const double *q, int n_sets, double *E, int *n_solutions);

/* ------------------------------------------------------------------ epipolar distances / 1-point voting
 * MotionEstimator::calcSampsonDistance (core/visual_odometry/motion_estimator.cpp:539-570; the F10 overload :572-600),
 * calcSymmetricEpipolarDistance (:621-653): per-correspondence distances for the motion X1 = R10 X0 + t10, pixels in,
 * F10 = Kinv^T skew(t10) R10 Kinv.  R10 [9] row-major, t10 [3], K4 = fx, fy, cx, cy, dist [n]. */
VO_API int vo_sampson_distance(vo_ctx *ctx, const float *pts0, const float *pts1, int n, const float *K4, const float *R10,
                        const float *t10, float *dist);
VO_API int vo_sampson_distance_F(vo_ctx *ctx, const float *pts0, const float *pts1, int n, const float *F10, float *dist);
VO_API int vo_symmetric_epipolar_distance(vo_ctx *ctx, const float *pts0, const float *pts1, int n, const float *K4,
                                   const float *R10, const float *t10, float *dist);
/* MotionEstimator::findInliers1PointHistogram (:471-537): planar one-point motion voting -- theta_i = -2 atan((x0 y1 - y0 x1) /
 * (y0 + y1)) on normalised coordinates, 400-bin histogram over [-0.5, 0.5] rad (core/util/histogram.h:11-35), the centre
 * of the fullest bin (first one on ties), mask_i = symmetric epipolar distance for that motion <= thres_1p^2.
 * theta_opt [1]; R10 [9], t10 [3], theta [n] nullable. */
VO_API int vo_inliers_1point_histogram(vo_ctx *ctx, const float *pts0, const float *pts1, int n, const float *K4, float thres_1p,
                                uint8_t *mask, float *theta_opt, float *R10, float *t10, float *theta);

/* ------------------------------------------------------------------ triangulation
 * mapping::triangulateDLT (core/util/triangulate_3d.cpp:5-130). K0/K1: fx,fy,cx,cy. */
VO_API int vo_triangulate_dlt(vo_ctx *ctx, const float *pts0, const float *pts1, int n, const float *R10,
                       const float *t10, const float *K0_4, const float *K1_4, float *X0, float *X1);
/* The same for points that do not share one relative pose: group[i] in [0, n_groups) selects R10s[group[i]] (9 floats,
 * row-major) / t10s[group[i]] (3 floats).  MonoVO's reconstructions (mono_vo.cpp:660-687, :1032-1076) call triangulateDLT
 * once per landmark with the pose of the frame of its first observation; this entry does the keyframe's batch in one launch. */
VO_API int vo_triangulate_dlt_grouped(vo_ctx *ctx, const float *pts0, const float *pts1, int n, const int *group, int n_groups,
                               const float *R10s, const float *t10s, const float *K0_4, const float *K1_4, float *X0, float *X1);

/* ------------------------------------------------------------------ depth filter
 * DepthFilter::updateNormalDistribution (standalone/depth_filter/depth_filter.cpp:3-13). */
VO_API int vo_depth_filter_normal(vo_ctx *ctx, const double *x_prev, const double *cov_prev,
                           const double *x_curr, const double *cov_curr, int n, double *x_upd,
                           double *cov_upd);
/* DepthFilter::updateStudentTDistribution (depth_filter.cpp:15-46; the reference body does
 * not compile -- semantics defined in DESIGN.md "D2").  a/b/x_min/x_max are in-out. */
VO_API int vo_depth_filter_student_t(vo_ctx *ctx, const double *x_prev, const double *cov_prev,
                              double *a_inout, double *b_inout, double *x_min_inout,
                              double *x_max_inout, const double *x_curr, const double *cov_curr,
                              int n, double *x_upd, double *cov_upd);

/* Device-resident forms of the two seed updates (all pointers are device pointers, asynchronous on the context's stream):
 * the seed state stays in HBM between frames; x_upd_d / cov_upd_d may alias x_prev_d / cov_prev_d (in-place update). */
VO_API int vo_depth_filter_normal_d(vo_ctx *ctx, const double *x_prev_d, const double *cov_prev_d,
                             const double *x_curr_d, const double *cov_curr_d, int n, double *x_upd_d,
                             double *cov_upd_d);
VO_API int vo_depth_filter_student_t_d(vo_ctx *ctx, const double *x_prev_d, const double *cov_prev_d,
                                double *a_inout_d, double *b_inout_d, double *x_min_inout_d,
                                double *x_max_inout_d, const double *x_curr_d, const double *cov_curr_d,
                                int n, double *x_upd_d, double *cov_upd_d);

/* ------------------------------------------------------------------ compaction
 * Stable mask compaction == LandmarkTracking(src, mask) (landmark.cpp:194-231, 291-332):
 * index_out[k] = index of the k-th set mask entry; *n_out = number kept. */
VO_API int vo_compact(vo_ctx *ctx, const uint8_t *mask, int n, int *index_out, int *n_out);

/* ------------------------------------------------------------------ local bundle adjustment
 * SparseBundleAdjustmentSolver::solveForFiniteIterations
 * (ba_solver/sparse_bundle_adjustment.cpp:150-768) on the problem SparseBAParameters packs
 * (ba_solver/sparse_ba_parameters.h:292-465): poses/points already in the reference
 * keyframe's frame and divided by the pose scale. FP64 throughout. */
typedef struct vo_lba_problem {
    int n_frames;            /* left keyframes in the window (fixed + optimisable) */
    int n_opt;               /* optimisable left keyframes */
    int n_points;            /* M */
    int n_obs;               /* total observations */
    const double *poses;     /* [n_frames][16] row-major T_jw (ref frame, scaled) */
    const int *opt_index;    /* [n_frames] index into the optimised block, or -1 if fixed */
    const double *points;    /* [n_points][3] */
    const int *obs_ptr;      /* [n_points+1] CSR offsets, observation order == reference order */
    const int *obs_frame;    /* [n_obs] left-keyframe index the observation belongs to */
    const uint8_t *obs_right;/* [n_obs] 1 if observed in the right image */
    const double *obs_px;    /* [n_obs][2] */
    double K_l[4], K_r[4];   /* fx, fy, cx, cy */
    double T_lr[16];         /* row-major, translation already scaled; identity in mono */
    int is_stereo;
    double huber;            /* 0.5 (motion_estimator.cpp:1232) */
    double lambda;           /* 1e-5 (sparse_bundle_adjustment.cpp:185) */
    int max_iter;            /* 10 (motion_estimator.cpp:1226) */
} vo_lba_problem;
/* poses_out [n_frames][16], points_out [n_points][3]; avg_err_out[max_iter] per-iteration
 * sqrt(err/n_obs) (sparse_bundle_adjustment.cpp:606). *success = last avg_err <= 1 px.
 * Argument checks: sizes, opt_index and obs_ptr are validated before anything is enqueued; the per-observation checks
 * (obs_frame range, no duplicate (landmark, keyframe, camera), left observations in ascending keyframe order) run on
 * the host while the GPU already iterates -- the kernels clamp the frame index they read -- and a violation returns
 * VO_ERR_INVALID_ARG with the outputs untouched.  VO_ERR_NAN mirrors the reference's "Local BA NAN!" exception. */
VO_API int vo_lba_solve(vo_ctx *ctx, const vo_lba_problem *prob, double *poses_out, double *points_out,
                 double *avg_err_out, int *success);
/* Device scratch of the local BA ahead of time (grow-only; vo_lba_solve sizes it on demand otherwise, which costs the first
 * keyframe with a local BA a cudaMalloc of 32 MB: ~1 ms).  StereoVO / MonoVO call it in their constructors. */
VO_API int vo_lba_reserve(vo_ctx *ctx, size_t bytes);

/* ------------------------------------------------------------------ oversize windows on several GPUs (NCCL over NVLink)
 * BASELINE.json north_star / SURVEY 8(e): the LANDMARKS of one local-BA window are partitioned over the ranks (one process +
 * one vo_ctx per GPU), the keyframe poses are replicated.  Every rank accumulates the Schur-complement reduced camera system
 * of its own landmarks (what sparse_bundle_adjustment.cpp:456-515 accumulates over all of them); one
 * ncclAllReduce(sum, float64) of (6 N_opt)(6 N_opt + 1) + 27 N_opt + 2 values per LM iteration, enqueued on the context's
 * stream between the build and the dense solve (:517-536), completes it; the solve and the pose retraction run redundantly
 * and identically on every rank, each rank back-substitutes its own landmarks.
 *   vo_dist_unique_id : rank 0 creates the 128-byte NCCL id; the host program ships it to the other ranks (any transport).
 *   vo_dist_init      : collective; ncclCommInitRank on the context's device.  libnccl.so.2 is loaded at run time.
 *   vo_lba_solve_dist : collective; `local_part` = ALL frames / poses (identical on every rank) + this rank's landmarks and
 *                       observations.  poses_out: the full pose set (identical on every rank); points_out: this rank's
 *                       landmarks; avg_err_out: the GLOBAL per-iteration error. */
#define VO_DIST_UNIQUE_ID_BYTES 128
VO_API int vo_dist_unique_id(void *id_out);
VO_API int vo_dist_init(vo_ctx *ctx, int rank, int world, const void *unique_id);
VO_API int vo_dist_finalize(vo_ctx *ctx);
VO_API int vo_lba_solve_dist(vo_ctx *ctx, const vo_lba_problem *local_part, double *poses_out, double *points_out,
                      double *avg_err_out, int *success);

#ifdef __cplusplus
}
#endif
#endif /* VO_B200_H_ */
