// vo_shim.h -- the reference's public hot-path classes, re-implemented on top of the C ABI
// (include/vo_b200.h).  Signatures, argument meaning and error behaviour (std::runtime_error with the
// reference's message) follow the reference headers cited on each class, so StereoVO / MonoVO and the
// ROS nodes compile against them unchanged.  No arithmetic happens here: every method marshals
// std::vector<bool> / cv::Point2f / Eigen containers to flat arrays and calls libvo_b200.so.
#pragma once
#include "../../include/vo_b200.h"
#include "vo_shim_types.h"

#include <mutex>

namespace vo_b200 {
// One process-wide device context shared by the shim objects (the reference keeps one VO object per
// node and is single-threaded, SURVEY 8b "Threading").  Created on first use; throws
// std::runtime_error if no CUDA device is present (there is no CPU fallback).
// The shim objects share that context's image slots 0-3 through ONE process-wide cache (image fingerprint -> slot), so
// several FeatureTracker objects never see each other's uploads as their own; every shim method that touches the context
// holds shared_mutex() for its duration (the shim may be called from several threads; calls are serialised).
vo_ctx *shared_context(int min_w = 0, int min_h = 0);
void release_shared_context();
void set_scale_faithful_borders(bool on);   // trackWithScale: the reference's stale sample buffers (vo_set_scale_mode), default off
std::recursive_mutex &shared_mutex();
[[noreturn]] void throw_status(vo_ctx *ctx, int status, const char *reference_message);
}  // namespace vo_b200

// core/visual_odometry/feature_tracker.h:44-104
class FeatureTracker {
public:
    FeatureTracker();
    ~FeatureTracker();
    void track(const cv::Mat &img0, const cv::Mat &img1, const PixelVec &pts0, int window_size, int max_pyr_lvl, float thres_err,
               PixelVec &pts_track, MaskVec &mask_valid);
    void trackBidirection(const cv::Mat &img0, const cv::Mat &img1, const PixelVec &pts0, int window_size, int max_pyr_lvl,
                          float thres_err, float thres_bidirection, PixelVec &pts_track, MaskVec &mask_valid);
    void trackBidirectionWithPrior(const cv::Mat &img0, const cv::Mat &img1, const PixelVec &pts0, int window_size, int max_pyr_lvl,
                                   float thres_err, float thres_bidirection, PixelVec &pts_track, MaskVec &mask_valid);
    void trackWithPrior(const cv::Mat &img0, const cv::Mat &img1, const PixelVec &pts0, int window_size, int max_pyr_lvl,
                        float thres_err, PixelVec &pts_track, MaskVec &mask_valid);
    void calcPrior(const PixelVec &pts0, const PointVec &Xw, const PoseSE3 &Tw1, const Eigen::Matrix3f &K, PixelVec &pts1_prior);
    // du0 / dv0: the reference passes cv::Sobel(img0, CV_32FC1, 1,0 / 0,1, 3) (stereo_vo.cpp:551-553, mono_vo.cpp:781-783).
    // The kernel evaluates those taps on the fly from img0, so derivative images that are NOT the 3x3 Sobel of img0
    // cannot be honoured: non-empty du0 / dv0 are verified on a pixel sample and a std::runtime_error is thrown if they
    // are anything else.  Empty Mats skip the check.
    void trackWithScale(const cv::Mat &img0, const cv::Mat &du0, const cv::Mat &dv0, const cv::Mat &img1, const PixelVec &pts0,
                        const std::vector<float> &scale_est, PixelVec &pts_track, MaskVec &mask_valid);
private:
    // images are re-uploaded only when the (data pointer, size, content hash) fingerprint changes; the fingerprint table
    // belongs to the shared context (vo_shim.cpp), not to this object
    int slotFor(const cv::Mat &img, int avoid);
};

// core/visual_odometry/feature_extractor.h:144-200: the bucketed extractor the VO classes call (initParams, resetWeightBin,
// updateWeightBin, extractORBwithBinning_fast; feature_extractor.cpp:26-98, 211-282).  The keypoints are cv::ORB::detect's,
// restated on the device (K-orb, bit-exact with cv2.ORB).  The per-bucket variant extractORBwithBinning / the descriptor
// functions / suppressCenterBins have no caller on the VO path and are not provided.
class FeatureExtractor {
public:
    FeatureExtractor();
    ~FeatureExtractor();
    void initParams(int n_cols, int n_rows, int n_bins_u, int n_bins_v, int THRES_FAST, int radius);
    void updateWeightBin(const PixelVec &pts);
    void resetWeightBin();
    void extractORBwithBinning_fast(const cv::Mat &img, PixelVec &pts_extracted, bool flag_nonmax);
private:
    int n_cols_ = 0, n_rows_ = 0, n_bins_u_ = 0, n_bins_v_ = 0, thres_fast_ = 20;
    PixelVec occupied_;
};

// core/visual_odometry/motion_estimator.h:107-147 (pose-only part) and
// standalone/motion_estimator/motion_estimator.h:22-35 (scalar-intrinsics overloads)
class MotionEstimator {
public:
    MotionEstimator(bool is_stereo_mode = false, const PoseSE3 &T_lr = PoseSE3::Identity());
    ~MotionEstimator();
    bool poseOnlyBundleAdjustment(const PointVec &X, const PixelVec &pts1, CameraConstPtr &cam, const int &thres_reproj_outlier,
                                  Rot3 &R01_true, Pos3 &t01_true, MaskVec &mask_inlier);
    bool poseOnlyBundleAdjustment_Stereo(const PointVec &X, const PixelVec &pts_l1, const PixelVec &pts_r1, CameraConstPtr &cam_left,
                                         CameraConstPtr &cam_right, const PoseSE3 &T_lr, float thres_reproj_outlier, PoseSE3 &T01,
                                         MaskVec &mask_inlier);
    // standalone overloads
    bool poseOnlyBundleAdjustment(const PointVec &X, const PixelVec &pts1, const float fx, const float fy, const float cx, const float cy,
                                  const int &thres_reproj_outlier, Rot3 &R01_true, Pos3 &t01_true, MaskVec &mask_inlier);
    bool poseOnlyBundleAdjustment_Stereo(const PointVec &X, const PixelVec &pts_l1, const PixelVec &pts_r1, const float fx_l,
                                         const float fy_l, const float cx_l, const float cy_l, const float fx_r, const float fy_r,
                                         const float cx_r, const float cy_r, const PoseSE3 &T_lr, float thres_reproj_outlier,
                                         PoseSE3 &T01, MaskVec &mask_inlier);
    // mono geometric front-end (core motion_estimator.h:110-115, 137-143)
    bool calcPose5PointsAlgorithm(const PixelVec &pts0, const PixelVec &pts1, CameraConstPtr &cam, Rot3 &R10_true, Pos3 &t10_true,
                                  PointVec &X0_true, MaskVec &mask_inlier);
    float findInliers1PointHistogram(const PixelVec &pts0, const PixelVec &pts1, CameraConstPtr &cam, MaskVec &maskvec_inlier);
    void calcSampsonDistance(const PixelVec &pts0, const PixelVec &pts1, CameraConstPtr &cam, const Rot3 &R10, const Pos3 &t10,
                             std::vector<float> &sampson_dist);
    void calcSampsonDistance(const PixelVec &pts0, const PixelVec &pts1, const Mat33 &F10, std::vector<float> &sampson_dist);
    float calcSampsonDistance(const Pixel &pt0, const Pixel &pt1, const Mat33 &F10);
    void calcSymmetricEpipolarDistance(const PixelVec &pts0, const PixelVec &pts1, CameraConstPtr &cam, const Rot3 &R10, const Pos3 &t10,
                                       std::vector<float> &sym_epi_dist);
    // five-point RANSAC knobs of this build (the reference's cv::findEssentialMat draws its own samples)
    void setRansac(int n_hypotheses, unsigned seed) { n_hypotheses_ = n_hypotheses; seed_ = seed; }
#ifndef VO_SHIM_USE_REAL_HEADERS
    // core motion_estimator.h:122-123, bodies motion_estimator.cpp:1090-1205 / :1207-1340: window -> SparseBAParameters
    // (fixed = the two oldest keyframes, the others optimised; stereo: left poses only, right frames carry residuals),
    // Huber 0.5, 10 iterations of SparseBundleAdjustmentSolver.  false when the window holds fewer than 3 keyframes.
    bool localBundleAdjustmentSparseSolver(const std::shared_ptr<Keyframes> &kfs, CameraConstPtr &cam);
    bool localBundleAdjustmentSparseSolver_Stereo(const std::shared_ptr<StereoKeyframes> &stkfs_window, CameraConstPtr &cam_left,
                                                  CameraConstPtr &cam_right, const PoseSE3 &T_lr);
#endif
    void setThres1p(float thres_1p) { thres_1p_ = thres_1p; }
    void setThres5p(float thres_5p) { thres_5p_ = thres_5p; }
private:
    bool monoImpl(const PointVec &X, const PixelVec &pts1, float fx, float fy, float cx, float cy, int thres, int standalone,
                  Rot3 &R01, Pos3 &t01, MaskVec &mask);
    bool stereoImpl(const PointVec &X, const PixelVec &pl, const PixelVec &pr, const float *Kl, const float *Kr, const PoseSE3 &T_lr,
                    float thres, PoseSE3 &T01, MaskVec &mask);
    bool is_stereo_mode_;
    PoseSE3 T_lr_;
    float thres_1p_ = 10.0f, thres_5p_ = 1.5f;                                          // motion_estimator.cpp:6-7
    int n_hypotheses_ = 0;
    unsigned seed_ = 0;
};

// core/util/triangulate_3d.h:16-30
namespace mapping {
void triangulateDLT(const PixelVec &pts0, const PixelVec &pts1, const Rot3 &R10, const Pos3 &t10, CameraConstPtr &cam, PointVec &X0,
                    PointVec &X1);
void triangulateDLT(const Pixel &pt0, const Pixel &pt1, const Rot3 &R10, const Pos3 &t10, CameraConstPtr &cam, Point &X0, Point &X1);
void triangulateDLT(const Pixel &pt0, const Pixel &pt1, const Rot3 &R10, const Pos3 &t10, CameraConstPtr &cam0, CameraConstPtr &cam1,
                    Point &X0, Point &X1);
Eigen::Matrix3f skew(const Eigen::Vector3f &vec);
}  // namespace mapping

// standalone/depth_filter/depth_filter.h:13-21 (+ batched forms, the natural unit of work on a GPU)
class DepthFilter {
public:
    DepthFilter() {}
    ~DepthFilter() {}
    void updateNormalDistribution(double x_prev, double cov_prev, double x_curr, double cov_curr, double &x_updated, double &cov_updated);
    void updateStudentTDistribution(double x_prev, double cov_prev, double a_prev, double b_prev, double x_min_prev, double x_max_prev,
                                    double x_curr, double cov_curr, double a_curr, double b_curr, double &x_updated, double &cov_updated,
                                    double &x_min_updated, double &x_max_updated);
    void updateNormalDistribution(const std::vector<double> &x_prev, const std::vector<double> &cov_prev, const std::vector<double> &x_curr,
                                  const std::vector<double> &cov_curr, std::vector<double> &x_updated, std::vector<double> &cov_updated);
};

#ifndef VO_SHIM_USE_REAL_HEADERS
// ba_solver/sparse_ba_parameters.h:292-465 -- window -> flat problem
class SparseBAParameters {
public:
    SparseBAParameters();
    SparseBAParameters(bool is_stereo, const PoseSE3 &T_stereo);
    void setPosesAndPoints(const FramePtrVec &frames, const std::vector<int> &idx_fix, const std::vector<int> &idx_optimize);
    int getNumOfAllFrames() const { return N_; }
    int getNumOfOptimizeFrames() const { return N_opt_; }
    int getNumOfOptimizeLandmarks() const { return M_; }
    int getNumOfObservations() const { return n_obs_; }
    bool isStereoMode() const { return is_stereo_mode_; }
    // flat problem (layout of vo_lba_problem)
    std::vector<double> poses, points, obs_px;
    std::vector<int> opt_index, obs_ptr, obs_frame;
    std::vector<uint8_t> obs_right;
    std::vector<FramePtr> left_frames;     // index -> left keyframe
    std::vector<FramePtr> right_frames;    // right frames seen in the window
    std::vector<LandmarkPtr> landmarks;    // index -> landmark
    double T_stereo[16];                   // row-major, translation scaled
    double Twj_ref[16], Tjw_ref[16];
    double pose_scale_, inv_pose_scale_;
private:
    int N_, N_opt_, N_nonopt_, M_, n_obs_;
    bool is_stereo_mode_;
};

// ba_solver/sparse_bundle_adjustment.h:105-130
class SparseBundleAdjustmentSolver {
public:
    SparseBundleAdjustmentSolver(bool is_stereo = false);
    void setBAParameters(const std::shared_ptr<SparseBAParameters> &ba_params);
    void setHuberThreshold(double thres_huber);
    void setCamera(const CameraPtr &cam);
    void setStereoCameras(const CameraPtr &cam0, const CameraPtr &cam1);
    bool solveForFiniteIterations(int MAX_ITER);
    void reset();
private:
    bool is_stereo_mode_;
    double thres_huber_;
    std::vector<CameraPtr> cams_;
    std::shared_ptr<SparseBAParameters> ba_params_;
};
#endif
