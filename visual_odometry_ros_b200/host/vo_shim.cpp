// vo_shim.cpp -- marshalling between the reference's C++ containers and the C ABI. No arithmetic
// on data happens here (only 4x4 parameter algebra for the local-BA packing / write-back, which the
// reference also does on the host: sparse_ba_parameters.h:204-256, sparse_bundle_adjustment.cpp:631-718).
#include "vo_shim.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <map>

namespace vo_b200 {
static vo_ctx *g_ctx = nullptr;
static int g_w = 0, g_h = 0;
// image fingerprint -> slot 0..3 of g_ctx; lives and dies with g_ctx
struct SlotFp { const unsigned char *data = nullptr; int rows = 0, cols = 0; size_t step = 0; unsigned long long sum = 0; };
static SlotFp g_fp[4];
static int g_next_slot = 0;
static void reset_slot_cache() { for (auto &f : g_fp) f = SlotFp(); g_next_slot = 0; }

std::recursive_mutex &shared_mutex()
{
    static std::recursive_mutex m;
    return m;
}

[[noreturn]] void throw_status(vo_ctx *ctx, int status, const char *reference_message)
{
    std::string msg = reference_message ? reference_message : "";
    if (msg.empty()) {
        msg = std::string("vo_b200: ") + vo_status_string(status);
        if (ctx) msg += std::string(" (") + vo_last_error(ctx) + ")";
    }
    throw std::runtime_error(msg);
}

static int g_scale_faithful = 0;

// FeatureTracker::trackWithScale: reproduce the reference's never-reset sample buffers for samples that leave the image
// (vo_set_scale_mode; off by default: the affected features are then processed in list order, as in the reference)
void set_scale_faithful_borders(bool on)
{
    std::lock_guard<std::recursive_mutex> lock(shared_mutex());
    g_scale_faithful = on ? 1 : 0;
    if (g_ctx) vo_set_scale_mode(g_ctx, g_scale_faithful);
}

vo_ctx *shared_context(int min_w, int min_h)
{
    std::lock_guard<std::recursive_mutex> lock(shared_mutex());
    if (g_ctx && min_w <= g_w && min_h <= g_h) return g_ctx;
    if (g_ctx) { vo_ctx_destroy(g_ctx); g_ctx = nullptr; }
    reset_slot_cache();                      // the new context's slots hold no image
    g_w = std::max(min_w, std::max(g_w, 1920));
    g_h = std::max(min_h, std::max(g_h, 1200));
    const int rc = vo_ctx_create(0, g_w, g_h, 5, 8192, nullptr, &g_ctx);      // slots 0-3: FeatureTracker, slot 4: FeatureExtractor
    if (rc != VO_OK) throw_status(nullptr, rc, nullptr);
    vo_set_scale_mode(g_ctx, g_scale_faithful);
    return g_ctx;
}

void release_shared_context()
{
    std::lock_guard<std::recursive_mutex> lock(shared_mutex());
    if (g_ctx) vo_ctx_destroy(g_ctx);
    g_ctx = nullptr; g_w = g_h = 0;
    reset_slot_cache();
}

// slot holding `img` (uploading it if no slot does); never the slot `avoid` (the other image of the pair)
static int shared_slot_for(const cv::Mat &img, int avoid)
{
    if (img.empty()) throw std::runtime_error("vo_b200: empty image");
    unsigned long long sum = 1469598103934665603ull;
    for (int y = 0; y < img.rows; ++y) {
        const unsigned char *row = img.data + (size_t)y * img.step;
        unsigned long long acc = 0;
        int x = 0;
        for (; x + 8 <= img.cols; x += 8) { unsigned long long w; memcpy(&w, row + x, 8); acc += w * (unsigned long long)(x + 1); }
        for (; x < img.cols; ++x) acc += row[x];
        sum = (sum ^ acc) * 1099511628211ull;
    }
    vo_ctx *ctx = shared_context(img.cols, img.rows);    // may recreate the context and clear the cache
    for (int s = 0; s < 4; ++s)
        if (g_fp[s].data == img.data && g_fp[s].rows == img.rows && g_fp[s].cols == img.cols && g_fp[s].step == img.step && g_fp[s].sum == sum)
            return s;
    int s = g_next_slot;
    if (s == avoid) s = (s + 1) % 4;
    g_next_slot = (s + 1) % 4;
    g_fp[s] = SlotFp();                                   // invalid until the upload succeeded
    const int rc = vo_upload_image(ctx, s, img.data, img.cols, img.rows, img.step);
    if (rc) throw_status(ctx, rc, nullptr);
    g_fp[s].data = img.data; g_fp[s].rows = img.rows; g_fp[s].cols = img.cols; g_fp[s].step = img.step; g_fp[s].sum = sum;
    return s;
}
}  // namespace vo_b200

using vo_b200::shared_context;
using vo_b200::throw_status;

// ------------------------------------------------------------------------------ helpers
static std::vector<uint8_t> unpack_mask(MaskVec &mask, size_t n)
{
    mask.resize(n, true);   // keeps pre-existing entries (feature_tracker.cpp:20,49,98,178,305)
    std::vector<uint8_t> m(n);
    for (size_t i = 0; i < n; ++i) m[i] = mask[i] ? 1 : 0;
    return m;
}
static void pack_mask(const std::vector<uint8_t> &m, MaskVec &mask)
{
    mask.resize(m.size());
    for (size_t i = 0; i < m.size(); ++i) mask[i] = m[i] != 0;
}
static void pose_to_rowmajor(const PoseSE3 &T, float *o) { for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) o[r * 4 + c] = T(r, c); }
static void rowmajor_to_pose(const float *o, PoseSE3 &T) { for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T(r, c) = o[r * 4 + c]; }
static const float *pix(const PixelVec &v) { return reinterpret_cast<const float *>(v.data()); }
static float *pix(PixelVec &v) { return reinterpret_cast<float *>(v.data()); }
static_assert(sizeof(Pixel) == 8, "cv::Point2f must be two packed floats");
static_assert(sizeof(Point) == 12, "Eigen::Vector3f must be three packed floats");

// ------------------------------------------------------------------------------ FeatureTracker
FeatureTracker::FeatureTracker() { printf(" - FEATURE_TRACKER is constructed.\n"); }
FeatureTracker::~FeatureTracker() { printf(" - FEATURE_TRACKER is deleted.\n"); }

int FeatureTracker::slotFor(const cv::Mat &img, int avoid) { return vo_b200::shared_slot_for(img, avoid); }

#define VO_SHIM_LOCK std::lock_guard<std::recursive_mutex> vo_shim_lock_(vo_b200::shared_mutex())

void FeatureTracker::track(const cv::Mat &img0, const cv::Mat &img1, const PixelVec &pts0, int window_size, int max_pyr_lvl,
                           float thres_err, PixelVec &pts_track, MaskVec &mask_valid)
{
    const size_t n = pts0.size();
    std::vector<uint8_t> m = unpack_mask(mask_valid, n);
    pts_track.resize(n);
    if (n == 0) return;
    VO_SHIM_LOCK;
    const int s0 = slotFor(img0, -1), s1 = slotFor(img1, s0);
    vo_ctx *ctx = shared_context();
    const int rc = vo_ft_track(ctx, s0, s1, pix(pts0), (int)n, window_size, max_pyr_lvl, thres_err, pix(pts_track), m.data());
    if (rc) throw_status(ctx, rc, nullptr);
    pack_mask(m, mask_valid);
}

void FeatureTracker::trackBidirection(const cv::Mat &img0, const cv::Mat &img1, const PixelVec &pts0, int window_size, int max_pyr_lvl,
                                      float thres_err, float thres_bidirection, PixelVec &pts_track, MaskVec &mask_valid)
{
    const size_t n = pts0.size();
    std::vector<uint8_t> m = unpack_mask(mask_valid, n);
    pts_track.resize(n);
    if (n == 0) return;
    VO_SHIM_LOCK;
    const int s0 = slotFor(img0, -1), s1 = slotFor(img1, s0);
    vo_ctx *ctx = shared_context();
    const int rc = vo_ft_track_bidirection(ctx, s0, s1, pix(pts0), (int)n, window_size, max_pyr_lvl, thres_err, thres_bidirection,
                                           pix(pts_track), m.data());
    if (rc) throw_status(ctx, rc, nullptr);
    pack_mask(m, mask_valid);
}

void FeatureTracker::trackBidirectionWithPrior(const cv::Mat &img0, const cv::Mat &img1, const PixelVec &pts0, int window_size,
                                               int max_pyr_lvl, float thres_err, float thres_bidirection, PixelVec &pts_track,
                                               MaskVec &mask_valid)
{
    const size_t n = pts0.size();
    std::vector<uint8_t> m = unpack_mask(mask_valid, n);
    if (pts_track.size() != n) throw std::runtime_error("vo_b200: prior pts_track.size() != pts0.size()");
    if (n == 0) return;
    VO_SHIM_LOCK;
    const int s0 = slotFor(img0, -1), s1 = slotFor(img1, s0);
    vo_ctx *ctx = shared_context();
    const int rc = vo_ft_track_bidirection_with_prior(ctx, s0, s1, pix(pts0), (int)n, window_size, max_pyr_lvl, thres_err,
                                                      thres_bidirection, pix(pts_track), m.data());
    if (rc) throw_status(ctx, rc, nullptr);
    pack_mask(m, mask_valid);
}

void FeatureTracker::trackWithPrior(const cv::Mat &img0, const cv::Mat &img1, const PixelVec &pts0, int window_size, int max_pyr_lvl,
                                    float thres_err, PixelVec &pts_track, MaskVec &mask_valid)
{
    const size_t n = pts0.size();
    std::vector<uint8_t> m = unpack_mask(mask_valid, n);
    if (pts_track.size() != n) throw std::runtime_error("vo_b200: prior pts_track.size() != pts0.size()");
    if (n == 0) return;
    VO_SHIM_LOCK;
    const int s0 = slotFor(img0, -1), s1 = slotFor(img1, s0);
    vo_ctx *ctx = shared_context();
    const int rc = vo_ft_track_with_prior(ctx, s0, s1, pix(pts0), (int)n, window_size, max_pyr_lvl, thres_err, pix(pts_track), m.data());
    if (rc) throw_status(ctx, rc, nullptr);
    pack_mask(m, mask_valid);
}

void FeatureTracker::calcPrior(const PixelVec &pts0, const PointVec &Xw, const PoseSE3 &Tw1, const Eigen::Matrix3f &K, PixelVec &pts1_prior)
{
    const size_t n_pts = Xw.size();
    pts1_prior.resize(pts0.size());
    std::copy(pts0.begin(), pts0.end(), pts1_prior.begin());   // feature_tracker.cpp:212-213
    if (n_pts == 0) return;
    if (pts0.size() < n_pts) throw std::runtime_error("vo_b200: calcPrior pts0.size() < Xw.size()");
    float T[16];
    pose_to_rowmajor(Tw1, T);
    const float K4[4] = {K(0, 0), K(1, 1), K(0, 2), K(1, 2)};
    VO_SHIM_LOCK;
    vo_ctx *ctx = shared_context();
    std::vector<float> out(n_pts * 2);
    const int rc = vo_ft_calc_prior(ctx, pix(pts0), reinterpret_cast<const float *>(Xw.data()), (int)n_pts, T, K4, out.data());
    if (rc) throw_status(ctx, rc, nullptr);
    memcpy(static_cast<void *>(pts1_prior.data()), out.data(), n_pts * 8);
}

// cv::Sobel(img, CV_32FC1, dx, dy, 3) with BORDER_REFLECT_101 at one pixel (exact in float: integer taps on u8 pixels)
static float sobel3_at(const cv::Mat &img, int x, int y, bool horizontal)
{
    auto px = [&](int xx, int yy) -> float {
        if (xx < 0) xx = -xx;
        if (xx >= img.cols) xx = 2 * img.cols - 2 - xx;
        if (yy < 0) yy = -yy;
        if (yy >= img.rows) yy = 2 * img.rows - 2 - yy;
        return (float)img.data[(size_t)yy * img.step + xx];
    };
    if (horizontal)
        return (px(x + 1, y - 1) - px(x - 1, y - 1)) + 2.f * (px(x + 1, y) - px(x - 1, y)) + (px(x + 1, y + 1) - px(x - 1, y + 1));
    return (px(x - 1, y + 1) - px(x - 1, y - 1)) + 2.f * (px(x, y + 1) - px(x, y - 1)) + (px(x + 1, y + 1) - px(x + 1, y - 1));
}
static void require_sobel_of(const cv::Mat &img0, const cv::Mat &d, bool horizontal)
{
    const char *msg = "vo_b200: trackWithScale needs du0 / dv0 == cv::Sobel(img0, CV_32FC1, 1,0 / 0,1, 3) (the device evaluates those "
                      "taps from img0); other derivative images are not supported";
    if (d.rows != img0.rows || d.cols != img0.cols || d.type() != CV_32FC1) throw std::runtime_error(msg);
    if (img0.rows < 2 || img0.cols < 2) return;
    // 23 x 17 pixel sample incl. the four borders
    for (int iy = 0; iy < 17; ++iy) {
        const int y = (int)((long long)iy * (img0.rows - 1) / 16);
        const float *row = reinterpret_cast<const float *>(d.data + (size_t)y * d.step);
        for (int ix = 0; ix < 23; ++ix) {
            const int x = (int)((long long)ix * (img0.cols - 1) / 22);
            if (row[x] != sobel3_at(img0, x, y, horizontal)) throw std::runtime_error(msg);
        }
    }
}

void FeatureTracker::trackWithScale(const cv::Mat &img0, const cv::Mat &du0, const cv::Mat &dv0, const cv::Mat &img1,
                                    const PixelVec &pts0, const std::vector<float> &scale_est, PixelVec &pts_track, MaskVec &mask_valid)
{
    if (pts_track.size() != pts0.size()) throw std::runtime_error("pts_track.size() != pts0.size()");   // feature_tracker.cpp:283
    if (!img0.empty() && !du0.empty()) require_sobel_of(img0, du0, true);
    if (!img0.empty() && !dv0.empty()) require_sobel_of(img0, dv0, false);
    const size_t n = pts0.size();
    std::vector<uint8_t> m = unpack_mask(mask_valid, n);
    if (n == 0) return;
    if (scale_est.size() < n) throw std::runtime_error("vo_b200: scale_est.size() < pts0.size()");
    VO_SHIM_LOCK;
    const int s0 = slotFor(img0, -1), s1 = slotFor(img1, s0);
    vo_ctx *ctx = shared_context();
    const int rc = vo_ft_track_with_scale(ctx, s0, s1, pix(pts0), scale_est.data(), (int)n, pix(pts_track), m.data());
    if (rc == VO_ERR_NAN) throw std::runtime_error("ax ay nan");                                        // feature_tracker.cpp:414
    if (rc) throw_status(ctx, rc, nullptr);
    pack_mask(m, mask_valid);
}

// ------------------------------------------------------------------------------ FeatureExtractor
FeatureExtractor::FeatureExtractor() {}
FeatureExtractor::~FeatureExtractor() {}

void FeatureExtractor::initParams(int n_cols, int n_rows, int n_bins_u, int n_bins_v, int THRES_FAST, int /*radius*/)
{
    n_cols_ = n_cols; n_rows_ = n_rows; n_bins_u_ = n_bins_u; n_bins_v_ = n_bins_v; thres_fast_ = THRES_FAST;
    occupied_.clear();
}
void FeatureExtractor::resetWeightBin() { occupied_.clear(); }                           // WeightBin::reset (feature_extractor.cpp:62-64)
void FeatureExtractor::updateWeightBin(const PixelVec &pts) { occupied_ = pts; }         // reset + update (:94-98)

void FeatureExtractor::extractORBwithBinning_fast(const cv::Mat &img, PixelVec &pts_extracted, bool /*flag_nonmax*/)
{
    // the reference ignores the argument and buckets on its member flag_nonmax_ = true (:211-282)
    if (img.empty()) throw std::runtime_error("vo_b200: empty image");
    if (n_bins_u_ <= 0 || n_bins_v_ <= 0) throw std::runtime_error("vo_b200: FeatureExtractor::initParams has not been called");
    VO_SHIM_LOCK;
    vo_ctx *ctx = shared_context(img.cols, img.rows);
    int rc = vo_upload_image(ctx, 4, img.data, img.cols, img.rows, img.step);
    if (rc) throw_status(ctx, rc, nullptr);
    rc = vo_set_detector(ctx, VO_DETECTOR_ORB, thres_fast_);
    if (rc) throw_status(ctx, rc, nullptr);
    const int cap = n_bins_u_ * n_bins_v_;
    pts_extracted.resize(cap);
    int n = 0;
    rc = vo_detect_bucketed(ctx, 4, occupied_.empty() ? nullptr : pix(occupied_), (int)occupied_.size(), n_bins_u_, n_bins_v_, 31, 0,
                            pix(pts_extracted), cap, &n);
    if (rc) throw_status(ctx, rc, nullptr);
    pts_extracted.resize(n);
}

// ------------------------------------------------------------------------------ MotionEstimator
MotionEstimator::MotionEstimator(bool is_stereo_mode, const PoseSE3 &T_lr) : is_stereo_mode_(is_stereo_mode), T_lr_(T_lr) {}
MotionEstimator::~MotionEstimator() {}

bool MotionEstimator::monoImpl(const PointVec &X, const PixelVec &pts1, float fx, float fy, float cx, float cy, int thres, int standalone,
                               Rot3 &R01, Pos3 &t01, MaskVec &mask)
{
    if (X.size() != pts1.size()) throw std::runtime_error("In 'poseOnlyBundleAdjustment()': X.size() != pts1.size().");
    const size_t n = X.size();
    mask.resize(n);                                                                     // motion_estimator.cpp:675
    std::vector<uint8_t> m(n, 1);
    float R[9], t[3];
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R[r * 3 + c] = R01(r, c); t[r] = t01(r); }
    int ok = 0;
    VO_SHIM_LOCK;
    vo_ctx *ctx = shared_context();
    const int rc = vo_pose_gn_mono(ctx, reinterpret_cast<const float *>(X.data()), pix(pts1), (int)n, fx, fy, cx, cy, thres, standalone, R, t,
                                   m.data(), &ok, nullptr);
    if (rc) throw_status(ctx, rc, nullptr);
    if (ok) for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R01(r, c) = R[r * 3 + c]; t01(r) = t[r]; }
    if (n) pack_mask(m, mask);
    return ok != 0;
}

bool MotionEstimator::stereoImpl(const PointVec &X, const PixelVec &pl, const PixelVec &pr, const float *Kl, const float *Kr,
                                 const PoseSE3 &T_lr, float thres, PoseSE3 &T01, MaskVec &mask)
{
    if (!is_stereo_mode_) throw std::runtime_error("In 'poseOnlyBundleAdjustment_Stereo()', is_stereo_mode_ == false");
    if (X.size() != pl.size() || X.size() != pr.size())
        throw std::runtime_error("In 'poseOnlyStereoBundleAdjustment()': X.size() != pts_l1.size() || X.size() != pts_r1.size().");
    const size_t n = X.size();
    mask.assign(n, true);                                                               // motion_estimator.cpp:878
    std::vector<uint8_t> m(n, 1);
    float Tlr[16], T[16];
    pose_to_rowmajor(T_lr, Tlr);
    pose_to_rowmajor(T01, T);
    int ok = 0;
    VO_SHIM_LOCK;
    vo_ctx *ctx = shared_context();
    const int rc = vo_pose_gn_stereo(ctx, reinterpret_cast<const float *>(X.data()), pix(pl), pix(pr), (int)n, Kl, Kr, Tlr, thres, T, m.data(),
                                     &ok, nullptr);
    if (rc) throw_status(ctx, rc, nullptr);
    if (ok) rowmajor_to_pose(T, T01);
    if (n) pack_mask(m, mask);
    return ok != 0;
}

bool MotionEstimator::poseOnlyBundleAdjustment(const PointVec &X, const PixelVec &pts1, CameraConstPtr &cam, const int &thres,
                                               Rot3 &R01_true, Pos3 &t01_true, MaskVec &mask_inlier)
{
    return monoImpl(X, pts1, cam->fx(), cam->fy(), cam->cx(), cam->cy(), thres, 0, R01_true, t01_true, mask_inlier);
}
bool MotionEstimator::poseOnlyBundleAdjustment(const PointVec &X, const PixelVec &pts1, const float fx, const float fy, const float cx,
                                               const float cy, const int &thres, Rot3 &R01_true, Pos3 &t01_true, MaskVec &mask_inlier)
{
    return monoImpl(X, pts1, fx, fy, cx, cy, thres, 1, R01_true, t01_true, mask_inlier);
}
bool MotionEstimator::poseOnlyBundleAdjustment_Stereo(const PointVec &X, const PixelVec &pts_l1, const PixelVec &pts_r1,
                                                      CameraConstPtr &cam_left, CameraConstPtr &cam_right, const PoseSE3 &T_lr,
                                                      float thres, PoseSE3 &T01, MaskVec &mask_inlier)
{
    const float Kl[4] = {cam_left->fx(), cam_left->fy(), cam_left->cx(), cam_left->cy()};
    const float Kr[4] = {cam_right->fx(), cam_right->fy(), cam_right->cx(), cam_right->cy()};
    return stereoImpl(X, pts_l1, pts_r1, Kl, Kr, T_lr, thres, T01, mask_inlier);
}
bool MotionEstimator::poseOnlyBundleAdjustment_Stereo(const PointVec &X, const PixelVec &pts_l1, const PixelVec &pts_r1, const float fx_l,
                                                      const float fy_l, const float cx_l, const float cy_l, const float fx_r,
                                                      const float fy_r, const float cx_r, const float cy_r, const PoseSE3 &T_lr,
                                                      float thres, PoseSE3 &T01, MaskVec &mask_inlier)
{
    const float Kl[4] = {fx_l, fy_l, cx_l, cy_l}, Kr[4] = {fx_r, fy_r, cx_r, cy_r};
    return stereoImpl(X, pts_l1, pts_r1, Kl, Kr, T_lr, thres, T01, mask_inlier);
}

// ---- mono geometric front-end
bool MotionEstimator::calcPose5PointsAlgorithm(const PixelVec &pts0, const PixelVec &pts1, CameraConstPtr &cam, Rot3 &R10_true,
                                               Pos3 &t10_true, PointVec &X0_true, MaskVec &mask_inlier)
{
    if (pts0.size() != pts1.size()) throw std::runtime_error("calcPose5PointsAlgorithm(): pts0.size() != pts1.size()");        // :28
    if (pts0.size() == 0) throw std::runtime_error("calcPose5PointsAlgorithm(): pts0.size() == pts1.size() == 0");             // :33
    const size_t n = pts0.size();
    mask_inlier.resize(n, true);
    std::vector<uint8_t> m(n, 0);
    X0_true.resize(n);
    const float K4[4] = {cam->fx(), cam->fy(), cam->cx(), cam->cy()};
    float R[9], t[3];
    VO_SHIM_LOCK;
    vo_ctx *ctx = shared_context();
    const int rc = vo_pose_5point(ctx, pix(pts0), pix(pts1), (int)n, K4, thres_5p_, n_hypotheses_, seed_, R, t,
                                  reinterpret_cast<float *>(X0_true.data()), m.data(), nullptr, nullptr);
    if (rc == VO_ERR_MODE) return false;                       // no model (cv::findEssentialMat would return an empty matrix)
    if (rc) throw_status(ctx, rc, nullptr);
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R10_true(r, c) = R[r * 3 + c]; t10_true(r) = t[r]; }
    pack_mask(m, mask_inlier);
    return true;
}

float MotionEstimator::findInliers1PointHistogram(const PixelVec &pts0, const PixelVec &pts1, CameraConstPtr &cam, MaskVec &maskvec_inlier)
{
    if (pts0.size() != pts1.size()) throw std::runtime_error("Error in 'fineInliers1PointHistogram()': pts0.size() != pts1.size()");   // :477
    const size_t n = pts0.size();
    maskvec_inlier.resize(n, false);
    std::vector<uint8_t> m(std::max<size_t>(n, 1), 0);
    const float K4[4] = {cam->fx(), cam->fy(), cam->cx(), cam->cy()};
    float th = 0.f;
    VO_SHIM_LOCK;
    vo_ctx *ctx = shared_context();
    const int rc = vo_inliers_1point_histogram(ctx, pix(pts0), pix(pts1), (int)n, K4, thres_1p_, m.data(), &th, nullptr, nullptr, nullptr);
    if (rc) throw_status(ctx, rc, nullptr);
    m.resize(n);
    if (n) pack_mask(m, maskvec_inlier);
    return th;
}

static void epi_impl(int which, const PixelVec &pts0, const PixelVec &pts1, const float *K4, const Rot3 *R10, const Pos3 *t10, const Mat33 *F10,
                     std::vector<float> &dist, const char *msg)
{
    if (pts0.size() != pts1.size()) throw std::runtime_error(msg);
    const size_t n = pts0.size();
    dist.resize(n);
    if (n == 0) return;
    VO_SHIM_LOCK;
    vo_ctx *ctx = shared_context();
    int rc;
    if (F10) {
        float F[9];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F[r * 3 + c] = (*F10)(r, c);
        rc = vo_sampson_distance_F(ctx, pix(pts0), pix(pts1), (int)n, F, dist.data());
    } else {
        float R[9], t[3];
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R[r * 3 + c] = (*R10)(r, c); t[r] = (*t10)(r); }
        rc = which == 0 ? vo_sampson_distance(ctx, pix(pts0), pix(pts1), (int)n, K4, R, t, dist.data())
                        : vo_symmetric_epipolar_distance(ctx, pix(pts0), pix(pts1), (int)n, K4, R, t, dist.data());
    }
    if (rc) throw_status(ctx, rc, nullptr);
}

void MotionEstimator::calcSampsonDistance(const PixelVec &pts0, const PixelVec &pts1, CameraConstPtr &cam, const Rot3 &R10, const Pos3 &t10,
                                          std::vector<float> &sampson_dist)
{
    const float K4[4] = {cam->fx(), cam->fy(), cam->cx(), cam->cy()};
    epi_impl(0, pts0, pts1, K4, &R10, &t10, nullptr, sampson_dist, "Error in 'fineInliers1PointHistogram()': pts0.size() != pts1.size()");   // :543
}
void MotionEstimator::calcSampsonDistance(const PixelVec &pts0, const PixelVec &pts1, const Mat33 &F10, std::vector<float> &sampson_dist)
{
    epi_impl(0, pts0, pts1, nullptr, nullptr, nullptr, &F10, sampson_dist, "Error in 'fineInliers1PointHistogram()': pts0.size() != pts1.size()");   // :576
}
float MotionEstimator::calcSampsonDistance(const Pixel &pt0, const Pixel &pt1, const Mat33 &F10)
{
    // one correspondence (:602-619): not worth a launch -- the same float arithmetic on the host
    const float F[9] = {F10(0, 0), F10(0, 1), F10(0, 2), F10(1, 0), F10(1, 1), F10(1, 2), F10(2, 0), F10(2, 1), F10(2, 2)};
    const float a0 = (F[0] * pt0.x + F[1] * pt0.y) + F[2] * 1.0f, a1 = (F[3] * pt0.x + F[4] * pt0.y) + F[5] * 1.0f,
                a2 = (F[6] * pt0.x + F[7] * pt0.y) + F[8] * 1.0f;
    const float b0 = (F[0] * pt1.x + F[3] * pt1.y) + F[6] * 1.0f, b1 = (F[1] * pt1.x + F[4] * pt1.y) + F[7] * 1.0f;
    float num = (pt1.x * a0 + pt1.y * a1) + 1.0f * a2;
    num *= num;
    return num / (((a0 * a0 + a1 * a1) + b0 * b0) + b1 * b1);
}
void MotionEstimator::calcSymmetricEpipolarDistance(const PixelVec &pts0, const PixelVec &pts1, CameraConstPtr &cam, const Rot3 &R10,
                                                    const Pos3 &t10, std::vector<float> &sym_epi_dist)
{
    const float K4[4] = {cam->fx(), cam->fy(), cam->cx(), cam->cy()};
    epi_impl(1, pts0, pts1, K4, &R10, &t10, nullptr, sym_epi_dist, "In 'calcSymmetricEpipolarDistance()', pts0.size() != pts1.size()");   // :626
}

// ------------------------------------------------------------------------------ mapping::triangulateDLT
namespace mapping {
static void tri_impl(const float *p0, const float *p1, int n, const Rot3 &R10, const Pos3 &t10, const Camera &c0, const Camera &c1,
                     float *X0, float *X1)
{
    float R[9], t[3];
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R[r * 3 + c] = R10(r, c); t[r] = t10(r); }
    const float K0[4] = {c0.fx(), c0.fy(), c0.cx(), c0.cy()}, K1[4] = {c1.fx(), c1.fy(), c1.cx(), c1.cy()};
    VO_SHIM_LOCK;
    vo_ctx *ctx = shared_context();
    const int rc = vo_triangulate_dlt(ctx, p0, p1, n, R, t, K0, K1, X0, X1);
    if (rc) throw_status(ctx, rc, nullptr);
}
void triangulateDLT(const PixelVec &pts0, const PixelVec &pts1, const Rot3 &R10, const Pos3 &t10, CameraConstPtr &cam, PointVec &X0,
                    PointVec &X1)
{
    if (pts0.size() != pts1.size()) throw std::runtime_error("pts0.size() != pts1.size()");            // triangulate_3d.cpp:10
    X0.resize(pts0.size());
    X1.resize(pts0.size());
    if (pts0.empty()) return;
    tri_impl(pix(pts0), pix(pts1), (int)pts0.size(), R10, t10, *cam, *cam, reinterpret_cast<float *>(X0.data()),
             reinterpret_cast<float *>(X1.data()));
}
void triangulateDLT(const Pixel &pt0, const Pixel &pt1, const Rot3 &R10, const Pos3 &t10, CameraConstPtr &cam, Point &X0, Point &X1)
{
    tri_impl(&pt0.x, &pt1.x, 1, R10, t10, *cam, *cam, X0.data(), X1.data());
}
void triangulateDLT(const Pixel &pt0, const Pixel &pt1, const Rot3 &R10, const Pos3 &t10, CameraConstPtr &cam0, CameraConstPtr &cam1,
                    Point &X0, Point &X1)
{
    tri_impl(&pt0.x, &pt1.x, 1, R10, t10, *cam0, *cam1, X0.data(), X1.data());
}
Eigen::Matrix3f skew(const Eigen::Vector3f &v)
{
    Eigen::Matrix3f m;
    m(0, 1) = -v(2); m(0, 2) = v(1); m(1, 0) = v(2); m(1, 2) = -v(0); m(2, 0) = -v(1); m(2, 1) = v(0);
    return m;
}
}  // namespace mapping

// ------------------------------------------------------------------------------ DepthFilter
void DepthFilter::updateNormalDistribution(double x_prev, double cov_prev, double x_curr, double cov_curr, double &x_updated,
                                           double &cov_updated)
{
    VO_SHIM_LOCK;
    vo_ctx *ctx = shared_context();
    const int rc = vo_depth_filter_normal(ctx, &x_prev, &cov_prev, &x_curr, &cov_curr, 1, &x_updated, &cov_updated);
    if (rc) throw_status(ctx, rc, nullptr);
}
void DepthFilter::updateNormalDistribution(const std::vector<double> &x_prev, const std::vector<double> &cov_prev,
                                           const std::vector<double> &x_curr, const std::vector<double> &cov_curr,
                                           std::vector<double> &x_updated, std::vector<double> &cov_updated)
{
    const size_t n = x_prev.size();
    if (cov_prev.size() != n || x_curr.size() != n || cov_curr.size() != n) throw std::runtime_error("vo_b200: depth filter size mismatch");
    x_updated.resize(n); cov_updated.resize(n);
    if (!n) return;
    VO_SHIM_LOCK;
    vo_ctx *ctx = shared_context();
    const int rc = vo_depth_filter_normal(ctx, x_prev.data(), cov_prev.data(), x_curr.data(), cov_curr.data(), (int)n, x_updated.data(),
                                          cov_updated.data());
    if (rc) throw_status(ctx, rc, nullptr);
}
void DepthFilter::updateStudentTDistribution(double x_prev, double cov_prev, double a_prev, double b_prev, double x_min_prev,
                                             double x_max_prev, double x_curr, double cov_curr, double /*a_curr*/, double /*b_curr*/,
                                             double &x_updated, double &cov_updated, double &x_min_updated, double &x_max_updated)
{
    double a = a_prev, b = b_prev, lo = x_min_prev, hi = x_max_prev;
    VO_SHIM_LOCK;
    vo_ctx *ctx = shared_context();
    const int rc = vo_depth_filter_student_t(ctx, &x_prev, &cov_prev, &a, &b, &lo, &hi, &x_curr, &cov_curr, 1, &x_updated, &cov_updated);
    if (rc) throw_status(ctx, rc, nullptr);
    x_min_updated = lo; x_max_updated = hi;
}

#ifndef VO_SHIM_USE_REAL_HEADERS
// ------------------------------------------------------------------------------ 4x4 parameter algebra (double, row-major)
static void mul4(const double *A, const double *B, double *C)
{
    double T[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j]; T[i * 4 + j] = s; }
    memcpy(C, T, sizeof(T));
}
static void inv_se3(const double *T, double *O)
{
    double R[16] = {0};
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R[i * 4 + j] = T[j * 4 + i];
        double s = 0; for (int k = 0; k < 3; ++k) s += T[k * 4 + i] * T[k * 4 + 3];
        R[i * 4 + 3] = -s;
    }
    R[15] = 1.0;
    memcpy(O, R, sizeof(R));
}
static void pose_f2d(const PoseSE3 &T, double *o) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) o[r * 4 + c] = T(r, c); o[12] = o[13] = o[14] = 0; o[15] = 1; }

void Frame::setPose(const PoseSE3 &Twc)
{
    Twc_ = Twc;
    // Tcw = inverse of a rigid transform (frame.cpp:44-48 uses geometry::inverseSE3_f)
    PoseSE3 Ti;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Ti(i, j) = Twc(j, i);
        float s = 0.f; for (int k = 0; k < 3; ++k) s += Twc(k, i) * Twc(k, 3);
        Ti(i, 3) = -s;
    }
    Ti(3, 3) = 1.f;
    Tcw_ = Ti;
}

// ------------------------------------------------------------------------------ SparseBAParameters
SparseBAParameters::SparseBAParameters() : pose_scale_(10.0), inv_pose_scale_(1.0 / 10.0), N_(0), N_opt_(0), N_nonopt_(0), M_(0), n_obs_(0), is_stereo_mode_(false)
{
    for (int i = 0; i < 16; ++i) T_stereo[i] = (i % 5 == 0) ? 1.0 : 0.0;
}
SparseBAParameters::SparseBAParameters(bool is_stereo, const PoseSE3 &T_st) : pose_scale_(10.0), inv_pose_scale_(1.0 / 10.0), N_(0), N_opt_(0), N_nonopt_(0), M_(0), n_obs_(0), is_stereo_mode_(is_stereo)
{
    if (!is_stereo_mode_) throw std::runtime_error("'is_stereo_mode_' should be set to 'true' when T_stereo is given!");
    pose_f2d(T_st, T_stereo);
}

void SparseBAParameters::setPosesAndPoints(const FramePtrVec &frames, const std::vector<int> &idx_fix, const std::vector<int> &idx_optimize)
{
    if (is_stereo_mode_) for (int r = 0; r < 3; ++r) T_stereo[r * 4 + 3] *= inv_pose_scale_;          // :306-310
    N_ = (int)frames.size(); N_nonopt_ = (int)idx_fix.size(); N_opt_ = (int)idx_optimize.size();
    if (is_stereo_mode_) { if (N_ != (N_nonopt_ + N_opt_) * 2) throw std::runtime_error("In 'SparseBAParameters::setPosesAndPoints()', stereo mode is set, but N != 2*N_fix + 2*N_opt "); }
    else if (N_ != (N_nonopt_ + N_opt_)) throw std::runtime_error("In 'SparseBAParameters::setPosesAndPoints()', monocular mode is set, but N != N_fix + N_opt ");
    // 1) window keyframes and their alive + triangulated landmarks (first-seen order; the reference's
    //    unordered_set order is address-hash dependent, Appendix B #10)
    std::map<const Frame *, int> window;          // frame -> position in `frames`
    for (int k = 0; k < N_; ++k) window[frames[k].get()] = k;
    std::vector<LandmarkPtr> lmset;
    std::map<const Landmark *, int> seen;
    for (const auto &f : frames)
        for (const auto &lm : f->getRelatedLandmarkPtr())
            if (lm->isTriangulated() && lm->isAlive() && !seen.count(lm.get())) { seen[lm.get()] = 1; lmset.push_back(lm); }
    // 1-1) reference pose = first window keyframe (:353-359)
    pose_f2d(frames.front()->getPose(), Twj_ref);
    inv_se3(Twj_ref, Tjw_ref);
    // left keyframes of the window, in window order
    left_frames.clear(); right_frames.clear();
    std::map<const Frame *, int> left_index;
    for (const auto &f : frames) if (!f->isRightImage()) { left_index[f.get()] = (int)left_frames.size(); left_frames.push_back(f); }
    // 2) landmarks with >= 2 window observations (:362-402)
    landmarks.clear(); points.clear(); obs_ptr.assign(1, 0); obs_frame.clear(); obs_right.clear(); obs_px.clear();
    std::map<const Frame *, int> frames_all;
    for (const auto &lm : lmset) {
        const Point &Xf = lm->get3DPoint();
        const double Xw[3] = {Xf(0), Xf(1), Xf(2)};
        double Xr[3];
        for (int r = 0; r < 3; ++r) Xr[r] = (Tjw_ref[r * 4] * Xw[0] + Tjw_ref[r * 4 + 1] * Xw[1] + Tjw_ref[r * 4 + 2] * Xw[2] + Tjw_ref[r * 4 + 3]) * inv_pose_scale_;
        std::vector<int> of; std::vector<uint8_t> orr; std::vector<double> opx; std::vector<const Frame *> kfs;
        const FramePtrVec &rel = lm->getRelatedKeyframePtr();
        const PixelVec &obs = lm->getObservationsOnKeyframes();
        for (size_t j = 0; j < rel.size(); ++j) {
            const FramePtr &kf = rel[j];
            if (!window.count(kf.get())) continue;
            const Frame *left = kf->isRightImage() ? kf->getLeftFramePtr().get() : kf.get();
            if (!left_index.count(left)) throw std::runtime_error("In 'getPose()', posemap_all_.find(frame) == posemap_all_.end()");
            of.push_back(left_index[left]); orr.push_back(kf->isRightImage() ? 1 : 0);
            opx.push_back(obs[j].x); opx.push_back(obs[j].y); kfs.push_back(kf.get());
        }
        if (of.size() < 2) continue;                                                                    // THRES_MINIMUM_SEEN
        landmarks.push_back(lm);
        points.insert(points.end(), Xr, Xr + 3);
        obs_frame.insert(obs_frame.end(), of.begin(), of.end());
        obs_right.insert(obs_right.end(), orr.begin(), orr.end());
        obs_px.insert(obs_px.end(), opx.begin(), opx.end());
        obs_ptr.push_back((int)obs_frame.size());
        for (const Frame *k : kfs) frames_all[k] = 1;
    }
    for (const auto &f : frames) if (f->isRightImage() && frames_all.count(f.get())) right_frames.push_back(f);
    // 3) sizes (:404-412)
    N_ = (int)frames_all.size();
    N_nonopt_ = is_stereo_mode_ ? (N_ - 2 * N_opt_) / 2 : N_ - N_opt_;
    M_ = (int)landmarks.size();
    // 4) left poses in the reference frame, translation scaled (:414-431)
    poses.assign(left_frames.size() * 16, 0.0);
    for (size_t k = 0; k < left_frames.size(); ++k) {
        double Tjw[16];
        pose_f2d(left_frames[k]->getPoseInv(), Tjw);
        mul4(Tjw, Twj_ref, Tjw);
        for (int r = 0; r < 3; ++r) Tjw[r * 4 + 3] *= inv_pose_scale_;
        memcpy(&poses[k * 16], Tjw, sizeof(Tjw));
    }
    // 5) optimisable index map (:433-444)
    opt_index.assign(left_frames.size(), -1);
    int cnt = 0;
    for (int j : idx_optimize) {
        if (j < 0 || j >= (int)frames.size()) throw std::runtime_error("vo_b200: idx_optimize out of range");
        if (!frames[j]->isRightImage()) opt_index[left_index[frames[j].get()]] = cnt++;
    }
    N_opt_ = cnt;
    n_obs_ = (int)obs_frame.size();
}

// ------------------------------------------------------------------------------ SparseBundleAdjustmentSolver
SparseBundleAdjustmentSolver::SparseBundleAdjustmentSolver(bool is_stereo) : is_stereo_mode_(is_stereo), thres_huber_(0) {}
void SparseBundleAdjustmentSolver::setBAParameters(const std::shared_ptr<SparseBAParameters> &ba_params)
{
    if (is_stereo_mode_ && !ba_params->isStereoMode())
        throw std::runtime_error("In SparseBundleAdjustmentSolver::setBAParameters(), 'ba_params' is not in stereo mode while 'is_stereo' of this module is set to 'true'.");
    ba_params_ = ba_params;
}
void SparseBundleAdjustmentSolver::setHuberThreshold(double t) { thres_huber_ = t; }
void SparseBundleAdjustmentSolver::setCamera(const CameraPtr &cam)
{
    if (is_stereo_mode_) throw std::runtime_error("In 'SparseBundleAdjustmentSolver::setCamera()': Before call this function, 'is_stereo' should be set to 'false'.");
    cams_.assign(1, cam);
}
void SparseBundleAdjustmentSolver::setStereoCameras(const CameraPtr &cam0, const CameraPtr &cam1)
{
    if (!is_stereo_mode_) throw std::runtime_error("In 'SparseBundleAdjustmentSolver::setStereoCameras()': Before call this function, 'is_stereo' should be set to 'true'.");
    cams_.clear(); cams_.push_back(cam0); cams_.push_back(cam1);
}
void SparseBundleAdjustmentSolver::reset() { ba_params_.reset(); cams_.clear(); thres_huber_ = 0; }

bool SparseBundleAdjustmentSolver::solveForFiniteIterations(int MAX_ITER)
{
    if (!ba_params_ || cams_.empty()) throw std::runtime_error("vo_b200: solver not configured");
    SparseBAParameters &P = *ba_params_;
    vo_lba_problem pr;
    memset(&pr, 0, sizeof(pr));
    pr.n_frames = (int)P.left_frames.size(); pr.n_opt = P.getNumOfOptimizeFrames(); pr.n_points = P.getNumOfOptimizeLandmarks();
    pr.n_obs = P.getNumOfObservations();
    pr.poses = P.poses.data(); pr.opt_index = P.opt_index.data(); pr.points = P.points.data(); pr.obs_ptr = P.obs_ptr.data();
    pr.obs_frame = P.obs_frame.data(); pr.obs_right = P.obs_right.data(); pr.obs_px = P.obs_px.data();
    const Camera &cl = *cams_[0], &cr = *cams_[is_stereo_mode_ ? 1 : 0];
    pr.K_l[0] = cl.fx(); pr.K_l[1] = cl.fy(); pr.K_l[2] = cl.cx(); pr.K_l[3] = cl.cy();
    pr.K_r[0] = cr.fx(); pr.K_r[1] = cr.fy(); pr.K_r[2] = cr.cx(); pr.K_r[3] = cr.cy();
    memcpy(pr.T_lr, P.T_stereo, sizeof(pr.T_lr));
    pr.is_stereo = is_stereo_mode_ ? 1 : 0; pr.huber = thres_huber_; pr.lambda = 0.00001; pr.max_iter = MAX_ITER;
    std::vector<double> poses_out(P.poses.size()), points_out(P.points.size() + 3), avg(MAX_ITER);
    int ok = 0;
    VO_SHIM_LOCK;
    vo_ctx *ctx = shared_context();
    const int rc = vo_lba_solve(ctx, &pr, poses_out.data(), points_out.data(), avg.data(), &ok);
    if (rc == VO_ERR_NAN) throw std::runtime_error("Local BA NAN!\n");                                 // sparse_bundle_adjustment.cpp:761
    if (rc) throw_status(ctx, rc, nullptr);
    // ---- write-back (sparse_bundle_adjustment.cpp:631-718)
    bool flag_large_update = false;
    for (size_t k = 0; k < P.left_frames.size(); ++k) {
        if (P.opt_index[k] < 0) continue;
        double Tjw[16], Twj0[16], dT[16];
        memcpy(Tjw, &poses_out[k * 16], sizeof(Tjw));
        for (int r = 0; r < 3; ++r) Tjw[r * 4 + 3] *= P.pose_scale_;                                   // recoverOriginalScalePose
        mul4(Tjw, P.Tjw_ref, Tjw);                                                                     // changeInvPoseRefToWorld
        pose_f2d(P.left_frames[k]->getPose(), Twj0);
        mul4(Twj0, Tjw, dT);
        if (std::sqrt(dT[3] * dT[3] + dT[7] * dT[7] + dT[11] * dT[11]) > 50) flag_large_update = true;
        PoseSE3 Tjw_f, Twj_f;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) Tjw_f(r, c) = (float)Tjw[r * 4 + c];
        Tjw_f(3, 3) = 1.f;
        for (int i = 0; i < 3; ++i) {                                                                  // inverseSE3_f
            for (int j = 0; j < 3; ++j) Twj_f(i, j) = Tjw_f(j, i);
            float s = 0.f; for (int q = 0; q < 3; ++q) s += Tjw_f(q, i) * Tjw_f(q, 3);
            Twj_f(i, 3) = -s;
        }
        Twj_f(3, 3) = 1.f;
        P.left_frames[k]->setPose(Twj_f);
    }
    if (is_stereo_mode_) {
        // :664-683 verbatim: f->setPose(f->getPoseInv() * T_lr_f) with the SCALED T_lr cast to float
        PoseSE3 Tlr_f;
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) Tlr_f(r, c) = (float)P.T_stereo[r * 4 + c];
        for (const auto &f : P.right_frames) {
            const PoseSE3 &Twj = f->getPoseInv();
            PoseSE3 Tn;
            for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { float s = 0.f; for (int q = 0; q < 4; ++q) s += Twj(r, q) * Tlr_f(q, c); Tn(r, c) = s; }
            f->setPose(Tn);
        }
    }
    for (size_t i = 0; i < P.landmarks.size(); ++i) {
        double X[3];
        for (int r = 0; r < 3; ++r) X[r] = points_out[3 * i + r] * P.pose_scale_;
        Point Xf;
        for (int r = 0; r < 3; ++r) Xf(r) = (float)(P.Twj_ref[r * 4] * X[0] + P.Twj_ref[r * 4 + 1] * X[1] + P.Twj_ref[r * 4 + 2] * X[2] + P.Twj_ref[r * 4 + 3]);
        P.landmarks[i]->set3DPoint(Xf);
        if (std::sqrt(Xf(0) * Xf(0) + Xf(1) * Xf(1) + Xf(2) * Xf(2)) <= 3000) P.landmarks[i]->setBundled();
        else P.landmarks[i]->setDead();
    }
    if (flag_large_update) throw std::runtime_error("large update!");                                   // :731
    return ok != 0;
}
#endif


#ifndef VO_SHIM_USE_REAL_HEADERS
// ------------------------------------------------------------------------------ local-BA drivers
// MotionEstimator::localBundleAdjustmentSparseSolver (motion_estimator.cpp:1090-1205)
bool MotionEstimator::localBundleAdjustmentSparseSolver(const std::shared_ptr<Keyframes> &kfs_window, CameraConstPtr &cam)
{
    constexpr int MAX_ITER = 10;
    constexpr double THRES_HUBER = 0.5;
    constexpr int NUM_MINIMUM_REQUIRED_KEYFRAMES = 3, NUM_FIX_KEYFRAMES_IN_WINDOW = 2;
    if (kfs_window->getCurrentNumOfKeyframes() < NUM_MINIMUM_REQUIRED_KEYFRAMES) return false;   // :1130-1134
    FramePtrVec frames;
    std::vector<int> idx_fix, idx_opt;
    for (const auto &kf : kfs_window->getList()) frames.push_back(kf);
    for (int j = 0; j < NUM_FIX_KEYFRAMES_IN_WINDOW; ++j) idx_fix.push_back(j);
    for (size_t j = NUM_FIX_KEYFRAMES_IN_WINDOW; j < frames.size(); ++j) idx_opt.push_back((int)j);
    auto ba_params = std::make_shared<SparseBAParameters>();
    ba_params->setPosesAndPoints(frames, idx_fix, idx_opt);
    SparseBundleAdjustmentSolver solver(false);                                                   // sparse_ba_solver_ (:18)
    solver.reset();
    solver.setCamera(cam);
    solver.setBAParameters(ba_params);
    solver.setHuberThreshold(THRES_HUBER);
    solver.solveForFiniteIterations(MAX_ITER);
    solver.reset();
    return true;
}

// MotionEstimator::localBundleAdjustmentSparseSolver_Stereo (motion_estimator.cpp:1207-1340)
bool MotionEstimator::localBundleAdjustmentSparseSolver_Stereo(const std::shared_ptr<StereoKeyframes> &stkfs_window, CameraConstPtr &cam_left,
                                                               CameraConstPtr &cam_right, const PoseSE3 &T_lr)
{
    if (!is_stereo_mode_)
        throw std::runtime_error("In 'MotionEstimator::localBundleAdjustmentSparseSolver_Stereo()', is_stereo_mode_ == false.");   // :1220
    constexpr int MAX_ITER = 10;
    constexpr double THRES_HUBER = 0.5;
    constexpr int NUM_MINIMUM_REQUIRED_KEYFRAMES = 3, NUM_FIX_KEYFRAMES_IN_WINDOW = 2;
    if (stkfs_window->getCurrentNumOfStereoKeyframes() < NUM_MINIMUM_REQUIRED_KEYFRAMES) return false;
    FramePtrVec frames_ba;                                     // all left frames, then all right frames (:1272-1276)
    std::vector<int> idx_fix, idx_opt;
    for (const auto &stkf : stkfs_window->getList()) frames_ba.push_back(stkf->getLeft());
    for (const auto &stkf : stkfs_window->getList()) frames_ba.push_back(stkf->getRight());
    for (int j = 0; j < NUM_FIX_KEYFRAMES_IN_WINDOW; ++j) idx_fix.push_back(j);
    for (size_t j = NUM_FIX_KEYFRAMES_IN_WINDOW; j < frames_ba.size(); ++j)
        if (!frames_ba[j]->isRightImage()) idx_opt.push_back((int)j);                            // left poses only (:1285-1293)
    auto ba_params = std::make_shared<SparseBAParameters>(is_stereo_mode_, T_lr);
    ba_params->setPosesAndPoints(frames_ba, idx_fix, idx_opt);
    SparseBundleAdjustmentSolver solver(true);
    solver.reset();
    solver.setStereoCameras(cam_left, cam_right);
    solver.setBAParameters(ba_params);
    solver.setHuberThreshold(THRES_HUBER);
    solver.solveForFiniteIterations(MAX_ITER);
    solver.reset();
    return true;
}
#endif
