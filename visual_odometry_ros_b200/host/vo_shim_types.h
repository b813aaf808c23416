// vo_shim_types.h -- the value types the reference's public API is written in
// (core/defines/define_type.h:15-64).
//
// On a ROS box define VO_SHIM_USE_REAL_HEADERS: the real Eigen3 / OpenCV 4 headers and the
// reference's own Camera / Frame / Landmark classes are used and the shim classes below are
// byte-for-byte drop-ins.  This container has neither Eigen nor OpenCV C++ headers (SURVEY
// Appendix A), so by default minimal layout-compatible stand-ins are provided: same member
// names, same memory layout (cv::Point2f = 2 floats, Eigen matrices column-major and dense),
// only the operations the shim needs.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef VO_SHIM_USE_REAL_HEADERS
#include <eigen3/Eigen/Dense>
#include "opencv4/opencv2/core.hpp"
#else
namespace cv {
struct Point2f {
    float x, y;
    Point2f() : x(0), y(0) {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};
enum { CV_8UC1_ = 0 };
#ifndef CV_8UC1
#define CV_8UC1 0
#define CV_32FC1 5
#endif
// Header over caller-owned pixels (like cv::Mat constructed from a ROS message buffer,
// ros2/visual_odometry/stereo_vo_ros2.cpp:96-99).  CV_8UC1, or CV_32FC1 through the float constructor
// (the Sobel derivative images of trackWithScale, stereo_vo.cpp:551-552).
struct Mat {
    int rows = 0, cols = 0;
    size_t step = 0;          // row pitch in BYTES, as in OpenCV
    unsigned char *data = nullptr;
    int type_ = CV_8UC1;
    Mat() {}
    Mat(int rows_, int cols_, unsigned char *data_, size_t step_ = 0) : rows(rows_), cols(cols_), step(step_ ? step_ : (size_t)cols_), data(data_) {}
    Mat(int rows_, int cols_, float *data_, size_t step_ = 0)
        : rows(rows_), cols(cols_), step(step_ ? step_ : (size_t)cols_ * sizeof(float)), data(reinterpret_cast<unsigned char *>(data_)), type_(CV_32FC1) {}
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
};
}  // namespace cv

namespace Eigen {
// Dense column-major fixed-size matrix with the accessors the shim uses.
template <typename T, int R, int C>
struct Matrix {
    T m[R * C];
    Matrix() { for (int i = 0; i < R * C; ++i) m[i] = T(0); }
    T &operator()(int r, int c) { return m[c * R + r]; }
    const T &operator()(int r, int c) const { return m[c * R + r]; }
    T &operator()(int i) { return m[i]; }
    const T &operator()(int i) const { return m[i]; }
    T *data() { return m; }
    const T *data() const { return m; }
    static Matrix Identity() { Matrix I; for (int i = 0; i < (R < C ? R : C); ++i) I(i, i) = T(1); return I; }
    static Matrix Zero() { return Matrix(); }
};
using Vector3f = Matrix<float, 3, 1>;
using Matrix3f = Matrix<float, 3, 3>;
using Matrix4f = Matrix<float, 4, 4>;
}  // namespace Eigen
#endif

// core/defines/define_type.h:15-48
using Pixel = cv::Point2f;
using Point = Eigen::Vector3f;
using PixelVec = std::vector<Pixel>;
using PointVec = std::vector<Point>;
using MaskVec = std::vector<bool>;
using Pos3 = Eigen::Vector3f;
using Rot3 = Eigen::Matrix3f;
using PoseSE3 = Eigen::Matrix4f;
using Mat33 = Eigen::Matrix3f;

#ifndef VO_SHIM_USE_REAL_HEADERS
// The subset of core/visual_odometry/camera.h the hot path reads (pinhole intrinsics).
class Camera {
public:
    Camera(float fx, float fy, float cx, float cy) : fx_(fx), fy_(fy), cx_(cx), cy_(cy) {}
    const float &fx() const { return fx_; }
    const float &fy() const { return fy_; }
    const float &cx() const { return cx_; }
    const float &cy() const { return cy_; }
    float fxinv() const { return 1.0f / fx_; }
    float fyinv() const { return 1.0f / fy_; }
    Eigen::Matrix3f K() const { Eigen::Matrix3f k; k(0, 0) = fx_; k(1, 1) = fy_; k(0, 2) = cx_; k(1, 2) = cy_; k(2, 2) = 1.f; return k; }
private:
    float fx_, fy_, cx_, cy_;
};
using CameraPtr = std::shared_ptr<Camera>;
using CameraConstPtr = const CameraPtr;

// The subset of Frame / Landmark (core/visual_odometry/frame.h, landmark.h) that the local-BA
// packing (sparse_ba_parameters.h:292-465) and write-back (sparse_bundle_adjustment.cpp:631-718) touch.
class Landmark;
class Frame;
using FramePtr = std::shared_ptr<Frame>;
using LandmarkPtr = std::shared_ptr<Landmark>;
using FramePtrVec = std::vector<FramePtr>;
using LandmarkPtrVec = std::vector<LandmarkPtr>;

class Frame {
public:
    explicit Frame(bool is_right = false) : is_right_(is_right) { Twc_ = PoseSE3::Identity(); Tcw_ = PoseSE3::Identity(); }
    void setPose(const PoseSE3 &Twc);           // frame.cpp:44-48 (defined in shim_geometry.cpp)
    const PoseSE3 &getPose() const { return Twc_; }
    const PoseSE3 &getPoseInv() const { return Tcw_; }
    bool isRightImage() const { return is_right_; }
    void setLeftFramePtr(const FramePtr &l) { left_ = l; }
    const FramePtr &getLeftFramePtr() const { return left_; }
    const LandmarkPtrVec &getRelatedLandmarkPtr() const { return lms_; }
    void addRelatedLandmark(const LandmarkPtr &lm) { lms_.push_back(lm); }
private:
    bool is_right_;
    PoseSE3 Twc_, Tcw_;
    FramePtr left_;
    LandmarkPtrVec lms_;
};

// core/visual_odometry/frame.h:96-113 / keyframes.h:24-88: the window containers the local-BA drivers walk
// (MotionEstimator::localBundleAdjustmentSparseSolver(_Stereo), motion_estimator.cpp:1090-1340).  Only the members those
// drivers touch; the keyframe RULE (keyframes.cpp:217-303) lives in StereoVO / MonoVO (host/stereo_vo.cpp, mono_vo.cpp).
class StereoFrame {
public:
    StereoFrame(const FramePtr &frame_left, const FramePtr &frame_right) : left_(frame_left), right_(frame_right) {}
    const FramePtr &getLeft() const { return left_; }
    const FramePtr &getRight() const { return right_; }
private:
    FramePtr left_, right_;
};
using StereoFramePtr = std::shared_ptr<StereoFrame>;

class Keyframes {
public:
    Keyframes() : n_max_(9) {}
    void setMaxKeyframes(int max_kf) { n_max_ = max_kf; }
    void addNewKeyframe(const FramePtr &frame)          // keyframes.cpp:122-140: drop the oldest beyond the window size
    {
        kfs_list_.push_back(frame);
        all_keyframes_.push_back(frame);
        if ((int)kfs_list_.size() > n_max_) kfs_list_.erase(kfs_list_.begin());
    }
    const std::vector<FramePtr> &getList() const { return kfs_list_; }
    int getCurrentNumOfKeyframes() const { return (int)kfs_list_.size(); }
    int getMaxNumOfKeyframes() const { return n_max_; }
private:
    std::vector<FramePtr> kfs_list_, all_keyframes_;
    int n_max_;
};

class StereoKeyframes {
public:
    StereoKeyframes() : n_max_(9) {}
    void setMaxStereoKeyframes(int max_kf) { n_max_ = max_kf; }
    void addNewStereoKeyframe(const StereoFramePtr &stframe)   // keyframes.cpp:305-323
    {
        list_.push_back(stframe);
        all_.push_back(stframe);
        if ((int)list_.size() > n_max_) list_.erase(list_.begin());
    }
    const std::vector<StereoFramePtr> &getList() const { return list_; }
    int getCurrentNumOfStereoKeyframes() const { return (int)list_.size(); }
    int getMaxNumOfStereoKeyframes() const { return n_max_; }
private:
    std::vector<StereoFramePtr> list_, all_;
    int n_max_;
};

class Landmark {
public:
    Landmark() : alive_(true), triangulated_(false), bundled_(false) {}
    void set3DPoint(const Point &X) { X_ = X; triangulated_ = true; }
    const Point &get3DPoint() const { return X_; }
    bool isTriangulated() const { return triangulated_; }
    bool isAlive() const { return alive_; }
    bool isBundled() const { return bundled_; }
    void setBundled() { bundled_ = true; }
    void setDead() { alive_ = false; }
    void addObservationOnKeyframe(const Pixel &p, const FramePtr &kf) { kf_obs_.push_back(p); kfs_.push_back(kf); }
    const FramePtrVec &getRelatedKeyframePtr() const { return kfs_; }
    const PixelVec &getObservationsOnKeyframes() const { return kf_obs_; }
private:
    Point X_;
    bool alive_, triangulated_, bundled_;
    FramePtrVec kfs_;
    PixelVec kf_obs_;
};
#endif
