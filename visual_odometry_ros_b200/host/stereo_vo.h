// stereo_vo.h -- StereoVO with the reference's public API (core/visual_odometry/stereo_vo/stereo_vo.h:233-249) on top
// of the C ABI.  The unchanged ROS nodes construct it with (mode, yaml directory), feed it image pairs and read
// getStatistics() / getDebugImage() (ros2/visual_odometry/stereo_vo_ros2.cpp:50,104,112;
// ros1/visual_odometry/stereo_vo_ros1.cpp:101,108,199).
//
// What moved: one device call per frame (vo_stereo_frame_step: both pyramids, prior, 2x trackWithPrior,
// trackWithScale, stereo pose-only GN, compactions, bucketed detection, bidirectional stereo match of the new
// features, depth gate -- one H2D, one D2H, one synchronisation), one more per keyframe (vo_stereo_reconstruct) and
// the local BA (vo_lba_solve).  What stays on the host, as in the reference: the landmark / frame / keyframe
// bookkeeping (landmark.cpp, frame.cpp, keyframes.cpp) and the window -> problem packing
// (sparse_ba_parameters.h:292-465) -- here over flat arrays instead of a shared_ptr graph with hash lookups.
// cv::imshow drawing, the /home/kch trajectory dump of the destructor and the console prints are not reproduced.
#pragma once
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "../../include/vo_b200.h"
#include "vo_shim_types.h"

class StereoVO {
public:
    // stereo_vo.h:106-186 (field-compatible with what the nodes read)
    struct AlgorithmStatistics {
        struct LandmarkStatistics {
            int n_initial = 0, n_pass_bidirection = 0, n_pass_1p = 0, n_pass_5p = 0, n_new = 0, n_final = 0;
            int max_age = 0, min_age = 0;
            float avg_age = 0.f;
            int n_ok_parallax = 0;
            float min_parallax = 0.f, max_parallax = 0.f, avg_parallax = 0.f;
        };
        struct FrameStatistics {
            PoseSE3 Twc = PoseSE3::Identity(), Tcw = PoseSE3::Identity(), dT_01 = PoseSE3::Identity(), dT_10 = PoseSE3::Identity();
            PointVec mappoints;
        };
        struct KeyframeStatistics {
            PoseSE3 Twc = PoseSE3::Identity();
            PointVec mappoints;
        };
        struct ExecutionStatistics {
            float time_track = 0.f, time_1p = 0.f, time_5p = 0.f, time_localba = 0.f, time_new = 0.f, time_total = 0.f;
        };
        std::vector<LandmarkStatistics> stats_landmark;
        std::vector<FrameStatistics> stats_frame;
        std::vector<KeyframeStatistics> stats_keyframe;
        std::vector<ExecutionStatistics> stats_execution;
    };

    // All user parameters of config/stereo/*.yaml that the step reads (stereo_vo.cpp:186-285), plus the K-det knobs.
    struct Parameters {
        int width = 1241, height = 376;
        float K_l[4] = {718.856f, 718.856f, 607.1928f, 185.2157f}, K_r[4] = {718.856f, 718.856f, 607.1928f, 185.2157f};
        float T_lr[16] = {1, 0, 0, 0.5371657189f, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};   // row-major
        float thres_error = 80.f, thres_bidirection = 0.5f, thres_sampson = 60.f;
        int window_size = 21, max_level = 6;
        int n_bins_u = 24, n_bins_v = 12;
        float thres_poseba_error = 3.f;
        float thres_alive_ratio = 0.6f, thres_trans = 10.f, thres_rotation_deg = 15.f;
        int n_max_keyframes_in_window = 9;
        int do_scale_refine = 1;
        int det_edge = 31;
        long long det_min_score = 0;
        int device = 0;
        // flagDoUndistortion (stereo_vo.cpp:239, 414-428): rectify both images on the device; K_l / K_r / T_lr above are
        // then the RAW pinhole-radtan cameras and D_l / D_r their distortion (k1, k2, p1, p2, k3; camera.cpp:30-35)
        int do_undistortion = 0;
        float D_l[5] = {0, 0, 0, 0, 0}, D_r[5] = {0, 0, 0, 0, 0};
        int collect_gate_counts = 0;     // 1: FrameInfo::counts (survivors after every gate; three extra tiny launches per frame)
        // keypoint extractor: VO_DETECTOR_HARRIS_SCHARR (K-det) or VO_DETECTOR_ORB (cv::ORB::detect restated, what the
        // reference runs; the yaml constructor selects it with feature_extractor.thres_fastscore, stereo_vo.cpp:31-36)
        int detector = VO_DETECTOR_HARRIS_SCHARR;
        int thres_fastscore = 20;
        // pose-only GN accumulation: 0 = VO_POSE_FAST (FP64 tree sums), 1 = VO_POSE_STRICT (sequential FP32 sums in point
        // order: the reference's arithmetic bit for bit, about 0.2 ms more per frame at 2000 landmarks); yaml key
        // motion_estimator.pose_strict (ours; the reference has no such key)
        int pose_strict = 0;
        // trackWithScale samples outside the image: 1 = the reference's stale sample buffers reproduced (vo_set_scale_mode),
        // 0 = masked out (default: the faithful mode serialises the features along an image edge, +0.1-0.4 ms per frame);
        // yaml key feature_tracker.scale_faithful_borders (ours)
        int scale_faithful_borders = 0;
    };

    StereoVO(std::string mode, std::string directory_intrinsic);   // stereo_vo.cpp:9-53 (yaml via a minimal parser)
    explicit StereoVO(const Parameters &prm);
    ~StereoVO();
    StereoVO(const StereoVO &) = delete;
    StereoVO &operator=(const StereoVO &) = delete;

    void trackStereoImages(const cv::Mat &img_left, const cv::Mat &img_right, const double &timestamp);
    const AlgorithmStatistics &getStatistics() const { return stat_; }
    const cv::Mat &getDebugImage() { return img_debug_; }

    // ---- introspection used by tests / bench (not part of the reference surface)
    struct FrameInfo {
        int frame = 0, keyframe = 0, n_in = 0, n_tracked = 0, n_detected = 0, n_new = 0, n_recon = 0;
        int lba_points = 0, lba_obs = 0, lba_ok = 0;
        int counts[5] = {0, 0, 0, 0, 0};
        // wall-clock milliseconds: frame step (device call incl. copies and the sync), host bookkeeping, keyframe
        // reconstruction, local-BA packing, local-BA solve + write-back, statistics refresh, total
        float ms_step = 0.f, ms_book = 0.f, ms_recon = 0.f, ms_lba_pack = 0.f, ms_lba_solve = 0.f, ms_stats = 0.f, ms_total = 0.f;
    };
    const FrameInfo &lastFrameInfo() const { return info_; }
    const std::vector<int> &currentLandmarkIds() const;
    const std::vector<float> &currentPtsLeft() const;
    const std::vector<float> &currentPtsRight() const;
    long long launchCount() const;
    bool statsConsistent() const;       // stats_keyframe equals the reference's full refresh (test hook)

private:
    struct FrameRec {
        int id = 0;
        int kf_index = -1;                  // position in all_keyframes_ / stats_keyframe
        float Twc[16], Tcw[16], dT01[16];
        std::vector<float> pts_l, pts_r;    // interleaved x, y
        std::vector<int> lm_ids;
    };
    using FrameRecPtr = std::shared_ptr<FrameRec>;

    void init();
    void setPose(FrameRec &f, const float *Twc);                 // frame.cpp:44-48
    int newLandmarks(int k, int frame_id);
    bool checkUpdateRule(const FrameRec &f) const;               // keyframes.cpp:217-303
    void addKeyframe(const FrameRecPtr &f);                      // keyframes.cpp:177-215
    void reconstruct(FrameRec &f, int n_first);                  // stereo_vo.cpp:767-797 / 911-941
    void localBundleAdjustment();                                // motion_estimator.cpp:1207-1340 + sparse_ba_parameters.h:292-465
    void pushStats(const FrameRec &f, bool keyframe);

    Parameters p_;
    float K_use_l_[4], K_use_r_[4], T_lr_use_[16];   // what the step uses: raw or rectified cameras
    vo_ctx *ctx_ = nullptr;
    AlgorithmStatistics stat_;
    cv::Mat img_debug_;
    FrameInfo info_;
    // landmark table (SoA)
    std::vector<float> lm_X_;
    std::vector<uint8_t> lm_tri_, lm_alive_, lm_bundled_;
    std::vector<int> lm_last_frame_;
    // where a landmark's point sits in stats_keyframe[k].mappoints, and the landmarks whose point changed in this frame:
    // the per-keyframe refresh of the reference (all keyframes x all their points, every keyframe) becomes incremental
    // (a singly linked list per landmark in one flat pool: appending is two sequential stores and one head update instead of a
    // push_back into one of half a million small vectors)
    struct KfSlot { int kf_index, slot, next; };
    std::vector<KfSlot> kf_slot_pool_;
    std::vector<int> lm_slot_head_;          // landmark -> newest pool entry, -1 = none
    std::vector<int> lm_first_kf_;           // landmark -> oldest keyframe (kf_index) that lists it, -1 = none
    std::vector<int> lm_dirty_stamp_;        // scratch of the statistics refresh
    int dirty_stamp_ = 0;
    std::vector<int> dirty_;
    // local-BA packing scratch (kept across keyframes)
    std::vector<int> lm_seen_stamp_, lm_lba_slot_, lba_lms_, lba_obs_ptr_, lba_obs_cursor_, lba_obs_frame_;
    std::vector<uint8_t> lba_obs_right_;
    std::vector<double> lba_points_, lba_obs_px_, lba_poses_out_, lba_points_out_;
    int seen_stamp_ = 0;
    FrameRecPtr prev_;
    std::deque<FrameRecPtr> window_;
    std::vector<FrameRecPtr> all_keyframes_;
    int n_frames_ = 0;           // committed frames (see trackStereoImages)
    bool poisoned_ = false;      // an exception escaped after the frame was committed
    // per-frame scratch
    std::vector<float> in_l0_, in_r0_, in_X_, out_l1_, out_r1_, new_l_, new_r_;
    std::vector<uint8_t> in_tri_;
    std::vector<int> in_ids_, out_idx_;
};

// C wrapper so that the tests and bench.py (ctypes) can drive the class.
extern "C" {
typedef struct vo_svo vo_svo;
VO_API int vo_svo_create(const StereoVO::Parameters *prm, vo_svo **out);
VO_API int vo_svo_create_from_yaml(const char *directory_intrinsic, vo_svo **out);
VO_API void vo_svo_destroy(vo_svo *s);
VO_API int vo_svo_track(vo_svo *s, const unsigned char *img_l, const unsigned char *img_r, int w, int h, size_t step, double timestamp);
VO_API int vo_svo_pose(const vo_svo *s, float *T_wc16);                      // row-major, last frame
VO_API int vo_svo_frame_info(const vo_svo *s, StereoVO::FrameInfo *out);
VO_API int vo_svo_tracks(const vo_svo *s, int cap, int *ids, float *pts_l, float *pts_r);   // returns the count
VO_API int vo_svo_keyframe_poses(const vo_svo *s, int cap, float *T_wc16);                    // returns the count
VO_API long long vo_svo_launch_count(const vo_svo *s);
VO_API int vo_svo_stats_consistent(const vo_svo *s);                             // 1 if stats_keyframe == the reference's full refresh
VO_API const char *vo_svo_last_error(void);
VO_API int vo_svo_struct_size(int which);                                       // 0 Parameters, 1 FrameInfo (binding layout check)
}
