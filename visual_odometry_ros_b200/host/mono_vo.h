// mono_vo.h -- MonoVO with the reference's public API (core/visual_odometry/mono_vo/mono_vo.h:235-267) on top of the
// C ABI.  The unchanged ROS node constructs it with (mode, yaml file), feeds it images and reads getStatistics() /
// getDebugImage() (ros1/visual_odometry/mono_vo_ros1.cpp:19,123,131,244).
//
// What moved: one device call per frame (vo_mono_frame_step: pyramid, prior, trackBidirectionWithPrior, trackWithScale,
// mono pose-only GN with the five-point fallback, Sampson gate, compactions, bucketed detection, bidirectional
// back-tracking of the new features -- one H2D, one D2H, one synchronisation; its init_mode is the second image's
// five-point initialisation), batched DLT for the reconstructions (vo_triangulate_dlt) and the local BA (vo_lba_solve).
// What stays on the host, as in the reference: landmark / frame / keyframe bookkeeping incl. the per-observation
// parallax (landmark.cpp:76-135) and the window -> problem packing (sparse_ba_parameters.h:292-465), over flat arrays.
// cv::imshow drawing, the /home/kch trajectory dump of the destructor, the console prints and the commented-out scale
// estimator thread are not reproduced.
#pragma once
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "../../include/vo_b200.h"
#include "vo_shim_types.h"

class MonoVO {
public:
    // mono_vo.h:118-205 (field-compatible with what the node reads)
    struct AlgorithmStatistics {
        struct LandmarkStatistics {
            int n_initial = 0, n_pass_bidirection = 0, n_pass_1p = 0, n_pass_5p = 0, n_new = 0, n_final = 0;
            int max_age = 0, min_age = 0;
            float avg_age = 0.f;
            int n_ok_parallax = 0;
            float min_parallax = 0.f, max_parallax = 0.f, avg_parallax = 0.f;
        };
        struct FrameStatistics {
            float steering_angle = 0.f;
            PoseSE3 Twc = PoseSE3::Identity(), Tcw = PoseSE3::Identity(), dT_01 = PoseSE3::Identity(), dT_10 = PoseSE3::Identity();
            PointVec mappoints;
        };
        struct KeyframeStatistics {
            float steering_angle = 0.f;
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

    // the user parameters of config/mono/*.yaml that the step reads (mono_vo.cpp:140-235), plus the K-det / RANSAC knobs
    struct Parameters {
        int width = 1241, height = 376;
        float K[4] = {718.856f, 718.856f, 607.1928f, 185.2157f};
        float thres_error = 60.f, thres_bidirection = 0.5f, thres_sampson = 1000.f;
        int window_size = 21, max_level = 6;
        float thres_parallax_deg = 1.0f;              // map_update.thres_parallax (degrees in the yaml, :218-219)
        int n_bins_u = 30, n_bins_v = 12;
        float thres_5p_error = 1.0f, thres_poseba_error = 5.0f;
        float thres_overlap_ratio = 0.6f, thres_translation = 4.f, thres_rotation_deg = 10.f;
        int n_max_keyframes_in_window = 9;
        int do_scale_refine = 1;
        int det_edge = 31;
        long long det_min_score = 0;
        int device = 0;
        int n_hypotheses = 0;                         // five-point RANSAC hypotheses (0 = 1024)
        unsigned seed = 0;                            // five-point sampling seed; frame k uses seed + k
        int collect_gate_counts = 0;
        int record_frame_mappoints = 0;               // 1: stats_frame[k].mappoints = all triangulated landmarks (:1166-1177; O(all landmarks) per frame)
        // keypoint extractor: VO_DETECTOR_HARRIS_SCHARR (K-det) or VO_DETECTOR_ORB (cv::ORB::detect restated, what the
        // reference runs; the yaml constructor selects it with feature_extractor.thres_fastscore, mono_vo.cpp:36-40)
        int detector = VO_DETECTOR_HARRIS_SCHARR;
        int thres_fastscore = 20;
        // flagDoUndistortion (mono_vo.cpp:509-513): undistort every image on the device with the map of camera.cpp:57-87;
        // D = k1 k2 p1 p2 k3 (camera.cpp:30-34).  The camera matrix is unchanged
        int do_undistortion = 0;
        float D[5] = {0, 0, 0, 0, 0};
        // pose-only GN accumulation: 0 = VO_POSE_FAST (FP64 tree sums), 1 = VO_POSE_STRICT (sequential FP32 sums in point
        // order: the reference's arithmetic bit for bit); yaml key motion_estimator.pose_strict (ours)
        int pose_strict = 0;
        // trackWithScale samples outside the image: 1 = the reference's stale sample buffers reproduced (vo_set_scale_mode),
        // 0 = masked out (default: the faithful mode serialises the features along an image edge, +0.1-0.4 ms per frame);
        // yaml key feature_tracker.scale_faithful_borders (ours)
        int scale_faithful_borders = 0;
    };

    MonoVO(std::string mode, std::string directory_intrinsic);     // mono_vo.cpp:11-55 (yaml via a minimal parser)
    explicit MonoVO(const Parameters &prm);
    ~MonoVO();
    MonoVO(const MonoVO &) = delete;
    MonoVO &operator=(const MonoVO &) = delete;

    void trackImage(const cv::Mat &img, const double &timestamp);
    const AlgorithmStatistics &getStatistics() const { return stat_; }
    const cv::Mat &getDebugImage() { return img_debug_; }

    // ---- introspection used by tests / bench (not part of the reference surface)
    struct FrameInfo {
        int frame = 0, keyframe = 0, n_in = 0, n_tracked = 0, n_detected = 0, n_new = 0, n_recon = 0, used_5point = 0;
        int lba_points = 0, lba_obs = 0, lba_ok = 0;
        int counts[5] = {0, 0, 0, 0, 0};
        float ms_step = 0.f, ms_book = 0.f, ms_recon = 0.f, ms_lba_pack = 0.f, ms_lba_solve = 0.f, ms_stats = 0.f, ms_total = 0.f;
    };
    const FrameInfo &lastFrameInfo() const { return info_; }
    const std::vector<int> &currentLandmarkIds() const;
    const std::vector<float> &currentPts() const;
    int framePose(int frame_id, float *T_wc16) const;   // current value (the local BA moves keyframe poses), row-major
    int numFrames() const { return (int)frames_.size(); }
    long long launchCount() const;
    bool statsConsistent() const;       // stats_keyframe equals the reference's full refresh (test hook)

private:
    struct FrameRec {
        int id = 0;
        int kf_index = -1;                  // position in all_keyframes_ / stats_keyframe
        bool is_keyframe = false;
        float Twc[16], Tcw[16], dT01[16], dT10[16];
        std::vector<float> pts;             // interleaved x, y (dropped once the frame is neither previous nor a keyframe)
        std::vector<int> lm_ids;
        std::vector<float> kf_px;           // keyframes: the landmarks' observations.back() at the time the keyframe was added
    };
    using FrameRecPtr = std::shared_ptr<FrameRec>;

    void init();
    void setPose(FrameRec &f, const float *Twc);                // frame.cpp:44-48
    void setPoseDiff10(FrameRec &f, const float *dT10);         // frame.cpp:50-54
    int newLandmarks(int k, const float *pts, const FrameRec &f);
    void addObservations(const int *ids, const float *pts, int k, const FrameRec &f);   // landmark.cpp:76-135
    bool checkUpdateRule(const FrameRec &f) const;              // keyframes.cpp:47-120
    void addKeyframe(const FrameRecPtr &f);                     // keyframes.cpp:30-45
    int reconstructInitial(const FrameRec &f);                  // mono_vo.cpp:660-687
    int reconstructKeyframe(const FrameRec &f);                 // mono_vo.cpp:1032-1076
    int dltGroups(const std::vector<int> &cand, const std::vector<float> &pt0, const std::vector<float> &pt1,
                  const std::vector<int> &f0, const FrameRec &f1, std::vector<float> &X0, std::vector<float> &X1);
    void localBundleAdjustment();                               // motion_estimator.cpp:1090-1205 + sparse_ba_parameters.h:292-465
    void pushStats(const FrameRec &f, bool keyframe);

    Parameters p_;
    vo_ctx *ctx_ = nullptr;
    AlgorithmStatistics stat_;
    cv::Mat img_debug_;
    FrameInfo info_;
    bool initialised_ = false;
    // landmark table (SoA)
    std::vector<float> lm_X_, lm_first_px_, lm_last_px_, lm_last_parallax_;
    std::vector<uint8_t> lm_tri_, lm_alive_, lm_bundled_;
    std::vector<int> lm_last_frame_, lm_first_frame_, lm_age_;
    // keyframe observations of a landmark: what the path reads is the first one, the last one and how many (mono_vo.cpp:1032-1050);
    // the window's observations for the local BA come from the window keyframes' own arrays
    std::vector<int> lm_kf_count_, lm_kf_first_id_;
    std::vector<float> lm_kf_first_px_, lm_kf_last_px_;
    // where a landmark's point sits in stats_keyframe[k].mappoints, and the landmarks whose point changed in this frame:
    // the per-keyframe refresh of the reference (all keyframes x all their points, every keyframe) becomes incremental
    // (a singly linked list per landmark in one flat pool instead of one small vector per landmark)
    struct KfSlot { int kf_index, slot, next; };
    std::vector<KfSlot> kf_slot_pool_;
    std::vector<int> lm_slot_head_;          // landmark -> newest pool entry, -1 = none
    std::vector<int> lm_first_kf_;           // landmark -> oldest keyframe (kf_index) that lists it, -1 = none
    std::vector<int> lm_dirty_stamp_;        // scratch of the statistics refresh
    int dirty_stamp_ = 0;
    std::vector<int> dirty_;
    // local-BA packing scratch (kept across keyframes)
    std::vector<int> lm_seen_stamp_, lm_lba_slot_, lba_cand_, lba_cnt_, lba_lms_, lba_obs_ptr_, lba_obs_cursor_, lba_obs_frame_;
    std::vector<uint8_t> lba_obs_right_;
    std::vector<double> lba_points_, lba_obs_px_, lba_poses_out_, lba_points_out_;
    int seen_stamp_ = 0;
    std::vector<FrameRecPtr> frames_;      // all frames (poses stay readable: parallax and reconstruction use them)
    FrameRecPtr prev_;
    bool poisoned_ = false;      // an exception escaped after the frame was committed (see trackImage)
    std::deque<FrameRecPtr> window_;
    std::vector<FrameRecPtr> all_keyframes_;
    // per-frame scratch
    std::vector<float> in_p0_, in_X_, out_p1_, new_p1_, new_p0_;
    std::vector<uint8_t> in_flags_;
    std::vector<int> in_ids_, out_idx_;
};

// C wrapper so that the tests and bench.py (ctypes) can drive the class.
extern "C" {
typedef struct vo_mvo vo_mvo;
VO_API int vo_mvo_create(const MonoVO::Parameters *prm, vo_mvo **out);
VO_API int vo_mvo_create_from_yaml(const char *directory_intrinsic, vo_mvo **out);
VO_API void vo_mvo_destroy(vo_mvo *s);
VO_API int vo_mvo_track(vo_mvo *s, const unsigned char *img, int w, int h, size_t step, double timestamp);
VO_API int vo_mvo_pose(const vo_mvo *s, float *T_wc16);                          // row-major, last frame
VO_API int vo_mvo_frame_pose(const vo_mvo *s, int frame_id, float *T_wc16);      // current value of any frame's pose
VO_API int vo_mvo_frame_info(const vo_mvo *s, MonoVO::FrameInfo *out);
VO_API int vo_mvo_tracks(const vo_mvo *s, int cap, int *ids, float *pts);        // returns the count
VO_API long long vo_mvo_launch_count(const vo_mvo *s);
VO_API int vo_mvo_stats_consistent(const vo_mvo *s);                             // 1 if stats_keyframe == the reference's full refresh
VO_API const char *vo_mvo_last_error(void);
VO_API int vo_mvo_struct_size(int which);                                       // 0 Parameters, 1 FrameInfo (binding layout check)
}
