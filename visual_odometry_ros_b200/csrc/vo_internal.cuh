// vo_internal.cuh -- shared device/host structures of libvo_b200 (sm_100a only).
#pragma once
#include <cuda.h>            // CUtensorMap (type only; the driver entry point is fetched at run time)
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/vo_b200.h"

#define VO_PAD 32           // reflect-101 border (px) kept around every pyramid level (>= win+1)
#define VO_MAX_WIN 31

// One pyramid level in HBM.  `img` points at pixel (0,0) of a padded, reflect-101-bordered
// u8 plane (row pitch `pitch` bytes, 128-B aligned, VO_PAD px of border on every side, so
// img[y*pitch + x] is valid for -VO_PAD <= x < w+VO_PAD, same for y).  `deriv` is the Scharr
// derivative plane (short2 = {Ix, Iy}) with the same pitch in ELEMENTS and a zero border --
// the layout cv::calcOpticalFlowPyrLK builds on the CPU (reflect-101 image ring, constant-0
// derivative ring), but resident and built once per image.
struct LevelDesc {
    uint8_t *img;
    short2 *deriv;
    int w, h;
    int pitch;
    int _pad;
};

// Slot ids of one batched launch travel by value in the kernel parameters (no staging copy).
#define VO_IDLIST_MAX 64
struct IdList {
    int id[VO_IDLIST_MAX];
};

struct SlotDesc {
    LevelDesc lv[VO_MAX_LEVELS];
    const uint8_t *raw;   // densely packed w x h upload target (DMA lands here at full PCIe rate)
};

struct Slot {
    SlotDesc desc;            // host copy of the device descriptor
    uint8_t *base = nullptr;  // one allocation: all image planes then all deriv planes
    size_t bytes = 0;
    int w = 0, h = 0;         // current image size (level 0)
    int levels_built = 0;     // pyramid levels valid (0 = only level 0 pixels uploaded)
    int deriv_built = 0;      // derivative levels valid
    bool border0 = false;     // level-0 border filled
    bool raw_pending = false; // pixels sit in the raw staging area, not yet ingested into level 0
};

struct vo_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;      // image DMA overlapped with compute in the batched entry points
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};   // extra compute streams of the batched entry points (chunk overlap)
    cudaEvent_t ev_aux[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> events;
    int max_w = 0, max_h = 0, n_slots = 0, max_feat = 0;
    std::vector<Slot> slots;
    SlotDesc *d_slots = nullptr;   // device mirror of all slot descriptors
    uint8_t *slot_pool = nullptr;  // one allocation: slot s lives at slot_pool + s * slot_stride
    size_t slot_stride = 0;
    void *klt_maps = nullptr;      // cache of TMA descriptor sets keyed by (w, h, window) (klt.cu)
    uint8_t *raw_base = nullptr;   // n_slots x raw_stride bytes: contiguous upload staging
    size_t raw_stride = 0;
    int max_levels = 0;            // levels allocated per slot
    // pinned + device staging for the host-pointer entry points
    uint8_t *h_stage = nullptr;
    uint8_t *d_stage = nullptr;
    size_t stage_bytes = 0;
    long long launches = 0;
    int pose_flags = 0;            // VO_POSE_FAST / VO_POSE_STRICT (vo_set_pose_mode)
    int scale_faithful = 0;        // trackWithScale: reproduce the reference's stale sample buffers (vo_set_scale_mode)
    uint8_t *d_ks = nullptr;       // scratch of that mode (k_klt_scale_fixup)
    size_t ks_bytes = 0;
    std::string last_error;
    // trackWithScale scratch (float image + Sobel derivatives), lazily allocated
    float *d_f32[4] = {nullptr, nullptr, nullptr, nullptr};
    int f32_slot[2] = {-1, -1};
    // K-det scratch (score plane + per-bin state)
    void *d_det = nullptr;
    size_t det_bytes = 0;
    // K-orb scratch (8-level pyramid, score planes, keypoint lists) and the detector the frame steps use
    void *d_orb = nullptr;
    size_t orb_bytes = 0;
    int detector = 0;              // VO_DETECTOR_*
    int orb_fast_threshold = 20;
    struct { const uint8_t *img, *score, *nms; int w, h, pitch; } orb_dbg[8] = {};   // planes of the last K-orb run (vo_orb_read_level)
    // rectification: four CV_32FC1 maps + one distorted-image scratch plane (rectify.cu)
    void *d_rect = nullptr;
    size_t rect_bytes = 0;
    int rect_w = 0, rect_h = 0;
    // five-point RANSAC scratch (normalised points, hypotheses)
    void *d_fp = nullptr;
    size_t fp_bytes = 0;
    // LBA scratch
    void *d_lba = nullptr;
    size_t lba_bytes = 0;
    // distributed local BA (vo_dist_init): NCCL communicator of the landmark-sharded solve
    void *nccl_comm = nullptr;
    int dist_rank = 0, dist_world = 1;
};

#define VO_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            ctx->last_error = std::string(#call) + ": " + cudaGetErrorString(e__);     \
            return VO_ERR_CUDA;                                                        \
        }                                                                              \
    } while (0)

#define VO_REQUIRE(cond, code, msg)                                                    \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            if (ctx) ctx->last_error = (msg);                                          \
            return (code);                                                             \
        }                                                                              \
    } while (0)

// Grow-only device+pinned staging; returns VO_OK or error.
int vo_stage_reserve(vo_ctx *ctx, size_t bytes);
// Size a slot for a w x h image whose pixels are about to be written into its raw plane by a kernel (api.cu).
int vo_slot_prepare(vo_ctx *ctx, int slot, int w, int h);

// pyramid.cu
int vo_ensure_pyramids(vo_ctx *ctx, const int *slots, int n, int n_levels, int with_deriv);

// klt.cu
void vo_klt_maps_free(vo_ctx *ctx);
struct KltPost {           // fused FeatureTracker post-filter (feature_tracker.cpp:33-34 etc.)
    int mode;              // 0 none, 1 track, 2 with_prior, 3 bidir-forward (no mask), 4 bidir-backward
    float thres_err;
    float thres_bi2;       // already squared (and x5 for the with-prior variant)
    int border;            // 3 for trackBidirection, 0 for the *WithPrior variants
    int skip_masked;       // 1: features whose mask entry is 0 are skipped entirely
    const float *ref_pts;      // bidir-backward: pts0 to compare the back-track with
    const float *fwd_pts;      // bidir-backward: forward-tracked points (border test)
    const uint8_t *fwd_status; // bidir-backward
    const float *fwd_err;
    uint8_t *mask;             // in-out (ANDed)
};
int vo_klt_launch(vo_ctx *ctx, int n_pairs, const int *slots0, const int *slots1, const float *pts0_d,
                  int n, int win, int max_level, int flags, float *pts1_d, uint8_t *status_d,
                  float *err_d, long long *counters_d, const KltPost *post);

// klt_scale.cu / pose_gn.cu: device-pointer launchers (asynchronous, no staging)
int vo_klt_scale_launch_d(vo_ctx *ctx, int slot0, int slot1, const float *pts0_d, const float *scale_d, int n,
                          float *pts_track_d, uint8_t *mask_d, int *nan_flag_d);
int vo_pose_launch_d(vo_ctx *ctx, int n_prob, const int *offsets_d, int n_single, const int *n_single_d, const float *X_d,
                     const float *pl_d, const float *pr_d, const float *Kl, const float *Kr, const float *T_lr, float thres,
                     int mono, int variant, float *T01_d, uint8_t *mask_d, int *success_d, int *iters_d);

int vo_pose_launch_ex_d(vo_ctx *ctx, int n_prob, const int *offsets_d, int n_single, const int *n_single_d, const float *X_d,
                        const float *pl_d, const float *pr_d, const float *Kl, const float *Kr, const float *T_lr, float thres,
                        int mono, int variant, float *T01_d, uint8_t *mask_d, int *success_d, int *iters_d, int flags,
                        int max_iter, float *trace_d);

int vo_track_chain_launch_d(vo_ctx *ctx, int slot_l0, int slot_l1, int slot_r1, const float *pts_l0_d, float *pts_l1_d, float *pts_r1_d,
                            const float *scale_d, uint8_t *mask_d, int *nan_flag_d, int n, int win, int max_level, float thres_err,
                            int do_scale);

int vo_bidir_chain_launch_d(vo_ctx *ctx, int slot0, int slot1, const float *pts0_d, float *pts1_d, float *back_d, uint8_t *st_d, uint8_t *stb_d,
                            float *err_d, float *errb_d, uint8_t *mask_d, int skip_masked, int n, int win, int max_level, float thres_err,
                            float thres_bi, int with_prior, const float *scale_d, int *nan_flag_d);

// orb.cu
int vo_orb_launch_d(vo_ctx *ctx, int slot, const float *occ_d, const int *n_occ_d, int n_occ, int n_bins_u, int n_bins_v, int edge,
                    float *out_d, uint8_t *out_mask_d, int *n_out_d, int max_out, float *all_pt_d, float *all_resp_d, int *all_octave_d,
                    int *n_all_d, int max_all);

// five_point.cu
int vo_5pt_launch_d(vo_ctx *ctx, const float *pts0_d, const float *pts1_d, int n, const int *n_d, const float *K, float thres_px,
                    int n_hyp, unsigned seed, float *R10_d, float *t10_d, float *X0_d, uint8_t *mask_d, float *E_d, int *info_d);

static inline int vo_div_up(int a, int b) { return (a + b - 1) / b; }
