// stereo_vo.cpp -- host glue of StereoVO::trackStereoImages (core/visual_odometry/stereo_vo/stereo_vo.cpp:392-989)
// over flat arrays; every numeric stage is a C-ABI call into libvo_b200.so (see stereo_vo.h).
#include "stereo_vo.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

#include "vo_host_util.h"
using namespace vo_host;

// ------------------------------------------------------------------------------ construction
StereoVO::StereoVO(const Parameters &prm) : p_(prm) { init(); }

// stereo_vo.cpp:186-285: "%YAML:1.0" key: value pairs + the T_lr opencv-matrix block
StereoVO::StereoVO(std::string mode, std::string directory_intrinsic)
{
    if (mode != "dataset" && mode != "rosbag") throw std::runtime_error("StereoVO - unknown mode.");     // stereo_vo.cpp:14-21
    std::ifstream f(directory_intrinsic);
    if (!f.is_open()) throw std::runtime_error("intrinsic file cannot be found!\n");                    // :190
    std::map<std::string, std::string> kv;
    std::string line, all;
    std::vector<float> tlr;
    bool in_tlr = false;
    while (std::getline(f, line)) {
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        const size_t c = line.find(':');
        if (c == std::string::npos) continue;
        std::string key = line.substr(0, c), val = line.substr(c + 1);
        key.erase(0, key.find_first_not_of(" \t")); key.erase(key.find_last_not_of(" \t") + 1);
        val.erase(0, val.find_first_not_of(" \t")); if (!val.empty()) val.erase(val.find_last_not_of(" \t\r") + 1);
        if (key == "T_lr") { in_tlr = true; continue; }
        if (in_tlr && key == "data") {
            std::string d = val;
            while (d.find(']') == std::string::npos && std::getline(f, line)) d += line;
            for (char &ch : d) if (ch == '[' || ch == ']' || ch == ',') ch = ' ';
            std::istringstream ss(d);
            float v;
            while (ss >> v) tlr.push_back(v);
            in_tlr = false;
            continue;
        }
        kv[key] = val;
    }
    auto num = [&](const char *k, double dflt) { auto it = kv.find(k); return it == kv.end() || it->second.empty() ? dflt : atof(it->second.c_str()); };
    p_.width = (int)num("Camera.left.width", p_.width); p_.height = (int)num("Camera.left.height", p_.height);
    const char *names[4] = {"fx", "fy", "cx", "cy"};
    for (int i = 0; i < 4; ++i) {
        p_.K_l[i] = (float)num((std::string("Camera.left.") + names[i]).c_str(), p_.K_l[i]);
        p_.K_r[i] = (float)num((std::string("Camera.right.") + names[i]).c_str(), p_.K_r[i]);
    }
    p_.do_undistortion = num("flagDoUndistortion", 0) != 0 ? 1 : 0;
    const char *dn[5] = {"k1", "k2", "p1", "p2", "k3"};               // cvD order, stereo_vo.cpp:158-164
    for (int i = 0; i < 5; ++i) {
        p_.D_l[i] = (float)num((std::string("Camera.left.") + dn[i]).c_str(), 0.0);
        p_.D_r[i] = (float)num((std::string("Camera.right.") + dn[i]).c_str(), 0.0);
    }
    if (tlr.size() == 16) memcpy(p_.T_lr, tlr.data(), 64);
    p_.thres_error = (float)num("feature_tracker.thres_error", p_.thres_error);
    p_.thres_bidirection = (float)num("feature_tracker.thres_bidirection", p_.thres_bidirection);
    p_.thres_sampson = (float)num("feature_tracker.thres_sampson", p_.thres_sampson);
    p_.window_size = (int)num("feature_tracker.window_size", p_.window_size);
    p_.max_level = (int)num("feature_tracker.max_level", p_.max_level);
    p_.n_bins_u = (int)num("feature_extractor.n_bins_u", p_.n_bins_u);
    p_.n_bins_v = (int)num("feature_extractor.n_bins_v", p_.n_bins_v);
    p_.detector = VO_DETECTOR_ORB;                                    // the reference's extractor
    p_.thres_fastscore = (int)num("feature_extractor.thres_fastscore", p_.thres_fastscore);   // initParams(..., int THRES_FAST, ...)
    p_.pose_strict = (int)num("motion_estimator.pose_strict", 1);     // yaml construction = drop-in use: the reference's arithmetic
    p_.scale_faithful_borders = (int)num("feature_tracker.scale_faithful_borders", 0);
    p_.thres_poseba_error = (float)num("motion_estimator.thres_poseba_error", p_.thres_poseba_error);
    p_.thres_alive_ratio = (float)num("keyframe_update.thres_alive_ratio", p_.thres_alive_ratio);
    p_.thres_trans = (float)num("keyframe_update.thres_trans", p_.thres_trans);
    p_.thres_rotation_deg = (float)num("keyframe_update.thres_rotation", p_.thres_rotation_deg);
    p_.n_max_keyframes_in_window = (int)num("keyframe_update.n_max_keyframes_in_window", p_.n_max_keyframes_in_window);
    init();
}

void StereoVO::init()
{
    const int nb = std::max(1, p_.n_bins_u * p_.n_bins_v);
    // staging for 256 k "feature units" (16 MB pinned + device) up front: the frame step and the local-BA problem
    // (a few thousand landmarks x ~10 observations) then never re-allocate pinned memory inside the frame loop
    const int rc = vo_ctx_create(p_.device, p_.width, p_.height, 4, std::max(4 * nb + 4096, 262144), nullptr, &ctx_);
    if (rc != VO_OK) fail(nullptr, rc);          // VO_ERR_NO_DEVICE: there is no CPU fallback
    const int rp = vo_set_pose_mode(ctx_, p_.pose_strict ? VO_POSE_STRICT : VO_POSE_FAST);
    if (rp) fail(ctx_, rp);
    const int rs = vo_set_scale_mode(ctx_, p_.scale_faithful_borders);
    if (rs) fail(ctx_, rs);
    const int rl = vo_lba_reserve(ctx_, (size_t)32 << 20);       // the first local BA then allocates nothing inside the frame loop
    if (rl) fail(ctx_, rl);
    const int rd = vo_set_detector(ctx_, p_.detector, p_.thres_fastscore);
    if (rd) fail(ctx_, rd);
    {   // landmark tables: room for 2^19 landmarks (several hundred frames) before the first reallocation, which would
        // otherwise move hundreds of thousands of per-landmark vectors inside a frame (a ~10 ms latency spike)
        const size_t cap = (size_t)1 << 19;
        lm_X_.reserve(cap * 3); lm_tri_.reserve(cap); lm_alive_.reserve(cap); lm_bundled_.reserve(cap); lm_last_frame_.reserve(cap);
        lm_slot_head_.reserve(cap); lm_seen_stamp_.reserve(cap); lm_lba_slot_.reserve(cap); kf_slot_pool_.reserve(cap);
    }
    memcpy(K_use_l_, p_.K_l, 16); memcpy(K_use_r_, p_.K_r, 16); memcpy(T_lr_use_, p_.T_lr, 64);
    if (p_.do_undistortion) {
        // stereo_vo.cpp:414-428: rectified images, the rectified camera for both sides, the rectified extrinsics
        float K_rect[4];
        const int rr = vo_rectify_init(ctx_, p_.K_l, p_.D_l, p_.K_r, p_.D_r, p_.T_lr, p_.width, p_.height, K_rect, T_lr_use_);
        if (rr) fail(ctx_, rr);
        memcpy(K_use_l_, K_rect, 16); memcpy(K_use_r_, K_rect, 16);
    }
}

StereoVO::~StereoVO() { if (ctx_) vo_ctx_destroy(ctx_); }

long long StereoVO::launchCount() const { return vo_ctx_launch_count(ctx_); }
const std::vector<int> &StereoVO::currentLandmarkIds() const { static const std::vector<int> e; return prev_ ? prev_->lm_ids : e; }
const std::vector<float> &StereoVO::currentPtsLeft() const { static const std::vector<float> e; return prev_ ? prev_->pts_l : e; }
const std::vector<float> &StereoVO::currentPtsRight() const { static const std::vector<float> e; return prev_ ? prev_->pts_r : e; }

// ------------------------------------------------------------------------------ bookkeeping
void StereoVO::setPose(FrameRec &f, const float *Twc)
{
    memcpy(f.Twc, Twc, 64);
    inv_se3_f(f.Twc, f.Tcw);
}

int StereoVO::newLandmarks(int k, int frame_id)
{
    const int base = (int)lm_tri_.size();
    lm_X_.resize((size_t)(base + k) * 3, 0.f);
    lm_tri_.resize(base + k, 0); lm_alive_.resize(base + k, 1); lm_bundled_.resize(base + k, 0);
    lm_last_frame_.resize(base + k, frame_id);
    lm_slot_head_.resize(base + k, -1);
    lm_first_kf_.resize(base + k, -1);
    return base;
}

bool StereoVO::checkUpdateRule(const FrameRec &f) const
{
    if (window_.empty()) return true;
    const FrameRec &kf = *window_.back();
    int cnt_tracked = 0;
    for (int id : kf.lm_ids) if (lm_last_frame_[id] == f.id) ++cnt_tracked;
    const float ratio = (float)cnt_tracked / (float)kf.lm_ids.size();
    if (ratio <= p_.thres_alive_ratio) return true;
    float dT[16];
    mul4_f(kf.Tcw, f.Twc, dT);
    float costheta = (dT[0] + dT[5] + dT[10] - 1.0f) * 0.5f;
    if (costheta >= 0.999999) costheta = 0.999999;
    if (costheta <= -0.999999) costheta = -0.999999;
    const float rot = acosf(costheta);
    const float dtrans = std::sqrt(dT[3] * dT[3] + dT[7] * dT[7] + dT[11] * dT[11]);
    return rot >= p_.thres_rotation_deg * D2R || dtrans >= p_.thres_trans;
}

void StereoVO::addKeyframe(const FrameRecPtr &f)
{
    all_keyframes_.push_back(f);
    if ((int)window_.size() == p_.n_max_keyframes_in_window) window_.pop_front();
    window_.push_back(f);
    // The window keyframes keep their own (landmark id, left pixel, right pixel) arrays: the local-BA packing reads those
    // (localBundleAdjustment), so no per-landmark observation list has to be maintained here.
    const size_t n = f->lm_ids.size();
    const int kf_index = (int)all_keyframes_.size() - 1;
    f->kf_index = kf_index;
    for (size_t i = 0; i < n; ++i) {
        const int id = f->lm_ids[i];
        if (lm_slot_head_[id] < 0) lm_first_kf_[id] = kf_index;
        kf_slot_pool_.push_back({kf_index, (int)i, lm_slot_head_[id]});
        lm_slot_head_[id] = (int)kf_slot_pool_.size() - 1;
    }
}

void StereoVO::reconstruct(FrameRec &f, int n_first)
{
    info_.n_recon = 0;
    if (n_first <= 0) return;
    std::vector<float> Xw((size_t)n_first * 3);
    std::vector<uint8_t> ok(n_first);
    const int rc = vo_stereo_reconstruct(ctx_, f.pts_l.data(), f.pts_r.data(), n_first, K_use_l_, K_use_r_, T_lr_use_, f.Twc, Xw.data(), ok.data());
    if (rc) fail(ctx_, rc);
    for (int i = 0; i < n_first; ++i) {
        if (!ok[i]) continue;
        const int id = f.lm_ids[i];
        memcpy(&lm_X_[(size_t)id * 3], &Xw[(size_t)i * 3], 12);      // Landmark::set3DPoint
        lm_tri_[id] = 1;
        dirty_.push_back(id);
        ++info_.n_recon;
    }
}

void StereoVO::localBundleAdjustment()
{
    const int NUM_MINIMUM_REQUIRED_KEYFRAMES = 3, NUM_FIX = 2;                     // motion_estimator.cpp:1245-1246
    info_.lba_points = info_.lba_obs = info_.lba_ok = 0;
    if ((int)window_.size() < NUM_MINIMUM_REQUIRED_KEYFRAMES) return;
    const auto t_pack = Clock::now();
    const int nf = (int)window_.size();
    // 1) alive + triangulated landmarks of the window in first-seen order (the reference's unordered_set order is
    //    address-hash dependent) and their observation counts, straight from the window keyframes'
    //    own arrays: every keyframe lists each of its landmarks once, with the left and the right pixel.
    //    (The first version walked a per-landmark vector of observations: 0.30 ms of pointer chasing per keyframe, plus
    //    0.15 ms per keyframe to maintain those vectors.)
    std::vector<int> &lms = lba_lms_, &obs_ptr = lba_obs_ptr_, &cursor = lba_obs_cursor_, &obs_frame = lba_obs_frame_;
    lms.clear(); obs_ptr.clear();
    {
        // stamp instead of a cleared flag array: no O(all landmarks) memset per keyframe
        lm_seen_stamp_.resize(lm_tri_.size(), 0);
        lm_lba_slot_.resize(lm_tri_.size(), 0);
        const int stamp = ++seen_stamp_;
        for (const auto &fr : window_)
            for (int id : fr->lm_ids) {
                if (!lm_tri_[id] || !lm_alive_[id]) continue;
                if (lm_seen_stamp_[id] != stamp) {
                    lm_seen_stamp_[id] = stamp;
                    lm_lba_slot_[id] = (int)lms.size();
                    lms.push_back(id);
                    obs_ptr.push_back(0);
                }
                obs_ptr[lm_lba_slot_[id]] += 2;          // left + right
            }
    }
    // THRES_MINIMUM_SEEN = 2 observations: a stereo keyframe contributes two, so every listed landmark qualifies
    if (lms.empty()) return;
    const int n_lm = (int)lms.size();
    cursor.resize(n_lm);
    {
        int run = 0;
        for (int j = 0; j < n_lm; ++j) { const int c = obs_ptr[j]; obs_ptr[j] = run; cursor[j] = run; run += c; }
        obs_ptr.push_back(run);
    }
    const int n_obs = obs_ptr[n_lm];
    double Twj_ref[16], Tjw_ref[16];
    for (int i = 0; i < 12; ++i) Twj_ref[i] = window_[0]->Twc[i];
    Twj_ref[12] = Twj_ref[13] = Twj_ref[14] = 0; Twj_ref[15] = 1;
    memset(Tjw_ref, 0, sizeof(Tjw_ref));
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Tjw_ref[i * 4 + j] = Twj_ref[j * 4 + i];
        double s = 0; for (int k = 0; k < 3; ++k) s += Twj_ref[k * 4 + i] * Twj_ref[k * 4 + 3];
        Tjw_ref[i * 4 + 3] = -s;
    }
    Tjw_ref[15] = 1;
    const double pose_scale = 10.0, inv_scale = 1.0 / pose_scale;
    // packing buffers live in the object: a few MB that would otherwise be mmap'ed, page-faulted and unmapped per keyframe
    std::vector<uint8_t> &obs_right = lba_obs_right_;
    std::vector<double> &points = lba_points_, &obs_px = lba_obs_px_;
    obs_frame.resize(n_obs); obs_right.resize(n_obs); obs_px.resize((size_t)2 * n_obs); points.resize((size_t)3 * n_lm);
    // 2) observations: keyframes in chronological order, left then right -- the order Landmark::related_keyframes_ /
    //    observations_on_keyframes_ have in the reference (landmark.cpp:105-124)
    for (int k = 0; k < nf; ++k) {
        const FrameRec &fr = *window_[k];
        const size_t n = fr.lm_ids.size();
        for (size_t i = 0; i < n; ++i) {
            const int id = fr.lm_ids[i];
            if (lm_seen_stamp_[id] != seen_stamp_) continue;          // stamped in pass 1 <=> triangulated and alive
            const int o = cursor[lm_lba_slot_[id]];
            cursor[lm_lba_slot_[id]] = o + 2;
            obs_frame[o] = k; obs_frame[o + 1] = k;
            obs_right[o] = 0; obs_right[o + 1] = 1;
            obs_px[2 * (size_t)o] = fr.pts_l[2 * i]; obs_px[2 * (size_t)o + 1] = fr.pts_l[2 * i + 1];
            obs_px[2 * (size_t)o + 2] = fr.pts_r[2 * i]; obs_px[2 * (size_t)o + 3] = fr.pts_r[2 * i + 1];
        }
    }
    for (int j = 0; j < n_lm; ++j) {
        const int id = lms[j];
        const double Xw[3] = {lm_X_[(size_t)id * 3], lm_X_[(size_t)id * 3 + 1], lm_X_[(size_t)id * 3 + 2]};
        for (int r = 0; r < 3; ++r)
            points[3 * (size_t)j + r] = (Tjw_ref[r * 4] * Xw[0] + Tjw_ref[r * 4 + 1] * Xw[1] + Tjw_ref[r * 4 + 2] * Xw[2] + Tjw_ref[r * 4 + 3]) * inv_scale;
    }
    std::vector<double> poses((size_t)nf * 16);
    for (int k = 0; k < nf; ++k) {
        double Tjw[16];
        for (int i = 0; i < 12; ++i) Tjw[i] = window_[k]->Tcw[i];
        Tjw[12] = Tjw[13] = Tjw[14] = 0; Tjw[15] = 1;
        mul4_d(Tjw, Twj_ref, Tjw);
        for (int r = 0; r < 3; ++r) Tjw[r * 4 + 3] *= inv_scale;
        memcpy(&poses[(size_t)k * 16], Tjw, sizeof(Tjw));
    }
    std::vector<int> opt_index(nf, -1);
    for (int k = NUM_FIX; k < nf; ++k) opt_index[k] = k - NUM_FIX;
    vo_lba_problem pr;
    memset(&pr, 0, sizeof(pr));
    pr.n_frames = nf; pr.n_opt = nf - NUM_FIX; pr.n_points = (int)lms.size(); pr.n_obs = (int)obs_frame.size();
    pr.poses = poses.data(); pr.opt_index = opt_index.data(); pr.points = points.data(); pr.obs_ptr = obs_ptr.data();
    pr.obs_frame = obs_frame.data(); pr.obs_right = obs_right.data(); pr.obs_px = obs_px.data();
    for (int i = 0; i < 4; ++i) { pr.K_l[i] = K_use_l_[i]; pr.K_r[i] = K_use_r_[i]; }
    for (int i = 0; i < 16; ++i) pr.T_lr[i] = T_lr_use_[i];
    for (int r = 0; r < 3; ++r) pr.T_lr[r * 4 + 3] *= inv_scale;
    pr.is_stereo = 1; pr.huber = 0.5; pr.lambda = 0.00001; pr.max_iter = 10;
    std::vector<double> &poses_out = lba_poses_out_, &points_out = lba_points_out_;
    poses_out.resize(poses.size()); points_out.resize(points.size());
    std::vector<double> avg(pr.max_iter);
    int ok = 0;
    info_.ms_lba_pack = ms_since(t_pack);
    const auto t_solve = Clock::now();
    const int rc = vo_lba_solve(ctx_, &pr, poses_out.data(), points_out.data(), avg.data(), &ok);
    if (rc == VO_ERR_NAN) throw std::runtime_error("Local BA NAN!\n");           // sparse_bundle_adjustment.cpp:761
    if (rc) fail(ctx_, rc);
    info_.lba_points = pr.n_points; info_.lba_obs = pr.n_obs; info_.lba_ok = ok;
    // write-back (sparse_bundle_adjustment.cpp:631-718)
    bool large_update = false;
    for (int k = 0; k < nf; ++k) {
        if (opt_index[k] < 0) continue;
        double Tjw[16], Twj0[16], dT[16];
        memcpy(Tjw, &poses_out[(size_t)k * 16], sizeof(Tjw));
        for (int r = 0; r < 3; ++r) Tjw[r * 4 + 3] *= pose_scale;
        mul4_d(Tjw, Tjw_ref, Tjw);
        for (int i = 0; i < 12; ++i) Twj0[i] = window_[k]->Twc[i];
        Twj0[12] = Twj0[13] = Twj0[14] = 0; Twj0[15] = 1;
        mul4_d(Twj0, Tjw, dT);
        if (std::sqrt(dT[3] * dT[3] + dT[7] * dT[7] + dT[11] * dT[11]) > 50) large_update = true;
        float Tjw_f[16], Twj_f[16];
        for (int i = 0; i < 12; ++i) Tjw_f[i] = (float)Tjw[i];
        Tjw_f[12] = Tjw_f[13] = Tjw_f[14] = 0.f; Tjw_f[15] = 1.f;
        inv_se3_f(Tjw_f, Twj_f);
        setPose(*window_[k], Twj_f);
    }
    for (size_t j = 0; j < lms.size(); ++j) {
        double X[3];
        for (int r = 0; r < 3; ++r) X[r] = points_out[3 * j + r] * pose_scale;
        float Xf[3];
        for (int r = 0; r < 3; ++r) Xf[r] = (float)(Twj_ref[r * 4] * X[0] + Twj_ref[r * 4 + 1] * X[1] + Twj_ref[r * 4 + 2] * X[2] + Twj_ref[r * 4 + 3]);
        const int id = lms[j];
        memcpy(&lm_X_[(size_t)id * 3], Xf, 12);
        lm_tri_[id] = 1;
        dirty_.push_back(id);
        if (std::sqrt(Xf[0] * Xf[0] + Xf[1] * Xf[1] + Xf[2] * Xf[2]) <= 3000) lm_bundled_[id] = 1;
        else lm_alive_[id] = 0;
    }
    info_.ms_lba_solve = ms_since(t_solve);
    if (large_update) throw std::runtime_error("large update!");                  // :731
}

void StereoVO::pushStats(const FrameRec &f, bool keyframe)
{
    if (keyframe) {
        // stereo_vo.cpp:814-822 refreshes the pose and every map point of EVERY keyframe ever made, each time.  Same values, incrementally:
        // the new keyframe in full, the poses of the window (the only ones the LBA moves), and the points that changed
        // in this frame (reconstruction, LBA) wherever they sit -- the cost no longer grows with the sequence length.
        stat_.stats_keyframe.emplace_back();
        if (stat_.stats_keyframe.size() == all_keyframes_.size()) {
            const FrameRec &nk = *all_keyframes_.back();
            PointVec &mp = stat_.stats_keyframe.back().mappoints;
            mp.resize(nk.lm_ids.size());
            for (size_t i = 0; i < nk.lm_ids.size(); ++i)
                for (int r = 0; r < 3; ++r) mp[i](r) = lm_X_[(size_t)nk.lm_ids[i] * 3 + r];
            for (const auto &kf : window_) rowmajor_to_pose(kf->Twc, stat_.stats_keyframe[kf->kf_index].Twc);
            // changed points: the window keyframes by a sequential scan of their own landmark lists (one cache-friendly pass
            // instead of a list walk per landmark), older keyframes -- only landmarks that outlived the window have any --
            // through the per-landmark list
            lm_dirty_stamp_.resize(lm_tri_.size(), 0);
            const int stamp = ++dirty_stamp_;
            for (int id : dirty_) lm_dirty_stamp_[id] = stamp;
            const int front = window_.front()->kf_index;
            for (const auto &kf : window_) {
                PointVec &kmp = stat_.stats_keyframe[kf->kf_index].mappoints;
                const size_t n_kf_lm = kf->lm_ids.size();
                for (size_t i = 0; i < n_kf_lm; ++i) {
                    const int id = kf->lm_ids[i];
                    if (lm_dirty_stamp_[id] != stamp) continue;
                    for (int r = 0; r < 3; ++r) kmp[i](r) = lm_X_[(size_t)id * 3 + r];
                }
            }
            for (int id : dirty_) {
                if (lm_first_kf_[id] < 0 || lm_first_kf_[id] >= front) continue;
                for (int e = lm_slot_head_[id]; e >= 0; e = kf_slot_pool_[e].next) {
                    const KfSlot &sl = kf_slot_pool_[e];
                    if (sl.kf_index >= front) continue;
                    for (int r = 0; r < 3; ++r) stat_.stats_keyframe[sl.kf_index].mappoints[sl.slot](r) = lm_X_[(size_t)id * 3 + r];
                }
            }
        }
        dirty_.clear();              // points that change on a non-keyframe (first-frame / initial reconstruction) wait for the next keyframe
    }
    stat_.stats_frame.emplace_back();
    rowmajor_to_pose(f.Twc, stat_.stats_frame.back().Twc);                       // :979-980
    // the ROS1 node reads stats_landmark / stats_execution .back(): keep them in step
    stat_.stats_landmark.emplace_back();
    stat_.stats_landmark.back().n_initial = info_.n_in;
    stat_.stats_landmark.back().n_new = info_.n_new;
    stat_.stats_landmark.back().n_final = (int)f.lm_ids.size();
    stat_.stats_execution.emplace_back();
}

// The reference's full refresh, recomputed and compared with the incrementally maintained statistics (test hook).
bool StereoVO::statsConsistent() const
{
    if (stat_.stats_keyframe.size() != all_keyframes_.size()) return false;
    for (size_t j = 0; j < all_keyframes_.size(); ++j) {
        const FrameRec &kf = *all_keyframes_[j];
        const auto &sk = stat_.stats_keyframe[j];
        if (sk.mappoints.size() != kf.lm_ids.size()) return false;
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) if (sk.Twc(r, c) != kf.Twc[r * 4 + c]) return false;
        for (size_t i = 0; i < kf.lm_ids.size(); ++i)
            for (int r = 0; r < 3; ++r) if (sk.mappoints[i](r) != lm_X_[(size_t)kf.lm_ids[i] * 3 + r]) return false;
    }
    return true;
}

// ------------------------------------------------------------------------------ the step
void StereoVO::trackStereoImages(const cv::Mat &img_left, const cv::Mat &img_right, const double & /*timestamp*/)
{
    if (img_left.empty() || img_right.empty() || img_left.cols != img_right.cols || img_left.rows != img_right.rows)
        throw std::runtime_error("vo_b200: bad stereo image pair");
    const int w = img_left.cols, h = img_left.rows;
    if (img_left.step != img_right.step) throw std::runtime_error("vo_b200: left/right row pitch differ");
    // A frame is COMMITTED (frame counter = image-slot parity, landmark tables) only once its device step has succeeded: a
    // step that throws before that point ("PoseOnlyStereoBA is failed!", a CUDA error) leaves the object as it was, so the
    // next call still tracks from prev_'s image slot.  An exception after that point (local BA NaN / "large update!")
    // leaves the tables half updated -- the reference dies there too -- and the object refuses further images.
    if (poisoned_) throw std::runtime_error("vo_b200: StereoVO state is inconsistent after a failed keyframe step; create a new object");
    const auto t_total = Clock::now();
    auto fr = std::make_shared<FrameRec>();
    fr->id = n_frames_;
    ident(fr->Twc); ident(fr->Tcw); ident(fr->dT01);
    info_ = FrameInfo();
    info_.frame = fr->id;
    const int k = fr->id;
    const int sl = 2 * (k % 2), sr = 2 * (k % 2) + 1, sp = 2 * ((k + 1) % 2);
    const int nb = p_.n_bins_u * p_.n_bins_v;

    vo_stereo_frame_params fp;
    memset(&fp, 0, sizeof(fp));
    fp.track.window_size = p_.window_size; fp.track.max_level = p_.max_level; fp.track.thres_error = p_.thres_error;
    fp.track.thres_poseba_error = p_.thres_poseba_error;
    memcpy(fp.track.K_l, K_use_l_, 16); memcpy(fp.track.K_r, K_use_r_, 16); memcpy(fp.track.T_lr, T_lr_use_, 64);
    fp.track.do_scale_refine = p_.do_scale_refine;
    // mask_sampson = dist < THRES_SAMPSON with dist = 100 for y > 660 (stereo_vo.cpp:657-668): the stub only fires
    // when the threshold is below 100
    fp.track.sampson_y = p_.thres_sampson > 100.f ? 3.0e38f : 660.f;
    fp.thres_bidirection = p_.thres_bidirection;
    fp.n_bins_u = p_.n_bins_u; fp.n_bins_v = p_.n_bins_v; fp.det_edge = p_.det_edge; fp.det_min_score = p_.det_min_score;

    new_l_.resize((size_t)std::max(nb, 1) * 2); new_r_.resize((size_t)std::max(nb, 1) * 2);
    float T_wc[16], dT[16];
    vo_stereo_frame_result res;
    memset(&res, 0, sizeof(res));
    res.T_wc = T_wc; res.dT_pc = dT; res.new_l1 = new_l_.data(); res.new_r1 = new_r_.data(); res.counts = p_.collect_gate_counts ? info_.counts : nullptr;

    const unsigned char *up_l = img_left.data, *up_r = img_right.data;
    if (p_.do_undistortion) {
        // stereo_vo.cpp:414-421: rectify on the device; the frame step then finds the images already in their slots
        int rr = vo_upload_image_rectified(ctx_, sl, 0, img_left.data, w, h, img_left.step);
        if (!rr) rr = vo_upload_image_rectified(ctx_, sr, 1, img_right.data, w, h, img_right.step);
        if (rr == VO_ERR_SIZE_MISMATCH) throw std::runtime_error(vo_last_error(ctx_));
        if (rr) fail(ctx_, rr);
        up_l = up_r = nullptr;
    }
    if (!prev_) {
        // ---- the very first image (stereo_vo.cpp:842-949)
        fp.new_depth_gate = 0;
        const int rc = vo_stereo_frame_step(ctx_, &fp, -1, sl, sr, up_l, up_r, w, h, img_left.step, 0, nullptr, nullptr,
                                            nullptr, nullptr, nullptr, nullptr, &res);
        if (rc) fail(ctx_, rc);
        ++n_frames_;                                 // commit point
        poisoned_ = true;
        const int m = res.n_new;
        const int base = newLandmarks(m, fr->id);
        fr->pts_l.assign(new_l_.begin(), new_l_.begin() + 2 * (size_t)m);
        fr->pts_r.assign(new_r_.begin(), new_r_.begin() + 2 * (size_t)m);
        fr->lm_ids.resize(m);
        for (int i = 0; i < m; ++i) fr->lm_ids[i] = base + i;
        info_.n_detected = res.n_detected; info_.n_new = m;
        reconstruct(*fr, m);                         // pose is identity: X_w = X_l (:937)
        pushStats(*fr, false);
        prev_ = fr;
        poisoned_ = false;
        return;
    }

    // ---- steady state. landmark.cpp:304: dead landmarks never survive the first compaction
    const FrameRec &pv = *prev_;
    const size_t n_prev = pv.lm_ids.size();
    in_ids_.clear(); in_l0_.clear(); in_r0_.clear(); in_X_.clear(); in_tri_.clear();
    for (size_t i = 0; i < n_prev; ++i) {
        const int id = pv.lm_ids[i];
        if (!lm_alive_[id]) continue;
        in_ids_.push_back(id);
        in_l0_.push_back(pv.pts_l[2 * i]); in_l0_.push_back(pv.pts_l[2 * i + 1]);
        in_r0_.push_back(pv.pts_r[2 * i]); in_r0_.push_back(pv.pts_r[2 * i + 1]);
        for (int r = 0; r < 3; ++r) in_X_.push_back(lm_X_[(size_t)id * 3 + r]);
        in_tri_.push_back(lm_tri_[id]);
    }
    const int n = (int)in_ids_.size();
    out_idx_.resize(std::max(n, 1)); out_l1_.resize((size_t)std::max(n, 1) * 2); out_r1_.resize((size_t)std::max(n, 1) * 2);
    res.index = out_idx_.data(); res.pts_l1 = out_l1_.data(); res.pts_r1 = out_r1_.data();
    fp.new_depth_gate = 1;
    const auto t_step = Clock::now();
    const int rc = vo_stereo_frame_step(ctx_, &fp, sp, sl, sr, up_l, up_r, w, h, img_left.step, n, in_l0_.data(), in_r0_.data(),
                                        in_X_.data(), in_tri_.data(), pv.Twc, pv.dT01, &res);
    if (rc) fail(ctx_, rc);
    ++n_frames_;                                     // commit point
    poisoned_ = true;
    const float ms_step = ms_since(t_step);
    setPose(*fr, T_wc);                              // :642
    float dT10[16];
    inv4_f(dT, dT10);                                // :643 dT_pc_poBA.inverse()
    inv_se3_f(dT10, fr->dT01);                       // frame.cpp:50-54
    const int nt = res.n_tracked, m = res.n_new;
    fr->lm_ids.resize((size_t)nt + m);
    fr->pts_l.resize(2 * ((size_t)nt + m)); fr->pts_r.resize(2 * ((size_t)nt + m));
    for (int i = 0; i < nt; ++i) {
        const int id = in_ids_[out_idx_[i]];
        fr->lm_ids[i] = id;
        lm_last_frame_[id] = fr->id;                 // [8] addObservationAndRelatedFrame
    }
    memcpy(fr->pts_l.data(), out_l1_.data(), (size_t)nt * 8);
    memcpy(fr->pts_r.data(), out_r1_.data(), (size_t)nt * 8);
    const int base = newLandmarks(m, fr->id);        // [10] :729-734
    for (int i = 0; i < m; ++i) fr->lm_ids[nt + i] = base + i;
    memcpy(fr->pts_l.data() + 2 * (size_t)nt, new_l_.data(), (size_t)m * 8);
    memcpy(fr->pts_r.data() + 2 * (size_t)nt, new_r_.data(), (size_t)m * 8);
    info_.n_in = n; info_.n_tracked = nt; info_.n_detected = res.n_detected; info_.n_new = m;
    // [12] keyframe (:755-827)
    const bool kf = checkUpdateRule(*fr);
    if (kf) {
        info_.keyframe = 1;
        addKeyframe(fr);
        const auto t_rec = Clock::now();
        reconstruct(*fr, nt);                        // only the tracked survivors (n_pts is fixed before the appends, :767)
        info_.ms_recon = ms_since(t_rec);
        localBundleAdjustment();
    }
    const auto t_stats = Clock::now();
    pushStats(*fr, kf);
    info_.ms_stats = ms_since(t_stats);
    info_.ms_step = ms_step;
    info_.ms_total = ms_since(t_total);
    info_.ms_book = info_.ms_total - info_.ms_step - info_.ms_recon - info_.ms_lba_pack - info_.ms_lba_solve - info_.ms_stats;
    // ExecutionStatistics as the reference declares them (stereo_vo.h:166-176)
    stat_.stats_execution.back().time_track = info_.ms_step;
    stat_.stats_execution.back().time_localba = info_.ms_lba_pack + info_.ms_lba_solve;
    stat_.stats_execution.back().time_new = info_.ms_recon;
    stat_.stats_execution.back().time_total = info_.ms_total;
    prev_ = fr;
    poisoned_ = false;
}

// ------------------------------------------------------------------------------ C wrapper
struct vo_svo { StereoVO *vo; };
static thread_local std::string g_svo_error;

extern "C" const char *vo_svo_last_error(void) { return g_svo_error.c_str(); }

extern "C" int vo_svo_create(const StereoVO::Parameters *prm, vo_svo **out)
{
    if (!prm || !out) return VO_ERR_INVALID_ARG;
    try { *out = new vo_svo{new StereoVO(*prm)}; return VO_OK; }
    catch (const std::exception &e) { g_svo_error = e.what(); *out = nullptr; return VO_ERR_NO_DEVICE; }
}
extern "C" int vo_svo_create_from_yaml(const char *dir, vo_svo **out)
{
    if (!dir || !out) return VO_ERR_INVALID_ARG;
    try { *out = new vo_svo{new StereoVO("dataset", dir)}; return VO_OK; }
    catch (const std::exception &e) { g_svo_error = e.what(); *out = nullptr; return VO_ERR_INVALID_ARG; }
}
extern "C" void vo_svo_destroy(vo_svo *s) { if (s) { delete s->vo; delete s; } }
extern "C" int vo_svo_track(vo_svo *s, const unsigned char *img_l, const unsigned char *img_r, int w, int h, size_t step, double timestamp)
{
    if (!s || !img_l || !img_r) return VO_ERR_INVALID_ARG;
    try {
        cv::Mat L(h, w, const_cast<unsigned char *>(img_l), step), R(h, w, const_cast<unsigned char *>(img_r), step);
        s->vo->trackStereoImages(L, R, timestamp);
        return VO_OK;
    } catch (const std::exception &e) { g_svo_error = e.what(); return VO_ERR_NAN; }
}
extern "C" int vo_svo_pose(const vo_svo *s, float *T)
{
    if (!s || !T || s->vo->getStatistics().stats_frame.empty()) return VO_ERR_INVALID_ARG;
    const PoseSE3 &P = s->vo->getStatistics().stats_frame.back().Twc;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T[r * 4 + c] = P(r, c);
    return VO_OK;
}
extern "C" int vo_svo_frame_info(const vo_svo *s, StereoVO::FrameInfo *out)
{
    if (!s || !out) return VO_ERR_INVALID_ARG;
    *out = s->vo->lastFrameInfo();
    return VO_OK;
}
extern "C" int vo_svo_tracks(const vo_svo *s, int cap, int *ids, float *pts_l, float *pts_r)
{
    if (!s) return VO_ERR_INVALID_ARG;
    const auto &id = s->vo->currentLandmarkIds();
    const int n = std::min<int>(cap, (int)id.size());
    if (ids) memcpy(ids, id.data(), (size_t)n * 4);
    if (pts_l) memcpy(pts_l, s->vo->currentPtsLeft().data(), (size_t)n * 8);
    if (pts_r) memcpy(pts_r, s->vo->currentPtsRight().data(), (size_t)n * 8);
    return (int)id.size();
}
extern "C" int vo_svo_keyframe_poses(const vo_svo *s, int cap, float *T)
{
    if (!s) return VO_ERR_INVALID_ARG;
    const auto &kf = s->vo->getStatistics().stats_keyframe;
    const int n = std::min<int>(cap, (int)kf.size());
    for (int k = 0; k < n && T; ++k)
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T[k * 16 + r * 4 + c] = kf[k].Twc(r, c);
    return (int)kf.size();
}
extern "C" int vo_svo_stats_consistent(const vo_svo *s) { return s ? (s->vo->statsConsistent() ? 1 : 0) : 0; }
extern "C" long long vo_svo_launch_count(const vo_svo *s) { return s ? s->vo->launchCount() : 0; }
// layout check for language bindings: 0 = sizeof(Parameters), 1 = sizeof(FrameInfo)
extern "C" int vo_svo_struct_size(int which) { return which == 0 ? (int)sizeof(StereoVO::Parameters) : (int)sizeof(StereoVO::FrameInfo); }
