// mono_vo.cpp -- host glue of MonoVO::trackImage (core/visual_odometry/mono_vo/mono_vo.cpp:496-1194) over flat arrays;
// every numeric stage is a C-ABI call into libvo_b200.so (see mono_vo.h).
#include "mono_vo.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>

#include "vo_host_util.h"
using namespace vo_host;

// ------------------------------------------------------------------------------ construction
MonoVO::MonoVO(const Parameters &prm) : p_(prm) { init(); }

// mono_vo.cpp:140-235: "%YAML:1.0" key: value pairs
MonoVO::MonoVO(std::string mode, std::string directory_intrinsic)
{
    if (mode == "dataset") throw std::runtime_error("dataset mode is not supported.");     // mono_vo.cpp:21
    if (mode != "rosbag") throw std::runtime_error("MonoVO - unknown mode.");              // :31
    std::ifstream f(directory_intrinsic);
    if (!f.is_open()) throw std::runtime_error("intrinsic file cannot be found!\n");
    std::map<std::string, std::string> kv;
    std::string line;
    while (std::getline(f, line)) {
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        const size_t c = line.find(':');
        if (c == std::string::npos) continue;
        std::string key = line.substr(0, c), val = line.substr(c + 1);
        key.erase(0, key.find_first_not_of(" \t")); key.erase(key.find_last_not_of(" \t") + 1);
        val.erase(0, val.find_first_not_of(" \t")); if (!val.empty()) val.erase(val.find_last_not_of(" \t\r") + 1);
        kv[key] = val;
    }
    auto num = [&](const char *k, double dflt) { auto it = kv.find(k); return it == kv.end() || it->second.empty() ? dflt : atof(it->second.c_str()); };
    p_.width = (int)num("Camera.width", p_.width); p_.height = (int)num("Camera.height", p_.height);
    const char *names[4] = {"fx", "fy", "cx", "cy"};
    for (int i = 0; i < 4; ++i) p_.K[i] = (float)num((std::string("Camera.") + names[i]).c_str(), p_.K[i]);
    p_.do_undistortion = num("flagDoUndistortion", 0) != 0 ? 1 : 0;
    const char *dn[5] = {"k1", "k2", "p1", "p2", "k3"};               // cvD order, mono_vo.cpp:166-172
    for (int i = 0; i < 5; ++i) p_.D[i] = (float)num((std::string("Camera.") + dn[i]).c_str(), 0.0);
    p_.thres_error = (float)num("feature_tracker.thres_error", p_.thres_error);
    p_.thres_bidirection = (float)num("feature_tracker.thres_bidirection", p_.thres_bidirection);
    p_.thres_sampson = (float)num("feature_tracker.thres_sampson", p_.thres_sampson);
    p_.window_size = (int)num("feature_tracker.window_size", p_.window_size);
    p_.max_level = (int)num("feature_tracker.max_level", p_.max_level);
    p_.thres_parallax_deg = (float)num("map_update.thres_parallax", p_.thres_parallax_deg);
    p_.n_bins_u = (int)num("feature_extractor.n_bins_u", p_.n_bins_u);
    p_.n_bins_v = (int)num("feature_extractor.n_bins_v", p_.n_bins_v);
    p_.detector = VO_DETECTOR_ORB;                                    // the reference's extractor
    p_.thres_fastscore = (int)num("feature_extractor.thres_fastscore", p_.thres_fastscore);   // initParams(..., int THRES_FAST, ...)
    p_.pose_strict = (int)num("motion_estimator.pose_strict", 1);     // yaml construction = drop-in use: the reference's arithmetic
    p_.scale_faithful_borders = (int)num("feature_tracker.scale_faithful_borders", 0);
    p_.thres_5p_error = (float)num("motion_estimator.thres_5p_error", p_.thres_5p_error);
    p_.thres_poseba_error = (float)num("motion_estimator.thres_poseba_error", p_.thres_poseba_error);
    p_.thres_overlap_ratio = (float)num("keyframe_update.thres_overlap_ratio", p_.thres_overlap_ratio);
    p_.thres_translation = (float)num("keyframe_update.thres_translation", p_.thres_translation);
    p_.thres_rotation_deg = (float)num("keyframe_update.thres_rotation", p_.thres_rotation_deg);
    p_.n_max_keyframes_in_window = (int)num("keyframe_update.n_max_keyframes_in_window", p_.n_max_keyframes_in_window);
    init();
}

void MonoVO::init()
{
    const int nb = std::max(1, p_.n_bins_u * p_.n_bins_v);
    const int rc = vo_ctx_create(p_.device, p_.width, p_.height, 2, std::max(4 * nb + 4096, 262144), nullptr, &ctx_);
    if (rc != VO_OK) fail(nullptr, rc);          // VO_ERR_NO_DEVICE: there is no CPU fallback
    const int rp = vo_set_pose_mode(ctx_, p_.pose_strict ? VO_POSE_STRICT : VO_POSE_FAST);
    if (rp) fail(ctx_, rp);
    const int rs = vo_set_scale_mode(ctx_, p_.scale_faithful_borders);
    if (rs) fail(ctx_, rs);
    const int rl = vo_lba_reserve(ctx_, (size_t)32 << 20);       // the first local BA then allocates nothing inside the frame loop
    if (rl) fail(ctx_, rl);
    const int rd = vo_set_detector(ctx_, p_.detector, p_.thres_fastscore);
    if (rd) fail(ctx_, rd);
    {   // landmark tables: room for 2^19 landmarks before the first reallocation (see StereoVO::init)
        const size_t cap = (size_t)1 << 19;
        lm_X_.reserve(cap * 3); lm_first_px_.reserve(cap * 2); lm_last_px_.reserve(cap * 2); lm_last_parallax_.reserve(cap);
        lm_tri_.reserve(cap); lm_alive_.reserve(cap); lm_bundled_.reserve(cap);
        lm_last_frame_.reserve(cap); lm_first_frame_.reserve(cap); lm_age_.reserve(cap);
        lm_kf_count_.reserve(cap); lm_kf_first_id_.reserve(cap); lm_kf_first_px_.reserve(2 * cap); lm_kf_last_px_.reserve(2 * cap);
        lm_slot_head_.reserve(cap); lm_seen_stamp_.reserve(cap); lm_lba_slot_.reserve(cap); kf_slot_pool_.reserve(cap);
        frames_.reserve(1 << 16);
    }
    {   // allocate the five-point scratch and load its kernels now, not inside the initialisation frame
        float q0[16], q1[16], R[9], t[3];
        uint8_t m[8];
        for (int i = 0; i < 8; ++i) { q0[2 * i] = 10.f + 37.f * i; q0[2 * i + 1] = 20.f + 11.f * (i % 3); q1[2 * i] = q0[2 * i] + 1.f + 0.3f * i; q1[2 * i + 1] = q0[2 * i + 1] + 0.5f; }
        (void)vo_pose_5point(ctx_, q0, q1, 8, p_.K, p_.thres_5p_error, p_.n_hypotheses, p_.seed, R, t, nullptr, m, nullptr, nullptr);
    }
    if (p_.do_undistortion) {
        const int ru = vo_undistort_init(ctx_, p_.K, p_.D, p_.width, p_.height);
        if (ru) fail(ctx_, ru);
    }
}

MonoVO::~MonoVO() { if (ctx_) vo_ctx_destroy(ctx_); }

long long MonoVO::launchCount() const { return vo_ctx_launch_count(ctx_); }
const std::vector<int> &MonoVO::currentLandmarkIds() const { static const std::vector<int> e; return prev_ ? prev_->lm_ids : e; }
const std::vector<float> &MonoVO::currentPts() const { static const std::vector<float> e; return prev_ ? prev_->pts : e; }
int MonoVO::framePose(int frame_id, float *T) const
{
    if (frame_id < 0 || frame_id >= (int)frames_.size() || !T) return VO_ERR_INVALID_ARG;
    memcpy(T, frames_[frame_id]->Twc, 64);
    return VO_OK;
}

// ------------------------------------------------------------------------------ bookkeeping
void MonoVO::setPose(FrameRec &f, const float *Twc)
{
    memcpy(f.Twc, Twc, 64);
    inv_se3_f(f.Twc, f.Tcw);
}
void MonoVO::setPoseDiff10(FrameRec &f, const float *dT10)
{
    memcpy(f.dT10, dT10, 64);
    inv_se3_f(f.dT10, f.dT01);
}

// Landmark(p, frame) (landmark.cpp:28-52): first observation, age 1
int MonoVO::newLandmarks(int k, const float *pts, const FrameRec &f)
{
    const int base = (int)lm_tri_.size();
    const size_t n = (size_t)base + k;
    lm_X_.resize(n * 3, 0.f);
    lm_tri_.resize(n, 0); lm_alive_.resize(n, 1); lm_bundled_.resize(n, 0);
    lm_last_frame_.resize(n, f.id); lm_first_frame_.resize(n, f.id); lm_age_.resize(n, 1);
    lm_last_parallax_.resize(n, 0.f);
    lm_first_px_.resize(n * 2); lm_last_px_.resize(n * 2);
    memcpy(&lm_first_px_[(size_t)base * 2], pts, (size_t)k * 8);
    memcpy(&lm_last_px_[(size_t)base * 2], pts, (size_t)k * 8);
    lm_kf_count_.resize(n, 0); lm_kf_first_id_.resize(n, -1); lm_kf_first_px_.resize(n * 2, 0.f); lm_kf_last_px_.resize(n * 2, 0.f);
    lm_slot_head_.resize(n, -1);
    lm_first_kf_.resize(n, -1);
    return base;
}

// Landmark::addObservationAndRelatedFrame (landmark.cpp:76-135): parallax of the new observation against the FIRST one,
// through the current poses of the two frames
void MonoVO::addObservations(const int *ids, const float *pts, int k, const FrameRec &f)
{
    const float fxinv = 1.0f / p_.K[0], fyinv = 1.0f / p_.K[1], cx = p_.K[2], cy = p_.K[3];
    int f0_cached = -1;
    float T01[16];
    for (int j = 0; j < k; ++j) {
        const int id = ids[j];
        ++lm_age_[id];
        lm_last_frame_[id] = f.id;
        lm_last_px_[2 * (size_t)id] = pts[2 * j]; lm_last_px_[2 * (size_t)id + 1] = pts[2 * j + 1];
        const int f0 = lm_first_frame_[id];
        if (f0 != f0_cached) { mul4_f(frames_[f0]->Tcw, f.Twc, T01); f0_cached = f0; }
        const float x0[3] = {(lm_first_px_[2 * (size_t)id] - cx) * fxinv, (lm_first_px_[2 * (size_t)id + 1] - cy) * fyinv, 1.0f};
        const float b[3] = {(pts[2 * j] - cx) * fxinv, (pts[2 * j + 1] - cy) * fyinv, 1.0f};
        float x1[3];
        for (int r = 0; r < 3; ++r) x1[r] = (T01[r * 4] * b[0] + T01[r * 4 + 1] * b[1]) + T01[r * 4 + 2] * b[2];
        const float dot = (x0[0] * x1[0] + x0[1] * x1[1]) + x0[2] * x1[2];
        const float n0 = std::sqrt((x0[0] * x0[0] + x0[1] * x0[1]) + x0[2] * x0[2]);
        const float n1 = std::sqrt((x1[0] * x1[0] + x1[1] * x1[1]) + x1[2] * x1[2]);
        float c = dot / (n0 * n1);
        if (c >= 1.0f) c = 0.99999f;
        if (c <= -1.0f) c = -0.99999f;
        lm_last_parallax_[id] = acosf(c);
    }
}

bool MonoVO::checkUpdateRule(const FrameRec &f) const
{
    if (window_.empty()) return true;
    const FrameRec &kf = *window_.back();
    int cnt_tracked = 0;
    for (int id : kf.lm_ids) if (lm_last_frame_[id] == f.id) ++cnt_tracked;
    const float ratio = (float)cnt_tracked / (float)kf.lm_ids.size();
    if (ratio <= p_.thres_overlap_ratio) return true;
    float dT[16];
    mul4_f(kf.Tcw, f.Twc, dT);
    float costheta = (dT[0] + dT[5] + dT[10] - 1.0f) * 0.5f;
    if (costheta >= 0.999999f) costheta = 0.999999f;
    if (costheta <= -0.999999f) costheta = -0.999999f;
    const float rot = acosf(costheta);
    const float dtrans = std::sqrt(dT[3] * dT[3] + dT[7] * dT[7] + dT[11] * dT[11]);
    return rot >= p_.thres_rotation_deg * D2R || dtrans >= p_.thres_translation;
}

void MonoVO::addKeyframe(const FrameRecPtr &f)
{
    f->is_keyframe = true;
    all_keyframes_.push_back(f);
    if ((int)window_.size() == p_.n_max_keyframes_in_window) window_.pop_front();
    window_.push_back(f);
    const size_t n = f->lm_ids.size();
    const int kf_index = (int)all_keyframes_.size() - 1;
    f->kf_index = kf_index;
    f->kf_px.resize(2 * n);
    for (size_t i = 0; i < n; ++i) {
        const int id = f->lm_ids[i];
        const float x = lm_last_px_[2 * (size_t)id], y = lm_last_px_[2 * (size_t)id + 1];     // observations.back()
        f->kf_px[2 * i] = x; f->kf_px[2 * i + 1] = y;
        if (lm_kf_count_[id]++ == 0) { lm_kf_first_id_[id] = f->id; lm_kf_first_px_[2 * (size_t)id] = x; lm_kf_first_px_[2 * (size_t)id + 1] = y; }
        lm_kf_last_px_[2 * (size_t)id] = x; lm_kf_last_px_[2 * (size_t)id + 1] = y;
        if (lm_slot_head_[id] < 0) lm_first_kf_[id] = kf_index;
        kf_slot_pool_.push_back({kf_index, (int)i, lm_slot_head_[id]});
        lm_slot_head_[id] = (int)kf_slot_pool_.size() - 1;
    }
}

// triangulateDLT of the candidates, one device call per distinct frame of the first point (T10 = T1w * Tw0)
int MonoVO::dltGroups(const std::vector<int> &cand, const std::vector<float> &pt0, const std::vector<float> &pt1,
                      const std::vector<int> &f0, const FrameRec &f1, std::vector<float> &X0, std::vector<float> &X1)
{
    const size_t n = cand.size();
    X0.assign(n * 3, 0.f); X1.assign(n * 3, 0.f);
    std::vector<int> groups(f0);
    std::sort(groups.begin(), groups.end());
    groups.erase(std::unique(groups.begin(), groups.end()), groups.end());
    // one device call for the whole batch: the relative pose T10 = T1w * Tw0 of every distinct first-observation frame, and per
    // point the index of its group (the first version made one call -- upload, launch, download, sync -- per distinct frame:
    // 0.3 ms per keyframe)
    std::vector<float> R10s(groups.size() * 9), t10s(groups.size() * 3);
    for (size_t gi = 0; gi < groups.size(); ++gi) {
        float T10[16];
        mul4_f(f1.Tcw, frames_[groups[gi]]->Twc, T10);
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R10s[gi * 9 + r * 3 + c] = T10[r * 4 + c]; t10s[gi * 3 + r] = T10[r * 4 + 3]; }
    }
    std::vector<int> gidx(n);
    for (size_t j = 0; j < n; ++j) gidx[j] = (int)(std::lower_bound(groups.begin(), groups.end(), f0[j]) - groups.begin());
    const int rc = vo_triangulate_dlt_grouped(ctx_, pt0.data(), pt1.data(), (int)n, gidx.data(), (int)groups.size(), R10s.data(), t10s.data(),
                                              p_.K, p_.K, X0.data(), X1.data());
    if (rc) fail(ctx_, rc);
    return (int)groups.size();
}

static inline void to_world(const float *Tw0, const float *X0, float *Xw)
{
    for (int r = 0; r < 3; ++r) Xw[r] = ((Tw0[r * 4] * X0[0] + Tw0[r * 4 + 1] * X0[1]) + Tw0[r * 4 + 2] * X0[2]) + Tw0[r * 4 + 3];
}

// mono_vo.cpp:660-687: first / last observation of every landmark of the frame with enough parallax; depth > 0 only
int MonoVO::reconstructInitial(const FrameRec &f)
{
    const float thr = p_.thres_parallax_deg * D2R;
    std::vector<int> cand, f0;
    std::vector<float> pt0, pt1, X0, X1;
    for (int id : f.lm_ids)
        if (!lm_tri_[id] && lm_last_parallax_[id] >= thr) {
            cand.push_back(id); f0.push_back(lm_first_frame_[id]);
            pt0.push_back(lm_first_px_[2 * (size_t)id]); pt0.push_back(lm_first_px_[2 * (size_t)id + 1]);
            pt1.push_back(lm_last_px_[2 * (size_t)id]); pt1.push_back(lm_last_px_[2 * (size_t)id + 1]);
        }
    if (cand.empty()) return 0;
    dltGroups(cand, pt0, pt1, f0, f, X0, X1);
    int n = 0;
    for (size_t j = 0; j < cand.size(); ++j)
        if (X0[3 * j + 2] > 0.f) {
            to_world(frames_[f0[j]]->Twc, &X0[3 * j], &lm_X_[(size_t)cand[j] * 3]);
            lm_tri_[cand[j]] = 1;
            dirty_.push_back(cand[j]);
            ++n;
        }
    return n;
}

// mono_vo.cpp:1032-1076: first / last KEYFRAME observation, more than two of them, 1-px^2 gates, both depths > 0
int MonoVO::reconstructKeyframe(const FrameRec &f)
{
    const float thr = p_.thres_parallax_deg * D2R;
    std::vector<int> cand, f0;
    std::vector<float> pt0, pt1, X0, X1;
    for (int id : f.lm_ids)
        if (lm_alive_[id] && !lm_tri_[id] && lm_last_parallax_[id] >= thr && lm_kf_count_[id] > 2) {
            cand.push_back(id); f0.push_back(lm_kf_first_id_[id]);
            pt0.push_back(lm_kf_first_px_[2 * (size_t)id]); pt0.push_back(lm_kf_first_px_[2 * (size_t)id + 1]);
            pt1.push_back(lm_kf_last_px_[2 * (size_t)id]); pt1.push_back(lm_kf_last_px_[2 * (size_t)id + 1]);
        }
    if (cand.empty()) return 0;
    dltGroups(cand, pt0, pt1, f0, f, X0, X1);
    int n = 0;
    const float fx = p_.K[0], fy = p_.K[1], cx = p_.K[2], cy = p_.K[3];
    for (size_t j = 0; j < cand.size(); ++j) {
        const float *a = &X0[3 * j], *b = &X1[3 * j];
        const float iz0 = 1.0f / a[2], iz1 = 1.0f / b[2];
        const float d0x = pt0[2 * j] - (fx * a[0] * iz0 + cx), d0y = pt0[2 * j + 1] - (fy * a[1] * iz0 + cy);
        if (d0x * d0x + d0y * d0y > 1.0) continue;
        const float d1x = pt1[2 * j] - (fx * b[0] * iz1 + cx), d1y = pt1[2 * j + 1] - (fy * b[1] * iz1 + cy);
        if (d1x * d1x + d1y * d1y > 1.0) continue;
        if (a[2] > 0.f && b[2] > 0.f) {
            to_world(frames_[f0[j]]->Twc, a, &lm_X_[(size_t)cand[j] * 3]);
            lm_tri_[cand[j]] = 1;
            dirty_.push_back(cand[j]);
            ++n;
        }
    }
    return n;
}

void MonoVO::localBundleAdjustment()
{
    const int NUM_MINIMUM_REQUIRED_KEYFRAMES = 3, NUM_FIX = 2;                     // motion_estimator.cpp:1126-1127
    info_.lba_points = info_.lba_obs = info_.lba_ok = 0;
    if ((int)window_.size() < NUM_MINIMUM_REQUIRED_KEYFRAMES) return;
    const auto t_pack = Clock::now();
    const int nf = (int)window_.size();
    // 1) alive + triangulated landmarks of the window in first-seen order with their observation counts, straight from the
    //    window keyframes' own arrays (every keyframe lists each of its landmarks once)
    std::vector<int> &cand = lba_cand_, &cnt = lba_cnt_, &lms = lba_lms_, &obs_ptr = lba_obs_ptr_, &cursor = lba_obs_cursor_, &obs_frame = lba_obs_frame_;
    cand.clear(); cnt.clear();
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
                    lm_lba_slot_[id] = (int)cand.size();
                    cand.push_back(id);
                    cnt.push_back(0);
                }
                ++cnt[lm_lba_slot_[id]];
            }
    }
    // THRES_MINIMUM_SEEN: fewer than two window observations -> not a BA landmark
    lms.clear(); obs_ptr.clear();
    int n_obs = 0;
    for (size_t j = 0; j < cand.size(); ++j) {
        if (cnt[j] < 2) { lm_lba_slot_[cand[j]] = -1; continue; }
        lm_lba_slot_[cand[j]] = (int)lms.size();
        lms.push_back(cand[j]);
        obs_ptr.push_back(n_obs);
        n_obs += cnt[j];
    }
    if (lms.empty()) return;
    const int n_lm = (int)lms.size();
    obs_ptr.push_back(n_obs);
    cursor.assign(obs_ptr.begin(), obs_ptr.begin() + n_lm);
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
    std::vector<double> &points = lba_points_, &obs_px = lba_obs_px_;
    obs_frame.resize(n_obs); obs_px.resize((size_t)2 * n_obs); points.resize((size_t)3 * n_lm);
    // 2) observations: keyframes in chronological order (landmark.cpp:105-124)
    for (int k = 0; k < nf; ++k) {
        const FrameRec &fr = *window_[k];
        const size_t n = fr.lm_ids.size();
        for (size_t i = 0; i < n; ++i) {
            const int id = fr.lm_ids[i];
            if (lm_seen_stamp_[id] != seen_stamp_) continue;          // stamped in pass 1 <=> triangulated and alive
            const int s = lm_lba_slot_[id];
            if (s < 0) continue;
            const int o = cursor[s]++;
            obs_frame[o] = k;
            obs_px[2 * (size_t)o] = fr.kf_px[2 * i]; obs_px[2 * (size_t)o + 1] = fr.kf_px[2 * i + 1];
        }
    }
    for (int j = 0; j < n_lm; ++j) {
        const int id = lms[j];
        const double Xw[3] = {lm_X_[(size_t)id * 3], lm_X_[(size_t)id * 3 + 1], lm_X_[(size_t)id * 3 + 2]};
        for (int r = 0; r < 3; ++r)
            points[3 * (size_t)j + r] = (Tjw_ref[r * 4] * Xw[0] + Tjw_ref[r * 4 + 1] * Xw[1] + Tjw_ref[r * 4 + 2] * Xw[2] + Tjw_ref[r * 4 + 3]) * inv_scale;
    }
    std::vector<uint8_t> &obs_right = lba_obs_right_;
    obs_right.assign(obs_frame.size(), 0);
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
    for (int i = 0; i < 4; ++i) { pr.K_l[i] = p_.K[i]; pr.K_r[i] = p_.K[i]; }
    for (int i = 0; i < 16; ++i) pr.T_lr[i] = (i % 5 == 0) ? 1.0 : 0.0;
    pr.is_stereo = 0; pr.huber = 0.5; pr.lambda = 0.00001; pr.max_iter = 10;
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
    bool large_update = false;
    for (int k = 0; k < nf; ++k) {                                                // write-back (:631-718)
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

void MonoVO::pushStats(const FrameRec &f, bool keyframe)
{
    if (keyframe) {
        // mono_vo.cpp:1142-1152 refreshes the pose and every map point of EVERY keyframe ever made, each time.  Same values, incrementally:
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
    AlgorithmStatistics::FrameStatistics &sf = stat_.stats_frame.back();
    rowmajor_to_pose(f.Twc, sf.Twc); rowmajor_to_pose(f.Tcw, sf.Tcw);              // :969-974
    rowmajor_to_pose(f.dT10, sf.dT_10); rowmajor_to_pose(f.dT01, sf.dT_01);
    if (p_.record_frame_mappoints)                                                 // :1166-1177
        for (size_t id = 0; id < lm_tri_.size(); ++id)
            if (lm_tri_[id]) {
                Point X;
                for (int r = 0; r < 3; ++r) X(r) = lm_X_[id * 3 + r];
                sf.mappoints.push_back(X);
            }
    // :1184-1185 refreshes every frame's pose; only the window's keyframes can have moved
    if (keyframe)
        for (const auto &kf : window_) rowmajor_to_pose(kf->Twc, stat_.stats_frame[kf->id].Twc);
    stat_.stats_landmark.emplace_back();
    stat_.stats_landmark.back().n_initial = info_.n_in;
    stat_.stats_landmark.back().n_pass_bidirection = info_.counts[0];
    stat_.stats_landmark.back().n_new = info_.n_new;
    stat_.stats_landmark.back().n_final = (int)f.lm_ids.size();
    stat_.stats_execution.emplace_back();
}

// The reference's full refresh, recomputed and compared with the incrementally maintained statistics (test hook).
bool MonoVO::statsConsistent() const
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
void MonoVO::trackImage(const cv::Mat &img, const double & /*timestamp*/)
{
    if (img.empty()) throw std::runtime_error("vo_b200: empty image");
    // A frame is COMMITTED (frame list, image-slot parity, landmark tables) only once its device step has succeeded: a
    // step that throws before that point ("calcPose5PointsAlgorithm() is failed.", a CUDA error) leaves the object exactly
    // as it was, so the caller may feed the next image.  An exception after that point (local BA NaN / "large update!")
    // leaves the tables half updated -- the reference dies there too -- and the object refuses further images.
    if (poisoned_) throw std::runtime_error("vo_b200: MonoVO state is inconsistent after a failed keyframe step; create a new object");
    const int w = img.cols, h = img.rows;
    const auto t_total = Clock::now();
    auto fr = std::make_shared<FrameRec>();
    fr->id = (int)frames_.size();
    ident(fr->Twc); ident(fr->Tcw); ident(fr->dT01); ident(fr->dT10);
    info_ = FrameInfo();
    info_.frame = fr->id;
    const int k = fr->id;
    const int s1 = k % 2, s0 = (k + 1) % 2;
    const int nb = p_.n_bins_u * p_.n_bins_v;

    vo_mono_frame_params fp;
    memset(&fp, 0, sizeof(fp));
    fp.window_size = p_.window_size; fp.max_level = p_.max_level; fp.thres_error = p_.thres_error;
    fp.thres_bidirection = p_.thres_bidirection; fp.thres_sampson = p_.thres_sampson; fp.thres_poseba_error = p_.thres_poseba_error;
    memcpy(fp.K, p_.K, 16);
    fp.use_bundled_only = (int)window_.size() > 5;               // mono_vo.cpp:800
    fp.do_scale_refine = p_.do_scale_refine;
    fp.n_bins_u = p_.n_bins_u; fp.n_bins_v = p_.n_bins_v; fp.det_edge = p_.det_edge; fp.det_min_score = p_.det_min_score;
    fp.thres_5p = p_.thres_5p_error; fp.n_hypotheses = p_.n_hypotheses; fp.seed = p_.seed + (unsigned)k;

    new_p1_.resize((size_t)std::max(nb, 1) * 2); new_p0_.resize((size_t)std::max(nb, 1) * 2);
    float T_wc[16], dT01[16], dT10[16];
    vo_mono_frame_result res;
    memset(&res, 0, sizeof(res));
    res.T_wc = T_wc; res.dT01 = dT01; res.dT10 = dT10; res.new_p1 = new_p1_.data(); res.new_p0 = new_p0_.data();
    res.counts = p_.collect_gate_counts ? info_.counts : nullptr;

    const unsigned char *up = img.data;
    if (p_.do_undistortion) {
        // mono_vo.cpp:509-513: Camera::undistortImage + convertTo(CV_8UC1), on the device; the step finds the image in its slot
        const int rr = vo_upload_image_rectified(ctx_, s1, 0, img.data, w, h, img.step);
        if (rr == VO_ERR_SIZE_MISMATCH) throw std::runtime_error("undistort image: provided image has not the same size as the camera model!\n");   // camera.cpp:166
        if (rr) fail(ctx_, rr);
        up = nullptr;
    }
    bool kf = false;
    if (!prev_) {
        // ---- the very first image (mono_vo.cpp:528-561): extraction only, identity pose, dT10 = [I | (0, 0, -1)]
        int n_det = 0;
        if (up) {
            const int rc0 = vo_upload_image(ctx_, s1, up, w, h, img.step);
            if (rc0) fail(ctx_, rc0);
        }
        const int rc = vo_detect_bucketed(ctx_, s1, nullptr, 0, p_.n_bins_u, p_.n_bins_v, p_.det_edge, p_.det_min_score, new_p1_.data(), std::max(nb, 1), &n_det);
        if (rc) fail(ctx_, rc);
        frames_.push_back(fr);                       // commit point
        poisoned_ = true;
        const int base = newLandmarks(n_det, new_p1_.data(), *fr);
        fr->pts.assign(new_p1_.begin(), new_p1_.begin() + 2 * (size_t)n_det);
        fr->lm_ids.resize(n_det);
        for (int i = 0; i < n_det; ++i) fr->lm_ids[i] = base + i;
        float T_init[16];
        ident(T_init);
        T_init[11] = -1.0f;
        setPoseDiff10(*fr, T_init);
        info_.n_detected = n_det; info_.n_new = n_det;
    } else {
        // ---- LandmarkTracking(pts, pts, lms) keeps the alive landmarks of the previous frame (landmark.cpp:233-270)
        const FrameRec &pv = *prev_;
        const size_t n_prev = pv.lm_ids.size();
        in_ids_.clear(); in_p0_.clear(); in_X_.clear(); in_flags_.clear();
        for (size_t i = 0; i < n_prev; ++i) {
            const int id = pv.lm_ids[i];
            if (!lm_alive_[id]) continue;
            in_ids_.push_back(id);
            in_p0_.push_back(pv.pts[2 * i]); in_p0_.push_back(pv.pts[2 * i + 1]);
            for (int r = 0; r < 3; ++r) in_X_.push_back(lm_X_[(size_t)id * 3 + r]);
            in_flags_.push_back((uint8_t)((lm_tri_[id] ? 1 : 0) | (lm_bundled_[id] ? 2 : 0)));
        }
        const int n = (int)in_ids_.size();
        out_idx_.resize(std::max(n, 1)); out_p1_.resize((size_t)std::max(n, 1) * 2);
        res.index = out_idx_.data(); res.pts1 = out_p1_.data();
        fp.init_mode = initialised_ ? 0 : 1;
        const auto t_step = Clock::now();
        const int rc = vo_mono_frame_step(ctx_, &fp, s0, s1, up, w, h, img.step, n, in_p0_.data(), in_X_.data(), in_flags_.data(),
                                          pv.Twc, pv.dT01, &res);
        if (rc == VO_ERR_MODE) throw std::runtime_error(vo_last_error(ctx_));      // "calcPose5PointsAlgorithm() is failed." (:590 / :940)
        if (rc) fail(ctx_, rc);
        frames_.push_back(fr);                       // commit point
        poisoned_ = true;
        info_.ms_step = ms_since(t_step);
        const int nt = res.n_tracked, m = res.n_new;
        fr->lm_ids.resize((size_t)nt + m);
        fr->pts.resize(2 * ((size_t)nt + m));
        for (int i = 0; i < nt; ++i) fr->lm_ids[i] = in_ids_[out_idx_[i]];
        memcpy(fr->pts.data(), out_p1_.data(), (size_t)nt * 8);
        if (!initialised_) {
            addObservations(fr->lm_ids.data(), out_p1_.data(), nt, *fr);           // :602-603: the frame still has the identity pose
            setPose(*fr, T_wc); setPoseDiff10(*fr, dT10);                          // :611-612
        } else {
            setPose(*fr, T_wc); setPoseDiff10(*fr, dT10);                          // :889-890 / :947-948
            addObservations(fr->lm_ids.data(), out_p1_.data(), nt, *fr);           // :966-967
        }
        // new features: born in the PREVIOUS frame at their back-tracked position, observed in this one (:638-657 / :993-1012)
        const int base = newLandmarks(m, new_p0_.data(), pv);
        for (int i = 0; i < m; ++i) fr->lm_ids[nt + i] = base + i;
        memcpy(fr->pts.data() + 2 * (size_t)nt, new_p1_.data(), (size_t)m * 8);
        addObservations(fr->lm_ids.data() + nt, new_p1_.data(), m, *fr);
        info_.n_in = n; info_.n_tracked = nt; info_.n_detected = res.n_detected; info_.n_new = m; info_.used_5point = res.used_5point;
        if (!initialised_) {
            const auto t_rec = Clock::now();
            info_.n_recon = reconstructInitial(*fr);
            info_.ms_recon = ms_since(t_rec);
            initialised_ = true;
        }
    }
    // ---- keyframe (:1021-1157)
    kf = checkUpdateRule(*fr);
    if (kf) {
        info_.keyframe = 1;
        addKeyframe(fr);
        const auto t_rec = Clock::now();
        info_.n_recon += reconstructKeyframe(*fr);
        info_.ms_recon += ms_since(t_rec);
        localBundleAdjustment();
    }
    const auto t_stats = Clock::now();
    pushStats(*fr, kf);
    info_.ms_stats = ms_since(t_stats);
    info_.ms_total = ms_since(t_total);
    info_.ms_book = info_.ms_total - info_.ms_step - info_.ms_recon - info_.ms_lba_pack - info_.ms_lba_solve - info_.ms_stats;
    stat_.stats_execution.back().time_track = info_.ms_step;
    stat_.stats_execution.back().time_localba = info_.ms_lba_pack + info_.ms_lba_solve;
    stat_.stats_execution.back().time_new = info_.ms_recon;
    stat_.stats_execution.back().time_total = info_.ms_total;
    // the previous frame's pixel / landmark lists are only needed again if it is a keyframe
    if (prev_ && !prev_->is_keyframe) {
        std::vector<float>().swap(prev_->pts);
        std::vector<int>().swap(prev_->lm_ids);
    }
    prev_ = fr;
    poisoned_ = false;
}

// ------------------------------------------------------------------------------ C wrapper
struct vo_mvo { MonoVO *vo; };
static thread_local std::string g_mvo_error;

extern "C" const char *vo_mvo_last_error(void) { return g_mvo_error.c_str(); }

extern "C" int vo_mvo_create(const MonoVO::Parameters *prm, vo_mvo **out)
{
    if (!prm || !out) return VO_ERR_INVALID_ARG;
    try { *out = new vo_mvo{new MonoVO(*prm)}; return VO_OK; }
    catch (const std::exception &e) { g_mvo_error = e.what(); *out = nullptr; return VO_ERR_NO_DEVICE; }
}
extern "C" int vo_mvo_create_from_yaml(const char *dir, vo_mvo **out)
{
    if (!dir || !out) return VO_ERR_INVALID_ARG;
    try { *out = new vo_mvo{new MonoVO("rosbag", dir)}; return VO_OK; }
    catch (const std::exception &e) { g_mvo_error = e.what(); *out = nullptr; return VO_ERR_INVALID_ARG; }
}
extern "C" void vo_mvo_destroy(vo_mvo *s) { if (s) { delete s->vo; delete s; } }
extern "C" int vo_mvo_track(vo_mvo *s, const unsigned char *img, int w, int h, size_t step, double timestamp)
{
    if (!s || !img) return VO_ERR_INVALID_ARG;
    try {
        cv::Mat I(h, w, const_cast<unsigned char *>(img), step);
        s->vo->trackImage(I, timestamp);
        return VO_OK;
    } catch (const std::exception &e) { g_mvo_error = e.what(); return VO_ERR_NAN; }
}
extern "C" int vo_mvo_pose(const vo_mvo *s, float *T)
{
    if (!s || !T || s->vo->getStatistics().stats_frame.empty()) return VO_ERR_INVALID_ARG;
    const PoseSE3 &P = s->vo->getStatistics().stats_frame.back().Twc;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T[r * 4 + c] = P(r, c);
    return VO_OK;
}
extern "C" int vo_mvo_frame_pose(const vo_mvo *s, int frame_id, float *T) { return s ? s->vo->framePose(frame_id, T) : VO_ERR_INVALID_ARG; }
extern "C" int vo_mvo_frame_info(const vo_mvo *s, MonoVO::FrameInfo *out)
{
    if (!s || !out) return VO_ERR_INVALID_ARG;
    *out = s->vo->lastFrameInfo();
    return VO_OK;
}
extern "C" int vo_mvo_tracks(const vo_mvo *s, int cap, int *ids, float *pts)
{
    if (!s) return VO_ERR_INVALID_ARG;
    const auto &id = s->vo->currentLandmarkIds();
    const int n = std::min<int>(cap, (int)id.size());
    if (ids) memcpy(ids, id.data(), (size_t)n * 4);
    if (pts) memcpy(pts, s->vo->currentPts().data(), (size_t)n * 8);
    return (int)id.size();
}
extern "C" int vo_mvo_stats_consistent(const vo_mvo *s) { return s ? (s->vo->statsConsistent() ? 1 : 0) : 0; }
extern "C" long long vo_mvo_launch_count(const vo_mvo *s) { return s ? s->vo->launchCount() : 0; }
// layout check for language bindings: 0 = sizeof(Parameters), 1 = sizeof(FrameInfo)
extern "C" int vo_mvo_struct_size(int which) { return which == 0 ? (int)sizeof(MonoVO::Parameters) : (int)sizeof(MonoVO::FrameInfo); }
