// stereo_step.cu -- S1: the steady-state tracking step of StereoVO::trackStereoImages, device resident.
//
// Restates core/visual_odometry/stereo_vo/stereo_vo.cpp:475-670 (steps [2]..[7]) as ONE asynchronous
// sequence on the context's stream -- one H2D of the two new images and the landmark state, one D2H
// of the pose and the surviving tracks, one synchronisation:
//   [2,3] constant-velocity prior T_wc = T_wp * dT_pc_prev; per-landmark prior pixels in l1 / r1 with the
//         3-px inImage margin and the z < 0.1 fallbacks, patch scale z_l0 / z_l1            (:475-522)
//   [4]   trackWithPrior(I0_l -> I1_l)                                                        (:533)
//   [4-1] trackWithScale(I0_l -> I1_l)                                                        (:553)
//   [5]   trackWithPrior(I1_l -> I1_r)                                                        (:566)
//   [6]   stereo pose-only GN on the triangulated survivors, Xp = T_pw * X                    (:595-646)
//   [7]   the "sampson" stub (drop y > 660)                                                   (:657-670)
// The reference physically compacts the track arrays after every gate (StereoLandmarkTracking(src, mask),
// landmark.cpp:291-332). Every stage is per-feature independent except the pose solve, whose input is the
// stable compaction of (mask AND triangulated); so the arrays keep their original indexing with a
// running mask, the pose input is compacted once (same order as the reference) and the final survivor
// list is compacted once at the end -- results and indexing are identical, with two scans instead of five
// reallocations.  Compiled with -fmad=false (FP32 glue arithmetic in the reference's operation order).
#include "vo_internal.cuh"
#include "step_device.cuh"
#include "tri_device.cuh"

#include <cstring>

struct StepDev {
    int n, w, h;
    const float2 *pts_l0, *pts_r0;
    const float *Xw;
    const uint8_t *tri;
    float2 *pts_l1, *pts_r1;     // priors, then tracked positions
    float *scale;
    uint8_t *mask;
    // pose-GN compacted inputs
    float *Xp;
    float2 *pl, *pr;
    int *idx_po, *n_po;
    uint8_t *mask_po;
    float *T01;                  // dT_pc (in-out of the GN)
    int *po_success;
    // outputs
    int *idx_out, *n_out;
    float2 *out_l1, *out_r1;
    float *T_wc;
    int *counts;                 // [5]: after l0l1, scale, l1r1, pose, final
    float T_cw_prior[12], T_pw[12], T_rl[12], T_wp[16];
    float K_l[4], K_r[4];
    float sampson_y;
};

// stereo_vo.cpp:485-522
__global__ void __launch_bounds__(256) k_step_prior(const StepDev d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    float2 p1 = d.pts_l0[i], q1 = d.pts_r0[i];
    float scale = 1.0f;
    if (d.tri[i]) {
        const float X[3] = {d.Xw[3 * i], d.Xw[3 * i + 1], d.Xw[3 * i + 2]};
        float Xl1[3], Xr1[3], Xl0[3];
        xform(d.T_cw_prior, X, Xl1);
        xform(d.T_rl, Xl1, Xr1);
        xform(d.T_pw, X, Xl0);
        scale = Xl0[2] / Xl1[2];
        const float izl = 1.0f / Xl1[2], izr = 1.0f / Xr1[2];
        const float2 pl = make_float2(d.K_l[0] * Xl1[0] * izl + d.K_l[2], d.K_l[1] * Xl1[1] * izl + d.K_l[3]);
        const float2 pr = make_float2(d.K_r[0] * Xr1[0] * izr + d.K_r[2], d.K_r[1] * Xr1[1] * izr + d.K_r[3]);
        const float off = 3.0f, cw = (float)d.w - off, ch = (float)d.h - off;
        const bool in_l = !(pl.x < off || pl.y < off || pl.x >= cw || pl.y >= ch);
        const bool in_r = !(pr.x < off || pr.y < off || pr.x >= cw || pr.y >= ch);
        if (!(!in_l || !in_r || Xl1[2] < 0.1 || Xr1[2] < 0.1)) { p1 = pl; q1 = pr; }
    }
    d.pts_l1[i] = p1;
    d.pts_r1[i] = q1;
    d.scale[i] = scale;
    d.mask[i] = 1;
}

// [6] input: stable compaction of (mask AND triangulated), Xp = T_pw * X  (stereo_vo.cpp:595-614)
__global__ void __launch_bounds__(1024) k_step_select(const StepDev d)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < d.n; c0 += 1024) {
        const int i = c0 + threadIdx.x;
        const bool keep = i < d.n && d.mask[i] && d.tri[i];
        const int pos = scan_chunk(keep, s_warp, &s_base);
        if (keep) {
            const float X[3] = {d.Xw[3 * i], d.Xw[3 * i + 1], d.Xw[3 * i + 2]};
            float Xp[3];
            xform(d.T_pw, X, Xp);
            d.Xp[3 * pos] = Xp[0]; d.Xp[3 * pos + 1] = Xp[1]; d.Xp[3 * pos + 2] = Xp[2];
            d.pl[pos] = d.pts_l1[i];
            d.pr[pos] = d.pts_r1[i];
            d.idx_po[pos] = i;
        }
    }
    if (threadIdx.x == 0) *d.n_po = s_base;
}

// [6] scatter of the inlier mask (:631-638), T_wc = T_wp * dT (:640), [7] the y > 660 stub (:657-668),
// final stable compaction of the survivors.
__global__ void __launch_bounds__(1024) k_step_finish(const StepDev d)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int tid = threadIdx.x;
    const int n_po = *d.n_po;
    for (int k = tid; k < n_po; k += 1024) d.mask[d.idx_po[k]] = d.mask_po[k] ? 1 : 0;
    if (tid == 0) {
        s_base = 0;
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                float s = 0.f;
                for (int k = 0; k < 4; ++k) s += d.T_wp[r * 4 + k] * d.T01[k * 4 + c];
                d.T_wc[r * 4 + c] = s;
            }
    }
    __syncthreads();
    for (int c0 = 0; c0 < d.n; c0 += 1024) {
        const int i = c0 + tid;
        bool keep = i < d.n && d.mask[i];
        if (keep && d.pts_l1[i].y > d.sampson_y) keep = false;
        const int pos = scan_chunk(keep, s_warp, &s_base);
        if (keep) { d.idx_out[pos] = i; d.out_l1[pos] = d.pts_l1[i]; d.out_r1[pos] = d.pts_r1[i]; }
    }
    if (tid == 0) { *d.n_out = s_base; d.counts[4] = s_base; d.counts[3] = n_po; }
}

// ------------------------------------------------------------------------------ new features / reconstruction
// [10] stereo_vo.cpp:706-739 (and the first frame, :871-903): after the bidirectional L->R match of the freshly
// detected points, keep those whose mask survived and -- in the steady state -- whose two DLT depths are positive;
// stable compaction.  One CTA (a frame adds at most n_bins points).
struct NewDev {
    const float2 *pl, *pr;
    const uint8_t *mask;
    const int *n_in;             // device count of detected points
    int depth_gate;
    float R_rl[9], t_rl[3], K_l[4], K_r[4];
    float2 *out_l, *out_r;
    int *n_out;
};

// depth gate of the new features (stereo_vo.cpp:723-725): one thread per candidate, many CTAs (a 4x4 Jacobi SVD per
// point is ~2 k dependent instructions; inside the single compaction CTA it was 22 us of the frame)
__global__ void __launch_bounds__(128) k_new_depth(const NewDev d, uint8_t *mask)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *d.n_in || !mask[i]) return;
    float Xl[3], Xr[3];
    tri_point(d.pl[i], d.pr[i], d.R_rl, d.t_rl, d.K_l, d.K_r, Xl, Xr);
    if (!(Xl[2] > 0.f && Xr[2] > 0.f)) mask[i] = 0;
}

__global__ void __launch_bounds__(1024) k_new_gate(const NewDev d)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int n = *d.n_in;
    for (int c0 = 0; c0 < n; c0 += 1024) {
        const int i = c0 + threadIdx.x;
        const bool keep = i < n && d.mask[i];
        const int pos = scan_chunk(keep, s_warp, &s_base);
        if (keep) { d.out_l[pos] = d.pl[i]; d.out_r[pos] = d.pr[i]; }
    }
    if (threadIdx.x == 0) *d.n_out = s_base;
}

// Keyframe / first-frame reconstruction (stereo_vo.cpp:767-797, :911-941): DLT, 1-px^2 reprojection gates on both
// images, both depths positive, X_w = T_wc * X_l.
struct ReconDev {
    const float2 *pl, *pr;
    int n;
    float R_rl[9], t_rl[3], K_l[4], K_r[4], T_wc[12];
    float *Xw;
    uint8_t *ok;
};

__global__ void __launch_bounds__(128) k_reconstruct(const ReconDev d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    const float2 p0 = d.pl[i], p1 = d.pr[i];
    float Xl[3], Xr[3];
    tri_point(p0, p1, d.R_rl, d.t_rl, d.K_l, d.K_r, Xl, Xr);
    bool ok = true;
    {
        const float invz = 1.0f / Xl[2];                                   // Camera::projectToPixel (camera.cpp:208-213)
        const float dx = p0.x - (d.K_l[0] * Xl[0] * invz + d.K_l[2]), dy = p0.y - (d.K_l[1] * Xl[1] * invz + d.K_l[3]);
        if (dx * dx + dy * dy > 1.0f) ok = false;
    }
    {
        const float invz = 1.0f / Xr[2];
        const float dx = p1.x - (d.K_r[0] * Xr[0] * invz + d.K_r[2]), dy = p1.y - (d.K_r[1] * Xr[1] * invz + d.K_r[3]);
        if (dx * dx + dy * dy > 1.0f) ok = false;
    }
    if (!(Xl[2] > 0.f && Xr[2] > 0.f)) ok = false;
    float Xw[3];
    xform(d.T_wc, Xl, Xw);
    d.Xw[3 * i] = Xw[0]; d.Xw[3 * i + 1] = Xw[1]; d.Xw[3 * i + 2] = Xw[2];
    d.ok[i] = ok ? 1 : 0;
}

static void inv_se3_f(const float *T, float *O)
{
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) O[i * 4 + j] = T[j * 4 + i];
        float s = 0.f;
        for (int k = 0; k < 3; ++k) s += T[k * 4 + i] * T[k * 4 + 3];
        O[i * 4 + 3] = -s;
    }
    O[12] = O[13] = O[14] = 0.f; O[15] = 1.f;
}
static void mul4_f(const float *A, const float *B, float *C)
{
    float T[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { float s = 0.f; for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j]; T[i * 4 + j] = s; }
    memcpy(C, T, sizeof(T));
}
static size_t a16(size_t v) { return (v + 15) / 16 * 16; }

int vo_detect_launch_d(vo_ctx *ctx, int slot, const float *occ_d, const int *n_occ_d, int n_occ, int n_bins_u, int n_bins_v,
                       int edge, long long min_score, float *out_d, uint8_t *out_mask_d, int *n_out_d, int max_out);

extern "C" int vo_stereo_reconstruct(vo_ctx *ctx, const float *pts_l, const float *pts_r, int n, const float *K_l4, const float *K_r4,
                                     const float *T_lr, const float *T_wc, float *Xw_out, uint8_t *ok_out)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0, VO_ERR_INVALID_ARG, "negative size");
    if (n == 0) return VO_OK;
    VO_REQUIRE(pts_l && pts_r && K_l4 && K_r4 && T_lr && T_wc && Xw_out && ok_out, VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    const size_t N = (size_t)n, o_l = 0, o_r = N * 8, o_X = N * 16, o_ok = o_X + a16(N * 12), total = o_ok + a16(N);
    int rc = vo_stage_reserve(ctx, total);
    if (rc) return rc;
    uint8_t *h = ctx->h_stage, *dv = ctx->d_stage;
    memcpy(h + o_l, pts_l, N * 8); memcpy(h + o_r, pts_r, N * 8);
    VO_CUDA(cudaMemcpyAsync(dv, h, N * 16, cudaMemcpyHostToDevice, ctx->stream));
    ReconDev d;
    d.pl = (const float2 *)(dv + o_l); d.pr = (const float2 *)(dv + o_r); d.n = n;
    float T_rl[16];
    inv_se3_f(T_lr, T_rl);
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) d.R_rl[r * 3 + c] = T_rl[r * 4 + c]; d.t_rl[r] = T_rl[r * 4 + 3]; }
    memcpy(d.K_l, K_l4, 16); memcpy(d.K_r, K_r4, 16); memcpy(d.T_wc, T_wc, 48);
    d.Xw = (float *)(dv + o_X); d.ok = dv + o_ok;
    k_reconstruct<<<vo_div_up(n, 128), 128, 0, ctx->stream>>>(d);
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    VO_CUDA(cudaMemcpyAsync(h + o_X, dv + o_X, total - o_X, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(Xw_out, h + o_X, N * 12);
    memcpy(ok_out, h + o_ok, N);
    return VO_OK;
}

extern "C" int vo_stereo_frame_step(vo_ctx *ctx, const vo_stereo_frame_params *fp, int slot_l0, int slot_l1, int slot_r1,
                                    const uint8_t *img_l1, const uint8_t *img_r1, int w, int h, size_t step, int n,
                                    const float *pts_l0, const float *pts_r0, const float *Xw, const uint8_t *triangulated,
                                    const float *T_wp, const float *dT_pc_prev, vo_stereo_frame_result *res)
{
    if (!ctx || !fp || !res) return VO_ERR_INVALID_ARG;
    const vo_stereo_step_params *prm = &fp->track;
    VO_REQUIRE(n >= 0, VO_ERR_INVALID_ARG, "negative size");
    VO_REQUIRE(n == 0 || (T_wp && dT_pc_prev && res->T_wc && res->dT_pc), VO_ERR_INVALID_ARG, "null pointer");
    VO_REQUIRE(n == 0 || (pts_l0 && pts_r0 && Xw && triangulated && res->index && res->pts_l1 && res->pts_r1), VO_ERR_INVALID_ARG, "null pointer");
    const int nb = fp->n_bins_u > 0 && fp->n_bins_v > 0 ? fp->n_bins_u * fp->n_bins_v : 0;
    VO_REQUIRE(nb == 0 || (res->new_l1 && res->new_r1), VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if (img_l1) { rc = vo_upload_image(ctx, slot_l1, img_l1, w, h, step); if (rc) return rc; }
    if (img_r1) { rc = vo_upload_image(ctx, slot_r1, img_r1, w, h, step); if (rc) return rc; }
    VO_REQUIRE(slot_l1 >= 0 && slot_l1 < ctx->n_slots && ctx->slots[slot_l1].w == w && ctx->slots[slot_l1].h == h &&
               slot_r1 >= 0 && slot_r1 < ctx->n_slots && ctx->slots[slot_r1].w == w && ctx->slots[slot_r1].h == h,
               VO_ERR_INVALID_ARG, "current-frame slots have no image of this size");
    if (n > 0)
        VO_REQUIRE(slot_l0 >= 0 && slot_l0 < ctx->n_slots && ctx->slots[slot_l0].w == w && ctx->slots[slot_l0].h == h, VO_ERR_INVALID_ARG,
                   "previous-left slot has no image of this size");
    const size_t N = (size_t)n, NB = (size_t)nb;
    // staging: inputs [pts_l0][pts_r0][Xw][tri] | work [pts_l1][pts_r1][scale][mask][Xp][pl][pr][idx_po][mask_po]
    //          | new-feature work [cand NB*8][cand_r NB*8][back NB*8][err NB*4][errb NB*4][st NB][stb NB][mask NB]
    //          | results [T01 16f][T_wc 16f][ints 16][idx_out][out_l1][out_r1][new_l NB*8][new_r NB*8]
    size_t o = 0;
    const size_t o_l0 = o; o += N * 8; const size_t o_r0 = o; o += N * 8; const size_t o_X = o; o += a16(N * 12);
    const size_t o_tri = o; o += a16(N);
    const size_t in_bytes = o;
    const size_t o_l1 = o; o += N * 8; const size_t o_r1 = o; o += N * 8; const size_t o_sc = o; o += a16(N * 4);
    const size_t o_m = o; o += a16(N); const size_t o_Xp = o; o += a16(N * 12); const size_t o_pl = o; o += N * 8;
    const size_t o_pr = o; o += N * 8; const size_t o_ip = o; o += a16(N * 4); const size_t o_mp = o; o += a16(N);
    const size_t o_c = o; o += NB * 8; const size_t o_cr = o; o += NB * 8; const size_t o_cb = o; o += NB * 8;
    const size_t o_ce = o; o += a16(NB * 4); const size_t o_ceb = o; o += a16(NB * 4); const size_t o_cs = o; o += a16(NB);
    const size_t o_csb = o; o += a16(NB); const size_t o_cm = o; o += a16(NB);
    const size_t o_res = o;
    const size_t o_T01 = o; o += 64; const size_t o_Twc = o; o += 64; const size_t o_int = o; o += 64;
    const size_t o_io = o; o += a16(N * 4); const size_t o_ol = o; o += N * 8; const size_t o_or = o; o += N * 8;
    const size_t o_nl = o; o += NB * 8; const size_t o_nr = o; o += NB * 8;
    const size_t total = o;
    rc = vo_stage_reserve(ctx, total);
    if (rc) return rc;
    uint8_t *hs = ctx->h_stage, *dv = ctx->d_stage;
    int *ints = (int *)(dv + o_int);      // 0 n_po, 1 po_success, 2 n_out, 3 nan, 4..8 counts, 9 n_detected, 10 n_new
    memset(hs + o_T01, 0, 192);
    if (n > 0) {
        memcpy(hs + o_l0, pts_l0, N * 8); memcpy(hs + o_r0, pts_r0, N * 8); memcpy(hs + o_X, Xw, N * 12); memcpy(hs + o_tri, triangulated, N);
        memcpy(hs + o_T01, dT_pc_prev, 64);           // the GN starts from the previous motion (:586)
        VO_CUDA(cudaMemcpyAsync(dv, hs, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    VO_CUDA(cudaMemcpyAsync(dv + o_T01, hs + o_T01, 192, cudaMemcpyHostToDevice, ctx->stream));

    {   // both new images' pyramids (and Scharr planes) in ONE batched launch set
        const int both[2] = {slot_l1, slot_r1};
        const int eff = vo_effective_max_level(w, h, prm->window_size, prm->max_level);
        rc = vo_ensure_pyramids(ctx, both, 2, (eff + 1 < ctx->max_levels ? eff + 1 : ctx->max_levels), 1);
        if (rc) return rc;
    }
    StepDev d;
    memset(&d, 0, sizeof(d));
    d.n = n; d.w = w; d.h = h;
    d.out_l1 = (float2 *)(dv + o_ol); d.out_r1 = (float2 *)(dv + o_or); d.n_out = ints + 2;
    if (n > 0) {
        d.pts_l0 = (const float2 *)(dv + o_l0); d.pts_r0 = (const float2 *)(dv + o_r0); d.Xw = (const float *)(dv + o_X); d.tri = dv + o_tri;
        d.pts_l1 = (float2 *)(dv + o_l1); d.pts_r1 = (float2 *)(dv + o_r1); d.scale = (float *)(dv + o_sc); d.mask = dv + o_m;
        d.Xp = (float *)(dv + o_Xp); d.pl = (float2 *)(dv + o_pl); d.pr = (float2 *)(dv + o_pr); d.idx_po = (int *)(dv + o_ip);
        d.mask_po = dv + o_mp; d.T01 = (float *)(dv + o_T01); d.T_wc = (float *)(dv + o_Twc);
        d.n_po = ints + 0; d.po_success = ints + 1; d.counts = ints + 4;
        int *nan_flag = ints + 3;
        d.idx_out = (int *)(dv + o_io);
        float T_wc_prior[16], T_cw_prior[16], T_pw[16], T_rl[16];
        mul4_f(T_wp, dT_pc_prev, T_wc_prior);          // :478
        inv_se3_f(T_wc_prior, T_cw_prior);             // :479 geometry::inverseSE3_f
        inv_se3_f(T_wp, T_pw);                         // Frame::getPoseInv()
        inv_se3_f(prm->T_lr, T_rl);
        memcpy(d.T_cw_prior, T_cw_prior, 48); memcpy(d.T_pw, T_pw, 48); memcpy(d.T_rl, T_rl, 48); memcpy(d.T_wp, T_wp, 64);
        memcpy(d.K_l, prm->K_l, 16); memcpy(d.K_r, prm->K_r, 16);
        d.sampson_y = prm->sampson_y;
        const bool want_counts = res->counts != nullptr;
        k_step_prior<<<vo_div_up(n, 256), 256, 0, ctx->stream>>>(d);
        ctx->launches++;
        if (!want_counts) {
            // [4], [4-1], [5] as ONE launch: each feature runs its three dependent stages back to back on its warp
            rc = vo_track_chain_launch_d(ctx, slot_l0, slot_l1, slot_r1, (const float *)d.pts_l0, (float *)d.pts_l1, (float *)d.pts_r1, d.scale,
                                         d.mask, nan_flag, n, prm->window_size, prm->max_level, prm->thres_error, prm->do_scale_refine);
            if (rc) return rc;
        } else {
            // [4] l0 -> l1
            KltPost post{};
            post.mode = 2; post.thres_err = prm->thres_error; post.mask = d.mask; post.skip_masked = 1;
            rc = vo_klt_launch(ctx, 1, &slot_l0, &slot_l1, (const float *)d.pts_l0, n, prm->window_size, prm->max_level, VO_KLT_USE_INITIAL_FLOW,
                               (float *)d.pts_l1, nullptr, nullptr, nullptr, &post);
            if (rc) return rc;
            if (want_counts) { k_step_count<<<1, 1024, 0, ctx->stream>>>(d.mask, n, d.counts + 0); ctx->launches++; }
            // [4-1] scale refinement
            if (prm->do_scale_refine) {
                rc = vo_klt_scale_launch_d(ctx, slot_l0, slot_l1, (const float *)d.pts_l0, d.scale, n, (float *)d.pts_l1, d.mask, nan_flag);
                if (rc) return rc;
            }
            if (want_counts) { k_step_count<<<1, 1024, 0, ctx->stream>>>(d.mask, n, d.counts + 1); ctx->launches++; }
            // [5] l1 -> r1
            rc = vo_klt_launch(ctx, 1, &slot_l1, &slot_r1, (const float *)d.pts_l1, n, prm->window_size, prm->max_level, VO_KLT_USE_INITIAL_FLOW,
                               (float *)d.pts_r1, nullptr, nullptr, nullptr, &post);
            if (rc) return rc;
            if (want_counts) { k_step_count<<<1, 1024, 0, ctx->stream>>>(d.mask, n, d.counts + 2); ctx->launches++; }
        }
        // [6] pose-only GN on the triangulated survivors
        k_step_select<<<1, 1024, 0, ctx->stream>>>(d);
        ctx->launches++;
        rc = vo_pose_launch_d(ctx, 1, nullptr, n, d.n_po, d.Xp, (const float *)d.pl, (const float *)d.pr, prm->K_l, prm->K_r, prm->T_lr,
                              prm->thres_poseba_error, 0, 0, d.T01, d.mask_po, d.po_success, nullptr);
        if (rc) return rc;
        k_step_finish<<<1, 1024, 0, ctx->stream>>>(d);
        ctx->launches++;
    }
    if (nb > 0) {
        // [10] new features from the empty bins: detect on l1, bidirectional l1 <-> r1 match, depth gate, compaction
        rc = vo_detect_launch_d(ctx, slot_l1, (const float *)d.out_l1, n > 0 ? d.n_out : nullptr, n, fp->n_bins_u, fp->n_bins_v, fp->det_edge,
                                fp->det_min_score, (float *)(dv + o_c), dv + o_cm, ints + 9, nb);
        if (rc) return rc;
        // trackBidirection(l1 -> r1) of the detected points (feature_tracker.cpp:39-86): forward + backward pass of a
        // feature back to back on its warp, validity test fused in the backward epilogue
        rc = vo_bidir_chain_launch_d(ctx, slot_l1, slot_r1, (const float *)(dv + o_c), (float *)(dv + o_cr), (float *)(dv + o_cb), dv + o_cs,
                                     dv + o_csb, (float *)(dv + o_ce), (float *)(dv + o_ceb), dv + o_cm, 1, nb, prm->window_size,
                                     prm->max_level, prm->thres_error, fp->thres_bidirection, 0, nullptr, nullptr);
        if (rc) return rc;
        NewDev nd;
        nd.pl = (const float2 *)(dv + o_c); nd.pr = (const float2 *)(dv + o_cr); nd.mask = dv + o_cm; nd.n_in = ints + 9;
        nd.depth_gate = fp->new_depth_gate;
        float T_rl[16];
        inv_se3_f(prm->T_lr, T_rl);
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) nd.R_rl[r * 3 + c] = T_rl[r * 4 + c]; nd.t_rl[r] = T_rl[r * 4 + 3]; }
        memcpy(nd.K_l, prm->K_l, 16); memcpy(nd.K_r, prm->K_r, 16);
        nd.out_l = (float2 *)(dv + o_nl); nd.out_r = (float2 *)(dv + o_nr); nd.n_out = ints + 10;
        if (nd.depth_gate) { k_new_depth<<<vo_div_up(nb, 128), 128, 0, ctx->stream>>>(nd, dv + o_cm); ctx->launches++; }
        k_new_gate<<<1, 1024, 0, ctx->stream>>>(nd);
        ctx->launches++;
    }
    VO_CUDA(cudaGetLastError());
    VO_CUDA(cudaMemcpyAsync(hs + o_res, dv + o_res, total - o_res, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    const int *hi = (const int *)(hs + o_int);
    res->n_tracked = 0; res->n_new = 0; res->n_detected = 0;
    if (n > 0) {
        if (res->counts) for (int k = 0; k < 5; ++k) res->counts[k] = hi[4 + k];
        if (hi[3]) { ctx->last_error = "ax ay nan (feature_tracker.cpp:414)"; return VO_ERR_NAN; }
        if (!hi[1]) { ctx->last_error = "PoseOnlyStereoBA is failed!"; return VO_ERR_NAN; }   // stereo_vo.cpp:624-627
        const int k_out = hi[2];
        memcpy(res->dT_pc, hs + o_T01, 64);
        memcpy(res->T_wc, hs + o_Twc, 64);
        res->n_tracked = k_out;
        memcpy(res->index, hs + o_io, (size_t)k_out * 4);
        memcpy(res->pts_l1, hs + o_ol, (size_t)k_out * 8);
        memcpy(res->pts_r1, hs + o_or, (size_t)k_out * 8);
    } else if (T_wp && dT_pc_prev && res->T_wc && res->dT_pc) {
        // nothing to track: the reference would run the GN on zero points and keep the prior
        mul4_f(T_wp, dT_pc_prev, res->T_wc);
        memcpy(res->dT_pc, dT_pc_prev, 64);
        if (res->counts) for (int k = 0; k < 5; ++k) res->counts[k] = 0;
    }
    if (nb > 0) {
        res->n_detected = hi[9];
        res->n_new = hi[10];
        memcpy(res->new_l1, hs + o_nl, (size_t)hi[10] * 8);
        memcpy(res->new_r1, hs + o_nr, (size_t)hi[10] * 8);
    }
    return VO_OK;
}

extern "C" int vo_stereo_track_step(vo_ctx *ctx, const vo_stereo_step_params *prm, int slot_l0, int slot_l1, int slot_r1,
                                    const uint8_t *img_l1, const uint8_t *img_r1, int w, int h, size_t step, int n,
                                    const float *pts_l0, const float *pts_r0, const float *Xw, const uint8_t *triangulated,
                                    const float *T_wp, const float *dT_pc_prev, float *T_wc_out, float *dT_pc_out, int *n_out,
                                    int *index_out, float *pts_l1_out, float *pts_r1_out, int *counts_out)
{
    if (!ctx || !prm) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(T_wp && dT_pc_prev && T_wc_out && dT_pc_out && n_out, VO_ERR_INVALID_ARG, "null pointer");
    vo_stereo_frame_params fp;
    memset(&fp, 0, sizeof(fp));
    fp.track = *prm;
    vo_stereo_frame_result res;
    memset(&res, 0, sizeof(res));
    res.T_wc = T_wc_out; res.dT_pc = dT_pc_out; res.index = index_out; res.pts_l1 = pts_l1_out; res.pts_r1 = pts_r1_out; res.counts = counts_out;
    const int rc = vo_stereo_frame_step(ctx, &fp, slot_l0, slot_l1, slot_r1, img_l1, img_r1, w, h, step, n, pts_l0, pts_r0, Xw, triangulated,
                                        T_wp, dT_pc_prev, &res);
    if (rc) return rc;
    *n_out = res.n_tracked;
    return VO_OK;
}
