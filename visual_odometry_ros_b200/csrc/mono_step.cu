// mono_step.cu -- S2: the steady-state tracking step of MonoVO::trackImage, device resident.
//
// Restates core/visual_odometry/mono_vo/mono_vo.cpp:724-992 as ONE asynchronous sequence on the context's stream
// (one H2D of the new image and the landmark state, one D2H of the pose and the surviving tracks, one sync):
//   prior      constant-velocity pose Twc_prev * dT01_prior; for BUNDLED landmarks the projected prior pixel (kept only
//              if the depth in the predicted frame is positive) and the patch scale z_prev / z_pred         (:739-761)
//   K4         trackBidirectionWithPrior(I0 -> I1)                                                        (:768)
//   K7         trackWithScale(I0 -> I1)                                                                   (:783)
//   select     landmarks for the pose-only BA: bundled (more than 5 keyframes) or triangulated, depth in the previous
//              frame > 0.1, stable order                                                                  (:799-827)
//   P2         mono pose-only GN from dT01_prior                                                          (:860-866)
//   finish     inlier-mask scatter, dT10 = inverseSE3_f(dT01), T_wc = Twc_prev * dT01, F10 = Kinv^T [t10]x R10 Kinv,
//              Sampson gate (motion_estimator.cpp:539-568), final stable compaction                       (:870-962)
//   new        bucketed detection on I1 with the survivors as occupancy, trackBidirection(I1 -> I0)       (:981-992)
//   fallback   when fewer than 11 points are selected or the GN fails, the reference runs calcPose5PointsAlgorithm on all
//              K7 survivors and scales the unit translation to the previous motion's length (:909-949): here the host
//              sees the GN outcome at the end-of-step sync and, in that rare case, queues five_point.cu + the same tail
//              (finish, new features) once more -- the common path keeps its single synchronisation
//   init       the second image of a sequence (:562-659): K1 track, the same five-point stage with |t| = 1, Sampson
//              gate, new features -- prm->init_mode
// Compiled with -fmad=false (FP32 glue arithmetic in the reference's order).
#include "vo_internal.cuh"
#include "step_device.cuh"

#include <cstring>

struct MonoDev {
    int n, w, h;
    const float2 *pts0;
    const float *Xw;
    const uint8_t *flags;        // bit 0 triangulated, bit 1 bundled
    float2 *pts1;                // priors, then tracked positions
    float *scale;
    uint8_t *mask;
    float *Xp;                   // compacted GN input
    float2 *p1;
    int *idx_po, *n_po;
    uint8_t *mask_po;
    float *T01;                  // dT01 (in-out of the GN), 4x4 row-major
    int *po_success;
    int *idx_out, *n_out;
    float2 *out1;
    float *T_wc, *dT10;
    int *counts;                 // [5]: after K4, after K7, GN points, after motion, final
    float Tcw_prev[12], Tcw_prior[12], Twc_prev[16], K[4];
    int use_bundled_only;
    float thres_sampson;
    int force;                   // tail after the five-point stage: dT01 / dT10 are given, the GN outcome is not consulted
    const float *R10, *t10;      // five-point outputs
    const int *fp_info;
    float t_scale;
};

// mono_vo.cpp:739-761
__global__ void __launch_bounds__(256) k_mono_prior(const MonoDev d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    float2 p1 = d.pts0[i];
    float scale = 1.0f;
    if (d.flags[i] & 2) {
        const float X[3] = {d.Xw[3 * i], d.Xw[3 * i + 1], d.Xw[3 * i + 2]};
        float Xp[3], Xc[3];
        xform(d.Tcw_prev, X, Xp);
        xform(d.Tcw_prior, X, Xc);
        scale = Xp[2] / Xc[2];
        if (Xc[2] > 0.f) {
            const float invz = 1.0f / Xc[2];
            p1 = make_float2(d.K[0] * Xc[0] * invz + d.K[2], d.K[1] * Xc[1] * invz + d.K[3]);
        }
    }
    d.pts1[i] = p1;
    d.scale[i] = scale;
}

// mono_vo.cpp:799-827 + :846-857
__global__ void __launch_bounds__(1024) k_mono_select(const MonoDev d)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int bit = d.use_bundled_only ? 2 : 1;
    for (int c0 = 0; c0 < d.n; c0 += 1024) {
        const int i = c0 + threadIdx.x;
        bool keep = i < d.n && d.mask[i] && (d.flags[i] & bit);
        float Xp[3] = {0.f, 0.f, 0.f};
        if (keep) {
            const float X[3] = {d.Xw[3 * i], d.Xw[3 * i + 1], d.Xw[3 * i + 2]};
            xform(d.Tcw_prev, X, Xp);
            keep = (double)Xp[2] > 0.1;
        }
        const int pos = scan_chunk(keep, s_warp, &s_base);
        if (keep) {
            d.Xp[3 * pos] = Xp[0]; d.Xp[3 * pos + 1] = Xp[1]; d.Xp[3 * pos + 2] = Xp[2];
            d.p1[pos] = d.pts1[i];
            d.idx_po[pos] = i;
        }
    }
    if (threadIdx.x == 0) { *d.n_po = s_base; d.counts[2] = s_base; }
}

__device__ __forceinline__ void mul3(const float *A, const float *B, float *C)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float s = 0.f;
            for (int k = 0; k < 3; ++k) s += A[i * 3 + k] * B[k * 3 + j];
            C[i * 3 + j] = s;
        }
}

__global__ void __launch_bounds__(1024) k_mono_finish(const MonoDev d)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    __shared__ float s_F[9];
    const int tid = threadIdx.x;
    const int n_po = *d.n_po;
    // the pose-only BA did not run or failed (:836, :868): the motion mask stays all-true for the fallback and nothing
    // of this pass is used
    const bool valid = d.force ? (d.fp_info[2] != 0) : (n_po > 10 && *d.po_success != 0);
    if (!valid) {
        if (tid == 0) { *d.n_out = 0; d.counts[3] = 0; d.counts[4] = 0; }
        return;
    }
    for (int k = tid; k < n_po; k += 1024) d.mask[d.idx_po[k]] = d.mask_po[k] ? 1 : 0;        // :872-879 / :937
    if (tid == 0) {
        s_base = 0;
        const float *T = d.T01;
        float T10[16];
        if (d.force) {
            for (int i = 0; i < 16; ++i) T10[i] = d.dT10[i];                                  // :944-945: dT10 is the primary
        } else {
            for (int i = 0; i < 3; ++i) {                                                     // geometry::inverseSE3_f (:882)
                for (int j = 0; j < 3; ++j) T10[i * 4 + j] = T[j * 4 + i];
                float s = 0.f;
                for (int k = 0; k < 3; ++k) s += T[k * 4 + i] * T[k * 4 + 3];
                T10[i * 4 + 3] = -s;
            }
            T10[12] = T10[13] = T10[14] = 0.f; T10[15] = 1.f;
            for (int i = 0; i < 16; ++i) d.dT10[i] = T10[i];
        }
        for (int r = 0; r < 4; ++r)                                                           // Twc_prev * dT01 (:890)
            for (int c = 0; c < 4; ++c) {
                float s = 0.f;
                for (int k = 0; k < 4; ++k) s += d.Twc_prev[r * 4 + k] * T[k * 4 + c];
                d.T_wc[r * 4 + c] = s;
            }
        // F10 = Kinv^T * (skew(t10) * R10) * Kinv (motion_estimator.cpp:549-551); Kinv = K.inverse() restated as
        // cofactors times 1/det (Eigen::Matrix3f::inverse, third-party)
        const float K[9] = {d.K[0], 0.f, d.K[2], 0.f, d.K[1], d.K[3], 0.f, 0.f, 1.f};
        float cf[9];
        cf[0] = K[4] * K[8] - K[5] * K[7]; cf[1] = K[2] * K[7] - K[1] * K[8]; cf[2] = K[1] * K[5] - K[2] * K[4];
        cf[3] = K[5] * K[6] - K[3] * K[8]; cf[4] = K[0] * K[8] - K[2] * K[6]; cf[5] = K[2] * K[3] - K[0] * K[5];
        cf[6] = K[3] * K[7] - K[4] * K[6]; cf[7] = K[1] * K[6] - K[0] * K[7]; cf[8] = K[0] * K[4] - K[1] * K[3];
        const float det = (K[0] * cf[0] + K[1] * cf[3]) + K[2] * cf[6];
        const float idet = 1.0f / det;
        float Kinv[9], KinvT[9];
        for (int i = 0; i < 9; ++i) Kinv[i] = cf[i] * idet;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) KinvT[i * 3 + j] = Kinv[j * 3 + i];
        const float tx = T10[3], ty = T10[7], tz = T10[11];
        const float S[9] = {0.f, -tz, ty, tz, 0.f, -tx, -ty, tx, 0.f};
        float R10[9], E[9], KE[9], F[9];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R10[i * 3 + j] = T10[i * 4 + j];
        mul3(S, R10, E);
        mul3(KinvT, E, KE);
        mul3(KE, Kinv, F);
        for (int i = 0; i < 9; ++i) s_F[i] = F[i];
    }
    __syncthreads();
    // count after the motion gate, then the Sampson gate and the final stable compaction
    int after_motion = 0;
    for (int c0 = 0; c0 < d.n; c0 += 1024) {
        const int i = c0 + tid;
        bool keep = i < d.n && d.mask[i];
        after_motion += keep ? 1 : 0;
        if (keep) {
            const float2 p0 = d.pts0[i], p1 = d.pts1[i];
            const float *F = s_F;
            const float a0 = (F[0] * p0.x + F[1] * p0.y) + F[2] * 1.0f;
            const float a1 = (F[3] * p0.x + F[4] * p0.y) + F[5] * 1.0f;
            const float a2 = (F[6] * p0.x + F[7] * p0.y) + F[8] * 1.0f;
            const float b0 = (F[0] * p1.x + F[3] * p1.y) + F[6] * 1.0f;
            const float b1 = (F[1] * p1.x + F[4] * p1.y) + F[7] * 1.0f;
            float num = (p1.x * a0 + p1.y * a1) + 1.0f * a2;
            num *= num;
            const float den = ((a0 * a0 + a1 * a1) + b0 * b0) + b1 * b1;
            keep = (num / den) < d.thres_sampson;
        }
        const int pos = scan_chunk(keep, s_warp, &s_base);
        if (keep) { d.idx_out[pos] = i; d.out1[pos] = d.pts1[i]; }
    }
    after_motion = __reduce_add_sync(0xffffffffu, after_motion);
    if ((tid & 31) == 0) atomicAdd(d.counts + 3, after_motion);
    if (tid == 0) { *d.n_out = s_base; d.counts[4] = s_base; }
}

// all survivors of the tracking stage, compacted, as the input of calcPose5PointsAlgorithm (:589 / :936-937)
__global__ void __launch_bounds__(1024) k_mono_pack(const MonoDev d, float2 *p0c)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < d.n; c0 += 1024) {
        const int i = c0 + threadIdx.x;
        const bool keep = i < d.n && d.mask[i];
        const int pos = scan_chunk(keep, s_warp, &s_base);
        if (keep) { p0c[pos] = d.pts0[i]; d.p1[pos] = d.pts1[i]; d.idx_po[pos] = i; }
    }
    if (threadIdx.x == 0) { *d.n_po = s_base; d.counts[3] = 0; }
}

// :606-609 (init: |t| = 1) / :943-945 (fallback: |t| = length of the previous motion): dT10 = [R10, s t10 / |t10|],
// dT01 = inverseSE3_f(dT10)
__global__ void k_mono_5pt_pose(const MonoDev d)
{
    if (threadIdx.x != 0 || !d.fp_info[2]) return;
    const float nrm = sqrtf((d.t10[0] * d.t10[0] + d.t10[1] * d.t10[1]) + d.t10[2] * d.t10[2]);
    float T10[16];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T10[i * 4 + j] = d.R10[i * 3 + j];
        T10[i * 4 + 3] = d.force == 2 ? d.t10[i] / nrm * d.t_scale : (d.t_scale / nrm) * d.t10[i];
    }
    T10[12] = T10[13] = T10[14] = 0.f; T10[15] = 1.f;
    for (int i = 0; i < 16; ++i) d.dT10[i] = T10[i];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) d.T01[i * 4 + j] = T10[j * 4 + i];
        float s = 0.f;
        for (int k = 0; k < 3; ++k) s += T10[k * 4 + i] * T10[k * 4 + 3];
        d.T01[i * 4 + 3] = -s;
    }
    d.T01[12] = d.T01[13] = d.T01[14] = 0.f; d.T01[15] = 1.f;
}

// compaction of the new features that survived the bidirectional back-tracking
__global__ void __launch_bounds__(1024) k_mono_new(const float2 *p1, const float2 *p0, const uint8_t *mask, const int *n_in, float2 *o1,
                                                   float2 *o0, int *n_out)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int n = *n_in;
    for (int c0 = 0; c0 < n; c0 += 1024) {
        const int i = c0 + threadIdx.x;
        const bool keep = i < n && mask[i];
        const int pos = scan_chunk(keep, s_warp, &s_base);
        if (keep) { o1[pos] = p1[i]; o0[pos] = p0[i]; }
    }
    if (threadIdx.x == 0) *n_out = s_base;
}

extern "C" int vo_mono_frame_step(vo_ctx *ctx, const vo_mono_frame_params *prm, int slot_0, int slot_1, const uint8_t *img_1, int w, int h,
                                  size_t step, int n, const float *pts0, const float *Xw, const uint8_t *flags, const float *T_wc_prev,
                                  const float *dT01_prior, vo_mono_frame_result *res)
{
    if (!ctx || !prm || !res) return VO_ERR_INVALID_ARG;
    const bool init = prm->init_mode != 0;
    VO_REQUIRE(n >= 0, VO_ERR_INVALID_ARG, "negative size");
    VO_REQUIRE(T_wc_prev && (init || dT01_prior) && res->T_wc && res->dT01 && res->dT10, VO_ERR_INVALID_ARG, "null pointer");
    VO_REQUIRE(n == 0 || (pts0 && (init || (Xw && flags)) && res->index && res->pts1), VO_ERR_INVALID_ARG, "null pointer");
    VO_REQUIRE(!init || n >= 5, VO_ERR_INVALID_ARG, "calcPose5PointsAlgorithm(): fewer than five correspondences");
    VO_REQUIRE(!init || prm->thres_5p > 0.f, VO_ERR_INVALID_ARG, "init_mode needs thres_5p");
    const int nb = prm->n_bins_u > 0 && prm->n_bins_v > 0 ? prm->n_bins_u * prm->n_bins_v : 0;
    VO_REQUIRE(nb == 0 || (res->new_p1 && res->new_p0), VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if (img_1) { rc = vo_upload_image(ctx, slot_1, img_1, w, h, step); if (rc) return rc; }
    VO_REQUIRE(slot_0 >= 0 && slot_0 < ctx->n_slots && ctx->slots[slot_0].w == w && ctx->slots[slot_0].h == h &&
               slot_1 >= 0 && slot_1 < ctx->n_slots && ctx->slots[slot_1].w == w && ctx->slots[slot_1].h == h,
               VO_ERR_INVALID_ARG, "image slots have no image of this size");
    const size_t N = (size_t)n, NB = (size_t)nb;
    // staging: inputs [pts0][Xw][flags] | work [pts1][back][err][errb][st][stb][scale][mask][Xp][p1][idx_po][mask_po]
    //          | new-feature work [cand][cand0][back][err][errb][st][stb][mask]
    //          | results [T01][T_wc][dT10][ints][R10 t10 fp_info][idx_out][out1][new1][new0]
    size_t o = 0;
    const size_t o_p0 = o; o += N * 8; const size_t o_X = o; o += step_a16(N * 12); const size_t o_fl = o; o += step_a16(N);
    const size_t in_bytes = o;
    const size_t o_p1 = o; o += N * 8; const size_t o_bk = o; o += N * 8; const size_t o_er = o; o += step_a16(N * 4);
    const size_t o_eb = o; o += step_a16(N * 4); const size_t o_st = o; o += step_a16(N); const size_t o_sb = o; o += step_a16(N);
    const size_t o_sc = o; o += step_a16(N * 4); const size_t o_m = o; o += step_a16(N); const size_t o_Xp = o; o += step_a16(N * 12);
    const size_t o_q1 = o; o += N * 8; const size_t o_ip = o; o += step_a16(N * 4); const size_t o_mp = o; o += step_a16(N);
    const size_t o_c = o; o += NB * 8; const size_t o_c0 = o; o += NB * 8; const size_t o_cb = o; o += NB * 8;
    const size_t o_ce = o; o += step_a16(NB * 4); const size_t o_ceb = o; o += step_a16(NB * 4); const size_t o_cs = o; o += step_a16(NB);
    const size_t o_csb = o; o += step_a16(NB); const size_t o_cm = o; o += step_a16(NB);
    const size_t o_res = o;
    const size_t o_T01 = o; o += 64; const size_t o_Twc = o; o += 64; const size_t o_T10 = o; o += 64; const size_t o_int = o; o += 64;
    const size_t o_fp = o; o += 64;
    const size_t o_io = o; o += step_a16(N * 4); const size_t o_o1 = o; o += N * 8; const size_t o_n1 = o; o += NB * 8;
    const size_t o_n0 = o; o += NB * 8;
    const size_t total = o;
    rc = vo_stage_reserve(ctx, total);
    if (rc) return rc;
    uint8_t *hs = ctx->h_stage, *dv = ctx->d_stage;
    int *ints = (int *)(dv + o_int);      // 0 n_po, 1 po_success, 2 n_out, 3 nan, 4..8 counts, 9 n_detected, 10 n_new
    float ident16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if (init) dT01_prior = ident16;
    memset(hs + o_T01, 0, 320);
    memcpy(hs + o_T01, dT01_prior, 64);   // the GN starts from the previous motion (:860-861)
    if (n > 0) {
        memcpy(hs + o_p0, pts0, N * 8);
        if (!init) { memcpy(hs + o_X, Xw, N * 12); memcpy(hs + o_fl, flags, N); }
        VO_CUDA(cudaMemcpyAsync(dv, hs, init ? N * 8 : in_bytes, cudaMemcpyHostToDevice, ctx->stream));
        VO_CUDA(cudaMemsetAsync(dv + o_m, 1, N, ctx->stream));
    }
    VO_CUDA(cudaMemcpyAsync(dv + o_T01, hs + o_T01, 320, cudaMemcpyHostToDevice, ctx->stream));
    {   // the new image's pyramid and Scharr planes (the previous image keeps its cached ones)
        const int both[2] = {slot_0, slot_1};
        const int eff = vo_effective_max_level(w, h, prm->window_size, prm->max_level);
        rc = vo_ensure_pyramids(ctx, both, 2, (eff + 1 < ctx->max_levels ? eff + 1 : ctx->max_levels), 1);
        if (rc) return rc;
    }
    MonoDev d;
    memset(&d, 0, sizeof(d));
    d.n = n; d.w = w; d.h = h;
    d.out1 = (float2 *)(dv + o_o1); d.n_out = ints + 2;
    d.pts0 = (const float2 *)(dv + o_p0); d.Xw = (const float *)(dv + o_X); d.flags = dv + o_fl;
    d.pts1 = (float2 *)(dv + o_p1); d.scale = (float *)(dv + o_sc); d.mask = dv + o_m;
    d.Xp = (float *)(dv + o_Xp); d.p1 = (float2 *)(dv + o_q1); d.idx_po = (int *)(dv + o_ip); d.mask_po = dv + o_mp;
    d.T01 = (float *)(dv + o_T01); d.T_wc = (float *)(dv + o_Twc); d.dT10 = (float *)(dv + o_T10);
    d.n_po = ints + 0; d.po_success = ints + 1; d.counts = ints + 4; d.idx_out = (int *)(dv + o_io);
    d.R10 = (const float *)(dv + o_fp); d.t10 = d.R10 + 9; d.fp_info = (const int *)(dv + o_fp + 48);
    memcpy(d.Twc_prev, T_wc_prev, 64);
    memcpy(d.K, prm->K, 16);
    d.use_bundled_only = prm->use_bundled_only; d.thres_sampson = prm->thres_sampson;
    int *nan_flag = ints + 3;

    // finish (mask scatter, pose products, Sampson gate, final compaction) + new features + read-back + the one sync
    auto tail = [&]() -> int {
        if (n > 0) {
            k_mono_finish<<<1, 1024, 0, ctx->stream>>>(d);
            ctx->launches++;
        }
        if (nb > 0) {
            // new features: detection on I1 (survivors = occupancy), back-tracking I1 -> I0 with trackBidirection (:985-992)
            int r = vo_detect_launch_d(ctx, slot_1, (const float *)d.out1, n > 0 ? d.n_out : nullptr, n, prm->n_bins_u, prm->n_bins_v,
                                       prm->det_edge, prm->det_min_score, (float *)(dv + o_c), dv + o_cm, ints + 9, nb);
            if (r) return r;
            r = vo_bidir_chain_launch_d(ctx, slot_1, slot_0, (const float *)(dv + o_c), (float *)(dv + o_c0), (float *)(dv + o_cb), dv + o_cs,
                                        dv + o_csb, (float *)(dv + o_ce), (float *)(dv + o_ceb), dv + o_cm, 1, nb, prm->window_size,
                                        prm->max_level, prm->thres_error, prm->thres_bidirection, 0, nullptr, nullptr);
            if (r) return r;
            k_mono_new<<<1, 1024, 0, ctx->stream>>>((const float2 *)(dv + o_c), (const float2 *)(dv + o_c0), dv + o_cm, ints + 9,
                                                    (float2 *)(dv + o_n1), (float2 *)(dv + o_n0), ints + 10);
            ctx->launches++;
        }
        VO_CUDA(cudaGetLastError());
        VO_CUDA(cudaMemcpyAsync(hs + o_res, dv + o_res, total - o_res, cudaMemcpyDeviceToHost, ctx->stream));
        VO_CUDA(cudaStreamSynchronize(ctx->stream));
        return VO_OK;
    };
    // calcPose5PointsAlgorithm on all survivors of the tracking stage, then the pose with the given translation length
    auto five_point = [&](int force, float t_scale) -> int {
        d.force = force; d.t_scale = t_scale;
        k_mono_pack<<<1, 1024, 0, ctx->stream>>>(d, (float2 *)d.Xp);
        ctx->launches++;
        const int r = vo_5pt_launch_d(ctx, d.Xp, (const float *)d.p1, n, d.n_po, prm->K, prm->thres_5p, prm->n_hypotheses, prm->seed,
                                      (float *)(dv + o_fp), (float *)(dv + o_fp) + 9, nullptr, d.mask_po, nullptr, (int *)(dv + o_fp + 48));
        if (r) return r;
        k_mono_5pt_pose<<<1, 32, 0, ctx->stream>>>(d);
        ctx->launches++;
        return VO_OK;
    };

    res->used_5point = 0; res->n_5p_ransac = 0;
    if (n > 0 && init) {
        // mono_vo.cpp:573-575: tracker_->track(I0, I1, pts0, ...) -- no prior, status && err <= thres
        KltPost post{};
        post.mode = 1; post.thres_err = prm->thres_error; post.mask = d.mask;
        rc = vo_klt_launch(ctx, 1, &slot_0, &slot_1, (const float *)d.pts0, n, prm->window_size, prm->max_level, 0, (float *)d.pts1,
                           dv + o_st, (float *)(dv + o_er), nullptr, &post);
        if (rc) return rc;
        k_step_count<<<1, 1024, 0, ctx->stream>>>(d.mask, n, d.counts + 0);
        ctx->launches++;
        rc = five_point(2, 1.0f);
        if (rc) return rc;
    } else if (n > 0) {
        float Tcw_prev[16], Twc_prior[16], Tcw_prior[16];
        step_inv_se3_f(T_wc_prev, Tcw_prev);                 // Frame::getPoseInv()
        step_mul4_f(T_wc_prev, dT01_prior, Twc_prior);       // :733
        step_inv_se3_f(Twc_prior, Tcw_prior);                // :734
        memcpy(d.Tcw_prev, Tcw_prev, 48); memcpy(d.Tcw_prior, Tcw_prior, 48);
        k_mono_prior<<<vo_div_up(n, 256), 256, 0, ctx->stream>>>(d);
        ctx->launches++;
        // K4 + K7: trackBidirectionWithPrior (forward pass from the prior, backward pass seeded with pts0, validity test
        // fused in its epilogue) and trackWithScale.  Without gate counts the three dependent stages of a feature run back
        // to back in ONE launch; with counts they are separate launches so that the mask can be counted in between.
        if (!res->counts) {
            rc = vo_bidir_chain_launch_d(ctx, slot_0, slot_1, (const float *)d.pts0, (float *)d.pts1, (float *)(dv + o_bk), dv + o_st, dv + o_sb,
                                         (float *)(dv + o_er), (float *)(dv + o_eb), d.mask, 0, n, prm->window_size, prm->max_level,
                                         prm->thres_error, prm->thres_bidirection, 1, prm->do_scale_refine ? d.scale : nullptr, nan_flag);
            if (rc) return rc;
        } else {
            rc = vo_bidir_chain_launch_d(ctx, slot_0, slot_1, (const float *)d.pts0, (float *)d.pts1, (float *)(dv + o_bk), dv + o_st, dv + o_sb,
                                         (float *)(dv + o_er), (float *)(dv + o_eb), d.mask, 0, n, prm->window_size, prm->max_level,
                                         prm->thres_error, prm->thres_bidirection, 1, nullptr, nullptr);
            if (rc) return rc;
            k_step_count<<<1, 1024, 0, ctx->stream>>>(d.mask, n, d.counts + 0);
            ctx->launches++;
            if (prm->do_scale_refine) {
                rc = vo_klt_scale_launch_d(ctx, slot_0, slot_1, (const float *)d.pts0, d.scale, n, (float *)d.pts1, d.mask, nan_flag);
                if (rc) return rc;
            }
        }
        k_step_count<<<1, 1024, 0, ctx->stream>>>(d.mask, n, d.counts + 1);
        k_mono_select<<<1, 1024, 0, ctx->stream>>>(d);
        ctx->launches += 2;
        // the reference passes the float threshold through a `const int &` parameter (motion_estimator.h:117): truncation
        rc = vo_pose_launch_d(ctx, 1, nullptr, n, d.n_po, d.Xp, (const float *)d.p1, nullptr, prm->K, prm->K, nullptr,
                              (float)(int)prm->thres_poseba_error, 1, 0, d.T01, d.mask_po, d.po_success, nullptr);
        if (rc) return rc;
    }
    rc = tail();
    if (rc) return rc;
    const int *hi = (const int *)(hs + o_int);
    const int *hfp = (const int *)(hs + o_fp + 48);
    if (n > 0 && !init && !hi[3] && (hi[0] <= 10 || !hi[1])) {
        // mono_vo.cpp:909-949: insufficient points or a failed pose-only BA -> five-point on the K7 survivors, translation
        // scaled to the length of the previous motion
        if (!(prm->thres_5p > 0.f)) {
            ctx->last_error = hi[0] <= 10 ? "insufficient points for the pose-only BA and the 5-point fallback is disabled (thres_5p <= 0)"
                                          : "pose-only BA failed and the 5-point fallback is disabled (thres_5p <= 0)";
            return VO_ERR_MODE;
        }
        const float scale = sqrtf((dT01_prior[3] * dT01_prior[3] + dT01_prior[7] * dT01_prior[7]) + dT01_prior[11] * dT01_prior[11]);   // :943
        rc = five_point(1, scale);
        if (rc) return rc;
        rc = tail();
        if (rc) return rc;
        res->used_5point = 1;
    }
    if (init) res->used_5point = 1;
    res->n_tracked = 0; res->n_new = 0; res->n_detected = 0;
    if (nb > 0) {
        res->n_detected = hi[9];
        res->n_new = hi[10];
        memcpy(res->new_p1, hs + o_n1, (size_t)hi[10] * 8);
        memcpy(res->new_p0, hs + o_n0, (size_t)hi[10] * 8);
    }
    if (n > 0) {
        if (res->counts) for (int k = 0; k < 5; ++k) res->counts[k] = hi[4 + k];
        if (hi[3]) { ctx->last_error = "ax ay nan (feature_tracker.cpp:414)"; return VO_ERR_NAN; }
        if (res->used_5point) {
            res->n_5p_ransac = hfp[0];
            if (!hfp[2]) {
                ctx->last_error = init ? "calcPose5PointsAlgorithm() is failed." : "'calcPose5PointsAlgorithm()' is failed. Terminate the algorithm.";   // :590 / :940
                return VO_ERR_MODE;
            }
        }
        const int k_out = hi[2];
        memcpy(res->dT01, hs + o_T01, 64);
        memcpy(res->T_wc, hs + o_Twc, 64);
        memcpy(res->dT10, hs + o_T10, 64);
        res->n_tracked = k_out;
        memcpy(res->index, hs + o_io, (size_t)k_out * 4);
        memcpy(res->pts1, hs + o_o1, (size_t)k_out * 8);
    } else {
        step_mul4_f(T_wc_prev, dT01_prior, res->T_wc);
        memcpy(res->dT01, dT01_prior, 64);
        step_inv_se3_f(dT01_prior, res->dT10);
        if (res->counts) for (int k = 0; k < 5; ++k) res->counts[k] = 0;
    }
    return VO_OK;
}
