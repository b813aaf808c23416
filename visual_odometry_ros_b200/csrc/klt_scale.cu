// klt_scale.cu -- K-klt-scale: FeatureTracker::trackWithScale, one warp per feature.
//
// Replaces core/visual_odometry/feature_tracker.cpp:236-504 (+ interpImageSameRatio /
// interpImage3SameRatio, core/util/image_processing.cpp:79-118, 268-331).  The reference
// converts both full images to float and takes full-image 3x3 Sobel derivatives of the
// previous image on every call (feature_tracker.cpp:297-299, stereo_vo.cpp:551-552); here
// nothing image-sized is materialised: the 264 checkerboard samples of each feature read the
// resident u8 level-0 planes directly, the float conversion is exact, and the Sobel taps are
// evaluated on the fly as exact small integers, so the sampled values are bit-identical to
// the reference's.  Kept quirks: every sample uses the PATCH CENTRE's bilinear fractions and
// (int)-truncated coordinates (Appendix B #9), fixed per-feature scale, stop rule evaluated
// only for iter > 1, RMS-error <= 30 acceptance.  Out-of-image samples are masked out (the
// intended semantics; the reference reuses stale buffer contents there -- see DESIGN.md).
// FP32 throughout like the reference, except that the 264-term sums are reduced in FP64
// across the warp and rounded once (the reference adds them sequentially in FP32).
// Compiled with -fmad=false.
#include "vo_internal.cuh"

#include <cstring>

#define KS_HALF 11
#define KS_NELEM 264
#define KS_PER_LANE 9      // ceil(264 / 32)

struct KltScaleArgs {
    const SlotDesc *slots;
    int slot0, slot1;
    const float2 *pts0;
    const float *scale;
    float2 *pts_track;
    uint8_t *mask;
    int *nan_flag;
    int *iters;     // nullable
    int n;
};

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float interp4(float I1, float I2, float I3, float I4, float ax, float ay, float axay)
{
    return axay * (I1 - I2 - I3 + I4) + ax * (-I1 + I2) + ay * (-I1 + I3) + I1;
}

__global__ void __launch_bounds__(128) k_klt_scale(const KltScaleArgs a)
{
    const int lane = threadIdx.x & 31;
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (f >= a.n) return;
    if (!a.mask[f]) { if (a.iters && lane == 0) a.iters[f] = 0; return; }
    const LevelDesc I0 = a.slots[a.slot0].lv[0];
    const LevelDesc I1 = a.slots[a.slot1].lv[0];
    const int n_cols = I0.w, n_rows = I0.h;
    const float2 pt0 = a.pts0[f];
    const float2 pt1 = a.pts_track[f];
    const float scale = a.scale[f];

    // checkerboard pattern: sample j -> (u, v) offsets (feature_tracker.cpp:308-320)
    float pu[KS_PER_LANE], pv[KS_PER_LANE];
#pragma unroll
    for (int k = 0; k < KS_PER_LANE; ++k) {
        const int j = lane + 32 * k;
        const int p = j / 23, r = j - 23 * p;
        const int v = r < 11 ? 2 * p : 2 * p + 1;
        const int u = r < 11 ? 1 + 2 * r : 2 * (r - 11);
        pu[k] = (float)(u - KS_HALF);
        pv[k] = (float)(v - KS_HALF);
    }

    float ax = pt0.x - floorf(pt0.x), ay = pt0.y - floorf(pt0.y), axay = ax * ay;
    if (ax < 0 || ax > 1 || ay < 0 || ay > 1) { if (lane == 0) a.mask[f] = 0; return; }
    if (isnan(ax + ay)) { if (lane == 0) { a.mask[f] = 0; atomicExch(a.nan_flag, 1); } return; }

    // ---- template: I0, du0, dv0 at the 264 samples (interpImage3SameRatio)
    float I0p[KS_PER_LANE], dup[KS_PER_LANE], dvp[KS_PER_LANE];
    unsigned m0 = 0;
    double sA11 = 0, sA12 = 0, sA22 = 0;
#pragma unroll
    for (int k = 0; k < KS_PER_LANE; ++k) {
        I0p[k] = 0.f; dup[k] = 0.f; dvp[k] = 0.f;
        if (lane + 32 * k >= KS_NELEM) continue;
        const float uc = pt0.x + pu[k], vc = pt0.y + pv[k];
        const int u0 = (int)uc, v0 = (int)vc;
        if (u0 < 1 || u0 >= n_cols - 2 || v0 < 1 || v0 >= n_rows - 2) continue;
        // 4x4 u8 neighbourhood rows v0-1..v0+2, cols u0-1..u0+2
        int p[4][4];
        const uint8_t *b = I0.img + (ptrdiff_t)(v0 - 1) * I0.pitch + (u0 - 1);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) p[r][c] = __ldg(b + r * I0.pitch + c);
        float gI[2][2], gU[2][2], gV[2][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                gI[r][c] = (float)p[r + 1][c + 1];
                gU[r][c] = (float)((p[r][c + 2] - p[r][c]) + 2 * (p[r + 1][c + 2] - p[r + 1][c]) + (p[r + 2][c + 2] - p[r + 2][c]));
                gV[r][c] = (float)((p[r + 2][c] - p[r][c]) + 2 * (p[r + 2][c + 1] - p[r][c + 1]) + (p[r + 2][c + 2] - p[r][c + 2]));
            }
        I0p[k] = interp4(gI[0][0], gI[0][1], gI[1][0], gI[1][1], ax, ay, axay);
        dup[k] = interp4(gU[0][0], gU[0][1], gU[1][0], gU[1][1], ax, ay, axay);
        dvp[k] = interp4(gV[0][0], gV[0][1], gV[1][0], gV[1][1], ax, ay, axay);
        m0 |= 1u << k;
        sA11 += (double)(dup[k] * dup[k]);
        sA12 += (double)(dup[k] * dvp[k]);
        sA22 += (double)(dvp[k] * dvp[k]);
    }
    const float A11 = (float)warp_sum_d(sA11), A12 = (float)warp_sum_d(sA12), A22 = (float)warp_sum_d(sA22);
    const float D = A11 * A22 - A12 * A12;
    if (D < 1e-4f) { if (lane == 0) { a.mask[f] = 0; if (a.iters) a.iters[f] = 0; } return; }
    const float invD = (float)(1.0 / (double)D);
    const float iD_A11 = A11 * invD, iD_A12 = A12 * invD, iD_A22 = A22 * invD;

    float err_curr = 0.f, err_prev = 1e12f;
    float tx = pt1.x - pt0.x, ty = pt1.y - pt0.y;
    bool nan_hit = false;
    int iter = 0;
    for (; iter < 30; ++iter) {
        const float pux = pt0.x + tx, puy = pt0.y + ty;
        ax = pux - floorf(pux); ay = puy - floorf(puy); axay = ax * ay;
        if (isnan(ax + ay)) { nan_hit = true; break; }
        double sb1 = 0, sb2 = 0, serr = 0;
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < KS_PER_LANE; ++k) {
            if (!((m0 >> k) & 1u)) continue;
            const float uc = pux + pu[k] * scale, vc = puy + pv[k] * scale;
            if (uc < 1 || uc >= (float)(n_cols - 2) || vc < 1 || vc >= (float)(n_rows - 2)) continue;
            const int u0 = (int)uc, v0 = (int)vc;
            const uint8_t *b = I1.img + (ptrdiff_t)v0 * I1.pitch + u0;
            const float J1 = (float)__ldg(b), J2 = (float)__ldg(b + 1), J3 = (float)__ldg(b + I1.pitch), J4 = (float)__ldg(b + I1.pitch + 1);
            const float r = interp4(J1, J2, J3, J4, ax, ay, axay) - I0p[k];
            sb1 += (double)(dup[k] * r);
            sb2 += (double)(dvp[k] * r);
            serr += (double)(r * r);
            ++cnt;
        }
        const float b1 = (float)warp_sum_d(sb1), b2 = (float)warp_sum_d(sb2);
        err_curr = (float)warp_sum_d(serr);
        const int cnt_valid = __reduce_add_sync(0xffffffffu, cnt);
        const float dtu = (-iD_A22 * b1 + iD_A12 * b2);
        const float dtv = (iD_A12 * b1 - iD_A11 * b2);
        if (isnan(dtu + dtv)) { nan_hit = true; break; }
        tx += dtu; ty += dtv;
        err_curr /= (float)cnt_valid;
        err_curr = sqrtf(err_curr);
        const float err_rate = fabsf(err_prev - err_curr) / err_prev;
        const float dt_norm = dtu * dtu + dtv * dtv;
        if (iter > 1 && (err_rate <= 1e-3f || dt_norm <= 1e-4f)) { ++iter; break; }
        err_prev = err_curr;
    }
    if (lane == 0) {
        if (nan_hit) { atomicExch(a.nan_flag, 1); a.mask[f] = 0; }
        else if (isnan(err_curr)) a.mask[f] = 0;
        else if (err_curr <= 30.f) { a.pts_track[f] = make_float2(pt0.x + tx, pt0.y + ty); a.mask[f] = 1; }
        else a.mask[f] = 0;
        if (a.iters) a.iters[f] = iter > 30 ? 30 : iter;
    }
}

int vo_klt_scale_launch_d(vo_ctx *ctx, int slot0, int slot1, const float *pts0_d, const float *scale_d, int n,
                          float *pts_track_d, uint8_t *mask_d, int *nan_flag_d)
{
    KltScaleArgs a;
    a.slots = ctx->d_slots; a.slot0 = slot0; a.slot1 = slot1;
    a.pts0 = (const float2 *)pts0_d; a.scale = scale_d;
    a.pts_track = (float2 *)pts_track_d; a.mask = mask_d; a.nan_flag = nan_flag_d; a.iters = nullptr; a.n = n;
    k_klt_scale<<<vo_div_up(n, 4), 128, 0, ctx->stream>>>(a);
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    return VO_OK;
}

extern "C" int vo_ft_track_with_scale(vo_ctx *ctx, int slot0, int slot1, const float *pts0, const float *scale_est, int n,
                                      float *pts_track_inout, uint8_t *mask_inout)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0, VO_ERR_INVALID_ARG, "negative size");
    if (n == 0) return VO_OK;
    VO_REQUIRE(pts0 && scale_est && pts_track_inout && mask_inout, VO_ERR_INVALID_ARG, "null pointer");
    VO_REQUIRE(slot0 >= 0 && slot0 < ctx->n_slots && slot1 >= 0 && slot1 < ctx->n_slots, VO_ERR_INVALID_ARG, "slot id out of range");
    VO_REQUIRE(ctx->slots[slot0].w > 0 && ctx->slots[slot0].w == ctx->slots[slot1].w && ctx->slots[slot0].h == ctx->slots[slot1].h,
               VO_ERR_SIZE_MISMATCH, "image pair sizes differ or missing");
    VO_CUDA(cudaSetDevice(ctx->device));
    // only level 0 is needed (border irrelevant: valid samples never leave the image)
    int rc = vo_ensure_pyramids(ctx, &slot0, 1, 1, 0);
    if (rc) return rc;
    rc = vo_ensure_pyramids(ctx, &slot1, 1, 1, 0);
    if (rc) return rc;
    const size_t N = (size_t)n;
    const size_t oP0 = 0, oS = oP0 + N * 8, oPt = oS + N * 4, oFlag = oPt + N * 8, oM = oFlag + 16, total = oM + N;
    rc = vo_stage_reserve(ctx, total);
    if (rc) return rc;
    uint8_t *h = ctx->h_stage, *d = ctx->d_stage;
    memcpy(h + oP0, pts0, N * 8);
    memcpy(h + oS, scale_est, N * 4);
    memcpy(h + oPt, pts_track_inout, N * 8);
    memset(h + oFlag, 0, 16);
    memcpy(h + oM, mask_inout, N);
    VO_CUDA(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, ctx->stream));
    rc = vo_klt_scale_launch_d(ctx, slot0, slot1, (const float *)(d + oP0), (const float *)(d + oS), n, (float *)(d + oPt), d + oM,
                               (int *)(d + oFlag));
    if (rc) return rc;
    VO_CUDA(cudaMemcpyAsync(h + oPt, d + oPt, total - oPt, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(pts_track_inout, h + oPt, N * 8);
    memcpy(mask_inout, h + oM, N);
    if (*(int *)(h + oFlag)) { ctx->last_error = "ax ay nan / dtu dtv nan (feature_tracker.cpp:414,465)"; return VO_ERR_NAN; }
    return VO_OK;
}
