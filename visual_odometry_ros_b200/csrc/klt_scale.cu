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
// only for iter > 1, RMS-error <= 30 acceptance.  Out-of-image samples: by default they are
// masked out (the intended semantics).  The reference reuses stale buffer contents there, a
// sequential dependence between features; vo_set_scale_mode(ctx, 1) reproduces that too
// (k_klt_scale_fixup below: the few features with an out-of-image sample are recomputed against
// the buffer state their predecessors would have left).
// FP32 throughout like the reference, except that the 264-term sums are reduced in FP64
// across the warp and rounded once (the reference adds them sequentially in FP32).
// Compiled with -fmad=false.
#include "vo_internal.cuh"
#include "klt_scale_device.cuh"

#include <cstring>

__global__ void __launch_bounds__(128) k_klt_scale(const KltScaleArgs a)
{
    const int lane = threadIdx.x & 31;
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (f >= a.n) return;
    klt_scale_feature(a, f, lane);
}

// Reference-faithful border mode, second pass.  A live feature that entered its iteration loop with every sample inside the
// image rewrites every entry of the reference's sample buffers (and every mask is true from then on): the buffer state
// behind such a "reset" feature does not depend on anything before it.  All other live features -- a sample outside the
// image, or rejected before the iterations -- read and/or carry stale entries; maximal runs of them between two reset
// features are CHAINS that the reference processes strictly in order.  One warp per chain (the warp of its first feature):
// it replays the reset feature in front of the chain, if any, against an empty state, then walks the chain, computing each
// feature exactly as the reference would and storing the results of those whose samples left the image.  The work is
// linear in the chain lengths; the latency is that of the longest chain (features along an image edge that are neighbours
// in the list, e.g. a row of border bins, form chains of a few dozen).
__global__ void __launch_bounds__(128) k_klt_scale_fixup(const KltScaleArgs a)
{
    const int lane = threadIdx.x & 31;
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (f >= a.n) return;
    auto live = [&](int i) { return (a.flags[i] & KS_FLAG_PROCESSED) != 0; };
    auto reset = [&](int i) { return (a.flags[i] & (KS_FLAG_PROCESSED | KS_FLAG_BORDER | KS_FLAG_RAN)) == (KS_FLAG_PROCESSED | KS_FLAG_RAN); };
    if (!live(f) || reset(f)) return;
    // nearest live feature in front of f
    int p = -1;
    for (int base = f - 1; base >= 0 && p < 0; base -= 32) {
        const int i = base - lane;
        const unsigned m = __ballot_sync(0xffffffffu, i >= 0 && live(i));
        if (m) p = base - (__ffs(m) - 1);
    }
    if (p >= 0 && !reset(p)) return;             // f is inside a chain; the chain's first feature owns it
    KsState st;
    ks_state_clear(st);
    bool ok, nan_hit;
    float2 out;
    if (p >= 0) klt_scale_feature_stale(a, p, lane, a.pre_pts[p], st, ok, out, nan_hit);
    for (int i = f; i < a.n; ++i) {
        if (!live(i)) continue;
        if (reset(i)) break;
        const float2 pt1 = a.pre_pts[i];
        klt_scale_feature_stale(a, i, lane, pt1, st, ok, out, nan_hit);
        if ((a.flags[i] & KS_FLAG_BORDER) && lane == 0) {
            if (nan_hit) atomicExch(a.nan_flag, 1);
            a.mask[i] = ok ? 1 : 0;
            a.pts_track[i] = ok ? out : pt1;      // a rejected feature keeps its input point (feature_tracker.cpp:489-500)
        }
    }
}

// grow-only scratch of the faithful mode: [flags n][pad][pre_pts n * 8]
static int ks_scratch(vo_ctx *ctx, int n, uint8_t **flags, float2 **pre)
{
    const size_t need = ((size_t)n + 15) / 16 * 16 + (size_t)n * 8;
    if (need > ctx->ks_bytes) {
        VO_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_ks) cudaFree(ctx->d_ks);
        ctx->d_ks = nullptr; ctx->ks_bytes = 0;
        size_t want = need * 2;
        if (want < ((size_t)1 << 16)) want = (size_t)1 << 16;
        VO_CUDA(cudaMalloc(&ctx->d_ks, want));
        ctx->ks_bytes = want;
    }
    *flags = ctx->d_ks;
    *pre = reinterpret_cast<float2 *>(ctx->d_ks + ((size_t)n + 15) / 16 * 16);
    return VO_OK;
}

// faithful mode: fills the record pointers of a scale stage (also used by the fused chains in klt.cu)
int vo_klt_scale_prepare(vo_ctx *ctx, KltScaleArgs &a)
{
    a.flags = nullptr; a.pre_pts = nullptr;
    if (!ctx->scale_faithful) return VO_OK;
    return ks_scratch(ctx, a.n, &a.flags, &a.pre_pts);
}

// faithful mode: the second pass over the records a scale stage left
int vo_klt_scale_fixup_launch(vo_ctx *ctx, const KltScaleArgs &a)
{
    if (!a.flags) return VO_OK;
    k_klt_scale_fixup<<<vo_div_up(a.n, 4), 128, 0, ctx->stream>>>(a);
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    return VO_OK;
}

int vo_klt_scale_launch_d(vo_ctx *ctx, int slot0, int slot1, const float *pts0_d, const float *scale_d, int n,
                          float *pts_track_d, uint8_t *mask_d, int *nan_flag_d)
{
    KltScaleArgs a;
    a.slots = ctx->d_slots; a.slot0 = slot0; a.slot1 = slot1;
    a.pts0 = (const float2 *)pts0_d; a.scale = scale_d;
    a.pts_track = (float2 *)pts_track_d; a.mask = mask_d; a.nan_flag = nan_flag_d; a.iters = nullptr; a.n = n;
    int rc = vo_klt_scale_prepare(ctx, a);
    if (rc) return rc;
    k_klt_scale<<<vo_div_up(n, 4), 128, 0, ctx->stream>>>(a);
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    return vo_klt_scale_fixup_launch(ctx, a);
}

extern "C" int vo_set_scale_mode(vo_ctx *ctx, int faithful_borders)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    ctx->scale_faithful = faithful_borders ? 1 : 0;
    return VO_OK;
}

extern "C" int vo_get_scale_mode(const vo_ctx *ctx) { return ctx ? ctx->scale_faithful : VO_ERR_INVALID_ARG; }

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
