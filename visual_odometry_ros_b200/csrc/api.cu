// api.cu -- context management and the image / KLT part of the C ABI (include/vo_b200.h).
#include "vo_internal.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#ifndef VO_VERSION
#define VO_VERSION "0.1.0"
#endif
#define VO_STR2(x) #x
#define VO_STR(x) VO_STR2(x)

extern "C" const char *vo_status_string(int status)
{
    switch (status) {
    case VO_OK: return "ok";
    case VO_ERR_INVALID_ARG: return "invalid argument";
    case VO_ERR_CUDA: return "CUDA error";
    case VO_ERR_SIZE_MISMATCH: return "size mismatch";
    case VO_ERR_NAN: return "NaN encountered";
    case VO_ERR_MODE: return "stereo/mono mode misuse";
    case VO_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    case VO_ERR_LARGE_UPDATE: return "large update!";
    default: return "unknown status";
    }
}

extern "C" const char *vo_last_error(const vo_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

extern "C" const char *vo_build_info(void)
{
    return "vo_b200 " VO_VERSION " sm_100a nvcc " VO_STR(__CUDACC_VER_MAJOR__) "." VO_STR(__CUDACC_VER_MINOR__);
}

extern "C" int vo_effective_max_level(int w, int h, int win, int max_level)
{
    // cv::buildOpticalFlowPyramid: stop when the next level's width or height <= window
    for (int level = 0; level < max_level; ++level) {
        w = (w + 1) / 2;
        h = (h + 1) / 2;
        if (w <= win || h <= win) return level;
    }
    return max_level;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" int vo_ctx_create(int device, int max_w, int max_h, int n_slots, int max_feat, void *stream, vo_ctx **out)
{
    if (!out) return VO_ERR_INVALID_ARG;
    *out = nullptr;
    if (max_w < 1 || max_h < 1 || n_slots < 0 || max_feat < 0) return VO_ERR_INVALID_ARG;
    // Load every kernel of the library when CUDA comes up instead of at its first launch: lazy loading costs tens of
    // milliseconds inside the first frames of a sequence.  Only effective if this is the first CUDA use of the process
    // (a node that links nothing else on CUDA); an explicit CUDA_MODULE_LOADING in the environment wins.
    static std::once_flag eager_once;       // once, before the first CUDA call of the library (setenv is not re-entrant)
    std::call_once(eager_once, [] { setenv("CUDA_MODULE_LOADING", "EAGER", 0); });
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return VO_ERR_NO_DEVICE;
    }
    vo_ctx *ctx = new (std::nothrow) vo_ctx();
    if (!ctx) return VO_ERR_INVALID_ARG;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return VO_ERR_CUDA; }
    if (stream) { ctx->stream = (cudaStream_t)stream; ctx->own_stream = false; }
    else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return VO_ERR_CUDA; }
        ctx->own_stream = true;
    }
    ctx->max_w = max_w; ctx->max_h = max_h; ctx->n_slots = n_slots; ctx->max_feat = max_feat;

    // level geometry for the maximum image; levels stop once a side would drop below 8 px
    int lw[VO_MAX_LEVELS], lh[VO_MAX_LEVELS], nl = 0;
    for (int w = max_w, h = max_h; nl < VO_MAX_LEVELS; ++nl) {
        lw[nl] = w; lh[nl] = h;
        const int nw = (w + 1) / 2, nh = (h + 1) / 2;
        if (nw < 8 || nh < 8) { ++nl; break; }
        w = nw; h = nh;
    }
    ctx->max_levels = nl;
    size_t img_bytes = 0, der_bytes = 0;
    for (int l = 0; l < nl; ++l) {
        const size_t pitch = align_up((size_t)lw[l] + 2 * VO_PAD, 128);
        const size_t rows = (size_t)lh[l] + 2 * VO_PAD;
        img_bytes += align_up(pitch * rows, 256);
        der_bytes += align_up(pitch * rows * sizeof(short2), 256);
    }
    // ONE allocation for all slots, uniform stride: the pyramid planes of level l of every slot are then one 3-D tensor
    // (x, y, slot) for the TMA descriptors of the LK kernels (klt.cu)
    ctx->slots.resize(n_slots);
    ctx->slot_stride = img_bytes + der_bytes;          // multiple of 256
    if (n_slots > 0) {
        if (cudaMalloc((void **)&ctx->slot_pool, ctx->slot_stride * (size_t)n_slots) != cudaSuccess) { vo_ctx_destroy(ctx); return VO_ERR_CUDA; }
        if (cudaMemsetAsync(ctx->slot_pool, 0, ctx->slot_stride * (size_t)n_slots, ctx->stream) != cudaSuccess) { vo_ctx_destroy(ctx); return VO_ERR_CUDA; }
    }
    for (int s = 0; s < n_slots; ++s) {
        Slot &S = ctx->slots[s];
        S.bytes = ctx->slot_stride;
        S.base = ctx->slot_pool + (size_t)s * ctx->slot_stride;
        memset(&S.desc, 0, sizeof(S.desc));
    }
    if (n_slots > 0 && cudaMalloc((void **)&ctx->d_slots, sizeof(SlotDesc) * n_slots) != cudaSuccess) {
        vo_ctx_destroy(ctx);
        return VO_ERR_CUDA;
    }
    ctx->raw_stride = (size_t)max_w * max_h;   // dense: consecutive slots are adjacent -> mergeable DMA
    if (n_slots > 0 && cudaMalloc((void **)&ctx->raw_base, ctx->raw_stride * n_slots + 256) != cudaSuccess) {
        vo_ctx_destroy(ctx);
        return VO_ERR_CUDA;
    }
    if (vo_stage_reserve(ctx, (size_t)(max_feat > 0 ? max_feat : 1) * 64) != VO_OK) { vo_ctx_destroy(ctx); return VO_ERR_CUDA; }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { vo_ctx_destroy(ctx); return VO_ERR_CUDA; }
    *out = ctx;
    return VO_OK;
}

extern "C" int vo_ctx_destroy(vo_ctx *ctx)
{
    if (!ctx) return VO_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->nccl_comm) vo_dist_finalize(ctx);
    vo_klt_maps_free(ctx);
    if (ctx->slot_pool) cudaFree(ctx->slot_pool);
    if (ctx->d_slots) cudaFree(ctx->d_slots);
    if (ctx->raw_base) cudaFree(ctx->raw_base);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    for (int i = 0; i < 4; ++i) if (ctx->d_f32[i]) cudaFree(ctx->d_f32[i]);
    if (ctx->d_lba) cudaFree(ctx->d_lba);
    if (ctx->d_ks) cudaFree(ctx->d_ks);
    if (ctx->d_det) cudaFree(ctx->d_det);
    if (ctx->d_fp) cudaFree(ctx->d_fp);
    if (ctx->d_orb) cudaFree(ctx->d_orb);
    if (ctx->d_rect) cudaFree(ctx->d_rect);
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < 3; ++i) if (ctx->side[i]) { cudaStreamSynchronize(ctx->side[i]); cudaStreamDestroy(ctx->side[i]); }
    for (int i = 0; i < 4; ++i) if (ctx->ev_aux[i]) cudaEventDestroy(ctx->ev_aux[i]);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return VO_OK;
}

extern "C" int vo_ctx_synchronize(vo_ctx *ctx)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    return VO_OK;
}

extern "C" long long vo_ctx_launch_count(const vo_ctx *ctx) { return ctx ? ctx->launches : 0; }

int vo_stage_reserve(vo_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->stage_bytes) return VO_OK;
    // grow geometrically (pinned allocations cost tens of milliseconds; callers' sizes creep up frame by frame)
    if (ctx->stage_bytes && bytes < ctx->stage_bytes * 2) bytes = ctx->stage_bytes * 2;
    bytes = align_up(bytes, 4096);
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    ctx->h_stage = nullptr; ctx->d_stage = nullptr; ctx->stage_bytes = 0;
    VO_CUDA(cudaMallocHost((void **)&ctx->h_stage, bytes));
    VO_CUDA(cudaMalloc((void **)&ctx->d_stage, bytes));
    ctx->stage_bytes = bytes;
    return VO_OK;
}

// (Re)derive the level descriptors of a slot for a w x h image and push them to the device.
static int slot_set_geometry(vo_ctx *ctx, int slot, int w, int h)
{
    Slot &S = ctx->slots[slot];
    if (S.w == w && S.h == h) return VO_OK;
    // image size changed: the zero ring of the derivative planes must be restored
    VO_CUDA(cudaMemsetAsync(S.base, 0, S.bytes, ctx->stream));
    size_t img_off = 0;
    // image planes first, then derivative planes (each plane 256-B aligned)
    int lw = w, lh = h;
    size_t offs_img[VO_MAX_LEVELS], offs_der[VO_MAX_LEVELS];
    int pitches[VO_MAX_LEVELS], ws[VO_MAX_LEVELS], hs[VO_MAX_LEVELS];
    int nl = 0;
    for (; nl < ctx->max_levels; ++nl) {
        const size_t pitch = align_up((size_t)lw + 2 * VO_PAD, 128);
        offs_img[nl] = img_off;
        pitches[nl] = (int)pitch; ws[nl] = lw; hs[nl] = lh;
        img_off += align_up(pitch * ((size_t)lh + 2 * VO_PAD), 256);
        const int nw = (lw + 1) / 2, nh = (lh + 1) / 2;
        if (nw < 8 || nh < 8) { ++nl; break; }
        lw = nw; lh = nh;
    }
    size_t der_off = img_off;
    for (int l = 0; l < nl; ++l) {
        offs_der[l] = der_off;
        der_off += align_up((size_t)pitches[l] * ((size_t)hs[l] + 2 * VO_PAD) * sizeof(short2), 256);
    }
    VO_REQUIRE(der_off <= S.bytes, VO_ERR_INVALID_ARG, "image larger than the context's max_w x max_h");
    memset(&S.desc, 0, sizeof(S.desc));
    for (int l = 0; l < nl; ++l) {
        LevelDesc &L = S.desc.lv[l];
        L.w = ws[l]; L.h = hs[l]; L.pitch = pitches[l];
        L.img = S.base + offs_img[l] + (size_t)VO_PAD * pitches[l] + VO_PAD;
        L.deriv = reinterpret_cast<short2 *>(S.base + offs_der[l]) + (size_t)VO_PAD * pitches[l] + VO_PAD;
    }
    S.desc.raw = ctx->raw_base + (size_t)slot * ctx->raw_stride;
    S.w = w; S.h = h;
    VO_CUDA(cudaMemcpyAsync(ctx->d_slots + slot, &S.desc, sizeof(SlotDesc), cudaMemcpyHostToDevice, ctx->stream));
    // the descriptor lives in pageable host memory inside the vector: make the copy complete
    // before anybody can move/modify it
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    return VO_OK;
}

static int set_image_common(vo_ctx *ctx, int slot, const uint8_t *data, int w, int h, size_t step, cudaMemcpyKind kind,
                            cudaStream_t st = nullptr)
{
    if (ctx && !st) st = ctx->stream;
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(slot >= 0 && slot < ctx->n_slots, VO_ERR_INVALID_ARG, "slot id out of range");
    VO_REQUIRE(data && w >= 8 && h >= 8 && step >= (size_t)w, VO_ERR_INVALID_ARG, "bad image arguments");
    VO_REQUIRE(w <= ctx->max_w + 0 && h <= ctx->max_h + 0, VO_ERR_INVALID_ARG, "image larger than the context's max_w x max_h");
    VO_CUDA(cudaSetDevice(ctx->device));
    int rc = slot_set_geometry(ctx, slot, w, h);
    if (rc) return rc;
    Slot &S = ctx->slots[slot];
    // DMA into the dense raw staging area (one contiguous copy when the source is dense);
    // the ingest kernel moves it into the padded level-0 plane when the pyramid is built.
    uint8_t *raw = ctx->raw_base + (size_t)slot * ctx->raw_stride;
    if (step == (size_t)w) VO_CUDA(cudaMemcpyAsync(raw, data, (size_t)w * h, kind, st));
    else VO_CUDA(cudaMemcpy2DAsync(raw, w, data, step, w, h, kind, st));
    S.levels_built = 0; S.deriv_built = 0; S.border0 = false; S.raw_pending = true;
    return VO_OK;
}

int vo_slot_prepare(vo_ctx *ctx, int slot, int w, int h)
{
    int rc = slot_set_geometry(ctx, slot, w, h);
    if (rc) return rc;
    Slot &S = ctx->slots[slot];
    S.levels_built = 0; S.deriv_built = 0; S.border0 = false; S.raw_pending = true;
    return VO_OK;
}

extern "C" int vo_upload_image(vo_ctx *ctx, int slot, const uint8_t *data, int w, int h, size_t step)
{
    return set_image_common(ctx, slot, data, w, h, step, cudaMemcpyHostToDevice);
}

extern "C" int vo_set_image_d(vo_ctx *ctx, int slot, const uint8_t *data_d, int w, int h, size_t step)
{
    return set_image_common(ctx, slot, data_d, w, h, step, cudaMemcpyDeviceToDevice);
}

extern "C" int vo_build_pyramids(vo_ctx *ctx, const int *slots, int n_slots, int n_levels, int with_deriv)
{
    if (!ctx || !slots) return VO_ERR_INVALID_ARG;
    VO_CUDA(cudaSetDevice(ctx->device));
    if (n_levels > ctx->max_levels) n_levels = ctx->max_levels;
    return vo_ensure_pyramids(ctx, slots, n_slots, n_levels, with_deriv);
}

extern "C" int vo_invalidate_pyramids(vo_ctx *ctx, const int *slots, int n_slots)
{
    if (!ctx || !slots) return VO_ERR_INVALID_ARG;
    for (int i = 0; i < n_slots; ++i) {
        VO_REQUIRE(slots[i] >= 0 && slots[i] < ctx->n_slots, VO_ERR_INVALID_ARG, "slot id out of range");
        Slot &S = ctx->slots[slots[i]];
        S.levels_built = 0; S.deriv_built = 0; S.border0 = false;
    }
    return VO_OK;
}

extern "C" int vo_read_pyramid_level(vo_ctx *ctx, int slot, int level, uint8_t *img, int16_t *deriv, int *w_l, int *h_l)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(slot >= 0 && slot < ctx->n_slots, VO_ERR_INVALID_ARG, "slot id out of range");
    const Slot &S = ctx->slots[slot];
    VO_REQUIRE(level >= 0 && level < S.levels_built, VO_ERR_INVALID_ARG, "level not built");
    const LevelDesc &L = S.desc.lv[level];
    if (w_l) *w_l = L.w;
    if (h_l) *h_l = L.h;
    if (img) VO_CUDA(cudaMemcpy2DAsync(img, L.w, L.img, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost, ctx->stream));
    if (deriv) {
        VO_REQUIRE(level < S.deriv_built, VO_ERR_INVALID_ARG, "derivative level not built");
        VO_CUDA(cudaMemcpy2DAsync(deriv, (size_t)L.w * 4, L.deriv, (size_t)L.pitch * 4, (size_t)L.w * 4, L.h,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    }
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    return VO_OK;
}

// ------------------------------------------------------------------------------------ KLT
extern "C" int vo_klt_track_batch_d(vo_ctx *ctx, int n_pairs, const int *slots0, const int *slots1, const float *pts0_d,
                                    int n, int win, int max_level, int flags, float *pts1_inout_d, uint8_t *status_d,
                                    float *err_d, long long *counters_d)
{
    if (!ctx || !slots0 || !slots1) return VO_ERR_INVALID_ARG;
    VO_CUDA(cudaSetDevice(ctx->device));
    return vo_klt_launch(ctx, n_pairs, slots0, slots1, pts0_d, n, win, max_level, flags, pts1_inout_d, status_d, err_d,
                         counters_d, nullptr);
}

// Staging layout for the host-pointer entry points (one H2D + one D2H per call):
//   [pts0 n*8][pts1 n*8][ptsb n*8][err n*4][errb n*4][status n][statusb n][mask n]
struct KltStage {
    size_t o_pts0, o_pts1, o_ptsb, o_err, o_errb, o_st, o_stb, o_mask, total;
    explicit KltStage(int n)
    {
        size_t o = 0;
        o_pts0 = o; o += (size_t)n * 8;
        o_pts1 = o; o += (size_t)n * 8;
        o_ptsb = o; o += (size_t)n * 8;
        o_err = o; o += (size_t)n * 4;
        o_errb = o; o += (size_t)n * 4;
        o_st = o; o += align_up(n, 16);
        o_stb = o; o += align_up(n, 16);
        o_mask = o; o += align_up(n, 16);
        total = align_up(o, 256);
    }
};

extern "C" int vo_klt_track(vo_ctx *ctx, int slot0, int slot1, const float *pts0, int n, int win, int max_level,
                            int flags, float *pts1_inout, uint8_t *status, float *err)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0, VO_ERR_INVALID_ARG, "negative size");
    if (n == 0) return VO_OK;
    VO_REQUIRE(pts0 && pts1_inout && status && err, VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    KltStage L(n);
    int rc = vo_stage_reserve(ctx, L.total);
    if (rc) return rc;
    uint8_t *h = ctx->h_stage, *d = ctx->d_stage;
    memcpy(h + L.o_pts0, pts0, (size_t)n * 8);
    size_t up = (size_t)n * 8;
    if (flags & VO_KLT_USE_INITIAL_FLOW) { memcpy(h + L.o_pts1, pts1_inout, (size_t)n * 8); up = (size_t)n * 16; }
    VO_CUDA(cudaMemcpyAsync(d, h, up, cudaMemcpyHostToDevice, ctx->stream));
    rc = vo_klt_launch(ctx, 1, &slot0, &slot1, (const float *)(d + L.o_pts0), n, win, max_level, flags,
                       (float *)(d + L.o_pts1), d + L.o_st, (float *)(d + L.o_err), nullptr, nullptr);
    if (rc) return rc;
    VO_CUDA(cudaMemcpyAsync(h + L.o_pts1, d + L.o_pts1, L.o_stb - L.o_pts1, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(pts1_inout, h + L.o_pts1, (size_t)n * 8);
    memcpy(err, h + L.o_err, (size_t)n * 4);
    memcpy(status, h + L.o_st, (size_t)n);
    return VO_OK;
}

// FeatureTracker front-ends. mode: 1 track, 2 trackWithPrior, 3 trackBidirection, 4 trackBidirectionWithPrior
static int ft_common(vo_ctx *ctx, int mode, int slot0, int slot1, const float *pts0, int n, int win, int max_lvl,
                     float thres_err, float thres_bi, float *pts_track, uint8_t *mask)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0, VO_ERR_INVALID_ARG, "negative size");
    if (n == 0) return VO_OK;
    VO_REQUIRE(pts0 && pts_track && mask, VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    KltStage L(n);
    int rc = vo_stage_reserve(ctx, L.total);
    if (rc) return rc;
    uint8_t *h = ctx->h_stage, *d = ctx->d_stage;
    const bool prior = (mode == 2 || mode == 4);
    memcpy(h + L.o_pts0, pts0, (size_t)n * 8);
    if (prior) memcpy(h + L.o_pts1, pts_track, (size_t)n * 8);
    memcpy(h + L.o_mask, mask, (size_t)n);
    VO_CUDA(cudaMemcpyAsync(d, h, prior ? (size_t)n * 16 : (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    VO_CUDA(cudaMemcpyAsync(d + L.o_mask, h + L.o_mask, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    const float *d_pts0 = (const float *)(d + L.o_pts0);
    float *d_pts1 = (float *)(d + L.o_pts1);
    KltPost post{};
    post.thres_err = thres_err;
    post.mask = d + L.o_mask;
    if (mode == 1 || mode == 2) {
        post.mode = mode;
        rc = vo_klt_launch(ctx, 1, &slot0, &slot1, d_pts0, n, win, max_lvl, prior ? VO_KLT_USE_INITIAL_FLOW : 0, d_pts1,
                           d + L.o_st, (float *)(d + L.o_err), nullptr, &post);
        if (rc) return rc;
    } else {
        // forward pass (no mask yet), then backward pass seeded with pts0 whose epilogue fuses
        // the bidirectional validity test (feature_tracker.cpp:57-83 / :105-149)
        post.mode = 3;
        rc = vo_klt_launch(ctx, 1, &slot0, &slot1, d_pts0, n, win, max_lvl, prior ? VO_KLT_USE_INITIAL_FLOW : 0, d_pts1,
                           d + L.o_st, (float *)(d + L.o_err), nullptr, &post);
        if (rc) return rc;
        VO_CUDA(cudaMemcpyAsync(d + L.o_ptsb, d + L.o_pts0, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        post.mode = 4;
        post.border = (mode == 3) ? 3 : 0;
        post.thres_bi2 = (mode == 3) ? thres_bi * thres_bi : (thres_bi * thres_bi) * 5;
        post.ref_pts = d_pts0;
        post.fwd_pts = d_pts1;
        post.fwd_status = d + L.o_st;
        post.fwd_err = (const float *)(d + L.o_err);
        const int back_lvl = (mode == 3) ? max_lvl - 1 : max_lvl;
        rc = vo_klt_launch(ctx, 1, &slot1, &slot0, d_pts1, n, win, back_lvl < 0 ? 0 : back_lvl, VO_KLT_USE_INITIAL_FLOW,
                           (float *)(d + L.o_ptsb), d + L.o_stb, (float *)(d + L.o_errb), nullptr, &post);
        if (rc) return rc;
    }
    VO_CUDA(cudaMemcpyAsync(h + L.o_pts1, d + L.o_pts1, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaMemcpyAsync(h + L.o_mask, d + L.o_mask, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(pts_track, h + L.o_pts1, (size_t)n * 8);
    memcpy(mask, h + L.o_mask, (size_t)n);
    return VO_OK;
}

extern "C" int vo_ft_track(vo_ctx *ctx, int slot0, int slot1, const float *pts0, int n, int window_size, int max_pyr_lvl,
                           float thres_err, float *pts_track, uint8_t *mask_inout)
{
    return ft_common(ctx, 1, slot0, slot1, pts0, n, window_size, max_pyr_lvl, thres_err, 0.f, pts_track, mask_inout);
}
extern "C" int vo_ft_track_with_prior(vo_ctx *ctx, int slot0, int slot1, const float *pts0, int n, int window_size,
                                      int max_pyr_lvl, float thres_err, float *pts_track_inout, uint8_t *mask_inout)
{
    return ft_common(ctx, 2, slot0, slot1, pts0, n, window_size, max_pyr_lvl, thres_err, 0.f, pts_track_inout, mask_inout);
}
extern "C" int vo_ft_track_bidirection(vo_ctx *ctx, int slot0, int slot1, const float *pts0, int n, int window_size,
                                       int max_pyr_lvl, float thres_err, float thres_bidirection, float *pts_track,
                                       uint8_t *mask_inout)
{
    return ft_common(ctx, 3, slot0, slot1, pts0, n, window_size, max_pyr_lvl, thres_err, thres_bidirection, pts_track,
                     mask_inout);
}
extern "C" int vo_ft_track_bidirection_with_prior(vo_ctx *ctx, int slot0, int slot1, const float *pts0, int n,
                                                  int window_size, int max_pyr_lvl, float thres_err,
                                                  float thres_bidirection, float *pts_track_inout, uint8_t *mask_inout)
{
    return ft_common(ctx, 4, slot0, slot1, pts0, n, window_size, max_pyr_lvl, thres_err, thres_bidirection,
                     pts_track_inout, mask_inout);
}

// Upload a list of images; consecutive slots fed from consecutive dense host images are merged
// into ONE contiguous DMA (the raw staging areas of consecutive slots are adjacent when the
// image fills max_w x max_h), which is what reaches full PCIe bandwidth.
static int upload_many(vo_ctx *ctx, int n, const int *slots, const uint8_t *const *imgs, int w, int h, size_t step,
                       cudaStream_t st)
{
    if (!imgs) return VO_OK;
    const size_t img_bytes = (size_t)w * h;
    const bool mergeable = (step == (size_t)w) && (img_bytes == ctx->raw_stride);
    int i = 0;
    while (i < n) {
        if (!imgs[i]) { ++i; continue; }
        int j = i + 1;
        while (mergeable && j < n && imgs[j] == imgs[j - 1] + img_bytes && slots[j] == slots[j - 1] + 1) ++j;
        if (j - i > 1) {
            for (int k = i; k < j; ++k) {
                VO_REQUIRE(slots[k] >= 0 && slots[k] < ctx->n_slots, VO_ERR_INVALID_ARG, "slot id out of range");
                int rc = slot_set_geometry(ctx, slots[k], w, h);
                if (rc) return rc;
                Slot &S = ctx->slots[slots[k]];
                S.levels_built = 0; S.deriv_built = 0; S.border0 = false; S.raw_pending = true;
            }
            VO_CUDA(cudaMemcpyAsync(ctx->raw_base + (size_t)slots[i] * ctx->raw_stride, imgs[i], img_bytes * (j - i),
                                    cudaMemcpyHostToDevice, st));
        } else {
            int rc = set_image_common(ctx, slots[i], imgs[i], w, h, step, cudaMemcpyHostToDevice, st);
            if (rc) return rc;
        }
        i = j;
    }
    return VO_OK;
}

extern "C" int vo_ft_track_batch(vo_ctx *ctx, int n_pairs, const int *slots0, const int *slots1,
                                 const uint8_t *const *imgs0, const uint8_t *const *imgs1, int w, int h, size_t step,
                                 const float *pts0, int n, int window_size, int max_pyr_lvl, float thres_err,
                                 int with_prior, float *pts_track_inout, uint8_t *mask_inout)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0 && n_pairs >= 0, VO_ERR_INVALID_ARG, "negative size");
    if (n == 0 || n_pairs == 0) return VO_OK;
    VO_REQUIRE(slots0 && slots1 && pts0 && pts_track_inout && mask_inout, VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    int rc;
    const size_t N = (size_t)n_pairs * n;
    // staging: [pts0 N*8][pts1 N*8][err N*4][status N][mask N]
    const size_t o_p0 = 0, o_p1 = o_p0 + N * 8, o_err = o_p1 + N * 8, o_st = o_err + N * 4, o_mask = o_st + align_up(N, 16);
    const size_t total = o_mask + align_up(N, 16);
    rc = vo_stage_reserve(ctx, total);
    if (rc) return rc;
    uint8_t *hs = ctx->h_stage, *d = ctx->d_stage;
    // Caller buffers that are page-locked (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory) are DMA'd
    // directly; pageable ones go through the context's pinned staging (two extra host copies of every array).
    auto pinned = [](const void *p) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeHost;
    };
    const bool pin_p0 = pinned(pts0), pin_pt = pinned(pts_track_inout), pin_m = pinned(mask_inout);
    if (pin_p0) VO_CUDA(cudaMemcpyAsync(d + o_p0, pts0, N * 8, cudaMemcpyHostToDevice, ctx->stream));
    else { memcpy(hs + o_p0, pts0, N * 8); VO_CUDA(cudaMemcpyAsync(d + o_p0, hs + o_p0, N * 8, cudaMemcpyHostToDevice, ctx->stream)); }
    if (with_prior) {
        if (pin_pt) VO_CUDA(cudaMemcpyAsync(d + o_p1, pts_track_inout, N * 8, cudaMemcpyHostToDevice, ctx->stream));
        else { memcpy(hs + o_p1, pts_track_inout, N * 8); VO_CUDA(cudaMemcpyAsync(d + o_p1, hs + o_p1, N * 8, cudaMemcpyHostToDevice, ctx->stream)); }
    }
    if (pin_m) VO_CUDA(cudaMemcpyAsync(d + o_mask, mask_inout, N, cudaMemcpyHostToDevice, ctx->stream));
    else { memcpy(hs + o_mask, mask_inout, N); VO_CUDA(cudaMemcpyAsync(d + o_mask, hs + o_mask, N, cudaMemcpyHostToDevice, ctx->stream)); }
    // Chunked software pipeline: the image DMA of chunk c+1 (copy stream) overlaps the pyramid + LK
    // kernels of chunk c (compute stream). Slots of different chunks are disjoint, so the only
    // dependency is "chunk c's kernels wait for chunk c's upload" (one event per chunk).
    if (!ctx->copy_stream) VO_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    // Uniform chunks of 16 pairs round-robin over FOUR compute streams (measured on 128 KITTI-size pairs: the round-1 ramp
    // 8, 16, 32 ... on two streams 3.69 ms; 16-pair chunks on 2 / 3 / 4 streams 3.58 / 3.49 / 3.47 ms; 8-pair chunks on four
    // streams 3.43 - 3.56 ms; against 3.0 - 3.1 ms for the same call with the images already resident).
    constexpr int CH = 16, NS = 4, SZ0 = 16;
    std::vector<int> chunk_begin;
    for (int c0 = 0, sz = SZ0; c0 < n_pairs; ) {
        chunk_begin.push_back(c0);
        c0 += sz < CH ? sz : CH;
        sz *= 2;
    }
    chunk_begin.push_back(n_pairs);
    const int n_chunks = (int)chunk_begin.size() - 1;
    // The two-stream / copy-stream overlap below is only valid when no slot is touched by more than one chunk.  A chained
    // batch (slots1[i] == slots0[i+1], allowed by the header) or any other reuse across chunks runs every chunk, uploads
    // included, in order on the context's stream instead.
    bool chained = false;
    {
        std::vector<int> owner(ctx->n_slots, -1);
        for (int c = 0; c < n_chunks && !chained; ++c)
            for (int i = chunk_begin[c]; i < chunk_begin[c + 1] && !chained; ++i)
                for (int sid : {slots0[i], slots1[i]}) {
                    VO_REQUIRE(sid >= 0 && sid < ctx->n_slots, VO_ERR_INVALID_ARG, "slot id out of range");
                    if (owner[sid] >= 0 && owner[sid] != c) { chained = true; break; }
                    owner[sid] = c;
                }
    }
    while ((int)ctx->events.size() < n_chunks + 1) {
        cudaEvent_t e;
        VO_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->events.push_back(e);
    }
    // the uploads must not overtake work already queued on the compute stream that still reads these slots
    VO_CUDA(cudaEventRecord(ctx->events[n_chunks], ctx->stream));
    VO_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->events[n_chunks], 0));
    KltPost post{};
    post.mode = with_prior ? 2 : 1;
    post.thres_err = thres_err;
    // Chunks go round-robin over the compute streams so that the tail wave of chunk c's LK kernel overlaps the pyramids /
    // heads of the next chunks (chunks touch disjoint slots and disjoint ranges of the point arrays).
    for (int i = 0; i < 3; ++i) if (!ctx->side[i]) VO_CUDA(cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking));
    if (!ctx->ev_aux[0]) for (int i = 0; i < 4; ++i) VO_CUDA(cudaEventCreateWithFlags(&ctx->ev_aux[i], cudaEventDisableTiming));
    cudaStream_t main_stream = ctx->stream;
    VO_CUDA(cudaEventRecord(ctx->ev_aux[0], main_stream));          // the point / mask uploads above
    for (int i = 0; i < 3; ++i) VO_CUDA(cudaStreamWaitEvent(ctx->side[i], ctx->ev_aux[0], 0));
    // error exits: nothing may stay queued on the side streams (vo_stage_reserve / the next call only order against ctx->stream)
    auto drain = [&](int code) {
        ctx->stream = main_stream;
        cudaStreamSynchronize(ctx->copy_stream);
        for (int i = 0; i < 3; ++i) cudaStreamSynchronize(ctx->side[i]);
        cudaStreamSynchronize(main_stream);
        return code;
    };
    for (int c = 0; c < n_chunks; ++c) {
        const int c0 = chunk_begin[c], nc = chunk_begin[c + 1] - c0;
        cudaStream_t up = chained ? main_stream : ctx->copy_stream;
        rc = upload_many(ctx, nc, slots0 + c0, imgs0 ? imgs0 + c0 : nullptr, w, h, step, up);
        if (rc) return drain(rc);
        rc = upload_many(ctx, nc, slots1 + c0, imgs1 ? imgs1 + c0 : nullptr, w, h, step, up);
        if (rc) return drain(rc);
        cudaStream_t cs = (!chained && (c % NS)) ? ctx->side[c % NS - 1] : main_stream;
        if (!chained) {
            if (cudaEventRecord(ctx->events[c], ctx->copy_stream) != cudaSuccess || cudaStreamWaitEvent(cs, ctx->events[c], 0) != cudaSuccess) {
                ctx->last_error = "vo_ft_track_batch: event record / wait failed";
                return drain(VO_ERR_CUDA);
            }
        }
        const size_t off = (size_t)c0 * n;
        post.mask = d + o_mask + off;
        ctx->stream = cs;                                            // the launch helpers enqueue on ctx->stream
        rc = vo_klt_launch(ctx, nc, slots0 + c0, slots1 + c0, (const float *)(d + o_p0) + 2 * off, n, window_size, max_pyr_lvl,
                           with_prior ? VO_KLT_USE_INITIAL_FLOW : 0, (float *)(d + o_p1) + 2 * off, d + o_st + off,
                           (float *)(d + o_err) + off, nullptr, &post);
        ctx->stream = main_stream;
        if (rc) return drain(rc);
    }
    bool join_ok = true;
    for (int i = 0; i < 3; ++i)
        join_ok = join_ok && cudaEventRecord(ctx->ev_aux[1 + i], ctx->side[i]) == cudaSuccess && cudaStreamWaitEvent(main_stream, ctx->ev_aux[1 + i], 0) == cudaSuccess;
    if (!join_ok) {
        ctx->last_error = "vo_ft_track_batch: joining the second compute stream failed";
        return drain(VO_ERR_CUDA);
    }
    VO_CUDA(cudaMemcpyAsync(pin_pt ? (void *)pts_track_inout : (void *)(hs + o_p1), d + o_p1, N * 8, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaMemcpyAsync(pin_m ? (void *)mask_inout : (void *)(hs + o_mask), d + o_mask, N, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!pin_pt) memcpy(pts_track_inout, hs + o_p1, N * 8);
    if (!pin_m) memcpy(mask_inout, hs + o_mask, N);
    return VO_OK;
}
