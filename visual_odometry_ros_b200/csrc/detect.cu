// detect.cu -- K-det: bucketed corner extraction on the resident Scharr derivative plane.
//
// Replaces FeatureExtractor::updateWeightBin + extractORBwithBinning_fast
// (core/visual_odometry/feature_extractor.cpp:94-98, 211-282; WeightBin feature_extractor.h:56-134) for the
// device-resident step: bins that already hold a tracked point get weight 0, every other bin returns its
// best-response keypoint, output in bin-index order.  The keypoint response of the reference is cv::ORB's
// (third-party OpenCV, not under /root/reference); here it is an exact-integer Harris response evaluated on
// the Scharr plane the KLT needs anyway (so detection costs no extra image pass and no image D2H):
//     a = sum_7x7 Ix^2 >> 10, b = sum_7x7 Ix*Iy >> 10, c = sum_7x7 Iy^2 >> 10
//     score = 25 * (a*c - b*b) - (a+c)^2          (int64; Harris with k = 1/25)
// Candidates are integer pixels with edge <= x < w-edge, edge <= y < h-edge, score > min_score; ties inside a
// bin go to the first pixel in raster order.  Bit-exact against its numpy restatement (test infrastructure).
//
// Launch sequence (all asynchronous on the context's stream):
//   k_det_reset  : weight = 1, best = min_score, arg = INT_MAX
//   k_det_mark   : WeightBin::update for the occupied points (count may live on the device)
//   k_det_score  : one thread per candidate pixel of a non-occupied bin: 49-tap sums from L1/L2, score plane
//                  store, per-bin atomicMax (pre-filtered by a plain read, so few atomics are issued)
//   k_det_pick   : pixels whose score equals their bin's maximum -> atomicMin of the raster index
//   k_det_emit   : one CTA, ordered scan over the bins -> point list, count, skip mask for the KLT
#include "vo_internal.cuh"

#include <climits>
#include <cstring>

struct DetArgs {
    LevelDesc L;                 // level 0 of the slot (with derivative plane)
    int n_bins_u, n_bins_v, u_step, v_step, edge;
    long long min_score;
    int *weight;                 // [bins]
    long long *best;             // [bins]
    int *arg;                    // [bins]
    long long *score;            // [w*h] scratch plane
    const float2 *occ;           // occupied points
    const int *n_occ_d;          // device count (nullable -> n_occ)
    int n_occ;
    float2 *out;                 // [max_out]
    uint8_t *out_mask;           // [max_out] 1 for the emitted points, 0 for the rest
    int *n_out;
    int max_out;
};

__global__ void __launch_bounds__(256) k_det_reset(const DetArgs a)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < a.n_bins_u * a.n_bins_v) { a.weight[b] = 1; a.best[b] = a.min_score; a.arg[b] = INT_MAX; }
}

// WeightBin::update (feature_extractor.h:119-131): floor((float)p.x / (float)u_step), no per-axis range check
__global__ void __launch_bounds__(256) k_det_mark(const DetArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = a.n_occ_d ? *a.n_occ_d : a.n_occ;
    if (i >= n) return;
    const float2 p = a.occ[i];
    const long long ui = (long long)floorf(__fdiv_rn(p.x, (float)a.u_step));
    const long long vi = (long long)floorf(__fdiv_rn(p.y, (float)a.v_step));
    const long long b = vi * a.n_bins_u + ui;
    if (b >= 0 && b < (long long)a.n_bins_u * a.n_bins_v) a.weight[b] = 0;
}

__device__ __forceinline__ int det_bin(const DetArgs &a, int x, int y)
{
    // (int)floor(ft.pt.x * inv_u_step_) (feature_extractor.cpp:254-255)
    const float inv_u = __fdiv_rn(1.0f, (float)a.u_step), inv_v = __fdiv_rn(1.0f, (float)a.v_step);
    const int u = (int)floorf(__fmul_rn((float)x, inv_u));
    const int v = (int)floorf(__fmul_rn((float)y, inv_v));
    if (u >= a.n_bins_u || v >= a.n_bins_v) return -1;
    return v * a.n_bins_u + u;
}

__global__ void __launch_bounds__(256) k_det_score(const DetArgs a)
{
    const int x = a.edge + blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = a.edge + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= a.L.w - a.edge || y >= a.L.h - a.edge) return;
    const int b = det_bin(a, x, y);
    if (b < 0 || a.weight[b] == 0) return;
    int sa = 0, sb = 0, sc = 0;       // |sum| <= 49 * 4080^2 < 2^31
    const short2 *row = a.L.deriv + (ptrdiff_t)(y - 3) * a.L.pitch + (x - 3);
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
#pragma unroll
        for (int dx = 0; dx < 7; ++dx) {
            const short2 d = __ldg(row + dx);
            sa += (int)d.x * d.x; sb += (int)d.x * d.y; sc += (int)d.y * d.y;
        }
        row += a.L.pitch;
    }
    const long long A = sa >> 10, B = sb >> 10, C = sc >> 10;
    const long long s = 25 * (A * C - B * B) - (A + C) * (A + C);
    a.score[(size_t)y * a.L.w + x] = s;
    if (s > a.best[b]) atomicMax(a.best + b, s);
}

__global__ void __launch_bounds__(256) k_det_pick(const DetArgs a)
{
    const int x = a.edge + blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = a.edge + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= a.L.w - a.edge || y >= a.L.h - a.edge) return;
    const int b = det_bin(a, x, y);
    if (b < 0 || a.weight[b] == 0) return;
    const long long s = a.score[(size_t)y * a.L.w + x];
    if (s > a.min_score && s == a.best[b]) atomicMin(a.arg + b, y * a.L.w + x);
}

__global__ void __launch_bounds__(1024) k_det_emit(const DetArgs a)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    const int nb = a.n_bins_u * a.n_bins_v;
    for (int c0 = 0; c0 < nb; c0 += 1024) {
        const int b = c0 + tid;
        const int r = b < nb ? a.arg[b] : INT_MAX;
        const bool keep = b < nb && r != INT_MAX && a.weight[b] > 0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const int within = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        if (wid == 0) {
            int v = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            s_warp[lane] = v;
        }
        __syncthreads();
        const int pos = s_base + (wid ? s_warp[wid - 1] : 0) + within;
        if (keep && pos < a.max_out) a.out[pos] = make_float2((float)(r % a.L.w), (float)(r / a.L.w));
        __syncthreads();
        if (tid == 0) s_base += s_warp[31];
        __syncthreads();
    }
    const int n = s_base < a.max_out ? s_base : a.max_out;
    if (tid == 0) *a.n_out = n;
    if (a.out_mask)
        for (int i = tid; i < a.max_out; i += 1024) a.out_mask[i] = i < n ? 1 : 0;
}

// Device-pointer launcher used by the fused frame step. occ_d / n_occ_d: occupied points (count on device,
// nullable -> n_occ); bins_d: scratch of n_bins*(4+8+4) bytes; out_d [max_out] float2, out_mask_d [max_out].
int vo_detect_launch_d(vo_ctx *ctx, int slot, const float *occ_d, const int *n_occ_d, int n_occ, int n_bins_u, int n_bins_v,
                       int edge, long long min_score, float *out_d, uint8_t *out_mask_d, int *n_out_d, int max_out)
{
    VO_REQUIRE(slot >= 0 && slot < ctx->n_slots && ctx->slots[slot].w > 0, VO_ERR_INVALID_ARG, "slot has no image");
    VO_REQUIRE(n_bins_u > 0 && n_bins_v > 0 && edge >= 3 && edge <= VO_PAD + 3, VO_ERR_INVALID_ARG, "bad detector arguments");
    if (ctx->detector == VO_DETECTOR_ORB)        // the reference's own keypoints (cv::ORB), orb.cu; min_score has no meaning there
        return vo_orb_launch_d(ctx, slot, occ_d, n_occ_d, n_occ, n_bins_u, n_bins_v, edge < 4 ? 4 : edge, out_d, out_mask_d, n_out_d, max_out,
                               nullptr, nullptr, nullptr, nullptr, 0);
    const Slot &S = ctx->slots[slot];
    VO_REQUIRE(S.w / n_bins_u >= 1 && S.h / n_bins_v >= 1, VO_ERR_INVALID_ARG, "more bins than pixels");
    int rc = vo_ensure_pyramids(ctx, &slot, 1, 1, 1);
    if (rc) return rc;
    const int nb = n_bins_u * n_bins_v;
    const size_t need = (size_t)S.w * S.h * 8 + (size_t)nb * 16 + 256;
    if (need > ctx->det_bytes) {
        VO_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_det) cudaFree(ctx->d_det);
        ctx->d_det = nullptr; ctx->det_bytes = 0;
        VO_CUDA(cudaMalloc(&ctx->d_det, need));
        ctx->det_bytes = need;
    }
    DetArgs a;
    a.L = S.desc.lv[0];
    a.n_bins_u = n_bins_u; a.n_bins_v = n_bins_v;
    a.u_step = S.w / n_bins_u;            // (int)floor((float)n_cols / (float)n_bins_u), exact for these magnitudes
    a.v_step = S.h / n_bins_v;
    a.edge = edge; a.min_score = min_score;
    uint8_t *p = (uint8_t *)ctx->d_det;
    a.score = (long long *)p; p += (size_t)S.w * S.h * 8;
    a.best = (long long *)p; p += (size_t)nb * 8;
    a.weight = (int *)p; p += (size_t)nb * 4;
    a.arg = (int *)p;
    a.occ = (const float2 *)occ_d; a.n_occ_d = n_occ_d; a.n_occ = n_occ;
    a.out = (float2 *)out_d; a.out_mask = out_mask_d; a.n_out = n_out_d; a.max_out = max_out;
    k_det_reset<<<vo_div_up(nb, 256), 256, 0, ctx->stream>>>(a);
    if (n_occ > 0) k_det_mark<<<vo_div_up(n_occ, 256), 256, 0, ctx->stream>>>(a);
    const int cw = S.w - 2 * edge, ch = S.h - 2 * edge;
    if (cw > 0 && ch > 0) {
        dim3 grd(vo_div_up(cw, 32), vo_div_up(ch, 8));
        k_det_score<<<grd, 256, 0, ctx->stream>>>(a);
        k_det_pick<<<grd, 256, 0, ctx->stream>>>(a);
        ctx->launches += 2;
    }
    k_det_emit<<<1, 1024, 0, ctx->stream>>>(a);
    ctx->launches += 2 + (n_occ > 0 ? 1 : 0);
    VO_CUDA(cudaGetLastError());
    return VO_OK;
}

extern "C" int vo_detect_bucketed(vo_ctx *ctx, int slot, const float *pts_occupied, int n_occupied, int n_bins_u, int n_bins_v,
                                  int edge, long long min_score, float *pts_out, int max_out, int *n_out)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n_occupied >= 0 && max_out >= 0 && n_out && (max_out == 0 || pts_out), VO_ERR_INVALID_ARG, "bad arguments");
    VO_REQUIRE(n_occupied == 0 || pts_occupied, VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    const size_t o_occ = 0, o_out = (size_t)(n_occupied + 1) * 8, o_n = o_out + (size_t)(max_out + 1) * 8, total = o_n + 64;
    int rc = vo_stage_reserve(ctx, total);
    if (rc) return rc;
    uint8_t *h = ctx->h_stage, *d = ctx->d_stage;
    if (n_occupied) {
        memcpy(h + o_occ, pts_occupied, (size_t)n_occupied * 8);
        VO_CUDA(cudaMemcpyAsync(d + o_occ, h + o_occ, (size_t)n_occupied * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    rc = vo_detect_launch_d(ctx, slot, (const float *)(d + o_occ), nullptr, n_occupied, n_bins_u, n_bins_v, edge, min_score,
                            (float *)(d + o_out), nullptr, (int *)(d + o_n), max_out);
    if (rc) return rc;
    VO_CUDA(cudaMemcpyAsync(h + o_out, d + o_out, total - o_out, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    const int n = *(const int *)(h + o_n);
    *n_out = n;
    if (n) memcpy(pts_out, h + o_out, (size_t)n * 8);
    return VO_OK;
}
