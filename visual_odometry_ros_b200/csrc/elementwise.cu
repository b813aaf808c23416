// elementwise.cu -- K-tri, K-df, K-prior, K-compact: the batched elementwise rows of the hot path.
//
//   K-tri     mapping::triangulateDLT          core/util/triangulate_3d.cpp:5-130
//   K-df      DepthFilter::update*Distribution standalone/depth_filter/depth_filter.cpp:3-46
//   K-prior   FeatureTracker::calcPrior        core/visual_odometry/feature_tracker.cpp:208-234
//   K-compact LandmarkTracking(src, mask)      core/visual_odometry/landmark.cpp:194-231, 291-332
//
// One thread per element, coalesced SoA-style streaming; all arithmetic in the reference's
// precision (FP32 for K-tri / K-prior, FP64 for K-df) with the reference's operation order --
// this file is compiled with -fmad=false so results match the non-FMA CPU build operation by
// operation.  These kernels are pure HBM-streaming work (16 B in + 24 B out per triangulated
// point, 48 B / 128 B per depth seed): they are launch-latency-bound at the reference's sizes.
#include "vo_internal.cuh"

#include <cfloat>
#include <cstring>

#include "tri_device.cuh"

struct TriArgs {
    const float2 *p0, *p1;
    float *X0, *X1;
    int n;
    float R10[9], t10[3], K0[4], K1[4];
};

__global__ void __launch_bounds__(128) k_triangulate(const TriArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    float X0[3], X1[3];
    tri_point(a.p0[i], a.p1[i], a.R10, a.t10, a.K0, a.K1, X0, X1);
#pragma unroll
    for (int r = 0; r < 3; ++r) { a.X0[3 * i + r] = X0[r]; a.X1[3 * i + r] = X1[r]; }
}

// the same with one relative pose per GROUP of points (MonoVO's keyframe reconstruction pairs every landmark with the frame
// of its first observation: a handful of distinct frames per keyframe, one launch instead of one call per frame)
struct TriGroupArgs {
    const float2 *p0, *p1;
    const int *group;          // [n]
    const float *Rt;           // [n_groups][12]: R10 row-major (9), t10 (3)
    float *X0, *X1;
    int n;
    float K0[4], K1[4];
};

__global__ void __launch_bounds__(128) k_triangulate_grouped(const TriGroupArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const float *rt = a.Rt + 12 * a.group[i];
    float R10[9], t10[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) R10[k] = rt[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) t10[k] = rt[9 + k];
    float X0[3], X1[3];
    tri_point(a.p0[i], a.p1[i], R10, t10, a.K0, a.K1, X0, X1);
#pragma unroll
    for (int r = 0; r < 3; ++r) { a.X0[3 * i + r] = X0[r]; a.X1[3 * i + r] = X1[r]; }
}

// ------------------------------------------------------------------------------ K-df
__global__ void __launch_bounds__(256)
k_df_normal(const double *xp, const double *cp, const double *__restrict__ xc,
            const double *__restrict__ cc, int n, double *xu, double *cu)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x0 = xp[i], c0 = cp[i], x1 = xc[i], c1 = cc[i];     // all reads first: xu / cu may alias xp / cp
    const double inv_cov_sum = 1.0 / (c0 + c1);
    cu[i] = (c0 * c1) * inv_cov_sum;
    xu[i] = (x0 * c1 + x1 * c0) * inv_cov_sum;
}

__global__ void __launch_bounds__(256)
k_df_student_t(const double *xp, const double *cp, double *__restrict__ a,
               double *__restrict__ b, double *__restrict__ xmin, double *__restrict__ xmax,
               const double *__restrict__ xc, const double *__restrict__ cc, int n, double *xu,
               double *cu)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double mu = xp[i], s2 = cp[i], t2 = cc[i], x = xc[i];
    const double ai = a[i], bi = b[i];
    const double inv_apb = 1.0 / (ai + bi);
    const double x_range = xmax[i] - xmin[i];
    const double sigma = sqrt(s2);
    double C1 = ai * inv_apb * 1.0 / sqrt(2.0 * 3.141592) / sigma * exp(-(x - mu) * (x - mu) / (2.0 * s2));
    double C2 = bi * inv_apb / x_range;
    const double invC = 1.0 / (C1 + C2);
    C1 *= invC;
    C2 *= invC;
    const double ss = 1.0 / (1.0 / s2 + 1.0 / t2);
    const double m = ss * (mu / s2 + x / t2);
    const double mu_new = C1 * m + C2 * mu;
    const double var_new = C1 * (ss + m * m) + C2 * (s2 + mu * mu) - mu_new * mu_new;
    const double F = C1 * (ai + 1.0) / (ai + bi + 1.0) + C2 * ai / (ai + bi + 1.0);
    const double E = C1 * (ai + 1.0) / (ai + bi + 1.0) * (ai + 2.0) / (ai + bi + 2.0) +
                     C2 * ai / (ai + bi + 1.0) * (ai + 1.0) / (ai + bi + 2.0);
    const double a_new = (E - F) / (F - E / F);
    a[i] = a_new;
    b[i] = a_new * (1.0 - F) / F;
    xu[i] = mu_new;
    cu[i] = var_new;
    xmin[i] = fmin(xmin[i], x);
    xmax[i] = fmax(xmax[i], x);
}

// ------------------------------------------------------------------------------ K-prior
struct PriorArgs {
    const float2 *pts0;
    const float *Xw;
    float2 *out;
    int n;
    float T1w[12];
    float K[4];
};

__global__ void __launch_bounds__(256) k_calc_prior(const PriorArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const float x0 = a.Xw[3 * i], x1 = a.Xw[3 * i + 1], x2 = a.Xw[3 * i + 2];
    float X1[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) X1[r] = ((a.T1w[r * 4 + 0] * x0 + a.T1w[r * 4 + 1] * x1) + a.T1w[r * 4 + 2] * x2) + a.T1w[r * 4 + 3];
    const float nrm = sqrtf((X1[0] * X1[0] + X1[1] * X1[1]) + X1[2] * X1[2]);
    float2 o = a.pts0[i];
    if (nrm > 0.f) {
        o.x = a.K[0] * X1[0] / X1[2] + a.K[2];
        o.y = a.K[1] * X1[1] / X1[2] + a.K[3];
    }
    a.out[i] = o;
}

// ------------------------------------------------------------------------------ K-compact
// Stable stream compaction by one CTA: ballot/popc scan inside warps, shared-memory scan
// across warps, running base across chunks (feature counts are a few thousand).
__global__ void __launch_bounds__(1024) k_compact(const uint8_t *__restrict__ mask, int n, int *__restrict__ index_out,
                                                   int *__restrict__ n_out)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += 1024) {
        const int i = c0 + tid;
        const bool keep = (i < n) && mask[i] != 0;
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
            s_warp[lane] = v;   // inclusive
        }
        __syncthreads();
        const int base = s_base + (wid ? s_warp[wid - 1] : 0);
        if (keep) index_out[base + within] = i;
        __syncthreads();
        if (tid == 0) s_base += s_warp[31];
        __syncthreads();
    }
    if (tid == 0) *n_out = s_base;
}

// ------------------------------------------------------------------------------ host side
static void inverse4_host(const float *m, float *out)
{
    float inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    const float id = 1.0f / det;
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * id;
}

extern "C" int vo_triangulate_dlt(vo_ctx *ctx, const float *pts0, const float *pts1, int n, const float *R10,
                                  const float *t10, const float *K0_4, const float *K1_4, float *X0, float *X1)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0, VO_ERR_INVALID_ARG, "negative size");
    if (n == 0) return VO_OK;
    VO_REQUIRE(pts0 && pts1 && R10 && t10 && K0_4 && K1_4 && X0 && X1, VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    const size_t o0 = 0, o1 = (size_t)n * 8, oX0 = (size_t)n * 16, oX1 = oX0 + (size_t)n * 12, total = oX1 + (size_t)n * 12;
    int rc = vo_stage_reserve(ctx, total);
    if (rc) return rc;
    uint8_t *h = ctx->h_stage, *d = ctx->d_stage;
    memcpy(h + o0, pts0, (size_t)n * 8);
    memcpy(h + o1, pts1, (size_t)n * 8);
    VO_CUDA(cudaMemcpyAsync(d, h, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
    TriArgs a;
    a.p0 = (const float2 *)(d + o0); a.p1 = (const float2 *)(d + o1);
    a.X0 = (float *)(d + oX0); a.X1 = (float *)(d + oX1); a.n = n;
    memcpy(a.R10, R10, 36); memcpy(a.t10, t10, 12); memcpy(a.K0, K0_4, 16); memcpy(a.K1, K1_4, 16);
    k_triangulate<<<vo_div_up(n, 128), 128, 0, ctx->stream>>>(a);
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    VO_CUDA(cudaMemcpyAsync(h + oX0, d + oX0, (size_t)n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(X0, h + oX0, (size_t)n * 12);
    memcpy(X1, h + oX1, (size_t)n * 12);
    return VO_OK;
}

extern "C" int vo_triangulate_dlt_grouped(vo_ctx *ctx, const float *pts0, const float *pts1, int n, const int *group, int n_groups,
                                          const float *R10s, const float *t10s, const float *K0_4, const float *K1_4, float *X0, float *X1)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0 && n_groups >= 0, VO_ERR_INVALID_ARG, "negative size");
    if (n == 0) return VO_OK;
    VO_REQUIRE(pts0 && pts1 && group && R10s && t10s && K0_4 && K1_4 && X0 && X1 && n_groups > 0, VO_ERR_INVALID_ARG, "null pointer");
    for (int i = 0; i < n; ++i) VO_REQUIRE(group[i] >= 0 && group[i] < n_groups, VO_ERR_INVALID_ARG, "group index out of range");
    VO_CUDA(cudaSetDevice(ctx->device));
    const size_t N = (size_t)n;
    const size_t o0 = 0, o1 = N * 8, oG = N * 16, oRt = oG + (N * 4 + 15) / 16 * 16, in_bytes = oRt + (size_t)n_groups * 48;
    const size_t oX0 = (in_bytes + 15) / 16 * 16, oX1 = oX0 + N * 12, total = oX1 + N * 12;
    int rc = vo_stage_reserve(ctx, total);
    if (rc) return rc;
    uint8_t *h = ctx->h_stage, *d = ctx->d_stage;
    memcpy(h + o0, pts0, N * 8);
    memcpy(h + o1, pts1, N * 8);
    memcpy(h + oG, group, N * 4);
    for (int g = 0; g < n_groups; ++g) { memcpy(h + oRt + (size_t)g * 48, R10s + 9 * g, 36); memcpy(h + oRt + (size_t)g * 48 + 36, t10s + 3 * g, 12); }
    VO_CUDA(cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    TriGroupArgs a;
    a.p0 = (const float2 *)(d + o0); a.p1 = (const float2 *)(d + o1); a.group = (const int *)(d + oG); a.Rt = (const float *)(d + oRt);
    a.X0 = (float *)(d + oX0); a.X1 = (float *)(d + oX1); a.n = n;
    memcpy(a.K0, K0_4, 16); memcpy(a.K1, K1_4, 16);
    k_triangulate_grouped<<<vo_div_up(n, 128), 128, 0, ctx->stream>>>(a);
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    VO_CUDA(cudaMemcpyAsync(h + oX0, d + oX0, N * 24, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(X0, h + oX0, N * 12);
    memcpy(X1, h + oX1, N * 12);
    return VO_OK;
}

extern "C" int vo_depth_filter_normal(vo_ctx *ctx, const double *x_prev, const double *cov_prev, const double *x_curr,
                                      const double *cov_curr, int n, double *x_upd, double *cov_upd)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0, VO_ERR_INVALID_ARG, "negative size");
    if (n == 0) return VO_OK;
    VO_REQUIRE(x_prev && cov_prev && x_curr && cov_curr && x_upd && cov_upd, VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    const size_t N = (size_t)n * 8;
    int rc = vo_stage_reserve(ctx, 6 * N);
    if (rc) return rc;
    uint8_t *h = ctx->h_stage, *d = ctx->d_stage;
    memcpy(h, x_prev, N); memcpy(h + N, cov_prev, N); memcpy(h + 2 * N, x_curr, N); memcpy(h + 3 * N, cov_curr, N);
    VO_CUDA(cudaMemcpyAsync(d, h, 4 * N, cudaMemcpyHostToDevice, ctx->stream));
    double *D = (double *)d;
    k_df_normal<<<vo_div_up(n, 256), 256, 0, ctx->stream>>>(D, D + n, D + 2 * n, D + 3 * n, n, D + 4 * n, D + 5 * n);
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    VO_CUDA(cudaMemcpyAsync(h + 4 * N, d + 4 * N, 2 * N, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(x_upd, h + 4 * N, N);
    memcpy(cov_upd, h + 5 * N, N);
    return VO_OK;
}

extern "C" int vo_depth_filter_student_t(vo_ctx *ctx, const double *x_prev, const double *cov_prev, double *a_inout,
                                         double *b_inout, double *x_min_inout, double *x_max_inout, const double *x_curr,
                                         const double *cov_curr, int n, double *x_upd, double *cov_upd)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0, VO_ERR_INVALID_ARG, "negative size");
    if (n == 0) return VO_OK;
    VO_REQUIRE(x_prev && cov_prev && a_inout && b_inout && x_min_inout && x_max_inout && x_curr && cov_curr && x_upd && cov_upd,
               VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    const size_t N = (size_t)n * 8;
    int rc = vo_stage_reserve(ctx, 10 * N);
    if (rc) return rc;
    uint8_t *h = ctx->h_stage, *d = ctx->d_stage;
    // layout: [a b xmin xmax | xu cu] (in-out + out, contiguous for one D2H) then inputs
    memcpy(h, a_inout, N); memcpy(h + N, b_inout, N); memcpy(h + 2 * N, x_min_inout, N); memcpy(h + 3 * N, x_max_inout, N);
    memcpy(h + 6 * N, x_prev, N); memcpy(h + 7 * N, cov_prev, N); memcpy(h + 8 * N, x_curr, N); memcpy(h + 9 * N, cov_curr, N);
    VO_CUDA(cudaMemcpyAsync(d, h, 4 * N, cudaMemcpyHostToDevice, ctx->stream));
    VO_CUDA(cudaMemcpyAsync(d + 6 * N, h + 6 * N, 4 * N, cudaMemcpyHostToDevice, ctx->stream));
    double *D = (double *)d;
    k_df_student_t<<<vo_div_up(n, 256), 256, 0, ctx->stream>>>(D + 6 * n, D + 7 * n, D, D + n, D + 2 * n, D + 3 * n,
                                                              D + 8 * n, D + 9 * n, n, D + 4 * n, D + 5 * n);
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    VO_CUDA(cudaMemcpyAsync(h, d, 6 * N, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(a_inout, h, N); memcpy(b_inout, h + N, N); memcpy(x_min_inout, h + 2 * N, N); memcpy(x_max_inout, h + 3 * N, N);
    memcpy(x_upd, h + 4 * N, N); memcpy(cov_upd, h + 5 * N, N);
    return VO_OK;
}

// Device-resident forms: the seed arrays stay in HBM between updates (a 20 000-seed update through host buffers is a
// 1 MB PCIe round trip around a 2 us kernel -- slower than one CPU core; resident, it is the kernel alone).
extern "C" int vo_depth_filter_normal_d(vo_ctx *ctx, const double *x_prev_d, const double *cov_prev_d, const double *x_curr_d,
                                        const double *cov_curr_d, int n, double *x_upd_d, double *cov_upd_d)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0, VO_ERR_INVALID_ARG, "negative size");
    if (n == 0) return VO_OK;
    VO_REQUIRE(x_prev_d && cov_prev_d && x_curr_d && cov_curr_d && x_upd_d && cov_upd_d, VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    k_df_normal<<<vo_div_up(n, 256), 256, 0, ctx->stream>>>(x_prev_d, cov_prev_d, x_curr_d, cov_curr_d, n, x_upd_d, cov_upd_d);
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    return VO_OK;
}

extern "C" int vo_depth_filter_student_t_d(vo_ctx *ctx, const double *x_prev_d, const double *cov_prev_d, double *a_inout_d,
                                           double *b_inout_d, double *x_min_inout_d, double *x_max_inout_d, const double *x_curr_d,
                                           const double *cov_curr_d, int n, double *x_upd_d, double *cov_upd_d)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0, VO_ERR_INVALID_ARG, "negative size");
    if (n == 0) return VO_OK;
    VO_REQUIRE(x_prev_d && cov_prev_d && a_inout_d && b_inout_d && x_min_inout_d && x_max_inout_d && x_curr_d && cov_curr_d && x_upd_d && cov_upd_d,
               VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    k_df_student_t<<<vo_div_up(n, 256), 256, 0, ctx->stream>>>(x_prev_d, cov_prev_d, a_inout_d, b_inout_d, x_min_inout_d, x_max_inout_d,
                                                              x_curr_d, cov_curr_d, n, x_upd_d, cov_upd_d);
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    return VO_OK;
}

extern "C" int vo_ft_calc_prior(vo_ctx *ctx, const float *pts0, const float *Xw, int n, const float *Tw1, const float *K4,
                                float *pts1_prior)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0, VO_ERR_INVALID_ARG, "negative size");
    if (n == 0) return VO_OK;
    VO_REQUIRE(pts0 && Xw && Tw1 && K4 && pts1_prior, VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    const size_t oP = 0, oX = (size_t)n * 8, oO = oX + (size_t)n * 12, total = oO + (size_t)n * 8;
    int rc = vo_stage_reserve(ctx, total);
    if (rc) return rc;
    uint8_t *h = ctx->h_stage, *d = ctx->d_stage;
    memcpy(h + oP, pts0, (size_t)n * 8);
    memcpy(h + oX, Xw, (size_t)n * 12);
    VO_CUDA(cudaMemcpyAsync(d, h, oO, cudaMemcpyHostToDevice, ctx->stream));
    PriorArgs a;
    a.pts0 = (const float2 *)(d + oP); a.Xw = (const float *)(d + oX); a.out = (float2 *)(d + oO); a.n = n;
    float T1w[16];
    inverse4_host(Tw1, T1w);   // Tw1.inverse(), feature_tracker.cpp:215 (a 4x4 parameter, not data)
    memcpy(a.T1w, T1w, 48);
    memcpy(a.K, K4, 16);
    k_calc_prior<<<vo_div_up(n, 256), 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    VO_CUDA(cudaMemcpyAsync(h + oO, d + oO, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(pts1_prior, h + oO, (size_t)n * 8);
    return VO_OK;
}

extern "C" int vo_compact(vo_ctx *ctx, const uint8_t *mask, int n, int *index_out, int *n_out)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0 && n_out, VO_ERR_INVALID_ARG, "bad arguments");
    if (n == 0) { *n_out = 0; return VO_OK; }
    VO_REQUIRE(mask && index_out, VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    const size_t oM = 0, oC = ((size_t)n + 15) / 16 * 16, oI = oC + 16, total = oI + (size_t)n * 4;
    int rc = vo_stage_reserve(ctx, total);
    if (rc) return rc;
    uint8_t *h = ctx->h_stage, *d = ctx->d_stage;
    memcpy(h + oM, mask, (size_t)n);
    VO_CUDA(cudaMemcpyAsync(d, h, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    k_compact<<<1, 1024, 0, ctx->stream>>>(d + oM, n, (int *)(d + oI), (int *)(d + oC));
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    VO_CUDA(cudaMemcpyAsync(h + oC, d + oC, total - oC, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    const int k = *(int *)(h + oC);
    *n_out = k;
    memcpy(index_out, h + oI, (size_t)k * 4);
    return VO_OK;
}
