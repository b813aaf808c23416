// epipolar.cu -- the mono geometric front-end around the pose solvers:
//   MotionEstimator::calcSampsonDistance            core/visual_odometry/motion_estimator.cpp:539-570
//   MotionEstimator::calcSymmetricEpipolarDistance  :621-653
//   MotionEstimator::findInliers1PointHistogram     :471-537 (+ core/util/histogram.h:11-35, histogram.cpp:4-28)
// FP32 in the reference's operation order (-fmad=false).  F10 = Kinv^T (skew(t10) R10) Kinv with Kinv = K.inverse()
// restated as cofactors / det (Eigen::Matrix3f::inverse, third-party).
#include "vo_internal.cuh"

#include <cstring>

namespace {

__host__ __device__ inline void ep_mul3(const float *A, const float *B, float *C)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float s = 0.f;
            for (int k = 0; k < 3; ++k) s += A[i * 3 + k] * B[k * 3 + j];
            C[i * 3 + j] = s;
        }
}

__host__ __device__ inline void ep_fundamental(const float *K4, const float *R10, const float *t10, float *F)
{
    const float K[9] = {K4[0], 0.f, K4[2], 0.f, K4[1], K4[3], 0.f, 0.f, 1.f};
    float cf[9];
    cf[0] = K[4] * K[8] - K[5] * K[7]; cf[1] = K[2] * K[7] - K[1] * K[8]; cf[2] = K[1] * K[5] - K[2] * K[4];
    cf[3] = K[5] * K[6] - K[3] * K[8]; cf[4] = K[0] * K[8] - K[2] * K[6]; cf[5] = K[2] * K[3] - K[0] * K[5];
    cf[6] = K[3] * K[7] - K[4] * K[6]; cf[7] = K[1] * K[6] - K[0] * K[7]; cf[8] = K[0] * K[4] - K[1] * K[3];
    const float det = (K[0] * cf[0] + K[1] * cf[3]) + K[2] * cf[6];
    const float idet = 1.0f / det;
    float Kinv[9], KinvT[9];
    for (int i = 0; i < 9; ++i) Kinv[i] = cf[i] * idet;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) KinvT[i * 3 + j] = Kinv[j * 3 + i];
    const float S[9] = {0.f, -t10[2], t10[1], t10[2], 0.f, -t10[0], -t10[1], t10[0], 0.f};
    float E[9], KE[9];
    ep_mul3(S, R10, E);
    ep_mul3(KinvT, E, KE);
    ep_mul3(KE, Kinv, F);
}

struct EpF { float F[9]; };

// mode 0: Sampson distance (:555-569); mode 1: symmetric epipolar distance (:638-652)
__device__ __forceinline__ float ep_distance(const float *F, float2 p0, float2 p1, int mode)
{
    const float a0 = (F[0] * p0.x + F[1] * p0.y) + F[2] * 1.0f;
    const float a1 = (F[3] * p0.x + F[4] * p0.y) + F[5] * 1.0f;
    const float a2 = (F[6] * p0.x + F[7] * p0.y) + F[8] * 1.0f;
    const float b0 = (F[0] * p1.x + F[3] * p1.y) + F[6] * 1.0f;
    const float b1 = (F[1] * p1.x + F[4] * p1.y) + F[7] * 1.0f;
    const float num = (p1.x * a0 + p1.y * a1) + 1.0f * a2;
    if (mode == 0) {
        const float den = ((a0 * a0 + a1 * a1) + b0 * b0) + b1 * b1;
        return (num * num) / den;
    }
    const float den = 1.0f / sqrtf(a0 * a0 + a1 * a1) + 1.0f / sqrtf(b0 * b0 + b1 * b1);
    return fabsf(num) * den;
}

__global__ void __launch_bounds__(256) k_ep_distance(const float2 *p0, const float2 *p1, int n, const EpF f, int mode, float *dist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dist[i] = ep_distance(f.F, p0[i], p1[i], mode);
}

#define EP_BINS 400
// findInliers1PointHistogram: planar-motion angle of every correspondence, 400-bin histogram on [-0.5, 0.5] rad, the
// fullest bin's centre (the first one on ties; std::sort leaves ties unspecified), R10 / t10 of that one-parameter
// motion, symmetric epipolar distance against thres_1p^2
__global__ void __launch_bounds__(1024) k_ep_1point(const float2 *p0, const float2 *p1, int n, const float4 K4, float thres2, uint8_t *mask,
                                                    float *theta_out, float *out /* th_opt, R10[9], t10[3] */)
{
    __shared__ int s_hist[EP_BINS];
    __shared__ float s_F[9];
    const int tid = threadIdx.x;
    for (int b = tid; b < EP_BINS; b += 1024) s_hist[b] = 0;
    __syncthreads();
    const float invfx = 1.0f / K4.x, invfy = 1.0f / K4.y, cx = K4.z, cy = K4.w;
    const float hist_min = -0.5f, hist_max = 0.5f;
    const float step = (hist_max - hist_min) / (float)EP_BINS;
    for (int i = tid; i < n; i += 1024) {
        const float x0 = (p0[i].x - cx) * invfx, y0 = (p0[i].y - cy) * invfy;
        const float x1 = (p1[i].x - cx) * invfx, y1 = (p1[i].y - cy) * invfy;
        const float val = (x0 * y1 - y0 * x1) / (y0 * 1.f + 1.f * y1);
        const float th = (float)(-2.0 * (double)atanf(val));
        if (theta_out) theta_out[i] = th;
        const float v = th - hist_min;
        const float q = floorf(v / step);
        if (q >= 0.f && q < (float)EP_BINS) atomicAdd(&s_hist[(int)q], 1);       // NaN fails both comparisons
    }
    __syncthreads();
    if (tid == 0) {
        int best = 0;
        for (int b = 1; b < EP_BINS; ++b)
            if (s_hist[b] > s_hist[best]) best = b;
        // histogram.h:19-24: centres accumulate `step` from hist_min; the last one is hist_max
        float c = hist_min;
        for (int b = 1; b <= best && b < EP_BINS - 1; ++b) c = c + step;
        if (best == EP_BINS - 1) c = hist_max;
        const float th_opt = c;
        const float costh = cosf(th_opt), sinth = sinf(th_opt);
        const float R10[9] = {costh, 0.f, sinth, 0.f, 1.f, 0.f, -sinth, 0.f, costh};
        const float t10[3] = {sinf(th_opt * 0.5f), 0.0f, cosf(th_opt * 0.5f)};
        const float Kv[4] = {K4.x, K4.y, K4.z, K4.w};
        float F[9];
        ep_fundamental(Kv, R10, t10, F);
        for (int i = 0; i < 9; ++i) s_F[i] = F[i];
        out[0] = th_opt;
        for (int i = 0; i < 9; ++i) out[1 + i] = R10[i];
        for (int i = 0; i < 3; ++i) out[10 + i] = t10[i];
    }
    __syncthreads();
    for (int i = tid; i < n; i += 1024) mask[i] = ep_distance(s_F, p0[i], p1[i], 1) <= thres2 ? 1 : 0;
}

int ep_distance_host(vo_ctx *ctx, const float *pts0, const float *pts1, int n, const float *K4, const float *R10, const float *t10,
                     const float *F10, int mode, float *dist)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0, VO_ERR_INVALID_ARG, "negative size");
    if (n == 0) return VO_OK;
    VO_REQUIRE(pts0 && pts1 && dist && (F10 || (K4 && R10 && t10)), VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    EpF f;
    if (F10) memcpy(f.F, F10, 36);
    else ep_fundamental(K4, R10, t10, f.F);
    const size_t N = (size_t)n;
    const int rc = vo_stage_reserve(ctx, N * 20);
    if (rc) return rc;
    uint8_t *hs = ctx->h_stage, *dv = ctx->d_stage;
    memcpy(hs, pts0, N * 8); memcpy(hs + N * 8, pts1, N * 8);
    VO_CUDA(cudaMemcpyAsync(dv, hs, N * 16, cudaMemcpyHostToDevice, ctx->stream));
    k_ep_distance<<<vo_div_up(n, 256), 256, 0, ctx->stream>>>((const float2 *)dv, (const float2 *)(dv + N * 8), n, f, mode, (float *)(dv + N * 16));
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    VO_CUDA(cudaMemcpyAsync(hs + N * 16, dv + N * 16, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(dist, hs + N * 16, N * 4);
    return VO_OK;
}

}  // namespace

extern "C" int vo_sampson_distance(vo_ctx *ctx, const float *pts0, const float *pts1, int n, const float *K4, const float *R10,
                                   const float *t10, float *dist)
{
    return ep_distance_host(ctx, pts0, pts1, n, K4, R10, t10, nullptr, 0, dist);
}

extern "C" int vo_sampson_distance_F(vo_ctx *ctx, const float *pts0, const float *pts1, int n, const float *F10, float *dist)
{
    if (ctx && !F10) { ctx->last_error = "null pointer"; return VO_ERR_INVALID_ARG; }
    return ep_distance_host(ctx, pts0, pts1, n, nullptr, nullptr, nullptr, F10, 0, dist);
}

extern "C" int vo_symmetric_epipolar_distance(vo_ctx *ctx, const float *pts0, const float *pts1, int n, const float *K4, const float *R10,
                                              const float *t10, float *dist)
{
    return ep_distance_host(ctx, pts0, pts1, n, K4, R10, t10, nullptr, 1, dist);
}

extern "C" int vo_inliers_1point_histogram(vo_ctx *ctx, const float *pts0, const float *pts1, int n, const float *K4, float thres_1p,
                                           uint8_t *mask, float *theta_opt, float *R10, float *t10, float *theta)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0, VO_ERR_INVALID_ARG, "negative size");
    VO_REQUIRE(K4 && theta_opt && (n == 0 || (pts0 && pts1 && mask)), VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    const size_t N = (size_t)n;
    const size_t o_out = (N * 16 + 15) / 16 * 16, o_m = o_out + 64, o_th = o_m + (N + 15) / 16 * 16, total = o_th + N * 4;
    const int rc = vo_stage_reserve(ctx, total);
    if (rc) return rc;
    uint8_t *hs = ctx->h_stage, *dv = ctx->d_stage;
    if (n > 0) {
        memcpy(hs, pts0, N * 8); memcpy(hs + N * 8, pts1, N * 8);
        VO_CUDA(cudaMemcpyAsync(dv, hs, N * 16, cudaMemcpyHostToDevice, ctx->stream));
    }
    k_ep_1point<<<1, 1024, 0, ctx->stream>>>((const float2 *)dv, (const float2 *)(dv + N * 8), n, make_float4(K4[0], K4[1], K4[2], K4[3]),
                                             thres_1p * thres_1p, dv + o_m, theta ? (float *)(dv + o_th) : nullptr, (float *)(dv + o_out));
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    VO_CUDA(cudaMemcpyAsync(hs + o_out, dv + o_out, total - o_out, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    const float *o = (const float *)(hs + o_out);
    *theta_opt = o[0];
    if (R10) memcpy(R10, o + 1, 36);
    if (t10) memcpy(t10, o + 10, 12);
    if (n > 0) memcpy(mask, hs + o_m, N);
    if (theta && n > 0) memcpy(theta, hs + o_th, N * 4);
    return VO_OK;
}
