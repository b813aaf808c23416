// rectify.cu -- stereo rectification: map generation and bilinear remap.
//
//   k_rectify_maps  StereoCamera::generateStereoImagesUndistortAndRectifyMaps
//                   (core/visual_odometry/camera.cpp:364-546): reference rotation from the two optical axes and the
//                   baseline, rectified intrinsics, per-pixel back-projection + radial/tangential distortion of both
//                   cameras -> four CV_32FC1 maps.  One thread per pixel, FP32 in the reference's operation order
//                   (this file is compiled with -fmad=false); the maps stay resident in HBM.
//   k_remap         StereoCamera::rectifyStereoImages + the convertTo(CV_8UC1) of StereoVO
//                   (camera.cpp:300-336, stereo_vo.cpp:416-421): cv::remap(CV_32FC1, INTER_LINEAR, BORDER_CONSTANT 0)
//                   restated -- map coordinates rounded to 1/32 px with round-half-even, float tap weights
//                   (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy fx, left-to-right sum, saturate_cast<uchar> -- bit-exact
//                   with OpenCV 4.13.  One thread per output pixel: 8 B of map + 4 byte gathers from the distorted
//                   image (L2 resident) in, 1 byte out, written straight into the slot's raw plane so that the fused
//                   pyramid kernel ingests it like an uploaded image.
#include "vo_internal.cuh"

#include <cstring>

struct RectMapArgs {
    int w, h;
    float M[9];            // R_0n * K_rect_inv
    float R_l0[9], R_r0[9];
    float K_l[4], K_r[4], D_l[5], D_r[5];   // D = k1 k2 p1 p2 k3
    float *map_lu, *map_lv, *map_ru, *map_rv;
};

__device__ __forceinline__ void distort_project(const float *R, const float *P0, const float *K, const float *D, float &mu, float &mv)
{
    float X[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) X[r] = (R[r * 3 + 0] * P0[0] + R[r * 3 + 1] * P0[1]) + R[r * 3 + 2] * P0[2];
    const float x = X[0] / X[2], y = X[1] / X[2];
    const float k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3], k3 = D[4];
    const float xx = x * x, yy = y * y, xy2 = x * y * 2.0f;
    const float r2 = xx + yy, r4 = r2 * r2, r6 = r4 * r2;
    const float r_radial = ((1.0f + k1 * r2) + k2 * r4) + k3 * r6;
    const float x_dist = (x * r_radial + p1 * xy2) + p2 * (r2 + 2.0f * xx);
    const float y_dist = (y * r_radial + p2 * xy2) + p1 * (r2 + 2.0f * yy);
    mu = (x_dist * K[0] + K[2]) - 1.0f;
    mv = (y_dist * K[1] + K[3]) - 1.0f;
}

__global__ void __launch_bounds__(256) k_rectify_maps(const RectMapArgs a)
{
    const int u = blockIdx.x * 32 + (threadIdx.x & 31), v = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (u >= a.w || v >= a.h) return;
    const float pu = (float)(u + 1), pv = (float)(v + 1);
    float P0[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) P0[r] = (a.M[r * 3 + 0] * pu + a.M[r * 3 + 1] * pv) + a.M[r * 3 + 2] * 1.0f;
    float mu, mv;
    const size_t o = (size_t)v * a.w + u;
    distort_project(a.R_l0, P0, a.K_l, a.D_l, mu, mv);
    a.map_lu[o] = mu; a.map_lv[o] = mv;
    distort_project(a.R_r0, P0, a.K_r, a.D_r, mu, mv);
    a.map_ru[o] = mu; a.map_rv[o] = mv;
}

// Camera::generateImageUndistortMaps (core/visual_odometry/camera.cpp:57-87): single-camera undistortion map, no rotation and
// the same intrinsics.  The reference mixes float variables with double literals (2.0, 1.0): promotions restated.
__global__ void __launch_bounds__(256) k_undistort_maps(int w, int h, const float4 K, const float k1, const float k2, const float p1,
                                                        const float p2, const float k3, float *map_u, float *map_v)
{
    const int u = blockIdx.x * 32 + (threadIdx.x & 31), v = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (u >= w || v >= h) return;
    const float fxinv = 1.0f / K.x, fyinv = 1.0f / K.y;
    const float y = ((float)v - K.w) * fyinv, x = ((float)u - K.z) * fxinv;
    const float xy2 = (float)((2.0 * (double)x) * (double)y);
    const float xx = x * x, yy = y * y;
    const float r2 = xx + yy, r4 = r2 * r2, r6 = r4 * r2;
    const float r_radial = (float)(((1.0 + (double)(k1 * r2)) + (double)(k2 * r4)) + (double)(k3 * r6));
    const float x_dist = (float)((double)(x * r_radial + p1 * xy2) + (double)p2 * ((double)r2 + 2.0 * (double)xx));
    const float y_dist = (float)(((double)(y * r_radial) + (double)p1 * ((double)r2 + 2.0 * (double)yy)) + (double)(p2 * xy2));
    const size_t o = (size_t)v * w + u;
    map_u[o] = K.z + x_dist * K.x;
    map_v[o] = K.w + y_dist * K.y;
}

__global__ void __launch_bounds__(256)
k_remap(const uint8_t *__restrict__ src, int w, int h, const float *__restrict__ map_u, const float *__restrict__ map_v, uint8_t *__restrict__ dst)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const size_t o = (size_t)y * w + x;
    const float fu = map_u[o] * 32.0f, fv = map_v[o] * 32.0f;
    // cvRound: round half to even; non-finite or out-of-int-range -> INT_MIN (cvtss2si), i.e. far outside the image
    const bool bad = !(fabsf(fu) < 2147483648.0f) || !(fabsf(fv) < 2147483648.0f);
    const int sx = bad ? INT_MIN : __float2int_rn(fu), sy = bad ? INT_MIN : __float2int_rn(fv);
    const float fx = (float)(sx & 31) * (1.0f / 32.0f), fy = (float)(sy & 31) * (1.0f / 32.0f);
    int ix = sx >> 5, iy = sy >> 5;
    ix = ix < -32768 ? -32768 : (ix > 32767 ? 32767 : ix);          // saturate_cast<short>
    iy = iy < -32768 ? -32768 : (iy > 32767 ? 32767 : iy);
    const float w00 = (1.0f - fy) * (1.0f - fx), w01 = (1.0f - fy) * fx, w10 = fy * (1.0f - fx), w11 = fy * fx;
    const bool x0 = ix >= 0 && ix < w, x1 = ix + 1 >= 0 && ix + 1 < w, y0 = iy >= 0 && iy < h, y1 = iy + 1 >= 0 && iy + 1 < h;
    const float s00 = (x0 && y0) ? (float)src[(size_t)iy * w + ix] : 0.f;
    const float s01 = (x1 && y0) ? (float)src[(size_t)iy * w + ix + 1] : 0.f;
    const float s10 = (x0 && y1) ? (float)src[(size_t)(iy + 1) * w + ix] : 0.f;
    const float s11 = (x1 && y1) ? (float)src[(size_t)(iy + 1) * w + ix + 1] : 0.f;
    const float val = ((s00 * w00 + s01 * w01) + s10 * w10) + s11 * w11;
    int r = __float2int_rn(val);
    r = r < 0 ? 0 : (r > 255 ? 255 : r);
    dst[o] = (uint8_t)r;
}

static void mul3h(const float *A, const float *B, float *C)
{
    float T[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { float s = 0.f; for (int k = 0; k < 3; ++k) s += A[i * 3 + k] * B[k * 3 + j]; T[i * 3 + j] = s; }
    memcpy(C, T, sizeof(T));
}
static void unit3(float *v) { const float n = sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); for (int i = 0; i < 3; ++i) v[i] = v[i] / n; }
static void cross3(const float *a, const float *b, float *c)
{
    const float t[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    memcpy(c, t, sizeof(t));
}

extern "C" int vo_rectify_init(vo_ctx *ctx, const float *K_l4, const float *D_l5, const float *K_r4, const float *D_r5, const float *T_lr,
                               int w, int h, float *K_rect4_out, float *T_lr_rect_out)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(K_l4 && D_l5 && K_r4 && D_r5 && T_lr && K_rect4_out && T_lr_rect_out, VO_ERR_INVALID_ARG, "null pointer");
    VO_REQUIRE(w >= 8 && h >= 8 && w <= ctx->max_w && h <= ctx->max_h, VO_ERR_INVALID_ARG, "image larger than the context's max_w x max_h");
    VO_CUDA(cudaSetDevice(ctx->device));
    // reference rotation (camera.cpp:372-399); the small 3x3 algebra runs on the host exactly as in the reference
    float R_0r[9], t_0r[3];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R_0r[i * 3 + j] = T_lr[i * 4 + j]; t_0r[i] = T_lr[i * 4 + 3]; }
    const float k_l[3] = {0.f, 0.f, 1.f}, k_r[3] = {R_0r[2], R_0r[5], R_0r[8]};
    float k_n[3] = {(k_l[0] + k_r[0]) * 0.5f, (k_l[1] + k_r[1]) * 0.5f, (k_l[2] + k_r[2]) * 0.5f};
    unit3(k_n);
    float i_n[3] = {t_0r[0], t_0r[1], t_0r[2]};
    unit3(i_n);
    float j_n[3];
    cross3(k_n, i_n, j_n); unit3(j_n);
    cross3(i_n, j_n, k_n); unit3(k_n);
    const float R_0n[9] = {i_n[0], j_n[0], k_n[0], i_n[1], j_n[1], k_n[1], i_n[2], j_n[2], k_n[2]};
    const float f_n = (K_l4[0] + K_r4[0]) * 0.5f, centu = (float)w * 0.5f, centv = (float)h * 0.5f;
    const float K[9] = {f_n, 0.f, centu, 0.f, f_n, centv, 0.f, 0.f, 1.f};
    float cf[9];                                   // K_rect.inverse(): cofactors times 1/det (Eigen, third-party)
    cf[0] = K[4] * K[8] - K[5] * K[7]; cf[1] = K[2] * K[7] - K[1] * K[8]; cf[2] = K[1] * K[5] - K[2] * K[4];
    cf[3] = K[5] * K[6] - K[3] * K[8]; cf[4] = K[0] * K[8] - K[2] * K[6]; cf[5] = K[2] * K[3] - K[0] * K[5];
    cf[6] = K[3] * K[7] - K[4] * K[6]; cf[7] = K[1] * K[6] - K[0] * K[7]; cf[8] = K[0] * K[4] - K[1] * K[3];
    const float det = (K[0] * cf[0] + K[1] * cf[3]) + K[2] * cf[6], idet = 1.0f / det;
    float Kinv[9];
    for (int i = 0; i < 9; ++i) Kinv[i] = cf[i] * idet;
    RectMapArgs a;
    a.w = w; a.h = h;
    mul3h(R_0n, Kinv, a.M);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { a.R_l0[i * 3 + j] = (i == j) ? 1.f : 0.f; a.R_r0[i * 3 + j] = R_0r[j * 3 + i]; }
    memcpy(a.K_l, K_l4, 16); memcpy(a.K_r, K_r4, 16); memcpy(a.D_l, D_l5, 20); memcpy(a.D_r, D_r5, 20);
    const size_t plane = (size_t)w * h;
    const size_t need = plane * 4 * 4 + plane + 256;
    if (need > ctx->rect_bytes) {
        VO_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_rect) cudaFree(ctx->d_rect);
        ctx->d_rect = nullptr; ctx->rect_bytes = 0;
        VO_CUDA(cudaMalloc(&ctx->d_rect, need));
        ctx->rect_bytes = need;
    }
    float *mp = (float *)ctx->d_rect;
    a.map_lu = mp; a.map_lv = mp + plane; a.map_ru = mp + 2 * plane; a.map_rv = mp + 3 * plane;
    ctx->rect_w = w; ctx->rect_h = h;
    k_rectify_maps<<<dim3(vo_div_up(w, 32), vo_div_up(h, 8)), 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    // rectified intrinsics and extrinsics (camera.cpp:401-412, 531-535)
    K_rect4_out[0] = f_n; K_rect4_out[1] = f_n; K_rect4_out[2] = centu; K_rect4_out[3] = centv;
    float R_ln[9], R_lnT[9];
    mul3h(a.R_l0, R_0n, R_ln);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R_lnT[i * 3 + j] = R_ln[j * 3 + i];
    for (int i = 0; i < 16; ++i) T_lr_rect_out[i] = (i % 5 == 0) ? 1.f : 0.f;
    for (int r = 0; r < 3; ++r) T_lr_rect_out[r * 4 + 3] = (R_lnT[r * 3 + 0] * t_0r[0] + R_lnT[r * 3 + 1] * t_0r[1]) + R_lnT[r * 3 + 2] * t_0r[2];
    return VO_OK;
}

extern "C" int vo_read_rectify_maps(vo_ctx *ctx, int right, float *map_u, float *map_v)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(ctx->d_rect && ctx->rect_w > 0, VO_ERR_INVALID_ARG, "vo_rectify_init has not been called");
    VO_REQUIRE(map_u && map_v, VO_ERR_INVALID_ARG, "null pointer");
    const size_t plane = (size_t)ctx->rect_w * ctx->rect_h;
    const float *mp = (const float *)ctx->d_rect + (right ? 2 : 0) * plane;
    VO_CUDA(cudaMemcpyAsync(map_u, mp, plane * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaMemcpyAsync(map_v, mp + plane, plane * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    return VO_OK;
}

extern "C" int vo_upload_image_rectified(vo_ctx *ctx, int slot, int right, const uint8_t *data, int w, int h, size_t step)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(ctx->d_rect && ctx->rect_w > 0, VO_ERR_INVALID_ARG, "vo_rectify_init has not been called");
    VO_REQUIRE(w == ctx->rect_w && h == ctx->rect_h, VO_ERR_SIZE_MISMATCH,
               "In 'rectifyStereoImages()': provided image has not the same size as the camera model!");      // camera.cpp:308
    VO_REQUIRE(data && step >= (size_t)w && slot >= 0 && slot < ctx->n_slots, VO_ERR_INVALID_ARG, "bad image arguments");
    VO_CUDA(cudaSetDevice(ctx->device));
    int rc = vo_slot_prepare(ctx, slot, w, h);
    if (rc) return rc;
    const size_t plane = (size_t)w * h;
    uint8_t *scratch = (uint8_t *)ctx->d_rect + plane * 16;
    if (step == (size_t)w) VO_CUDA(cudaMemcpyAsync(scratch, data, plane, cudaMemcpyHostToDevice, ctx->stream));
    else VO_CUDA(cudaMemcpy2DAsync(scratch, w, data, step, w, h, cudaMemcpyHostToDevice, ctx->stream));
    const float *mp = (const float *)ctx->d_rect + (right ? 2 : 0) * plane;
    k_remap<<<dim3(vo_div_up(w, 32), vo_div_up(h, 8)), 256, 0, ctx->stream>>>(scratch, w, h, mp, mp + plane,
                                                                             ctx->raw_base + (size_t)slot * ctx->raw_stride);
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    return VO_OK;
}

// MonoVO with flagDoUndistortion (mono_vo.cpp:509-513): Camera::undistortImage (camera.cpp:163-183) = cv::remap with the
// maps of generateImageUndistortMaps (:57-87) + convertTo(CV_8UC1).  Fills map set 0; vo_upload_image_rectified(slot, 0, ...)
// then undistorts on upload.  D5 = k1 k2 p1 p2 k3 (cvD order, camera.cpp:30-34).
extern "C" int vo_undistort_init(vo_ctx *ctx, const float *K4, const float *D5, int w, int h)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(K4 && D5, VO_ERR_INVALID_ARG, "null pointer");
    VO_REQUIRE(w >= 8 && h >= 8 && w <= ctx->max_w && h <= ctx->max_h, VO_ERR_INVALID_ARG, "image larger than the context's max_w x max_h");
    VO_CUDA(cudaSetDevice(ctx->device));
    const size_t plane = (size_t)w * h;
    const size_t need = plane * 4 * 4 + plane + 256;
    if (need > ctx->rect_bytes) {
        VO_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_rect) cudaFree(ctx->d_rect);
        ctx->d_rect = nullptr; ctx->rect_bytes = 0;
        VO_CUDA(cudaMalloc(&ctx->d_rect, need));
        ctx->rect_bytes = need;
    }
    float *mp = (float *)ctx->d_rect;
    ctx->rect_w = w; ctx->rect_h = h;
    k_undistort_maps<<<dim3(vo_div_up(w, 32), vo_div_up(h, 8)), 256, 0, ctx->stream>>>(w, h, make_float4(K4[0], K4[1], K4[2], K4[3]), D5[0], D5[1],
                                                                                    D5[2], D5[3], D5[4], mp, mp + plane);
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    return VO_OK;
}
