// tri_device.cuh -- device code of mapping::triangulateDLT (core/util/triangulate_3d.cpp:91-130), shared by
// K-tri (elementwise.cu) and the fused stereo frame step (stereo_step.cu). Include only from translation units
// compiled with -fmad=false (FP32 in the reference's operation order).
#pragma once
#include <cfloat>

// ------------------------------------------------------------------------------ K-tri
// 4x4 FP32 two-sided Jacobi SVD (the algorithm of Eigen::JacobiSVD for a square real matrix):
// returns the right-singular vector of the smallest singular value.
__device__ __forceinline__ void make_jacobi(float x, float y, float z, float &c, float &s)
{
    const float deno = 2.f * fabsf(y);
    if (deno < FLT_MIN) { c = 1.f; s = 0.f; return; }
    const float tau = (x - z) / deno;
    const float w = sqrtf(tau * tau + 1.f);
    const float t = (tau > 0.f) ? 1.f / (tau + w) : 1.f / (tau - w);
    const float sign_t = t > 0.f ? 1.f : -1.f;
    const float n = 1.f / sqrtf(t * t + 1.f);
    s = -sign_t * (y / fabsf(y)) * fabsf(t) * n;
    c = n;
}

static __device__ void svd4_null(const float *M, float *v4)
{
    float W[16], V[16];
    float scale = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) scale = fmaxf(scale, fabsf(M[i]));
    if (scale == 0.f) scale = 1.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { W[i] = M[i] / scale; V[i] = (i % 5 == 0) ? 1.f : 0.f; }
    float maxDiag = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) maxDiag = fmaxf(maxDiag, fabsf(W[i * 5]));
    const float precision = 2.f * FLT_EPSILON;
    bool finished = false;
    int sweeps = 0;
    while (!finished && sweeps < 64) {
        finished = true;
        ++sweeps;
#pragma unroll
        for (int p = 1; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < p; ++q) {
                const float thr = fmaxf(FLT_MIN, precision * maxDiag);
                if (fabsf(W[p * 4 + q]) > thr || fabsf(W[q * 4 + p]) > thr) {
                    finished = false;
                    const float m00 = W[p * 4 + p], m01 = W[p * 4 + q], m10 = W[q * 4 + p], m11 = W[q * 4 + q];
                    const float t = m00 + m11, d = m10 - m01;
                    float c1, s1;
                    if (fabsf(d) < FLT_MIN) { s1 = 0.f; c1 = 1.f; }
                    else { const float u = t / d; const float tmp = sqrtf(1.f + u * u); s1 = 1.f / tmp; c1 = u / tmp; }
                    const float a00 = c1 * m00 + s1 * m10, a01 = c1 * m01 + s1 * m11;
                    const float a11 = -s1 * m01 + c1 * m11;
                    float cr, sr;
                    make_jacobi(a00, a01, a11, cr, sr);
                    const float cl = c1 * cr - s1 * (-sr);
                    const float sl = c1 * (-sr) + s1 * cr;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {   // rows p,q of W
                        const float xi = W[p * 4 + i], yi = W[q * 4 + i];
                        W[p * 4 + i] = cl * xi + sl * yi;
                        W[q * 4 + i] = -sl * xi + cl * yi;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {   // columns p,q of W and V
                        const float xi = W[i * 4 + p], yi = W[i * 4 + q];
                        W[i * 4 + p] = cr * xi - sr * yi;
                        W[i * 4 + q] = sr * xi + cr * yi;
                        const float xv = V[i * 4 + p], yv = V[i * 4 + q];
                        V[i * 4 + p] = cr * xv - sr * yv;
                        V[i * 4 + q] = sr * xv + cr * yv;
                    }
                    maxDiag = fmaxf(maxDiag, fmaxf(fabsf(W[p * 4 + p]), fabsf(W[q * 4 + q])));
                }
            }
    }
    int k = 0;
    float best = fabsf(W[0]);
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (fabsf(W[i * 5]) < best) { best = fabsf(W[i * 5]); k = i; }
#pragma unroll
    for (int i = 0; i < 4; ++i) v4[i] = V[i * 4 + k];
}


// Two-camera scalar DLT (triangulate_3d.cpp:91-130): X0 in camera 0, X1 = R10 * X0 + t10.
__device__ __forceinline__ void tri_point(float2 q0, float2 q1, const float *R10, const float *t10, const float *K0, const float *K1,
                                          float *X0, float *X1)
{
    const float fx0 = K0[0], fy0 = K0[1], cx0 = K0[2], cy0 = K0[3];
    const float fx1 = K1[0], fy1 = K1[1], cx1 = K1[2], cy1 = K1[3];
    float P10[12];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        P10[0 * 4 + j] = (fx1 * R10[0 * 3 + j] + 0.f * R10[1 * 3 + j]) + cx1 * R10[2 * 3 + j];
        P10[1 * 4 + j] = (0.f * R10[0 * 3 + j] + fy1 * R10[1 * 3 + j]) + cy1 * R10[2 * 3 + j];
        P10[2 * 4 + j] = (0.f * R10[0 * 3 + j] + 0.f * R10[1 * 3 + j]) + 1.f * R10[2 * 3 + j];
    }
    P10[0 * 4 + 3] = (fx1 * t10[0] + 0.f * t10[1]) + cx1 * t10[2];
    P10[1 * 4 + 3] = (0.f * t10[0] + fy1 * t10[1]) + cy1 * t10[2];
    P10[2 * 4 + 3] = (0.f * t10[0] + 0.f * t10[1]) + 1.f * t10[2];
    float M[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) M[j] = 0.f;
    M[0] = -fx0; M[2] = q0.x - cx0;
    M[5] = -fy0; M[6] = q0.y - cy0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        M[8 + j] = q1.x * P10[8 + j] - P10[0 + j];
        M[12 + j] = q1.y * P10[8 + j] - P10[4 + j];
    }
    float v[4];
    svd4_null(M, v);
    X0[0] = v[0] / v[3]; X0[1] = v[1] / v[3]; X0[2] = v[2] / v[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) X1[r] = ((R10[r * 3 + 0] * X0[0] + R10[r * 3 + 1] * X0[1]) + R10[r * 3 + 2] * X0[2]) + t10[r];
}
