// pose_gn.cu -- K-pose: Huber 3D-2D pose-only Gauss-Newton, whole iteration loop on device.
//
// Replaces MotionEstimator::poseOnlyBundleAdjustment (core/visual_odometry/motion_estimator.cpp:665-861,
// standalone/motion_estimator/motion_estimator.cpp:4-193) and ..._Stereo (core :863-1088,
// standalone :195-411).  One CTA per problem; every thread evaluates the residual / Jacobian rows
// of its points in FP32 with the reference's operation order (this file is compiled with
// -fmad=false so no multiply-add is fused), accumulates the 21 upper-triangular JtWJ entries,
// the 6 entries of -JtWr and the error in FP64 registers, reduces warp-level by shuffles then
// block-level through shared memory, and thread 0 rounds the sums to FP32 and performs the
// reference's damped 6x6 pivoted LDLT solve, se3Exp update and stopping test.  No host
// round trip inside the <=100-iteration loop.  Reference quirks kept: Huber gate '>=' on the
// (mean-)L1 residual, masks rewritten each iteration with outliers still weighted in
// (motion_estimator.cpp:955-967), the right-camera Jacobian reusing the left closed form
// (:1008-1031), int outlier threshold in mono (motion_estimator.h:117), err without sqrt in mono.
#include "vo_internal.cuh"

#include <cooperative_groups.h>

#include <cstdlib>

#include <cstring>
#include <mutex>

#define POSE_MAX_ITER 100
#define NACC 28   // 21 (upper JtWJ) + 6 (mJtWr) + 1 (err)

struct PoseArgs {
    const int *offsets;   // [n_prob+1] device, or null (single problem of n points)
    int n_single;
    const int *n_single_d;  // nullable: point count read from device memory (device-resident pipelines)
    const float *X, *pl, *pr;
    float Kl[4], Kr[4];
    float T_rl[16];       // row-major
    float thres;
    int mono;             // 1: mono (2 rows / point)
    int variant;          // mono: 0 core, 1 standalone error accounting
    float *T01;           // [n_prob][16] in-out (mono: packed as 4x4 too)
    uint8_t *mask;
    int *success;
    int *iters;
    int max_iter;         // <= POSE_MAX_ITER
    int no_early_stop;    // 1: ignore the stop test (fixed-point / trace comparisons)
    float *trace;         // nullable: [n_prob][max_iter][24] = {T10 after update (16), err, dxi (6), delta_err}
};

__device__ __forceinline__ void inv_se3(const float *T, float *O)
{
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) O[i * 4 + j] = T[j * 4 + i];
        float s = 0.f;
        for (int k = 0; k < 3; ++k) s += T[k * 4 + i] * T[k * 4 + 3];
        O[i * 4 + 3] = -s;
    }
    O[12] = 0.f; O[13] = 0.f; O[14] = 0.f; O[15] = 1.f;
}

__device__ void inverse4(const float *m, float *out)
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

// geometry::se3Exp_f (core/util/geometry_library.cpp:386-440): sin/cos evaluated in double on
// the float angle, coefficients rounded to float, small-angle branch below 1e-7.
__device__ void se3exp_f(const float *xi, float *T)
{
    const float v0 = xi[0], v1 = xi[1], v2 = xi[2], w0 = xi[3], w1 = xi[4], w2 = xi[5];
    const float theta = sqrtf(w0 * w0 + w1 * w1 + w2 * w2);
    const float wx[9] = {0.f, -w2, w1, w2, 0.f, -w0, -w1, w0, 0.f};
    float wx2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float s = 0.f;
            for (int k = 0; k < 3; ++k) s += wx[i * 3 + k] * wx[k * 3 + j];
            wx2[i * 3 + j] = s;
        }
    float a, b, c;
    if (theta < 1e-7) {
        a = 1.f; b = 0.5f; c = 0.33333333333333333333333333f;
    } else {
        const double th = (double)theta;
        double sn, cs;
        sincos(th, &sn, &cs);
        a = (float)(sn / th);
        b = (float)((1 - cs) / (double)(theta * theta));
        c = (float)((th - sn) / (double)(theta * theta * theta));
    }
    float V[9];
    for (int i = 0; i < 9; ++i) {
        const float id = (i == 0 || i == 4 || i == 8) ? 1.f : 0.f;
        const int r = i / 3, cc = i - 3 * r;
        T[r * 4 + cc] = (id + a * wx[i]) + b * wx2[i];
        V[i] = (id + b * wx[i]) + c * wx2[i];
    }
    for (int i = 0; i < 3; ++i) T[i * 4 + 3] = (V[i * 3 + 0] * v0 + V[i * 3 + 1] * v1) + V[i * 3 + 2] * v2;
    T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
}

// Eigen::LDLT<Matrix<float,6,6>,Lower> (diagonal pivoting, unblocked) + solve, with the matrix held
// in REGISTERS: every loop has compile-time bounds and the data-dependent pivot swap is expressed as
// a chain of compile-time (k, b) swap routines, so no element is ever addressed dynamically (the
// first version kept the matrix in local memory and its serial latency chain cost ~15 us per GN
// iteration). The arithmetic and its order are exactly those of the oracle's pivoted LDLT.
struct Sym6 {
    float m[36];   // full 6x6 row-major; only the lower triangle is referenced after the swaps
};

template <int K, int B>
__device__ __forceinline__ void swap_rc(Sym6 &S)
{
    // Eigen ldlt_inplace<Lower>: symmetric swap of rows/cols K and B (B > K) in lower storage
#define M(i, j) S.m[(i) * 6 + (j)]
#pragma unroll
    for (int j = 0; j < K; ++j) { const float t = M(K, j); M(K, j) = M(B, j); M(B, j) = t; }
#pragma unroll
    for (int i = B + 1; i < 6; ++i) { const float t = M(i, K); M(i, K) = M(i, B); M(i, B) = t; }
    { const float t = M(K, K); M(K, K) = M(B, B); M(B, B) = t; }
#pragma unroll
    for (int i = K + 1; i < B; ++i) { const float t = M(i, K); M(i, K) = M(B, i); M(B, i) = t; }
#undef M
}

template <int K>
__device__ __forceinline__ int ldlt_step(Sym6 &S)
{
#define M(i, j) S.m[(i) * 6 + (j)]
    // pivot: first index of the largest |diagonal| in K..5
    int big = K;
    float bv = fabsf(M(K, K));
#pragma unroll
    for (int i = K + 1; i < 6; ++i) { const float v = fabsf(M(i, i)); if (v > bv) { bv = v; big = i; } }
    if (K + 1 < 6 && big == K + 1) swap_rc<K, (K + 1 < 6 ? K + 1 : 5)>(S);
    if (K + 2 < 6 && big == K + 2) swap_rc<K, (K + 2 < 6 ? K + 2 : 5)>(S);
    if (K + 3 < 6 && big == K + 3) swap_rc<K, (K + 3 < 6 ? K + 3 : 5)>(S);
    if (K + 4 < 6 && big == K + 4) swap_rc<K, (K + 4 < 6 ? K + 4 : 5)>(S);
    if (K + 5 < 6 && big == K + 5) swap_rc<K, (K + 5 < 6 ? K + 5 : 5)>(S);
    if (K > 0) {
        float temp[K > 0 ? K : 1];
#pragma unroll
        for (int j = 0; j < K; ++j) temp[j] = M(j, j) * M(K, j);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < K; ++j) s += M(K, j) * temp[j];
        M(K, K) -= s;
#pragma unroll
        for (int i = K + 1; i < 6; ++i) {
            float s2 = 0.f;
#pragma unroll
            for (int j = 0; j < K; ++j) s2 += M(i, j) * temp[j];
            M(i, K) -= s2;
        }
    }
    const float akk = M(K, K);
    if (K < 5 && fabsf(akk) > 0.f) {
#pragma unroll
        for (int i = K + 1; i < 6; ++i) M(i, K) /= akk;
    }
    return big;
#undef M
}

__device__ __forceinline__ void swap_dyn(float *y, int k, int b)
{
    // y[k] <-> y[b] without dynamic register indexing
    float yk = 0.f, yb = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) { if (i == k) yk = y[i]; if (i == b) yb = y[i]; }
#pragma unroll
    for (int i = 0; i < 6; ++i) { if (i == k) y[i] = yb; else if (i == b) y[i] = yk; }
}

__device__ void ldlt6_solve(float *m_in /*6x6 row-major sym*/, const float *b, float *x)
{
    Sym6 S;
#pragma unroll
    for (int i = 0; i < 36; ++i) S.m[i] = m_in[i];
    int tr[6];
    tr[0] = ldlt_step<0>(S); tr[1] = ldlt_step<1>(S); tr[2] = ldlt_step<2>(S);
    tr[3] = ldlt_step<3>(S); tr[4] = ldlt_step<4>(S); tr[5] = ldlt_step<5>(S);
#define M(i, j) S.m[(i) * 6 + (j)]
    float y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = b[i];
#pragma unroll
    for (int k = 0; k < 6; ++k) if (tr[k] != k) swap_dyn(y, k, tr[k]);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        float s = y[i];
#pragma unroll
        for (int j = 0; j < i; ++j) s -= M(i, j) * y[j];
        y[i] = s;
    }
    const float tol = 1.0f / 3.402823466e+38f;
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = fabsf(M(i, i)) > tol ? y[i] / M(i, i) : 0.f;
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        float s = y[i];
#pragma unroll
        for (int j = i + 1; j < 6; ++j) s -= M(j, i) * y[j];
        y[i] = s;
    }
#pragma unroll
    for (int k = 5; k >= 0; --k) if (tr[k] != k) swap_dyn(y, k, tr[k]);
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = y[i];
#undef M
}

// One residual row: weighted rank-1 update with the row's structural zero (calcJtWJ_x / _y).
// ZERO is the index of the zero Jacobian entry (1 for an x row, 0 for a y row).
template <int ZERO>
__device__ __forceinline__ void acc_row(double *acc, const float *Jt, float w, float r, bool weighted)
{
    // FP64 fused multiply-adds on the converted FP32 factors: 12 conversions + 26 DFMA per row.  (The first version rounded every
    // product to FP32 as the reference does and added it in FP64 -- FMUL + F2F + DADD per entry, 40 % of the batched kernel's
    // instructions, profiles/r2_pose_batch_*; this mode does not reproduce the reference's sums anyway, VO_POSE_STRICT does.)
    double dJ[6], dW[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { dJ[i] = (double)Jt[i]; dW[i] = (double)(weighted ? w * Jt[i] : Jt[i]); }
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i; j < 6; ++j, ++idx)
            if (i != ZERO && j != ZERO) acc[idx] = __fma_rn(dW[i], dJ[j], acc[idx]);
    const double nwr = -(double)(weighted ? w * r : r);
#pragma unroll
    for (int i = 0; i < 6; ++i)
        if (i != ZERO) acc[21 + i] = __fma_rn(nwr, dJ[i], acc[21 + i]);
}

// Residuals / Jacobian rows of ONE point handed to a sink, one call per row in the reference's row order (per-point
// arithmetic in the reference's FP32 operation order; motion_estimator.cpp:720-800 mono, :915-1050 stereo).
// sink.row<ZERO>(Jt, weight, r, weighted, e): ZERO = index of the structurally zero Jacobian entry of the row
// (calcJtWJ_x / _y), e = the row's error term.
template <class Sink>
__device__ __forceinline__ void pose_point(const PoseArgs &a, const float *__restrict__ X, const float *__restrict__ pl,
                                           const float *__restrict__ pr, uint8_t *__restrict__ mask, int i, const float *T10,
                                           Sink &sink)
{
    const float fx_l = a.Kl[0], fy_l = a.Kl[1], cx_l = a.Kl[2], cy_l = a.Kl[3];
    const float fx_r = a.Kr[0], fy_r = a.Kr[1], cx_r = a.Kr[2], cy_r = a.Kr[3];
    const float THRES_HUBER = 0.5f;
    const float x0 = X[3 * i], x1 = X[3 * i + 1], x2 = X[3 * i + 2];
    float Xl[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) Xl[r] = ((T10[r * 4 + 0] * x0 + T10[r * 4 + 1] * x1) + T10[r * 4 + 2] * x2) + T10[r * 4 + 3];
    const float iz_l = 1.0f / Xl[2], xiz_l = Xl[0] * iz_l, yiz_l = Xl[1] * iz_l;
    const float fxxiz_l = fx_l * xiz_l, fyyiz_l = fy_l * yiz_l;
    const float rx_l = (fxxiz_l + cx_l) - pl[2 * i], ry_l = (fyyiz_l + cy_l) - pl[2 * i + 1];
    float Jt[6];
    if (a.mono) {
        float weight = 1.0f;
        bool fw = false;
        const float absrxry = fabsf(rx_l) + fabsf(ry_l);
        if (absrxry >= THRES_HUBER) { weight = THRES_HUBER / absrxry; fw = true; }
        mask[i] = (absrxry >= a.thres) ? 0 : 1;
        Jt[0] = fx_l * iz_l; Jt[1] = 0.f; Jt[2] = -fxxiz_l * iz_l; Jt[3] = -fxxiz_l * yiz_l;
        Jt[4] = fx_l * (1.0f + xiz_l * xiz_l); Jt[5] = -fx_l * yiz_l;
        sink.template row<1>(Jt, weight, rx_l, fw, rx_l * rx_l);
        Jt[0] = 0.f; Jt[1] = fy_l * iz_l; Jt[2] = -fyyiz_l * iz_l; Jt[3] = -fy_l * (1.0f + yiz_l * yiz_l);
        Jt[4] = fyyiz_l * xiz_l; Jt[5] = fy_l * xiz_l;
        sink.template row<0>(Jt, weight, ry_l, fw, (fw && a.variant == 0) ? (weight * ry_l) * ry_l : ry_l * ry_l);
    } else {
        float Xr[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            Xr[r] = ((a.T_rl[r * 4 + 0] * Xl[0] + a.T_rl[r * 4 + 1] * Xl[1]) + a.T_rl[r * 4 + 2] * Xl[2]) + a.T_rl[r * 4 + 3];
        const float iz_r = 1.0f / Xr[2], xiz_r = Xr[0] * iz_r, yiz_r = Xr[1] * iz_r;
        const float fxxiz_r = fx_r * xiz_r, fyyiz_r = fy_r * yiz_r;
        const float rx_r = (fxxiz_r + cx_r) - pr[2 * i], ry_r = (fyyiz_r + cy_r) - pr[2 * i + 1];
        float weight = 1.0f;
        float absrxry = fabsf(rx_l) + fabsf(ry_l) + fabsf(rx_r) + fabsf(ry_r);
        absrxry *= 0.5f;
        if (absrxry >= THRES_HUBER) weight = THRES_HUBER / absrxry;
        mask[i] = (absrxry >= a.thres) ? 0 : 1;
        Jt[0] = fx_l * iz_l; Jt[1] = 0.f; Jt[2] = -fxxiz_l * iz_l; Jt[3] = -fxxiz_l * yiz_l;
        Jt[4] = fx_l * (1.0f + xiz_l * xiz_l); Jt[5] = -fx_l * yiz_l;
        sink.template row<1>(Jt, weight, rx_l, true, rx_l * rx_l);
        Jt[0] = 0.f; Jt[1] = fy_l * iz_l; Jt[2] = -fyyiz_l * iz_l; Jt[3] = -fy_l * (1.0f + yiz_l * yiz_l);
        Jt[4] = fyyiz_l * xiz_l; Jt[5] = fy_l * xiz_l;
        sink.template row<0>(Jt, weight, ry_l, true, ry_l * ry_l);
        Jt[0] = fx_r * iz_r; Jt[1] = 0.f; Jt[2] = -fxxiz_r * iz_r; Jt[3] = -fxxiz_r * yiz_r;
        Jt[4] = fx_r * (1.0f + xiz_r * xiz_r); Jt[5] = -fx_r * yiz_r;
        sink.template row<1>(Jt, weight, rx_r, true, rx_r * rx_r);
        Jt[0] = 0.f; Jt[1] = fy_r * iz_r; Jt[2] = -fyyiz_r * iz_r; Jt[3] = -fy_r * (1.0f + yiz_r * yiz_r);
        Jt[4] = fyyiz_r * xiz_r; Jt[5] = fy_r * xiz_r;
        sink.template row<0>(Jt, weight, ry_r, true, ry_r * ry_r);
    }
}

// Default ("fast") sink: 28 FP64 partials per thread, tree-reduced afterwards.
struct AccSink {
    double *acc;
    template <int ZERO>
    __device__ __forceinline__ void row(const float *Jt, float w, float r, bool weighted, float e)
    {
        acc_row<ZERO>(acc, Jt, w, r, weighted);
        acc[27] += (double)e;
    }
};

__device__ __forceinline__ void pose_points(const PoseArgs &a, const float *__restrict__ X, const float *__restrict__ pl,
                                            const float *__restrict__ pr, uint8_t *__restrict__ mask, int n, int first, int stride,
                                            const float *T10, double *acc)
{
    AccSink sink{acc};
    for (int i = first; i < n; i += stride) pose_point(a, X, pl, pr, mask, i, T10, sink);
}

// The serial part of one GN iteration (one thread): normal equations from the reduced partials, damped 6x6 LDLT,
// se3Exp_f, left-multiplication of T10, stop test (motion_estimator.cpp:803-840 / :1052-1070).
__device__ __forceinline__ void pose_solve(const PoseArgs &a, const double *red, int n, float *s_T10, float *s_err_prev, int *s_stop,
                                           int prob, int iter)
{
            float H[36], g[6];
            int idx = 0;
            for (int i = 0; i < 6; ++i)
                for (int j = i; j < 6; ++j, ++idx) { const float v = (float)red[idx]; H[i * 6 + j] = v; H[j * 6 + i] = v; }
            for (int i = 0; i < 6; ++i) g[i] = (float)red[21 + i];
            float err_curr = (float)red[27];
            const float inv_npts = 1.0f / (float)n;
            err_curr *= (inv_npts * 0.5f);
            if (!a.mono) err_curr = sqrtf(err_curr);
            const float delta_err = fabsf(err_curr - (*s_err_prev));
            for (int i = 0; i < 6; ++i) H[i * 6 + i] *= (1.0f + 0.00001f);
            float dxi[6], dT[16], Tn[16];
            ldlt6_solve(H, g, dxi);
            se3exp_f(dxi, dT);
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) {
                    float s = 0.f;
                    for (int k = 0; k < 4; ++k) s += dT[i * 4 + k] * s_T10[k * 4 + j];
                    Tn[i * 4 + j] = s;
                }
            for (int i = 0; i < 16; ++i) s_T10[i] = Tn[i];
            (*s_err_prev) = err_curr;
            float nrm = 0.f;
            for (int i = 0; i < 6; ++i) nrm += dxi[i] * dxi[i];
            nrm = sqrtf(nrm);
            *s_stop = (!a.no_early_stop && (nrm < 1e-6f || delta_err < 1e-7f)) ? 1 : 0;
            if (a.trace) {
                float *tr = a.trace + ((size_t)prob * a.max_iter + iter) * 24;
                for (int i = 0; i < 16; ++i) tr[i] = Tn[i];
                tr[16] = err_curr;
                for (int i = 0; i < 6; ++i) tr[17 + i] = dxi[i];
                tr[23] = delta_err;
            }
}

// Transposing butterfly over a warp: in, every lane's NACC partials; out (return value), on lane L < NACC the warp total of
// accumulator L.  At every level a lane keeps one half of its values and hands the other half to its partner: 31 exchanges
// instead of NACC x 5.
__device__ __forceinline__ double warp_transpose_sum(const double *acc, const int lane)
{
    double v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = k < NACC ? acc[k] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < o; ++k) {
            const double send = up ? v[k] : v[k + o];
            const double keep = up ? v[k + o] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

// THREADS per problem: 32 for the batched small problems -- one warp per solve, up to 20 resident CTAs per SM at 96 registers,
// whose point loops overlap the serial 6 x 6 solve that one thread of another CTA runs (measured on 4096 x 500 points:
// 128 threads, 4 CTAs/SM at 126 registers 8.76 M solves/s; 5 CTAs at 96 registers 8.97 M; 6 CTAs at 80 registers spill: 8.65 M;
// with the fused FP64 accumulation 10.9 M at 128 threads, 11.5 M at 64, 11.7 M at 32); 512 for one large problem (2000+
// points, 4 per thread) when the cluster kernel below is not used.
template <int POSE_THREADS>
__global__ void __launch_bounds__(POSE_THREADS, POSE_THREADS == 32 ? 20 : 1)
k_pose_gn(const PoseArgs a)
{
    __shared__ double s_part[POSE_THREADS / 32][NACC];
    __shared__ float s_T10[16];
    __shared__ int s_stop;
    __shared__ float s_err_prev;

    const int prob = blockIdx.x;
    const int beg = a.offsets ? a.offsets[prob] : 0;
    const int n = a.offsets ? a.offsets[prob + 1] - beg : (a.n_single_d ? *a.n_single_d : a.n_single);
    const float *X = a.X + 3 * (size_t)beg;
    const float *pl = a.pl + 2 * (size_t)beg;
    const float *pr = a.mono ? nullptr : a.pr + 2 * (size_t)beg;
    uint8_t *mask = a.mask + beg;
    float *T01 = a.T01 + 16 * (size_t)prob;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    if (tid == 0) {
        float T[16];
        for (int i = 0; i < 16; ++i) T[i] = T01[i];
        float Ti[16];
        if (a.mono) inverse4(T, Ti);       // Matrix4f::inverse() (motion_estimator.cpp:700)
        else inv_se3(T, Ti);               // explicit R^T, -R^T t (:903-904)
        for (int i = 0; i < 16; ++i) s_T10[i] = Ti[i];
        s_stop = 0;
        s_err_prev = 1e10f;
    }
    __syncthreads();

    int iter = 0;
    for (; iter < a.max_iter; ++iter) {
        float T10[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) T10[i] = s_T10[i];
        double acc[NACC];
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = 0.0;

        pose_points(a, X, pl, pr, mask, n, tid, POSE_THREADS, T10, acc);
        // warp-level then block-level reduction of the 28 FP64 partials
        {
            const double tot = warp_transpose_sum(acc, lane);
            if (lane < NACC) s_part[wid][lane] = tot;
        }
        __syncthreads();
        if (tid < NACC) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < POSE_THREADS / 32; ++w) v += s_part[w][tid];
            s_part[0][tid] = v;
        }
        __syncthreads();
        if (tid == 0) pose_solve(a, &s_part[0][0], n, s_T10, &s_err_prev, &s_stop, prob, iter);
        __syncthreads();
        if (s_stop) { ++iter; break; }
    }
    if (tid == 0) {
        float nrm2 = 0.f;
        for (int i = 0; i < 16; ++i) nrm2 += s_T10[i] * s_T10[i];
        const bool ok = !isnan(nrm2);
        if (ok) {
            float T10[16], Ti[16];
            for (int i = 0; i < 16; ++i) T10[i] = s_T10[i];
            inv_se3(T10, Ti);
            for (int i = 0; i < 16; ++i) T01[i] = Ti[i];
        }
        if (a.success) a.success[prob] = ok ? 1 : 0;
        if (a.iters) a.iters[prob] = iter;
    }
}

// One LARGE problem (the frame step: 2000+ points) on a thread-block cluster: 8 CTAs x 256 threads share the point
// loop (one point per thread at 2048 points instead of 8 per thread on one SM); the 28 FP64 partials go warp -> CTA
// -> rank 0 through distributed shared memory, rank 0 solves and broadcasts the new pose and the stop flag back
// through DSMEM; two cluster barriers per iteration.  Summation order is fixed (lanes by shuffle tree, warps 0..7,
// ranks 0..7), so the result is deterministic.
#define POSE_CLUSTER 8
__global__ void __cluster_dims__(POSE_CLUSTER, 1, 1) __launch_bounds__(256)
k_pose_gn_cluster(const PoseArgs a)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    __shared__ double s_part[8][NACC];
    __shared__ double s_cta[2][POSE_CLUSTER][NACC];   // [iteration parity][source CTA]: every CTA receives every CTA's partials
    __shared__ float s_T10[16];
    __shared__ int s_stop;
    __shared__ float s_err_prev;

    const int n = a.n_single_d ? *a.n_single_d : a.n_single;
    const float *X = a.X, *pl = a.pl, *pr = a.mono ? nullptr : a.pr;
    uint8_t *mask = a.mask;
    float *T01 = a.T01;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    if (tid == 0) {
        float T[16], Ti[16];
        for (int i = 0; i < 16; ++i) T[i] = T01[i];
        if (a.mono) inverse4(T, Ti);
        else inv_se3(T, Ti);
        for (int i = 0; i < 16; ++i) s_T10[i] = Ti[i];
        s_stop = 0;
        s_err_prev = 1e10f;
    }
    // only rank 0 records the per-iteration trace; the solve itself runs on every CTA (identical inputs in identical order
    // give identical bits), which replaces the pose broadcast and the second cluster barrier of an iteration
    PoseArgs a_solve = a;
    if (rank != 0) a_solve.trace = nullptr;
    cluster.sync();
    int iter = 0;
    for (; iter < a.max_iter; ++iter) {
        float T10[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) T10[i] = s_T10[i];
        double acc[NACC];
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
        pose_points(a, X, pl, pr, mask, n, rank * 256 + tid, POSE_CLUSTER * 256, T10, acc);
        {
            const double tot = warp_transpose_sum(acc, lane);
            if (lane < NACC) s_part[wid][lane] = tot;
        }
        __syncthreads();
        if (tid < NACC) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += s_part[w][tid];
            s_part[0][tid] = v;                      // column tid is read and written by this thread only
        }
        __syncthreads();
        // all-gather through distributed shared memory: this CTA's 28 sums into row `rank` of every CTA's table
        if (tid < NACC * POSE_CLUSTER) {
            const int r = tid / NACC, k = tid - r * NACC;
            *(cluster.map_shared_rank(&s_cta[iter & 1][0][0], r) + rank * NACC + k) = s_part[0][k];
        }
        cluster.sync();
        if (tid < NACC) {
            double v = 0.0;
#pragma unroll
            for (int r = 0; r < POSE_CLUSTER; ++r) v += s_cta[iter & 1][r][tid];
            s_part[1][tid] = v;
        }
        __syncthreads();
        if (tid == 0) pose_solve(a_solve, &s_part[1][0], n, s_T10, &s_err_prev, &s_stop, 0, iter);
        __syncthreads();
        if (s_stop) { ++iter; break; }
    }
    // nobody may leave while a peer can still write into its shared memory
    cluster.sync();
    if (rank == 0 && tid == 0) {
        float nrm2 = 0.f;
        for (int i = 0; i < 16; ++i) nrm2 += s_T10[i] * s_T10[i];
        const bool ok = !isnan(nrm2);
        if (ok) {
            float T10[16], Ti[16];
            for (int i = 0; i < 16; ++i) T10[i] = s_T10[i];
            inv_se3(T10, Ti);
            for (int i = 0; i < 16; ++i) T01[i] = Ti[i];
        }
        if (a.success) a.success[0] = ok ? 1 : 0;
        if (a.iters) a.iters[0] = iter;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// STRICT-ORDER mode: the reference adds every row's contribution to JtWJ / mJtWr / err_curr one after the other in
// FP32, in point order (motion_estimator.cpp:770-800, :990-1040) -- and its stop test `delta_err < 1e-7` on an FP32
// error of order 1 px only fires when that sum repeats (almost) bit for bit, so the stop iteration depends on the
// summation order.  This kernel reproduces the order exactly: producer warps evaluate the rows of a chunk of points
// (one point per lane, same FP32 arithmetic as the default kernel) and park {Jt, w*Jt, -w*r, e} per row in shared
// memory; ONE consumer warp owns the 28 sums, one per lane, and walks the rows sequentially
// (acc = acc + a*b, un-fused), i.e. 28 independent dependent-add chains of 4N (stereo) / 2N (mono) links -- the
// chain latency (4 cycles per row) is the cost: about 16 us per iteration at N = 2000.  Chunks are double-buffered so
// the producers stay ahead of the consumer.  Structurally zero entries contribute a*0 = +-0, which leaves an FP32 sum
// unchanged, exactly like the `+= JtJ_tmp` of the zero-filled temporary in the reference.
// Shared-memory layout: rows are stored in BLOCKS of four (= one stereo point, two mono points), field-major: a block is
// 16 fields x 16 bytes, field f = {f of row 0, row 1, row 2, row 3}; fields [0..5] Jt, [6..11] w*Jt, [12] -(w*r), [13] e,
// [14] 1, [15] 0.  A consumer lane therefore fetches its two operands for FOUR rows with two 128-bit loads (the first
// version used two 32-bit loads per row and was bound by the shared-memory instruction rate, not by the add chain:
// profiles/r2_strict_*).  Blocks are 17 slots (272 bytes) apart, so the producers' 128-bit stores of eight neighbouring
// points fall on eight different bank groups while the consumer's field offsets stay compile-time immediates.
#define STRICT_PTS_OF(THREADS) ((THREADS) == 64 ? 32 : ((THREADS) == 96 ? 64 : 128))      // points per chunk = producer warps x 32 lanes
// warp 0 = consumer; producers: THREADS = 192 (one large problem): warps 1, 2, 3, 5, and warp 4 idles, so that the consumer has its
// scheduler (warp % 4) to itself -- the solve is bound by the consumer's chain of dependent adds (N = 2000: 0.278 -> 0.265 ms);
// THREADS = 64 (batched solves): ONE producer warp and chunks of 32 points, 18 KB of shared memory instead of 70 KB, so
// seven CTAs fit an SM instead of three (4096 x 500 points: 4.67 -> 6.08 M solves/s; two producers per consumer: 5.66 M;
// capping the registers at 96 for ten CTAs: 6.00 M).
#define STRICT_BLK 68                  // floats per block of four rows (16 fields x 4 + one slot of padding)
#define STRICT_SMEM(RPP, THREADS) (2 * (STRICT_PTS_OF(THREADS) / (4 / (RPP)) + 2) * STRICT_BLK * 4)     // two buffers, each padded by two blocks (prefetch overrun)

template <int RPP>
struct RecSink {
    float Jt_[RPP][6], wJ_[RPP][6], nwr_[RPP], e_[RPP];
    int k;           // next row
    template <int ZERO>
    __device__ __forceinline__ void row(const float *Jt, float w, float r, bool weighted, float e)
    {
#pragma unroll
        for (int q = 0; q < RPP; ++q)
            if (q == k) {
#pragma unroll
                for (int i = 0; i < 6; ++i) { Jt_[q][i] = Jt[i]; wJ_[q][i] = weighted ? w * Jt[i] : Jt[i]; }
                nwr_[q] = -(weighted ? w * r : r);
                e_[q] = e;
            }
        ++k;
    }
    // block: this point's block; half = which half of the block (mono)
    __device__ __forceinline__ void store(float *block, int half) const
    {
#pragma unroll
        for (int f = 0; f < 16; ++f) {
            float v[RPP];
#pragma unroll
            for (int q = 0; q < RPP; ++q)
                v[q] = f < 6 ? Jt_[q][f] : (f < 12 ? wJ_[q][f - 6] : (f == 12 ? nwr_[q] : (f == 13 ? e_[q] : (f == 14 ? 1.0f : 0.f))));
            float *dst = block + 4 * f;
            if (RPP == 4) *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[RPP - 1]);
            else *reinterpret_cast<float2 *>(dst + 2 * half) = make_float2(v[0], v[RPP - 1]);
        }
    }
};

template <int RPP, int THREADS>   // rows per point: 2 mono, 4 stereo
__global__ void __launch_bounds__(THREADS)
k_pose_gn_strict(const PoseArgs a)
{
    constexpr int PPB = 4 / RPP;                                   // points per block of four rows
    constexpr int STRICT_PTS = STRICT_PTS_OF(THREADS);
    constexpr int BLOCKS = STRICT_PTS / PPB;                       // blocks per chunk
    extern __shared__ __align__(16) float s_rec_dyn[];             // [2][(BLOCKS + 2) * STRICT_BLK]
    float (*s_rec)[(BLOCKS + 2) * STRICT_BLK] = reinterpret_cast<float (*)[(BLOCKS + 2) * STRICT_BLK]>(s_rec_dyn);
    __shared__ double s_red[NACC];
    __shared__ float s_T10[16];
    __shared__ int s_stop;
    __shared__ float s_err_prev;

    const int prob = blockIdx.x;
    const int beg = a.offsets ? a.offsets[prob] : 0;
    const int n = a.offsets ? a.offsets[prob + 1] - beg : (a.n_single_d ? *a.n_single_d : a.n_single);
    const float *X = a.X + 3 * (size_t)beg;
    const float *pl = a.pl + 2 * (size_t)beg;
    const float *pr = a.mono ? nullptr : a.pr + 2 * (size_t)beg;
    uint8_t *mask = a.mask + beg;
    float *T01 = a.T01 + 16 * (size_t)prob;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    if (tid == 0) {
        float T[16], Ti[16];
        for (int i = 0; i < 16; ++i) T[i] = T01[i];
        if (a.mono) inverse4(T, Ti);
        else inv_se3(T, Ti);
        for (int i = 0; i < 16; ++i) s_T10[i] = Ti[i];
        s_stop = 0;
        s_err_prev = 1e10f;
    }
    __syncthreads();

    // consumer lane -> the two fields it multiplies: H(i,j) = sum (w*Jt[i]) * Jt[j]; g(i) = sum (-(w*r)) * Jt[i]; err = sum e * 1
    int wa = 15, wb = 15;
    if (lane < 21) {
        int idx = 0;
        for (int i = 0; i < 6; ++i)
            for (int j = i; j < 6; ++j, ++idx)
                if (idx == lane) { wa = 6 + i; wb = j; }
    } else if (lane < 27) { wa = 12; wb = lane - 21; }
    else if (lane == 27) { wa = 13; wb = 14; }

    const int n_chunks = (n + STRICT_PTS - 1) / STRICT_PTS;
    const int pw = THREADS <= 96 ? (wid >= 1 ? wid - 1 : -1)
                                 : ((wid >= 1 && wid <= 3) ? wid - 1 : (wid == (THREADS == 192 ? 5 : 4) ? 3 : -1));      // producer index, -1 = not a producer
    int iter = 0;
    for (; iter < a.max_iter; ++iter) {
        float T10[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) T10[i] = s_T10[i];
        float acc = 0.f;
        for (int c = 0; c <= n_chunks; ++c) {
            if (pw >= 0) {
                if (c < n_chunks) {
                    // producer: point (c * STRICT_PTS + p), p = lane of producer 0 .. 3
                    const int p = pw * 32 + lane;
                    const int i = c * STRICT_PTS + p;
                    const int blk = p / PPB;
                    float *block = &s_rec[c & 1][blk * STRICT_BLK];
                    if (i < n) {
                        RecSink<RPP> sink;
                        sink.k = 0;
                        pose_point(a, X, pl, pr, mask, i, T10, sink);
                        sink.store(block, p % PPB);
                    } else if (PPB == 2 && i == n && (n & 1)) {
                        // mono, odd point count: the second half of the last block is all-zero rows (+0 leaves every sum unchanged)
#pragma unroll
                        for (int f = 0; f < 16; ++f) *reinterpret_cast<float2 *>(block + 4 * f + 2) = make_float2(0.f, 0.f);
                    }
                }
            } else if (wid == 0 && c > 0) {
                // consumer: chunk c-1, one block (four rows) per step, software-pipelined two blocks deep: the adds of block b
                // run while the products of block b+1 are formed and the operands of block b+2 are loaded.
                const int cc = c - 1;
                const int npts = min(STRICT_PTS, n - cc * STRICT_PTS);
                const int blocks = (npts + PPB - 1) / PPB;
                // (blocks past the chunk's end are read -- the buffers are padded -- but their products are never added)
                const float *qa = s_rec[cc & 1] + 4 * wa, *qb = s_rec[cc & 1] + 4 * wb;
                float4 pa = *reinterpret_cast<const float4 *>(qa), pb = *reinterpret_cast<const float4 *>(qb);
                float4 pr0 = make_float4(pa.x * pb.x, pa.y * pb.y, pa.z * pb.z, pa.w * pb.w);
                float4 la = *reinterpret_cast<const float4 *>(qa + STRICT_BLK), lb = *reinterpret_cast<const float4 *>(qb + STRICT_BLK);
                qa += 2 * STRICT_BLK; qb += 2 * STRICT_BLK;
#pragma unroll 4
                for (int b = 0; b < blocks; ++b) {
                    const float4 na = *reinterpret_cast<const float4 *>(qa), nb = *reinterpret_cast<const float4 *>(qb);
                    qa += STRICT_BLK; qb += STRICT_BLK;
                    acc = acc + pr0.x; acc = acc + pr0.y; acc = acc + pr0.z; acc = acc + pr0.w;
                    pr0 = make_float4(la.x * lb.x, la.y * lb.y, la.z * lb.z, la.w * lb.w);
                    la = na; lb = nb;
                }
            }
            __syncthreads();
        }
        if (wid == 0 && lane < NACC) s_red[lane] = (double)acc;
        __syncthreads();
        if (tid == 0) pose_solve(a, s_red, n, s_T10, &s_err_prev, &s_stop, prob, iter);
        __syncthreads();
        if (s_stop) { ++iter; break; }
    }
    if (tid == 0) {
        float nrm2 = 0.f;
        for (int i = 0; i < 16; ++i) nrm2 += s_T10[i] * s_T10[i];
        const bool ok = !isnan(nrm2);
        if (ok) {
            float T10[16], Ti[16];
            for (int i = 0; i < 16; ++i) T10[i] = s_T10[i];
            inv_se3(T10, Ti);
            for (int i = 0; i < 16; ++i) T01[i] = Ti[i];
        }
        if (a.success) a.success[prob] = ok ? 1 : 0;
        if (a.iters) a.iters[prob] = iter;
    }
}

static void host_inv_se3(const float *T, float *O)
{
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) O[i * 4 + j] = T[j * 4 + i];
        float s = 0.f;
        for (int k = 0; k < 3; ++k) s += T[k * 4 + i] * T[k * 4 + 3];
        O[i * 4 + 3] = -s;
    }
    O[12] = O[13] = O[14] = 0.f; O[15] = 1.f;
}

int vo_pose_launch_ex_d(vo_ctx *ctx, int n_prob, const int *offsets_d, int n_single, const int *n_single_d, const float *X_d,
                        const float *pl_d, const float *pr_d, const float *Kl, const float *Kr, const float *T_lr, float thres,
                        int mono, int variant, float *T01_d, uint8_t *mask_d, int *success_d, int *iters_d, int flags,
                        int max_iter, float *trace_d)
{
    PoseArgs a;
    a.offsets = offsets_d; a.n_single = n_single; a.n_single_d = n_single_d;
    a.X = X_d; a.pl = pl_d; a.pr = pr_d;
    for (int i = 0; i < 4; ++i) { a.Kl[i] = Kl[i]; a.Kr[i] = Kr ? Kr[i] : Kl[i]; }
    if (T_lr) host_inv_se3(T_lr, a.T_rl);
    else { for (int i = 0; i < 16; ++i) a.T_rl[i] = (i % 5 == 0) ? 1.f : 0.f; }
    a.thres = thres; a.mono = mono; a.variant = variant;
    a.T01 = T01_d; a.mask = mask_d; a.success = success_d; a.iters = iters_d;
    a.max_iter = (max_iter > 0 && max_iter < POSE_MAX_ITER) ? max_iter : POSE_MAX_ITER;
    a.no_early_stop = (flags & VO_POSE_NO_EARLY_STOP) ? 1 : 0;
    a.trace = trace_d;
    ctx->launches++;
    if (flags & VO_POSE_STRICT) {
        // sequential FP32 sums in point order: one CTA per problem whatever its size (the consumer chain is the cost)
        {   // 32 / 64 KB of dynamic shared memory (double-buffered row records); the attribute is process-wide: set once
            static std::once_flag once;
            std::call_once(once, [] {
                cudaFuncSetAttribute(k_pose_gn_strict<2, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, STRICT_SMEM(2, 64));
                cudaFuncSetAttribute(k_pose_gn_strict<4, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, STRICT_SMEM(4, 64));
                cudaFuncSetAttribute(k_pose_gn_strict<2, 192>, cudaFuncAttributeMaxDynamicSharedMemorySize, STRICT_SMEM(2, 192));
                cudaFuncSetAttribute(k_pose_gn_strict<4, 192>, cudaFuncAttributeMaxDynamicSharedMemorySize, STRICT_SMEM(4, 192));
            });
        }
        if (n_prob == 1) {
            if (mono) k_pose_gn_strict<2, 192><<<1, 192, STRICT_SMEM(2, 192), ctx->stream>>>(a);
            else k_pose_gn_strict<4, 192><<<1, 192, STRICT_SMEM(4, 192), ctx->stream>>>(a);
        } else {
            if (mono) k_pose_gn_strict<2, 64><<<n_prob, 64, STRICT_SMEM(2, 64), ctx->stream>>>(a);
            else k_pose_gn_strict<4, 64><<<n_prob, 64, STRICT_SMEM(4, 64), ctx->stream>>>(a);
        }
        VO_CUDA(cudaGetLastError());
        return VO_OK;
    }
    // one large problem (n_single = exact count or upper bound): thread-block cluster of 8 CTAs
    if (n_prob == 1 && !offsets_d && n_single >= 1024) {
        k_pose_gn_cluster<<<POSE_CLUSTER, 256, 0, ctx->stream>>>(a);
        VO_CUDA(cudaGetLastError());
        return VO_OK;
    }
    if (n_prob > 1) k_pose_gn<32><<<n_prob, 32, 0, ctx->stream>>>(a);
    else k_pose_gn<512><<<n_prob, 512, 0, ctx->stream>>>(a);
    VO_CUDA(cudaGetLastError());
    return VO_OK;
}

int vo_pose_launch_d(vo_ctx *ctx, int n_prob, const int *offsets_d, int n_single, const int *n_single_d, const float *X_d,
                     const float *pl_d, const float *pr_d, const float *Kl, const float *Kr, const float *T_lr, float thres,
                     int mono, int variant, float *T01_d, uint8_t *mask_d, int *success_d, int *iters_d)
{
    return vo_pose_launch_ex_d(ctx, n_prob, offsets_d, n_single, n_single_d, X_d, pl_d, pr_d, Kl, Kr, T_lr, thres, mono, variant,
                               T01_d, mask_d, success_d, iters_d, ctx->pose_flags & VO_POSE_STRICT, 0, nullptr);
}

extern "C" int vo_set_pose_mode(vo_ctx *ctx, int flags)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE((flags & ~VO_POSE_STRICT) == 0, VO_ERR_INVALID_ARG, "vo_set_pose_mode: VO_POSE_FAST or VO_POSE_STRICT");
    ctx->pose_flags = flags;
    return VO_OK;
}

extern "C" int vo_get_pose_mode(const vo_ctx *ctx) { return ctx ? ctx->pose_flags : VO_ERR_INVALID_ARG; }

extern "C" int vo_pose_gn_stereo_batch_ex_d(vo_ctx *ctx, int n_prob, const int *offsets_d, const float *X_d,
                                            const float *pts_l1_d, const float *pts_r1_d, const float *K_l4,
                                            const float *K_r4, const float *T_lr, float thres_reproj_outlier,
                                            float *T01_inout_d, uint8_t *mask_inlier_d, int *success_d, int *iters_d,
                                            int flags, int max_iter, float *trace_d)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n_prob >= 0, VO_ERR_INVALID_ARG, "negative size");
    if (n_prob == 0) return VO_OK;
    VO_REQUIRE(offsets_d && X_d && pts_l1_d && pts_r1_d && K_l4 && K_r4 && T_lr && T01_inout_d && mask_inlier_d,
               VO_ERR_INVALID_ARG, "null pointer");
    VO_REQUIRE((flags & ~(VO_POSE_STRICT | VO_POSE_NO_EARLY_STOP)) == 0 && max_iter >= 0, VO_ERR_INVALID_ARG, "bad pose flags");
    VO_CUDA(cudaSetDevice(ctx->device));
    return vo_pose_launch_ex_d(ctx, n_prob, offsets_d, 0, nullptr, X_d, pts_l1_d, pts_r1_d, K_l4, K_r4, T_lr, thres_reproj_outlier, 0, 0,
                               T01_inout_d, mask_inlier_d, success_d, iters_d, flags, max_iter, trace_d);
}

extern "C" int vo_pose_gn_stereo_batch_d(vo_ctx *ctx, int n_prob, const int *offsets_d, const float *X_d,
                                         const float *pts_l1_d, const float *pts_r1_d, const float *K_l4,
                                         const float *K_r4, const float *T_lr, float thres_reproj_outlier,
                                         float *T01_inout_d, uint8_t *mask_inlier_d, int *success_d, int *iters_d)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    return vo_pose_gn_stereo_batch_ex_d(ctx, n_prob, offsets_d, X_d, pts_l1_d, pts_r1_d, K_l4, K_r4, T_lr, thres_reproj_outlier,
                                        T01_inout_d, mask_inlier_d, success_d, iters_d, ctx->pose_flags & VO_POSE_STRICT, 0, nullptr);
}

// Host-pointer single-problem entry points: one H2D, one launch, one D2H.
static int pose_host(vo_ctx *ctx, const float *X, const float *pl, const float *pr, int n, const float *Kl,
                     const float *Kr, const float *T_lr, float thres, int mono, int variant, float *T01_inout,
                     uint8_t *mask, int *success, int *iters_out, int flags, int max_iter, float *trace)
{
    VO_REQUIRE((flags & ~(VO_POSE_STRICT | VO_POSE_NO_EARLY_STOP)) == 0 && max_iter >= 0, VO_ERR_INVALID_ARG, "bad pose flags");
    VO_CUDA(cudaSetDevice(ctx->device));
    const int mi = (max_iter > 0 && max_iter < POSE_MAX_ITER) ? max_iter : POSE_MAX_ITER;
    const size_t trace_bytes = trace ? (size_t)mi * 24 * sizeof(float) : 0;
    const size_t oX = 0, oPl = oX + (size_t)n * 12, oPr = oPl + (size_t)n * 8, oT = oPr + (size_t)n * 8;
    const size_t oFlags = oT + 64, oMask = oFlags + 16, oTrace = (oMask + (size_t)n + 63) & ~(size_t)63, total = oTrace + trace_bytes + 64;
    int rc = vo_stage_reserve(ctx, total);
    if (rc) return rc;
    uint8_t *h = ctx->h_stage, *d = ctx->d_stage;
    memcpy(h + oX, X, (size_t)n * 12);
    memcpy(h + oPl, pl, (size_t)n * 8);
    if (pr) memcpy(h + oPr, pr, (size_t)n * 8);
    memcpy(h + oT, T01_inout, 64);
    VO_CUDA(cudaMemcpyAsync(d, h, oFlags, cudaMemcpyHostToDevice, ctx->stream));
    if (trace) VO_CUDA(cudaMemsetAsync(d + oTrace, 0, trace_bytes, ctx->stream));
    rc = vo_pose_launch_ex_d(ctx, 1, nullptr, n, nullptr, (const float *)(d + oX), (const float *)(d + oPl), (const float *)(d + oPr), Kl, Kr,
                             T_lr, thres, mono, variant, (float *)(d + oT), d + oMask, (int *)(d + oFlags), (int *)(d + oFlags + 4), flags, mi,
                             trace ? (float *)(d + oTrace) : nullptr);
    if (rc) return rc;
    VO_CUDA(cudaMemcpyAsync(h + oT, d + oT, total - oT, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    const int ok = *(int *)(h + oFlags);
    if (ok) memcpy(T01_inout, h + oT, 64);
    memcpy(mask, h + oMask, (size_t)n);
    if (trace) memcpy(trace, h + oTrace, trace_bytes);
    if (success) *success = ok;
    if (iters_out) *iters_out = *(int *)(h + oFlags + 4);
    return VO_OK;
}

extern "C" int vo_pose_gn_stereo(vo_ctx *ctx, const float *X, const float *pts_l1, const float *pts_r1, int n,
                                 const float *K_l4, const float *K_r4, const float *T_lr, float thres_reproj_outlier,
                                 float *T01_inout, uint8_t *mask_inlier, int *success, int *iters_out)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    return vo_pose_gn_stereo_ex(ctx, X, pts_l1, pts_r1, n, K_l4, K_r4, T_lr, thres_reproj_outlier, T01_inout, mask_inlier, success,
                                iters_out, ctx->pose_flags & VO_POSE_STRICT, 0, nullptr);
}

extern "C" int vo_pose_gn_stereo_ex(vo_ctx *ctx, const float *X, const float *pts_l1, const float *pts_r1, int n,
                                    const float *K_l4, const float *K_r4, const float *T_lr, float thres_reproj_outlier,
                                    float *T01_inout, uint8_t *mask_inlier, int *success, int *iters_out, int flags,
                                    int max_iter, float *trace)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0, VO_ERR_INVALID_ARG, "negative size");
    VO_REQUIRE(K_l4 && K_r4 && T_lr && T01_inout, VO_ERR_INVALID_ARG, "null pointer");
    if (n == 0) { if (success) *success = 1; if (iters_out) *iters_out = 1; return VO_OK; }
    VO_REQUIRE(X && pts_l1 && pts_r1 && mask_inlier, VO_ERR_INVALID_ARG, "null pointer");
    return pose_host(ctx, X, pts_l1, pts_r1, n, K_l4, K_r4, T_lr, thres_reproj_outlier, 0, 0, T01_inout, mask_inlier,
                     success, iters_out, flags, max_iter, trace);
}

extern "C" int vo_pose_gn_mono(vo_ctx *ctx, const float *X, const float *pts1, int n, float fx, float fy, float cx,
                               float cy, int thres_reproj_outlier, int standalone_variant, float *R01_inout,
                               float *t01_inout, uint8_t *mask_inlier, int *success, int *iters_out)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    return vo_pose_gn_mono_ex(ctx, X, pts1, n, fx, fy, cx, cy, thres_reproj_outlier, standalone_variant, R01_inout, t01_inout,
                              mask_inlier, success, iters_out, ctx->pose_flags & VO_POSE_STRICT, 0, nullptr);
}

extern "C" int vo_pose_gn_mono_ex(vo_ctx *ctx, const float *X, const float *pts1, int n, float fx, float fy, float cx,
                                  float cy, int thres_reproj_outlier, int standalone_variant, float *R01_inout,
                                  float *t01_inout, uint8_t *mask_inlier, int *success, int *iters_out, int flags,
                                  int max_iter, float *trace)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(n >= 0, VO_ERR_INVALID_ARG, "negative size");
    VO_REQUIRE(R01_inout && t01_inout, VO_ERR_INVALID_ARG, "null pointer");
    if (n == 0) { if (success) *success = 1; if (iters_out) *iters_out = 1; return VO_OK; }
    VO_REQUIRE(X && pts1 && mask_inlier, VO_ERR_INVALID_ARG, "null pointer");
    float T[16] = {0};
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T[i * 4 + j] = R01_inout[i * 3 + j]; T[i * 4 + 3] = t01_inout[i]; }
    T[15] = 1.f;
    const float K[4] = {fx, fy, cx, cy};
    int ok = 0;
    int rc = pose_host(ctx, X, pts1, nullptr, n, K, K, nullptr, (float)thres_reproj_outlier, 1, standalone_variant ? 1 : 0, T, mask_inlier, &ok,
                       iters_out, flags, max_iter, trace);
    if (rc) return rc;
    if (ok) for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R01_inout[i * 3 + j] = T[i * 4 + j]; t01_inout[i] = T[i * 4 + 3]; }
    if (success) *success = ok;
    return VO_OK;
}
