/*
 * oracle/pose_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Eigen-free CPU restatement, in the reference's precision (FP32) and summation order
 * (sequential over points), of
 *   MotionEstimator::poseOnlyBundleAdjustment        core/visual_odometry/motion_estimator.cpp:665-861
 *                                                    standalone/motion_estimator/motion_estimator.cpp:4-193
 *   MotionEstimator::poseOnlyBundleAdjustment_Stereo core :863-1088 == standalone :195-411
 *   calcJtJ_x/_y, calcJtWJ_x/_y                      core :1342-1576 == standalone :413-647
 *   geometry::se3Exp_f / inverseSE3_f                core/util/geometry_library.cpp:386-440, 554-560
 *   Eigen::LDLT<Matrix<float,6,6>>::solve            (third-party Eigen3, unpinned, absent here:
 *        restated from the published algorithm -- diagonal-pivoted, unblocked, lower)
 *
 * PARITY UNPINNED by the reference: it ships no golden vectors for this path
 * and cannot be compiled here (no Eigen / OpenCV headers).  The restatement is anchored on
 * the cited lines and validated against analytic ground truth (tests/test_oracle_pose.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* ------------------------------------------------------------------ small helpers */
static void mat4_mul_f(const float *A, const float *B, float *C)
{
    float T[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j];
            T[i * 4 + j] = s;
        }
    memcpy(C, T, sizeof(T));
}

/* geometry::inverseSE3_f (geometry_library.cpp:554-560); row-major 4x4 */
void orc_inverse_se3_f(const float *T, float *Ti)
{
    float R[9], t[3];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R[i * 3 + j] = T[i * 4 + j]; t[i] = T[i * 4 + 3]; }
    float O[16] = {0};
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) O[i * 4 + j] = R[j * 3 + i];
        float s = 0.f;
        for (int k = 0; k < 3; ++k) s += R[k * 3 + i] * t[k];
        O[i * 4 + 3] = -s;
    }
    O[15] = 1.f;
    memcpy(Ti, O, sizeof(O));
}

/* General 4x4 inverse (Matrix4f::inverse(), used by the mono variant at motion_estimator.cpp:700
 * and by StereoVO at stereo_vo.cpp:643) -- cofactor expansion in float. */
void orc_inverse4_f(const float *m, float *out)
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
    float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    float id = 1.0f / det;
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * id;
}

/* geometry::se3Exp_f (geometry_library.cpp:386-440). xi = (v, w); T row-major 4x4.
 * sin/cos are the unqualified C calls on a float (promoted to double); the scalar
 * coefficients are rounded to float before they multiply the float matrices. */
void orc_se3exp_f(const float *xi, float *T)
{
    float v[3] = {xi[0], xi[1], xi[2]}, w[3] = {xi[3], xi[4], xi[5]};
    float theta = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    float wx[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    float wx2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float s = 0.f;
            for (int k = 0; k < 3; ++k) s += wx[i * 3 + k] * wx[k * 3 + j];
            wx2[i * 3 + j] = s;
        }
    float a, b, c; /* R = I + a wx + b wx2 ; V = I + b' wx + c wx2 */
    float bV;
    if (theta < 1e-7) {
        a = 1.f; b = (float)0.5; bV = (float)0.5; c = 0.33333333333333333333333333f;
    } else {
        double th = (double)theta;
        a = (float)(sin(th) / th);
        b = (float)((1 - cos(th)) / (double)(theta * theta));
        bV = b;
        c = (float)((th - sin(th)) / (double)(theta * theta * theta));
    }
    float R[9], V[9];
    for (int i = 0; i < 9; ++i) {
        float id = (i == 0 || i == 4 || i == 8) ? 1.f : 0.f;
        R[i] = (id + a * wx[i]) + b * wx2[i];
        V[i] = (id + bV * wx[i]) + c * wx2[i];
    }
    memset(T, 0, sizeof(float) * 16);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T[i * 4 + j] = R[i * 3 + j];
        T[i * 4 + 3] = (V[i * 3 + 0] * v[0] + V[i * 3 + 1] * v[1]) + V[i * 3 + 2] * v[2];
    }
    T[15] = 1.f;
}

/* Eigen::LDLT<Matrix<float,6,6>, Lower>::compute + solve, unblocked with diagonal pivoting. */
void orc_ldlt6_solve_f(const float *A_in /*6x6 row-major, symmetric*/, const float *b, float *x)
{
    const int n = 6;
    float m[36];
    int tr[6];
    memcpy(m, A_in, sizeof(m));
#define M(i, j) m[(i) * 6 + (j)]
    for (int k = 0; k < n; ++k) {
        int big = k;
        float bv = fabsf(M(k, k));
        for (int i = k + 1; i < n; ++i)
            if (fabsf(M(i, i)) > bv) { bv = fabsf(M(i, i)); big = i; }
        tr[k] = big;
        if (big != k) {
            int s = n - big - 1;
            for (int j = 0; j < k; ++j) { float t = M(k, j); M(k, j) = M(big, j); M(big, j) = t; }
            for (int i = 0; i < s; ++i) { float t = M(big + 1 + i, k); M(big + 1 + i, k) = M(big + 1 + i, big); M(big + 1 + i, big) = t; }
            { float t = M(k, k); M(k, k) = M(big, big); M(big, big) = t; }
            for (int i = k + 1; i < big; ++i) { float t = M(i, k); M(i, k) = M(big, i); M(big, i) = t; }
        }
        int rs = n - k - 1;
        if (k > 0) {
            float temp[6];
            for (int j = 0; j < k; ++j) temp[j] = M(j, j) * M(k, j);
            float s = 0.f;
            for (int j = 0; j < k; ++j) s += M(k, j) * temp[j];
            M(k, k) -= s;
            for (int i = 0; i < rs; ++i) {
                float s2 = 0.f;
                for (int j = 0; j < k; ++j) s2 += M(k + 1 + i, j) * temp[j];
                M(k + 1 + i, k) -= s2;
            }
        }
        float akk = M(k, k);
        if (rs > 0 && fabsf(akk) > 0.f)
            for (int i = 0; i < rs; ++i) M(k + 1 + i, k) /= akk;
    }
    float y[6];
    memcpy(y, b, sizeof(y));
    for (int k = 0; k < n; ++k) if (tr[k] != k) { float t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
    for (int i = 0; i < n; ++i) { float s = y[i]; for (int j = 0; j < i; ++j) s -= M(i, j) * y[j]; y[i] = s; }
    const float tol = 1.0f / 3.402823466e+38f;
    for (int i = 0; i < n; ++i) y[i] = fabsf(M(i, i)) > tol ? y[i] / M(i, i) : 0.f;
    for (int i = n - 1; i >= 0; --i) { float s = y[i]; for (int j = i + 1; j < n; ++j) s -= M(j, i) * y[j]; y[i] = s; }
    for (int k = n - 1; k >= 0; --k) if (tr[k] != k) { float t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
    memcpy(x, y, sizeof(y));
#undef M
}

/* Upper-triangle rank-1 accumulate with the structural zero of each row
 * (calcJtWJ_x / calcJtWJ_y: wJt = w*Jt first, then wJt(i)*Jt(j)). H is 6x6 row-major, upper. */
static inline void acc_row(float *H, float *g, const float *Jt, float w, float r, int zero_idx, int weighted)
{
    float wJ[6];
    for (int i = 0; i < 6; ++i) wJ[i] = weighted ? w * Jt[i] : Jt[i];
    for (int i = 0; i < 6; ++i) {
        if (i == zero_idx) continue;
        for (int j = i; j < 6; ++j) {
            if (j == zero_idx) continue;
            H[i * 6 + j] += wJ[i] * Jt[j];
        }
    }
    float wr = weighted ? w * r : r;
    for (int i = 0; i < 6; ++i) g[i] -= wr * Jt[i];
}

#define MAX_ITER 100
#define THRES_HUBER 0.5f
#define THRES_DELTA_XI 1e-6f
#define THRES_DELTA_ERROR 1e-7f
#define LAMBDA 0.00001f

static int is_nan_mat(const float *T) { float s = 0; for (int i = 0; i < 16; ++i) s += T[i] * T[i]; return isnan(s); }

static int finish_iteration(float *H, float *g, float *T10, float err_curr, float *err_prev, float *trace, int iter)
{
    for (int i = 0; i < 6; ++i) for (int j = 0; j < i; ++j) H[i * 6 + j] = H[j * 6 + i];
    float delta_err = fabsf(err_curr - *err_prev);
    for (int i = 0; i < 6; ++i) H[i * 6 + i] *= (1.0f + LAMBDA);
    float dxi[6], dT[16];
    orc_ldlt6_solve_f(H, g, dxi);
    orc_se3exp_f(dxi, dT);
    mat4_mul_f(dT, T10, T10);
    *err_prev = err_curr;
    if (trace) { memcpy(trace + iter * 24, T10, 64); trace[iter * 24 + 16] = err_curr; memcpy(trace + iter * 24 + 17, dxi, 24); trace[iter * 24 + 23] = delta_err; }
    float nrm = 0.f;
    for (int i = 0; i < 6; ++i) nrm += dxi[i] * dxi[i];
    nrm = sqrtf(nrm);
    return (nrm < THRES_DELTA_XI || delta_err < THRES_DELTA_ERROR);
}

/* Stereo. X n*3, pl/pr n*2, K = fx,fy,cx,cy, T_lr / T01 row-major 4x4 (T01 in-out).
 * trace (nullable): MAX_ITER x 24 floats {T10 after update (16), err, dxi(6), delta_err}.
 * Returns is_success; *iters_out = number of iterations executed. */
int orc_pose_gn_stereo_ex(const float *X, const float *pl, const float *pr, int n, const float *Kl, const float *Kr,
                          const float *T_lr, float thres_reproj_outlier, float *T01, uint8_t *mask, int *iters_out,
                          float *trace, int max_iter, int no_early_stop);
int orc_pose_gn_stereo(const float *X, const float *pl, const float *pr, int n, const float *Kl, const float *Kr,
                       const float *T_lr, float thres_reproj_outlier, float *T01, uint8_t *mask, int *iters_out,
                       float *trace)
{
    return orc_pose_gn_stereo_ex(X, pl, pr, n, Kl, Kr, T_lr, thres_reproj_outlier, T01, mask, iters_out, trace, MAX_ITER, 0);
}

/* _ex: max_iter <= MAX_ITER iterations; no_early_stop = 1 ignores the stop test (test-only: fixed-point and
 * iterate-by-iterate comparisons with the kernel). */
int orc_pose_gn_stereo_ex(const float *X, const float *pl, const float *pr, int n, const float *Kl, const float *Kr,
                          const float *T_lr, float thres_reproj_outlier, float *T01, uint8_t *mask, int *iters_out,
                          float *trace, int max_iter, int no_early_stop)
{
    float T_rl[16];
    orc_inverse_se3_f(T_lr, T_rl);
    const float fx_l = Kl[0], fy_l = Kl[1], cx_l = Kl[2], cy_l = Kl[3];
    const float fx_r = Kr[0], fy_r = Kr[1], cx_r = Kr[2], cy_r = Kr[3];
    for (int i = 0; i < n; ++i) mask[i] = 1;
    float err_prev = 1e10f;
    float T10[16];
    orc_inverse_se3_f(T01, T10); /* :903-904: explicit R^T, -R^T t */
    int iter = 0;
    if (max_iter <= 0 || max_iter > MAX_ITER) max_iter = MAX_ITER;
    for (; iter < max_iter; ++iter) {
        float H[36] = {0}, g[6] = {0};
        float err_curr = 0.f;
        float inv_npts = 1.0f / (float)n;
        for (int i = 0; i < n; ++i) {
            const float *Xi = X + 3 * i;
            float Xl[3], Xr[3];
            for (int r = 0; r < 3; ++r)
                Xl[r] = ((T10[r * 4 + 0] * Xi[0] + T10[r * 4 + 1] * Xi[1]) + T10[r * 4 + 2] * Xi[2]) + T10[r * 4 + 3];
            for (int r = 0; r < 3; ++r)
                Xr[r] = ((T_rl[r * 4 + 0] * Xl[0] + T_rl[r * 4 + 1] * Xl[1]) + T_rl[r * 4 + 2] * Xl[2]) + T_rl[r * 4 + 3];
            float iz_l = 1.0f / Xl[2], xiz_l = Xl[0] * iz_l, yiz_l = Xl[1] * iz_l;
            float fxxiz_l = fx_l * xiz_l, fyyiz_l = fy_l * yiz_l;
            float rx_l = (fxxiz_l + cx_l) - pl[2 * i], ry_l = (fyyiz_l + cy_l) - pl[2 * i + 1];
            float iz_r = 1.0f / Xr[2], xiz_r = Xr[0] * iz_r, yiz_r = Xr[1] * iz_r;
            float fxxiz_r = fx_r * xiz_r, fyyiz_r = fy_r * yiz_r;
            float rx_r = (fxxiz_r + cx_r) - pr[2 * i], ry_r = (fyyiz_r + cy_r) - pr[2 * i + 1];
            float weight = 1.0f;
            float absrxry = fabsf(rx_l) + fabsf(ry_l) + fabsf(rx_r) + fabsf(ry_r);
            absrxry *= 0.5f;
            if (absrxry >= THRES_HUBER) weight = THRES_HUBER / absrxry;
            mask[i] = (absrxry >= thres_reproj_outlier) ? 0 : 1;
            float Jt[6];
            Jt[0] = fx_l * iz_l; Jt[1] = 0.f; Jt[2] = -fxxiz_l * iz_l; Jt[3] = -fxxiz_l * yiz_l;
            Jt[4] = fx_l * (1.0f + xiz_l * xiz_l); Jt[5] = -fx_l * yiz_l;
            acc_row(H, g, Jt, weight, rx_l, 1, 1); err_curr += rx_l * rx_l;
            Jt[0] = 0.f; Jt[1] = fy_l * iz_l; Jt[2] = -fyyiz_l * iz_l; Jt[3] = -fy_l * (1.0f + yiz_l * yiz_l);
            Jt[4] = fyyiz_l * xiz_l; Jt[5] = fy_l * xiz_l;
            acc_row(H, g, Jt, weight, ry_l, 0, 1); err_curr += ry_l * ry_l;
            Jt[0] = fx_r * iz_r; Jt[1] = 0.f; Jt[2] = -fxxiz_r * iz_r; Jt[3] = -fxxiz_r * yiz_r;
            Jt[4] = fx_r * (1.0f + xiz_r * xiz_r); Jt[5] = -fx_r * yiz_r;
            acc_row(H, g, Jt, weight, rx_r, 1, 1); err_curr += rx_r * rx_r;
            Jt[0] = 0.f; Jt[1] = fy_r * iz_r; Jt[2] = -fyyiz_r * iz_r; Jt[3] = -fy_r * (1.0f + yiz_r * yiz_r);
            Jt[4] = fyyiz_r * xiz_r; Jt[5] = fy_r * xiz_r;
            acc_row(H, g, Jt, weight, ry_r, 0, 1); err_curr += ry_r * ry_r;
        }
        err_curr *= (inv_npts * 0.5f);
        err_curr = sqrtf(err_curr);
        int stop = finish_iteration(H, g, T10, err_curr, &err_prev, trace, iter);
        if (stop && !no_early_stop) { ++iter; break; }
    }
    if (iters_out) *iters_out = iter;
    if (!is_nan_mat(T10)) { orc_inverse_se3_f(T10, T01); return 1; }
    return 0;
}

/* Mono. variant 0 = core (weighted y-row adds w*ry^2 to err, motion_estimator.cpp:794-799),
 *       variant 1 = standalone (adds ry^2, standalone/.../motion_estimator.cpp:135).
 * R01 (3x3 row-major) and t01 in-out. thres is an int (motion_estimator.h:117). */
int orc_pose_gn_mono_ex(const float *X, const float *p1, int n, float fx, float fy, float cx, float cy, int thres,
                        float *R01, float *t01, uint8_t *mask, int variant, int *iters_out, float *trace, int max_iter,
                        int no_early_stop);
int orc_pose_gn_mono(const float *X, const float *p1, int n, float fx, float fy, float cx, float cy, int thres,
                     float *R01, float *t01, uint8_t *mask, int variant, int *iters_out, float *trace)
{
    return orc_pose_gn_mono_ex(X, p1, n, fx, fy, cx, cy, thres, R01, t01, mask, variant, iters_out, trace, MAX_ITER, 0);
}

int orc_pose_gn_mono_ex(const float *X, const float *p1, int n, float fx, float fy, float cx, float cy, int thres,
                        float *R01, float *t01, uint8_t *mask, int variant, int *iters_out, float *trace, int max_iter,
                        int no_early_stop)
{
    const float THRES_REPROJ_ERROR = (float)thres;
    float T01[16] = {0}, T10[16];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T01[i * 4 + j] = R01[i * 3 + j]; T01[i * 4 + 3] = t01[i]; }
    T01[15] = 1.f;
    orc_inverse4_f(T01, T10); /* :700 general Matrix4f::inverse() */
    float err_prev = 1e10f;
    int iter = 0;
    if (max_iter <= 0 || max_iter > MAX_ITER) max_iter = MAX_ITER;
    for (; iter < max_iter; ++iter) {
        float H[36] = {0}, g[6] = {0};
        float err_curr = 0.f;
        float inv_npts = 1.0f / (float)n;
        for (int i = 0; i < n; ++i) {
            const float *Xi = X + 3 * i;
            float Xw[3];
            for (int r = 0; r < 3; ++r)
                Xw[r] = ((T10[r * 4 + 0] * Xi[0] + T10[r * 4 + 1] * Xi[1]) + T10[r * 4 + 2] * Xi[2]) + T10[r * 4 + 3];
            float iz = 1.0f / Xw[2], xiz = Xw[0] * iz, yiz = Xw[1] * iz;
            float fxxiz = fx * xiz, fyyiz = fy * yiz;
            float rx = (fxxiz + cx) - p1[2 * i], ry = (fyyiz + cy) - p1[2 * i + 1];
            float weight = 1.0f;
            int flag_weight = 0;
            float absrxry = fabsf(rx) + fabsf(ry);
            if (absrxry >= THRES_HUBER) { weight = THRES_HUBER / absrxry; flag_weight = 1; }
            mask[i] = (absrxry >= THRES_REPROJ_ERROR) ? 0 : 1;
            float Jt[6];
            Jt[0] = fx * iz; Jt[1] = 0.f; Jt[2] = -fxxiz * iz; Jt[3] = -fxxiz * yiz;
            Jt[4] = fx * (1.0f + xiz * xiz); Jt[5] = -fx * yiz;
            acc_row(H, g, Jt, weight, rx, 1, flag_weight);
            err_curr += rx * rx;
            Jt[0] = 0.f; Jt[1] = fy * iz; Jt[2] = -fyyiz * iz; Jt[3] = -fy * (1.0f + yiz * yiz);
            Jt[4] = fyyiz * xiz; Jt[5] = fy * xiz;
            acc_row(H, g, Jt, weight, ry, 0, flag_weight);
            if (flag_weight && variant == 0) err_curr += (weight * ry) * ry;
            else err_curr += ry * ry;
        }
        err_curr *= (inv_npts * 0.5f);
        int stop = finish_iteration(H, g, T10, err_curr, &err_prev, trace, iter);
        if (stop && !no_early_stop) { ++iter; break; }
    }
    if (iters_out) *iters_out = iter;
    if (!is_nan_mat(T10)) {
        float T01u[16];
        orc_inverse_se3_f(T10, T01u);
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R01[i * 3 + j] = T01u[i * 4 + j]; t01[i] = T01u[i * 4 + 3]; }
        return 1;
    }
    return 0;
}
