/*
 * oracle/misc_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Eigen-free CPU restatement of the batched-elementwise rows of the hot path:
 *   mapping::triangulateDLT            core/util/triangulate_3d.cpp:5-130
 *       (Eigen::JacobiSVD<MatrixXf> is third-party Eigen3, unpinned, absent here: restated
 *        from the published two-sided Jacobi algorithm, FP32, square case = no preconditioner)
 *   DepthFilter::updateNormalDistribution      standalone/depth_filter/depth_filter.cpp:3-13
 *   DepthFilter::updateStudentTDistribution    standalone/depth_filter/depth_filter.cpp:15-46
 *       (the reference body does not compile; semantics defined in DESIGN.md "D2")
 *   FeatureTracker::calcPrior                  core/visual_odometry/feature_tracker.cpp:208-234
 *   LandmarkTracking(src, mask) compaction     core/visual_odometry/landmark.cpp:194-231, 291-332
 * PARITY UNPINNED by the reference (no golden vectors, cannot be compiled here).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

void orc_inverse4_f(const float *m, float *out); /* pose_oracle.c */

/* ------------------------------------------------------------------ 4x4 FP32 Jacobi SVD */
/* Plane rotation applied like Eigen's apply_rotation_in_the_plane(x, y, j):
 *   x' = c x + s y ; y' = -s x + c y */
static void rot_rows(float *W, int p, int q, float c, float s)
{
    for (int i = 0; i < 4; ++i) {
        float xi = W[p * 4 + i], yi = W[q * 4 + i];
        W[p * 4 + i] = c * xi + s * yi;
        W[q * 4 + i] = -s * xi + c * yi;
    }
}
/* applyOnTheRight(p, q, j) == rotation j.transpose() = (c, -s) on columns p, q */
static void rot_cols(float *W, int p, int q, float c, float s)
{
    for (int i = 0; i < 4; ++i) {
        float xi = W[i * 4 + p], yi = W[i * 4 + q];
        W[i * 4 + p] = c * xi - s * yi;
        W[i * 4 + q] = s * xi + c * yi;
    }
}

static void make_jacobi(float x, float y, float z, float *c, float *s)
{
    float deno = 2.f * fabsf(y);
    if (deno < FLT_MIN) { *c = 1.f; *s = 0.f; return; }
    float tau = (x - z) / deno;
    float w = sqrtf(tau * tau + 1.f);
    float t = (tau > 0.f) ? 1.f / (tau + w) : 1.f / (tau - w);
    float sign_t = t > 0.f ? 1.f : -1.f;
    float n = 1.f / sqrtf(t * t + 1.f);
    *s = -sign_t * (y / fabsf(y)) * fabsf(t) * n;
    *c = n;
}

/* Right-singular vector of the smallest singular value of a 4x4 FP32 matrix (row-major). */
void orc_svd4_null_f(const float *M, float *v4)
{
    float W[16], V[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    float scale = 0.f;
    for (int i = 0; i < 16; ++i) scale = fmaxf(scale, fabsf(M[i]));
    if (scale == 0.f) scale = 1.f;
    for (int i = 0; i < 16; ++i) W[i] = M[i] / scale;
    float maxDiag = 0.f;
    for (int i = 0; i < 4; ++i) maxDiag = fmaxf(maxDiag, fabsf(W[i * 5]));
    const float precision = 2.f * FLT_EPSILON;
    int finished = 0, sweeps = 0;
    while (!finished && sweeps < 64) {
        finished = 1;
        ++sweeps;
        for (int p = 1; p < 4; ++p)
            for (int q = 0; q < p; ++q) {
                float thr = fmaxf(FLT_MIN, precision * maxDiag);
                if (fabsf(W[p * 4 + q]) > thr || fabsf(W[q * 4 + p]) > thr) {
                    finished = 0;
                    /* real_2x2_jacobi_svd */
                    float m00 = W[p * 4 + p], m01 = W[p * 4 + q], m10 = W[q * 4 + p], m11 = W[q * 4 + q];
                    float t = m00 + m11, d = m10 - m01;
                    float c1, s1;
                    if (fabsf(d) < FLT_MIN) { s1 = 0.f; c1 = 1.f; }
                    else { float u = t / d; float tmp = sqrtf(1.f + u * u); s1 = 1.f / tmp; c1 = u / tmp; }
                    /* m.applyOnTheLeft(0,1,rot1) */
                    float a00 = c1 * m00 + s1 * m10, a01 = c1 * m01 + s1 * m11;
                    float a11 = -s1 * m01 + c1 * m11;
                    float cr, sr;
                    make_jacobi(a00, a01, a11, &cr, &sr);
                    /* j_left = rot1 * j_right^T : (c1,s1)*(cr,-sr) */
                    /* JacobiRotation product (c,s)*(c',s') = (c c' - s s', c s' + s c') for real */
                    float cl = c1 * cr - s1 * (-sr);
                    float sl = c1 * (-sr) + s1 * cr;
                    rot_rows(W, p, q, cl, sl);
                    rot_cols(W, p, q, cr, sr);
                    rot_cols(V, p, q, cr, sr);
                    maxDiag = fmaxf(maxDiag, fmaxf(fabsf(W[p * 4 + p]), fabsf(W[q * 4 + q])));
                }
            }
    }
    int k = 0;
    float best = fabsf(W[0]);
    for (int i = 1; i < 4; ++i)
        if (fabsf(W[i * 5]) < best) { best = fabsf(W[i * 5]); k = i; }
    for (int i = 0; i < 4; ++i) v4[i] = V[i * 4 + k];
}

/* mapping::triangulateDLT, two-camera form (triangulate_3d.cpp:91-130); the same-camera
 * forms (:5-89) are the K0 == K1 case. R10 row-major 3x3, K = fx,fy,cx,cy. */
void orc_triangulate_dlt(const float *pts0, const float *pts1, int n, const float *R10, const float *t10,
                         const float *K0, const float *K1, float *X0, float *X1)
{
    const float fx0 = K0[0], fy0 = K0[1], cx0 = K0[2], cy0 = K0[3];
    const float fx1 = K1[0], fy1 = K1[1], cx1 = K1[2], cy1 = K1[3];
    float P10[12];
    for (int j = 0; j < 3; ++j) {
        P10[0 * 4 + j] = (fx1 * R10[0 * 3 + j] + 0.f * R10[1 * 3 + j]) + cx1 * R10[2 * 3 + j];
        P10[1 * 4 + j] = (0.f * R10[0 * 3 + j] + fy1 * R10[1 * 3 + j]) + cy1 * R10[2 * 3 + j];
        P10[2 * 4 + j] = (0.f * R10[0 * 3 + j] + 0.f * R10[1 * 3 + j]) + 1.f * R10[2 * 3 + j];
    }
    P10[0 * 4 + 3] = (fx1 * t10[0] + 0.f * t10[1]) + cx1 * t10[2];
    P10[1 * 4 + 3] = (0.f * t10[0] + fy1 * t10[1]) + cy1 * t10[2];
    P10[2 * 4 + 3] = (0.f * t10[0] + 0.f * t10[1]) + 1.f * t10[2];
    for (int i = 0; i < n; ++i) {
        const float u0 = pts0[2 * i], v0 = pts0[2 * i + 1], u1 = pts1[2 * i], v1 = pts1[2 * i + 1];
        float M[16] = {0};
        M[0] = -fx0; M[2] = u0 - cx0;
        M[5] = -fy0; M[6] = v0 - cy0;
        for (int j = 0; j < 4; ++j) {
            M[8 + j] = u1 * P10[8 + j] - P10[0 + j];
            M[12 + j] = v1 * P10[8 + j] - P10[4 + j];
        }
        float v[4];
        orc_svd4_null_f(M, v);
        float x0[3] = {v[0] / v[3], v[1] / v[3], v[2] / v[3]};
        for (int r = 0; r < 3; ++r) {
            X0[3 * i + r] = x0[r];
            X1[3 * i + r] = ((R10[r * 3 + 0] * x0[0] + R10[r * 3 + 1] * x0[1]) + R10[r * 3 + 2] * x0[2]) + t10[r];
        }
    }
}

/* ------------------------------------------------------------------ depth filter */
void orc_depth_filter_normal(const double *x_prev, const double *cov_prev, const double *x_curr,
                             const double *cov_curr, int n, double *x_upd, double *cov_upd)
{
    for (int i = 0; i < n; ++i) {
        double inv_cov_sum = 1.0 / (cov_prev[i] + cov_curr[i]);
        cov_upd[i] = (cov_prev[i] * cov_curr[i]) * inv_cov_sum;
        x_upd[i] = (x_prev[i] * cov_curr[i] + x_curr[i] * cov_prev[i]) * inv_cov_sum;
    }
}

/* D2 (semantics DEFINED BY US from the non-compiling sketch, depth_filter.cpp:15-46):
 * sigma^2 := cov_prev, tau^2 := cov_curr, x_ := x_curr, seed{mu,a,b} := x_prev, a, b.
 * Outputs: x_upd = mu', cov_upd = sigma'^2 (variance), a/b updated, x_min/x_max = min/max(prev, x_curr)
 * (legacy/matlab/stereoDisparityTemporal.m:276-280). */
void orc_depth_filter_student_t(const double *x_prev, const double *cov_prev, double *a, double *b, double *x_min,
                                double *x_max, const double *x_curr, const double *cov_curr, int n, double *x_upd,
                                double *cov_upd)
{
    for (int i = 0; i < n; ++i) {
        const double mu = x_prev[i], s2 = cov_prev[i], t2 = cov_curr[i], x = x_curr[i];
        const double ai = a[i], bi = b[i];
        const double inv_apb = 1.0 / (ai + bi);
        const double x_range = x_max[i] - x_min[i];
        const double sigma = sqrt(s2);
        double C1 = ai * inv_apb * 1.0 / sqrt(2.0 * 3.141592) / sigma * exp(-(x - mu) * (x - mu) / (2.0 * s2));
        double C2 = bi * inv_apb / x_range;
        const double invC = 1.0 / (C1 + C2);
        C1 *= invC;
        C2 *= invC;
        const double ss = 1.0 / (1.0 / s2 + 1.0 / t2); /* s^2 */
        const double m = ss * (mu / s2 + x / t2);
        const double mu_new = C1 * m + C2 * mu;
        const double var_new = C1 * (ss + m * m) + C2 * (s2 + mu * mu) - mu_new * mu_new;
        const double F = C1 * (ai + 1.0) / (ai + bi + 1.0) + C2 * ai / (ai + bi + 1.0);
        const double E = C1 * (ai + 1.0) / (ai + bi + 1.0) * (ai + 2.0) / (ai + bi + 2.0) +
                         C2 * ai / (ai + bi + 1.0) * (ai + 1.0) / (ai + bi + 2.0);
        const double a_new = (E - F) / (F - E / F);
        a[i] = a_new;
        b[i] = a_new * (1.0 - F) / F;
        x_upd[i] = mu_new;
        cov_upd[i] = var_new;
        x_min[i] = fmin(x_min[i], x);
        x_max[i] = fmax(x_max[i], x);
    }
}

/* ------------------------------------------------------------------ calcPrior */
void orc_calc_prior(const float *pts0, const float *Xw, int n, const float *Tw1 /*row-major*/, const float *K4,
                    float *pts1_prior)
{
    float T1w[16];
    orc_inverse4_f(Tw1, T1w); /* Tw1.inverse() :215 */
    for (int i = 0; i < n; ++i) {
        const float *X = Xw + 3 * i;
        float X1[3];
        for (int r = 0; r < 3; ++r)
            X1[r] = ((T1w[r * 4 + 0] * X[0] + T1w[r * 4 + 1] * X[1]) + T1w[r * 4 + 2] * X[2]) + T1w[r * 4 + 3];
        float nrm = sqrtf((X1[0] * X1[0] + X1[1] * X1[1]) + X1[2] * X1[2]);
        if (nrm > 0) {
            pts1_prior[2 * i] = K4[0] * X1[0] / X1[2] + K4[2];
            pts1_prior[2 * i + 1] = K4[1] * X1[1] / X1[2] + K4[3];
        } else {
            pts1_prior[2 * i] = pts0[2 * i];
            pts1_prior[2 * i + 1] = pts0[2 * i + 1];
        }
    }
}

/* ------------------------------------------------------------------ stable compaction */
int orc_compact(const uint8_t *mask, int n, int *index_out)
{
    int k = 0;
    for (int i = 0; i < n; ++i)
        if (mask[i]) index_out[k++] = i;
    return k;
}
