/*
 * oracle/klt_scale_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement (FP32, sequential sums, same operation order) of
 *   FeatureTracker::trackWithScale                 core/visual_odometry/feature_tracker.cpp:236-504
 *   image_processing::interpImageSameRatio         core/util/image_processing.cpp:79-118
 *   image_processing::interpImage3SameRatio        core/util/image_processing.cpp:268-331
 *   cv::Sobel(u8 -> CV_32F, 3x3, BORDER_DEFAULT)    call sites stereo_vo.cpp:551-552, mono_vo.cpp:781-782
 *     (third-party OpenCV; restated here and pinned bit-exactly against cv2.Sobel in tests)
 *
 * Reference defect kept selectable (`faithful`): the per-sample buffers and masks are allocated once
 * outside the feature loop and `resize(n, value)` never resets them, so a sample that falls outside
 * the image silently reuses the value/mask left by an earlier feature or iteration
 * (feature_tracker.cpp:324-333 + image_processing.cpp:88-89, 275-278).  faithful=1 reproduces that
 * (inherently sequential across features); faithful=0 is the intended semantics (out-of-image samples
 * are masked out), which is what the CUDA kernel implements.  The two agree whenever every sample of
 * every feature stays inside the image.
 * PARITY UNPINNED by the reference (no golden vectors, cannot be compiled here).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int reflect101i(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * (len - 1) - p; }
    return p;
}

/* cv::Sobel(src u8, dst CV_32F, dx, dy, ksize 3, scale 1, delta 0, BORDER_DEFAULT). */
void orc_sobel3_f32(const uint8_t *src, int w, int h, int step, float *du, float *dv)
{
    for (int y = 0; y < h; ++y) {
        int ym = reflect101i(y - 1, h), yp = reflect101i(y + 1, h);
        for (int x = 0; x < w; ++x) {
            int xm = reflect101i(x - 1, w), xp = reflect101i(x + 1, w);
            int a00 = src[ym * step + xm], a01 = src[ym * step + x], a02 = src[ym * step + xp];
            int a10 = src[y * step + xm], a12 = src[y * step + xp];
            int a20 = src[yp * step + xm], a21 = src[yp * step + x], a22 = src[yp * step + xp];
            du[y * w + x] = (float)((a02 - a00) + 2 * (a12 - a10) + (a22 - a20));
            dv[y * w + x] = (float)((a20 - a00) + 2 * (a21 - a01) + (a22 - a02));
        }
    }
}

static inline float interp4(const float *p, int n_cols, float ax, float ay, float axay)
{
    const float I1 = p[0], I2 = p[1], I4 = p[1 + n_cols], I3 = p[n_cols];
    return axay * (I1 - I2 - I3 + I4) + ax * (-I1 + I2) + ay * (-I1 + I3) + I1;
}

#define HALF_WIN 11
#define WIN_LEN 23
#define MAX_ELEM (WIN_LEN * WIN_LEN)

/* Returns 0, or -4 if the reference would have thrown a NaN runtime_error.
 * iters_out (nullable): iterations executed per feature (0 if skipped). */
int orc_track_with_scale(const uint8_t *img0, const float *du0, const float *dv0, const uint8_t *img1, int w, int h,
                         int step0, int step1, const float *pts0, const float *scale_est, int n, float *pts_track,
                         uint8_t *mask_valid, int faithful, int *iters_out)
{
    const int MAX_ITER = 30;
    const float EPS_ERR_RATE = 1e-3f, EPS_UPDATE = 1e-4f, minEigThreshold = 1e-4f;
    const int n_cols = w, n_rows = h;
    float *I0 = (float *)malloc(sizeof(float) * w * h), *I1 = (float *)malloc(sizeof(float) * w * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) { I0[y * w + x] = (float)img0[y * step0 + x]; I1[y * w + x] = (float)img1[y * step1 + x]; }

    float pattx[MAX_ELEM], patty[MAX_ELEM], pattsx[MAX_ELEM], pattsy[MAX_ELEM];
    int n_elem = 0;
    for (int v = 0; v < WIN_LEN; ++v)
        for (int u = !(v & 0x01); u < WIN_LEN; u += 2) { pattx[n_elem] = (float)(u - HALF_WIN); patty[n_elem] = (float)(v - HALF_WIN); ++n_elem; }

    float I0p[MAX_ELEM] = {0}, du0p[MAX_ELEM] = {0}, dv0p[MAX_ELEM] = {0}, I1p[MAX_ELEM] = {0};
    uint8_t mI0[MAX_ELEM] = {0}, mI1[MAX_ELEM] = {0};
    int rc = 0;

    for (int i = 0; i < n; ++i) {
        if (iters_out) iters_out[i] = 0;
        if (!mask_valid[i]) continue;
        const float pt0x = pts0[2 * i], pt0y = pts0[2 * i + 1];
        const float pt1x = pts_track[2 * i], pt1y = pts_track[2 * i + 1];
        const float scale = scale_est[i];
        for (int j = 0; j < n_elem; ++j) { pattsx[j] = pattx[j] * scale; pattsy[j] = patty[j] * scale; }
        float ax = pt0x - floorf(pt0x), ay = pt0y - floorf(pt0y), axay = ax * ay;
        if (ax < 0 || ax > 1 || ay < 0 || ay > 1) { mask_valid[i] = 0; continue; }
        if (!faithful) { memset(mI0, 0, sizeof(mI0)); memset(mI1, 0, sizeof(mI1)); }
        /* interpImage3SameRatio */
        for (int j = 0; j < n_elem; ++j) {
            const float uc = pt0x + pattx[j], vc = pt0y + patty[j];
            const int u0 = (int)uc, v0 = (int)vc;
            if (u0 < 1 || u0 >= n_cols - 2 || v0 < 1 || v0 >= n_rows - 2) continue;
            const int idx = v0 * n_cols + u0;
            I0p[j] = interp4(I0 + idx, n_cols, ax, ay, axay);
            du0p[j] = interp4(du0 + idx, n_cols, ax, ay, axay);
            dv0p[j] = interp4(dv0 + idx, n_cols, ax, ay, axay);
            mI0[j] = 1;
        }
        float A11 = 0, A12 = 0, A22 = 0;
        for (int j = 0; j < n_elem; ++j)
            if (mI0[j]) { A11 += du0p[j] * du0p[j]; A12 += du0p[j] * dv0p[j]; A22 += dv0p[j] * dv0p[j]; }
        const float D = A11 * A22 - A12 * A12;
        if (D < minEigThreshold) { mask_valid[i] = 0; continue; }
        const float invD = (float)(1.0 / D);
        const float iD_A11 = A11 * invD, iD_A12 = A12 * invD, iD_A22 = A22 * invD;
        float err_curr = 0, err_prev = 1e12f;
        float tx = pt1x - pt0x, ty = pt1y - pt0y;
        int iter;
        for (iter = 0; iter < MAX_ITER; ++iter) {
            const float pux = pt0x + tx, puy = pt0y + ty;
            ax = pux - floorf(pux); ay = puy - floorf(puy); axay = ax * ay;
            if (ax < 0 || ax > 1 || ay < 0 || ay > 1) { mask_valid[i] = 0; break; }
            if (isnan(ax + ay)) { rc = -4; goto done; }
            /* interpImageSameRatio */
            for (int j = 0; j < n_elem; ++j) {
                const float uc = pux + pattsx[j], vc = puy + pattsy[j];
                if (uc < 1 || uc >= (float)(n_cols - 2) || vc < 1 || vc >= (float)(n_rows - 2)) continue;
                const int u0 = (int)uc, v0 = (int)vc;
                I1p[j] = interp4(I1 + v0 * n_cols + u0, n_cols, ax, ay, axay);
                mI1[j] = 1;
            }
            float b1 = 0, b2 = 0;
            int cnt_valid = 0;
            err_curr = 0;
            for (int j = 0; j < n_elem; ++j)
                if (mI0[j] && mI1[j]) {
                    if (isnan(I0p[j]) || isnan(I1p[j]) || isnan(du0p[j]) || isnan(dv0p[j])) { rc = -4; goto done; }
                    const float r = I1p[j] - I0p[j];
                    b1 += du0p[j] * r; b2 += dv0p[j] * r; err_curr += r * r;
                    ++cnt_valid;
                }
            const float dtu = (-iD_A22 * b1 + iD_A12 * b2);
            const float dtv = (iD_A12 * b1 - iD_A11 * b2);
            if (isnan(dtu + dtv)) { rc = -4; goto done; }
            tx += dtu; ty += dtv;
            err_curr /= (float)cnt_valid;
            err_curr = sqrtf(err_curr);
            const float err_rate = fabsf(err_prev - err_curr) / err_prev;
            const float dt_norm = dtu * dtu + dtv * dtv;
            if (iter > 1 && (err_rate <= EPS_ERR_RATE || dt_norm <= EPS_UPDATE)) { ++iter; break; }
            err_prev = err_curr;
            if (!faithful) memset(mI1, 0, sizeof(mI1));
        }
        if (iters_out) iters_out[i] = iter > MAX_ITER ? MAX_ITER : iter;
        if (isnan(err_curr)) mask_valid[i] = 0;
        else if (err_curr <= 30) { pts_track[2 * i] = pt0x + tx; pts_track[2 * i + 1] = pt0y + ty; mask_valid[i] = 1; }
        else mask_valid[i] = 0;
    }
done:
    free(I0); free(I1);
    return rc;
}
