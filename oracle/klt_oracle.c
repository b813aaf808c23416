/*
 * oracle/klt_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Scalar CPU restatement of the pyramidal Lucas-Kanade tracker the reference
 * delegates to: cv::calcOpticalFlowPyrLK, called from
 *   core/visual_odometry/feature_tracker.cpp:29,60,69,108,117,186
 * The arithmetic lives in OpenCV (third-party, unpinned "OpenCV 4" in
 * core/CMakeLists.txt:12; this container ships opencv-python-headless 4.13.0.92),
 * not under /root/reference.  This file restates the published algorithm of
 * modules/video/src/lkpyramid.cpp (buildOpticalFlowPyramid + calcScharrDeriv +
 * LKTrackerInvoker, scalar path) and modules/imgproc/src/pyramids.cpp (pyrDown 8u).
 *
 * Pinning: tests/test_oracle_klt.py checks every function below against
 * cv2 4.13.0 itself (cv2.pyrDown / cv2.Scharr / cv2.buildOpticalFlowPyramid are
 * bit-exact; cv2.calcOpticalFlowPyrLK agrees to float-summation-order noise).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define W_BITS 14
#define DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))

static inline int reflect101(int p, int len)
{
    /* BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba ; valid for |overshoot| < len */
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

/* pyrDown, 8-bit single channel. dst size = ((w+1)/2, (h+1)/2).
 * 5x5 [1 4 6 4 1]x[1 4 6 4 1] at even coordinates, reflect-101, (sum+128)>>8. */
void orc_pyrdown_u8(const uint8_t *src, int w, int h, int sstep,
                    uint8_t *dst, int dstep)
{
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
    for (int y = 0; y < dh; ++y) {
        for (int x = 0; x < dw; ++x) {
            int acc = 0;
            static const int k[5] = {1, 4, 6, 4, 1};
            for (int dy = -2; dy <= 2; ++dy) {
                int sy = reflect101(2 * y + dy, h);
                int row = 0;
                for (int dx = -2; dx <= 2; ++dx) {
                    int sx = reflect101(2 * x + dx, w);
                    row += k[dx + 2] * src[sy * sstep + sx];
                }
                acc += k[dy + 2] * row;
            }
            dst[y * dstep + x] = (uint8_t)((acc + 128) >> 8);
        }
    }
}

/* Scharr derivative, int16 interleaved (dx, dy), unnormalised [3 10 3]x[-1 0 1],
 * reflect-101 at the image edge (== cv2.Scharr(..., CV_16S) BORDER_DEFAULT). */
void orc_scharr_s16(const uint8_t *src, int w, int h, int sstep,
                    int16_t *dst /* 2*w per row */, int dstep_elems)
{
    for (int y = 0; y < h; ++y) {
        int ym = reflect101(y - 1, h), yp = reflect101(y + 1, h);
        for (int x = 0; x < w; ++x) {
            int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            int a00 = src[ym * sstep + xm], a01 = src[ym * sstep + x], a02 = src[ym * sstep + xp];
            int a10 = src[y * sstep + xm], a12 = src[y * sstep + xp];
            int a20 = src[yp * sstep + xm], a21 = src[yp * sstep + x], a22 = src[yp * sstep + xp];
            int dx = 3 * (a02 - a00) + 10 * (a12 - a10) + 3 * (a22 - a20);
            int dy = 3 * (a20 - a00) + 10 * (a21 - a01) + 3 * (a22 - a02);
            dst[y * dstep_elems + 2 * x] = (int16_t)dx;
            dst[y * dstep_elems + 2 * x + 1] = (int16_t)dy;
        }
    }
}

/* Effective number of levels-1 (cv::buildOpticalFlowPyramid return value). */
int orc_effective_max_level(int w, int h, int win, int max_level)
{
    int level = 0;
    for (level = 0; level < max_level; ++level) {
        w = (w + 1) / 2;
        h = (h + 1) / 2;
        if (w <= win || h <= win) return level;
    }
    return max_level;
}

typedef struct {
    int w, h;
    int pad;           /* padding on each side */
    int pitch;         /* padded row pitch in pixels */
    uint8_t *img;      /* padded, reflect-101 border */
    int16_t *deriv;    /* padded (2 shorts / px), zero border */
} orc_level;

static void level_alloc(orc_level *L, int w, int h, int pad)
{
    L->w = w; L->h = h; L->pad = pad; L->pitch = w + 2 * pad;
    L->img = (uint8_t *)malloc((size_t)L->pitch * (h + 2 * pad));
    L->deriv = NULL;
}

static void level_fill_border(orc_level *L)
{
    int p = L->pad;
    for (int y = -p; y < L->h + p; ++y) {
        int sy = reflect101(y, L->h);
        for (int x = -p; x < L->w + p; ++x) {
            if (x >= 0 && x < L->w && y >= 0 && y < L->h) continue;
            int sx = reflect101(x, L->w);
            L->img[(y + p) * L->pitch + (x + p)] = L->img[(sy + p) * L->pitch + (sx + p)];
        }
    }
}

static void level_make_deriv(orc_level *L)
{
    int p = L->pad;
    size_t n = (size_t)L->pitch * (L->h + 2 * p) * 2;
    L->deriv = (int16_t *)calloc(n, sizeof(int16_t));
    orc_scharr_s16(L->img + p * L->pitch + p, L->w, L->h, L->pitch,
                   L->deriv + ((size_t)p * L->pitch + p) * 2, L->pitch * 2);
}

typedef struct {
    int nlevels; /* = effective max_level + 1 */
    orc_level lv[16];
} orc_pyramid;

orc_pyramid *orc_pyramid_build(const uint8_t *img, int w, int h, int step,
                               int win, int max_level, int with_deriv)
{
    orc_pyramid *P = (orc_pyramid *)calloc(1, sizeof(orc_pyramid));
    int eff = orc_effective_max_level(w, h, win, max_level);
    P->nlevels = eff + 1;
    int pad = win + 1;
    for (int l = 0; l <= eff; ++l) {
        orc_level *L = &P->lv[l];
        if (l == 0) {
            level_alloc(L, w, h, pad);
            for (int y = 0; y < h; ++y)
                memcpy(L->img + (y + pad) * L->pitch + pad, img + (size_t)y * step, w);
        } else {
            orc_level *S = &P->lv[l - 1];
            level_alloc(L, (S->w + 1) / 2, (S->h + 1) / 2, pad);
            orc_pyrdown_u8(S->img + S->pad * S->pitch + S->pad, S->w, S->h, S->pitch,
                           L->img + pad * L->pitch + pad, L->pitch);
        }
        level_fill_border(L);
        if (with_deriv) level_make_deriv(L);
    }
    return P;
}

void orc_pyramid_free(orc_pyramid *P)
{
    if (!P) return;
    for (int l = 0; l < P->nlevels; ++l) { free(P->lv[l].img); free(P->lv[l].deriv); }
    free(P);
}

int orc_pyramid_nlevels(const orc_pyramid *P) { return P->nlevels; }
void orc_pyramid_level_size(const orc_pyramid *P, int l, int *w, int *h) { *w = P->lv[l].w; *h = P->lv[l].h; }
void orc_pyramid_get_level(const orc_pyramid *P, int l, uint8_t *dst /* w*h */)
{
    const orc_level *L = &P->lv[l];
    for (int y = 0; y < L->h; ++y)
        memcpy(dst + (size_t)y * L->w, L->img + (y + L->pad) * L->pitch + L->pad, L->w);
}
void orc_pyramid_get_deriv(const orc_pyramid *P, int l, int16_t *dst /* h*w*2 */)
{
    const orc_level *L = &P->lv[l];
    for (int y = 0; y < L->h; ++y)
        memcpy(dst + (size_t)y * L->w * 2,
               L->deriv + ((size_t)(y + L->pad) * L->pitch + L->pad) * 2, sizeof(int16_t) * 2 * L->w);
}

static inline int cv_round(float v) { return (int)lrintf(v); } /* round-half-even, as cvRound/SSE */
static inline int cv_floor(float v) { return (int)floorf(v); }

#define OPTFLOW_USE_INITIAL_FLOW 4

/* One pyramid level of LKTrackerInvoker for all points (scalar path).
 * iters_out (optional): per point, number of iterations executed at this level. */
static void lk_level(const orc_level *I, const orc_level *J, int level, int max_level,
                     const float *prev_pts, float *next_pts, uint8_t *status, float *err,
                     int n, int win, int max_count, double epsilon2, int flags, float min_eig_thr,
                     int *iters_out)
{
    const float halfWin = (win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    int16_t *Iwin = (int16_t *)malloc(sizeof(int16_t) * win * win * 3);
    int16_t *dIwin = Iwin + win * win;
    const int stepI = I->pitch, stepJ = J->pitch, dstep = I->pitch * 2;
    const uint8_t *Ibase = I->img + I->pad * I->pitch + I->pad;
    const uint8_t *Jbase = J->img + J->pad * J->pitch + J->pad;
    const int16_t *Dbase = I->deriv + ((size_t)I->pad * I->pitch + I->pad) * 2;

    for (int pt = 0; pt < n; ++pt) {
        float prevx = prev_pts[2 * pt] * (float)(1. / (1 << level));
        float prevy = prev_pts[2 * pt + 1] * (float)(1. / (1 << level));
        float nextx, nexty;
        if (level == max_level) {
            if (flags & OPTFLOW_USE_INITIAL_FLOW) {
                nextx = next_pts[2 * pt] * (float)(1. / (1 << level));
                nexty = next_pts[2 * pt + 1] * (float)(1. / (1 << level));
            } else { nextx = prevx; nexty = prevy; }
        } else {
            nextx = next_pts[2 * pt] * 2.f;
            nexty = next_pts[2 * pt + 1] * 2.f;
        }
        next_pts[2 * pt] = nextx; next_pts[2 * pt + 1] = nexty;
        if (iters_out) iters_out[pt] = 0;

        prevx -= halfWin; prevy -= halfWin;
        int ipx = cv_floor(prevx), ipy = cv_floor(prevy);
        if (ipx < -win || ipx >= I->w || ipy < -win || ipy >= I->h) {
            if (level == 0) { status[pt] = 0; err[pt] = 0; }
            continue;
        }
        float a = prevx - ipx, b = prevy - ipy;
        int iw00 = cv_round((1.f - a) * (1.f - b) * (1 << W_BITS));
        int iw01 = cv_round(a * (1.f - b) * (1 << W_BITS));
        int iw10 = cv_round((1.f - a) * b * (1 << W_BITS));
        int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
        float iA11 = 0, iA12 = 0, iA22 = 0;
        for (int y = 0; y < win; ++y) {
            const uint8_t *src = Ibase + (y + ipy) * stepI + ipx;
            const int16_t *dsrc = Dbase + (y + ipy) * dstep + ipx * 2;
            int16_t *Iptr = Iwin + y * win;
            int16_t *dIptr = dIwin + y * win * 2;
            for (int x = 0; x < win; ++x, dsrc += 2, dIptr += 2) {
                int ival = DESCALE(src[x] * iw00 + src[x + 1] * iw01 + src[x + stepI] * iw10 + src[x + stepI + 1] * iw11, W_BITS - 5);
                int ixval = DESCALE(dsrc[0] * iw00 + dsrc[2] * iw01 + dsrc[dstep] * iw10 + dsrc[dstep + 2] * iw11, W_BITS);
                int iyval = DESCALE(dsrc[1] * iw00 + dsrc[3] * iw01 + dsrc[dstep + 1] * iw10 + dsrc[dstep + 3] * iw11, W_BITS);
                Iptr[x] = (int16_t)ival; dIptr[0] = (int16_t)ixval; dIptr[1] = (int16_t)iyval;
                iA11 += (float)(ixval * ixval);
                iA12 += (float)(ixval * iyval);
                iA22 += (float)(iyval * iyval);
            }
        }
        float A11 = iA11 * FLT_SCALE, A12 = iA12 * FLT_SCALE, A22 = iA22 * FLT_SCALE;
        float D = A11 * A22 - A12 * A12;
        float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
        if (minEig < min_eig_thr || D < FLT_EPSILON) {
            if (level == 0) status[pt] = 0;
            continue;
        }
        D = 1.f / D;
        nextx -= halfWin; nexty -= halfWin;
        float pdx = 0, pdy = 0;
        int j;
        for (j = 0; j < max_count; ++j) {
            int inx = cv_floor(nextx), iny = cv_floor(nexty);
            if (inx < -win || inx >= J->w || iny < -win || iny >= J->h) {
                if (level == 0) status[pt] = 0;
                break;
            }
            if (iters_out) iters_out[pt] = j + 1;
            a = nextx - inx; b = nexty - iny;
            iw00 = cv_round((1.f - a) * (1.f - b) * (1 << W_BITS));
            iw01 = cv_round(a * (1.f - b) * (1 << W_BITS));
            iw10 = cv_round((1.f - a) * b * (1 << W_BITS));
            iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
            float ib1 = 0, ib2 = 0;
            for (int y = 0; y < win; ++y) {
                const uint8_t *Jptr = Jbase + (y + iny) * stepJ + inx;
                const int16_t *Iptr = Iwin + y * win;
                const int16_t *dIptr = dIwin + y * win * 2;
                for (int x = 0; x < win; ++x, dIptr += 2) {
                    int diff = DESCALE(Jptr[x] * iw00 + Jptr[x + 1] * iw01 + Jptr[x + stepJ] * iw10 + Jptr[x + stepJ + 1] * iw11, W_BITS - 5) - Iptr[x];
                    ib1 += (float)(diff * dIptr[0]);
                    ib2 += (float)(diff * dIptr[1]);
                }
            }
            float b1 = ib1 * FLT_SCALE, b2 = ib2 * FLT_SCALE;
            float dx = (float)((A12 * b2 - A22 * b1) * D);
            float dy = (float)((A12 * b1 - A11 * b2) * D);
            nextx += dx; nexty += dy;
            next_pts[2 * pt] = nextx + halfWin; next_pts[2 * pt + 1] = nexty + halfWin;
            if ((double)dx * dx + (double)dy * dy <= epsilon2) break;
            if (j > 0 && fabs(dx + pdx) < 0.01 && fabs(dy + pdy) < 0.01) {
                next_pts[2 * pt] -= dx * 0.5f; next_pts[2 * pt + 1] -= dy * 0.5f;
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (status[pt] && level == 0) {
            float npx = next_pts[2 * pt] - halfWin, npy = next_pts[2 * pt + 1] - halfWin;
            int inx = cv_floor(npx), iny = cv_floor(npy);
            if (inx < -win || inx >= J->w || iny < -win || iny >= J->h) { status[pt] = 0; continue; }
            float aa = npx - inx, bb = npy - iny;
            iw00 = cv_round((1.f - aa) * (1.f - bb) * (1 << W_BITS));
            iw01 = cv_round(aa * (1.f - bb) * (1 << W_BITS));
            iw10 = cv_round((1.f - aa) * bb * (1 << W_BITS));
            iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
            float errval = 0.f;
            for (int y = 0; y < win; ++y) {
                const uint8_t *Jptr = Jbase + (y + iny) * stepJ + inx;
                const int16_t *Iptr = Iwin + y * win;
                for (int x = 0; x < win; ++x) {
                    int diff = DESCALE(Jptr[x] * iw00 + Jptr[x + 1] * iw01 + Jptr[x + stepJ] * iw10 + Jptr[x + stepJ + 1] * iw11, W_BITS - 5) - Iptr[x];
                    errval += fabsf((float)diff);
                }
            }
            err[pt] = errval * 1.f / (32 * win * win);
        }
    }
    free(Iwin);
}

/* cv::calcOpticalFlowPyrLK restatement with the default criteria the reference
 * uses (COUNT+EPS, 30, 0.01), minEigThreshold 1e-4.
 * next_pts is in/out (used as the prior when flags has USE_INITIAL_FLOW).
 * iters (optional): [nlevels][n] iterations executed, for the algorithmic-bytes model.
 * Returns number of levels actually used. */
int orc_calc_optical_flow_pyr_lk(const uint8_t *img0, const uint8_t *img1, int w, int h, int step0, int step1,
                                 const float *prev_pts, float *next_pts, int n, int win, int max_level, int flags,
                                 uint8_t *status, float *err, int *iters)
{
    orc_pyramid *P0 = orc_pyramid_build(img0, w, h, step0, win, max_level, 1);
    orc_pyramid *P1 = orc_pyramid_build(img1, w, h, step1, win, max_level, 0);
    int eff = P0->nlevels - 1;
    for (int i = 0; i < n; ++i) status[i] = 1;
    for (int i = 0; i < n; ++i) err[i] = 0.f;
    if (!(flags & OPTFLOW_USE_INITIAL_FLOW))
        memcpy(next_pts, prev_pts, sizeof(float) * 2 * n);
    for (int level = eff; level >= 0; --level)
        lk_level(&P0->lv[level], &P1->lv[level], level, eff, prev_pts, next_pts, status, err, n, win,
                 30, 0.01 * 0.01, flags, 1e-4f, iters ? iters + (size_t)level * n : NULL);
    orc_pyramid_free(P0);
    orc_pyramid_free(P1);
    return eff + 1;
}
