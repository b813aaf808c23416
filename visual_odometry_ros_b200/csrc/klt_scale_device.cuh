// klt_scale_device.cuh -- device code of FeatureTracker::trackWithScale (see klt_scale.cu), shared with the fused
// tracking chain in klt.cu.  Include only from translation units compiled with -fmad=false.
#pragma once
#include "vo_internal.cuh"

#define KS_HALF 11
#define KS_NELEM 264
#define KS_PER_LANE 9      // ceil(264 / 32)

// Per-feature record for the reference-faithful border mode (k_klt_scale_fixup, klt_scale.cu)
#define KS_FLAG_PROCESSED 1    // the feature was live at entry: the reference touched its sample buffers
#define KS_FLAG_BORDER 2       // a sample of the template or of an iteration fell outside the image
#define KS_FLAG_RAN 4          // the iteration loop was entered

struct KltScaleArgs {
    const SlotDesc *slots;
    int slot0, slot1;
    const float2 *pts0;
    const float *scale;
    float2 *pts_track;
    uint8_t *mask;
    int *nan_flag;
    int *iters;     // nullable
    int n;
    uint8_t *flags = nullptr;     // nullable [n]: KS_FLAG_* of this pass (faithful mode)
    float2 *pre_pts = nullptr;    // nullable [n]: pts_track as it was on entry (faithful mode replays features from it)
};

// klt_scale.cu, faithful border mode: record pointers of a scale stage / second pass over the records
int vo_klt_scale_prepare(vo_ctx *ctx, KltScaleArgs &a);
int vo_klt_scale_fixup_launch(vo_ctx *ctx, const KltScaleArgs &a);

static __device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

static __device__ __forceinline__ float interp4(float I1, float I2, float I3, float I4, float ax, float ay, float axay)
{
    return axay * (I1 - I2 - I3 + I4) + ax * (-I1 + I2) + ay * (-I1 + I3) + I1;
}

// One feature on one warp (all 32 lanes call it with the same f).
static __device__ __forceinline__ void klt_scale_feature(const KltScaleArgs &a, const int f, const int lane)
{
    if (!a.mask[f]) { if (lane == 0) { if (a.iters) a.iters[f] = 0; if (a.flags) a.flags[f] = 0; } return; }
    const LevelDesc I0 = a.slots[a.slot0].lv[0];
    const LevelDesc I1 = a.slots[a.slot1].lv[0];
    const int n_cols = I0.w, n_rows = I0.h;
    const float2 pt0 = a.pts0[f];
    const float2 pt1 = a.pts_track[f];
    const float scale = a.scale[f];
    if (lane == 0 && a.pre_pts) a.pre_pts[f] = pt1;
    bool border = false;

    // checkerboard pattern: sample j -> (u, v) offsets (feature_tracker.cpp:308-320)
    float pu[KS_PER_LANE], pv[KS_PER_LANE];
#pragma unroll
    for (int k = 0; k < KS_PER_LANE; ++k) {
        const int j = lane + 32 * k;
        const int p = j / 23, r = j - 23 * p;
        const int v = r < 11 ? 2 * p : 2 * p + 1;
        const int u = r < 11 ? 1 + 2 * r : 2 * (r - 11);
        pu[k] = (float)(u - KS_HALF);
        pv[k] = (float)(v - KS_HALF);
    }

    float ax = pt0.x - floorf(pt0.x), ay = pt0.y - floorf(pt0.y), axay = ax * ay;
    if (ax < 0 || ax > 1 || ay < 0 || ay > 1) { if (lane == 0) { a.mask[f] = 0; if (a.flags) a.flags[f] = 0; } return; }
    if (isnan(ax + ay)) { if (lane == 0) { a.mask[f] = 0; atomicExch(a.nan_flag, 1); if (a.flags) a.flags[f] = 0; } return; }

    // ---- template: I0, du0, dv0 at the 264 samples (interpImage3SameRatio)
    float I0p[KS_PER_LANE], dup[KS_PER_LANE], dvp[KS_PER_LANE];
    unsigned m0 = 0;
    double sA11 = 0, sA12 = 0, sA22 = 0;
#pragma unroll
    for (int k = 0; k < KS_PER_LANE; ++k) {
        I0p[k] = 0.f; dup[k] = 0.f; dvp[k] = 0.f;
        if (lane + 32 * k >= KS_NELEM) continue;
        const float uc = pt0.x + pu[k], vc = pt0.y + pv[k];
        const int u0 = (int)uc, v0 = (int)vc;
        if (u0 < 1 || u0 >= n_cols - 2 || v0 < 1 || v0 >= n_rows - 2) { border = true; continue; }
        // 4x4 u8 neighbourhood rows v0-1..v0+2, cols u0-1..u0+2
        int p[4][4];
        const uint8_t *b = I0.img + (ptrdiff_t)(v0 - 1) * I0.pitch + (u0 - 1);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) p[r][c] = __ldg(b + r * I0.pitch + c);
        float gI[2][2], gU[2][2], gV[2][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                gI[r][c] = (float)p[r + 1][c + 1];
                gU[r][c] = (float)((p[r][c + 2] - p[r][c]) + 2 * (p[r + 1][c + 2] - p[r + 1][c]) + (p[r + 2][c + 2] - p[r + 2][c]));
                gV[r][c] = (float)((p[r + 2][c] - p[r][c]) + 2 * (p[r + 2][c + 1] - p[r][c + 1]) + (p[r + 2][c + 2] - p[r][c + 2]));
            }
        I0p[k] = interp4(gI[0][0], gI[0][1], gI[1][0], gI[1][1], ax, ay, axay);
        dup[k] = interp4(gU[0][0], gU[0][1], gU[1][0], gU[1][1], ax, ay, axay);
        dvp[k] = interp4(gV[0][0], gV[0][1], gV[1][0], gV[1][1], ax, ay, axay);
        m0 |= 1u << k;
        sA11 += (double)(dup[k] * dup[k]);
        sA12 += (double)(dup[k] * dvp[k]);
        sA22 += (double)(dvp[k] * dvp[k]);
    }
    const float A11 = (float)warp_sum_d(sA11), A12 = (float)warp_sum_d(sA12), A22 = (float)warp_sum_d(sA22);
    const float D = A11 * A22 - A12 * A12;
    if (D < 1e-4f) {
        const bool any_border = __any_sync(0xffffffffu, border);
        if (lane == 0) { a.mask[f] = 0; if (a.iters) a.iters[f] = 0; if (a.flags) a.flags[f] = KS_FLAG_PROCESSED | (any_border ? KS_FLAG_BORDER : 0); }
        return;
    }
    const float invD = (float)(1.0 / (double)D);
    const float iD_A11 = A11 * invD, iD_A12 = A12 * invD, iD_A22 = A22 * invD;

    float err_curr = 0.f, err_prev = 1e12f;
    float tx = pt1.x - pt0.x, ty = pt1.y - pt0.y;
    bool nan_hit = false;
    int iter = 0;
    for (; iter < 30; ++iter) {
        const float pux = pt0.x + tx, puy = pt0.y + ty;
        ax = pux - floorf(pux); ay = puy - floorf(puy); axay = ax * ay;
        if (isnan(ax + ay)) { nan_hit = true; break; }
        double sb1 = 0, sb2 = 0, serr = 0;
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < KS_PER_LANE; ++k) {
            if (!((m0 >> k) & 1u)) continue;
            const float uc = pux + pu[k] * scale, vc = puy + pv[k] * scale;
            if (uc < 1 || uc >= (float)(n_cols - 2) || vc < 1 || vc >= (float)(n_rows - 2)) { border = true; continue; }
            const int u0 = (int)uc, v0 = (int)vc;
            const uint8_t *b = I1.img + (ptrdiff_t)v0 * I1.pitch + u0;
            const float J1 = (float)__ldg(b), J2 = (float)__ldg(b + 1), J3 = (float)__ldg(b + I1.pitch), J4 = (float)__ldg(b + I1.pitch + 1);
            const float r = interp4(J1, J2, J3, J4, ax, ay, axay) - I0p[k];
            sb1 += (double)(dup[k] * r);
            sb2 += (double)(dvp[k] * r);
            serr += (double)(r * r);
            ++cnt;
        }
        const float b1 = (float)warp_sum_d(sb1), b2 = (float)warp_sum_d(sb2);
        err_curr = (float)warp_sum_d(serr);
        const int cnt_valid = __reduce_add_sync(0xffffffffu, cnt);
        const float dtu = (-iD_A22 * b1 + iD_A12 * b2);
        const float dtv = (iD_A12 * b1 - iD_A11 * b2);
        if (isnan(dtu + dtv)) { nan_hit = true; break; }
        tx += dtu; ty += dtv;
        err_curr /= (float)cnt_valid;
        err_curr = sqrtf(err_curr);
        const float err_rate = fabsf(err_prev - err_curr) / err_prev;
        const float dt_norm = dtu * dtu + dtv * dtv;
        if (iter > 1 && (err_rate <= 1e-3f || dt_norm <= 1e-4f)) { ++iter; break; }
        err_prev = err_curr;
    }
    const bool any_border = __any_sync(0xffffffffu, border);
    if (lane == 0) {
        if (a.flags) a.flags[f] = KS_FLAG_PROCESSED | KS_FLAG_RAN | (any_border ? KS_FLAG_BORDER : 0);
        if (nan_hit) { atomicExch(a.nan_flag, 1); a.mask[f] = 0; }
        else if (isnan(err_curr)) a.mask[f] = 0;
        else if (err_curr <= 30.f) { a.pts_track[f] = make_float2(pt0.x + tx, pt0.y + ty); a.mask[f] = 1; }
        else a.mask[f] = 0;
        if (a.iters) a.iters[f] = iter > 30 ? 30 : iter;
    }
}



// ---------------------------------------------------------------------------------------------------------------
// Reference-faithful border mode.  In the reference the per-sample buffers (I0_patt, du0_patt, dv0_patt, I1_patt) and
// their masks are allocated once outside the feature loop and never reset (feature_tracker.cpp:324-333,
// image_processing.cpp:88-89, 275-278): a sample that falls outside the image keeps whatever the most recent feature /
// iteration wrote at that sample index, and its mask stays true once it has ever been true.  That is a sequential
// dependence between features.  KsState is that buffer state, distributed over the warp like the samples;
// klt_scale_feature_stale() processes ONE feature against it exactly as the reference would.
struct KsState {
    float I0p[KS_PER_LANE], dup[KS_PER_LANE], dvp[KS_PER_LANE], I1p[KS_PER_LANE];
    unsigned m0, m1;
};

static __device__ __forceinline__ void ks_state_clear(KsState &st)
{
#pragma unroll
    for (int k = 0; k < KS_PER_LANE; ++k) { st.I0p[k] = 0.f; st.dup[k] = 0.f; st.dvp[k] = 0.f; st.I1p[k] = 0.f; }
    st.m0 = 0; st.m1 = 0;
}

// pt1_in: the feature's pts_track on entry.  Returns the reference's verdict (ok) and, if ok, the refined point.
static __device__ __noinline__ void klt_scale_feature_stale(const KltScaleArgs &a, const int f, const int lane, const float2 pt1, KsState &st,
                                                            bool &ok, float2 &out, bool &nan_hit)
{
    ok = false; nan_hit = false; out = pt1;
    const LevelDesc I0 = a.slots[a.slot0].lv[0];
    const LevelDesc I1 = a.slots[a.slot1].lv[0];
    const int n_cols = I0.w, n_rows = I0.h;
    const float2 pt0 = a.pts0[f];
    const float scale = a.scale[f];
    float pu[KS_PER_LANE], pv[KS_PER_LANE];
#pragma unroll
    for (int k = 0; k < KS_PER_LANE; ++k) {
        const int j = lane + 32 * k;
        const int p = j / 23, r = j - 23 * p;
        const int v = r < 11 ? 2 * p : 2 * p + 1;
        const int u = r < 11 ? 1 + 2 * r : 2 * (r - 11);
        pu[k] = (float)(u - KS_HALF);
        pv[k] = (float)(v - KS_HALF);
    }
    float ax = pt0.x - floorf(pt0.x), ay = pt0.y - floorf(pt0.y), axay = ax * ay;
    if (ax < 0 || ax > 1 || ay < 0 || ay > 1) return;
    if (isnan(ax + ay)) { nan_hit = true; return; }
    double sA11 = 0, sA12 = 0, sA22 = 0;
#pragma unroll
    for (int k = 0; k < KS_PER_LANE; ++k) {
        if (lane + 32 * k >= KS_NELEM) continue;
        const float uc = pt0.x + pu[k], vc = pt0.y + pv[k];
        const int u0 = (int)uc, v0 = (int)vc;
        if (!(u0 < 1 || u0 >= n_cols - 2 || v0 < 1 || v0 >= n_rows - 2)) {
            int p[4][4];
            const uint8_t *b = I0.img + (ptrdiff_t)(v0 - 1) * I0.pitch + (u0 - 1);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) p[r][c] = __ldg(b + r * I0.pitch + c);
            float gI[2][2], gU[2][2], gV[2][2];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    gI[r][c] = (float)p[r + 1][c + 1];
                    gU[r][c] = (float)((p[r][c + 2] - p[r][c]) + 2 * (p[r + 1][c + 2] - p[r + 1][c]) + (p[r + 2][c + 2] - p[r + 2][c]));
                    gV[r][c] = (float)((p[r + 2][c] - p[r][c]) + 2 * (p[r + 2][c + 1] - p[r][c + 1]) + (p[r + 2][c + 2] - p[r][c + 2]));
                }
            st.I0p[k] = interp4(gI[0][0], gI[0][1], gI[1][0], gI[1][1], ax, ay, axay);
            st.dup[k] = interp4(gU[0][0], gU[0][1], gU[1][0], gU[1][1], ax, ay, axay);
            st.dvp[k] = interp4(gV[0][0], gV[0][1], gV[1][0], gV[1][1], ax, ay, axay);
            st.m0 |= 1u << k;
        }
        if ((st.m0 >> k) & 1u) {
            sA11 += (double)(st.dup[k] * st.dup[k]);
            sA12 += (double)(st.dup[k] * st.dvp[k]);
            sA22 += (double)(st.dvp[k] * st.dvp[k]);
        }
    }
    const float A11 = (float)warp_sum_d(sA11), A12 = (float)warp_sum_d(sA12), A22 = (float)warp_sum_d(sA22);
    const float D = A11 * A22 - A12 * A12;
    if (D < 1e-4f) return;
    const float invD = (float)(1.0 / (double)D);
    const float iD_A11 = A11 * invD, iD_A12 = A12 * invD, iD_A22 = A22 * invD;
    float err_curr = 0.f, err_prev = 1e12f;
    float tx = pt1.x - pt0.x, ty = pt1.y - pt0.y;
    for (int iter = 0; iter < 30; ++iter) {
        const float pux = pt0.x + tx, puy = pt0.y + ty;
        ax = pux - floorf(pux); ay = puy - floorf(puy); axay = ax * ay;
        if (isnan(ax + ay)) { nan_hit = true; return; }
        double sb1 = 0, sb2 = 0, serr = 0;
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < KS_PER_LANE; ++k) {
            if (lane + 32 * k >= KS_NELEM) continue;
            const float uc = pux + pu[k] * scale, vc = puy + pv[k] * scale;
            if (!(uc < 1 || uc >= (float)(n_cols - 2) || vc < 1 || vc >= (float)(n_rows - 2))) {
                const int u0 = (int)uc, v0 = (int)vc;
                const uint8_t *b = I1.img + (ptrdiff_t)v0 * I1.pitch + u0;
                const float J1 = (float)__ldg(b), J2 = (float)__ldg(b + 1), J3 = (float)__ldg(b + I1.pitch), J4 = (float)__ldg(b + I1.pitch + 1);
                st.I1p[k] = interp4(J1, J2, J3, J4, ax, ay, axay);
                st.m1 |= 1u << k;
            }
            if (((st.m0 & st.m1) >> k) & 1u) {
                const float r = st.I1p[k] - st.I0p[k];
                sb1 += (double)(st.dup[k] * r);
                sb2 += (double)(st.dvp[k] * r);
                serr += (double)(r * r);
                ++cnt;
            }
        }
        const float b1 = (float)warp_sum_d(sb1), b2 = (float)warp_sum_d(sb2);
        err_curr = (float)warp_sum_d(serr);
        const int cnt_valid = __reduce_add_sync(0xffffffffu, cnt);
        const float dtu = (-iD_A22 * b1 + iD_A12 * b2);
        const float dtv = (iD_A12 * b1 - iD_A11 * b2);
        if (isnan(dtu + dtv)) { nan_hit = true; return; }
        tx += dtu; ty += dtv;
        err_curr /= (float)cnt_valid;
        err_curr = sqrtf(err_curr);
        const float err_rate = fabsf(err_prev - err_curr) / err_prev;
        const float dt_norm = dtu * dtu + dtv * dtv;
        if (iter > 1 && (err_rate <= 1e-3f || dt_norm <= 1e-4f)) break;
        err_prev = err_curr;
    }
    if (isnan(err_curr)) return;
    if (err_curr <= 30.f) { ok = true; out = make_float2(pt0.x + tx, pt0.y + ty); }
}
