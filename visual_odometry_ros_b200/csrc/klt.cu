// klt.cu -- K-klt: pyramidal Lucas-Kanade, one warp per feature, all levels in one launch.
//
// Device restatement of what cv::calcOpticalFlowPyrLK does for the reference's
// FeatureTracker (core/visual_odometry/feature_tracker.cpp:29,60,69,108,117,186):
// OpenCV's LKTrackerInvoker semantics are reproduced exactly in the integer domain
// (W_BITS=14 fixed-point bilinear weights, cvRound = round-half-even, int16 template with
// 5 fractional bits, status/err only at level 0, epsilon/oscillation stopping rules).
// The 2x2 normal-equation sums are accumulated EXACTLY: per-lane int32 partials are
// reduced with redux.sync (two 16-bit halves -> int64), then rounded once to float --
// OpenCV's float accumulators differ from this only by their summation-order rounding.
//
// Work decomposition: one warp per (pair, feature); lane L owns window pixels
// idx = L + 32*k (k < NPX).  The template (I, Ix, Iy) lives in registers for the whole
// level; every iteration gathers the 2x2 bilinear neighbourhood of the next image from
// L1/L2 (the padded pyramids make the gathers branch-free), so a level costs
// (w+1)^2 * (1 + 4 + k) algorithmic bytes, which is the figure bench.py reports against.
#include "vo_internal.cuh"
#include "klt_scale_device.cuh"

#include <cfloat>
#include <map>
#include <mutex>
#include <cstdlib>

#define W_BITS 14
#ifndef KLT2_MIN_BLOCKS
#define KLT2_MIN_BLOCKS 5
#endif

struct KltArgs {
    const SlotDesc *slots;
    int sid0, sid1;                 // image slots of the pair (single-pair launches: the fused chains)
    const float2 *pts0;
    float2 *pts1;
    uint8_t *status;
    float *err;
    unsigned long long *counters;   // [2*VO_MAX_LEVELS] or null
    const uint8_t *skip_mask;       // nullable: features whose entry is 0 are not tracked (outputs untouched)
    int n;
    int win;
    int top_level;      // effective maxLevel
    int flags;
    int max_count;
    float min_eig;
    double eps2;
    KltPost post;
};
// batched launches: up to VO_IDLIST_MAX pairs, slot ids by value (kept apart from KltArgs so that the fused chain kernel,
// which carries three KltArgs and the TMA descriptors, stays below 4 KB of kernel parameters)
struct KltBatchIds {
    IdList s0, s1;
};

__device__ __forceinline__ long long warp_sum_exact(int v)
{
    // exact 32-lane sum of int32 partials without overflow: split in 16-bit halves
    const int hi = v >> 16;
    const int lo = v & 0xffff;
    const int shi = __reduce_add_sync(0xffffffffu, hi);
    const int slo = __reduce_add_sync(0xffffffffu, lo);
    return ((long long)shi << 16) + (long long)slo;
}

// The same exact 32-lane sum, rounded ONCE to float: |hi-sum| <= 2^20 and lo-sum < 2^21 are exact in FP32 and the
// fused multiply-add rounds the exact value hi*2^16 + lo once, i.e. identically to I2F.S64 of the 64-bit sum
// (which costs an IMAD.WIDE and a multi-cycle 64-bit conversion per sum in the iteration tail).
__device__ __forceinline__ float warp_sum_exact_f(int v)
{
    const int shi = __reduce_add_sync(0xffffffffu, v >> 16);
    const int slo = __reduce_add_sync(0xffffffffu, v & 0xffff);
    return __fmaf_rn((float)shi, 65536.f, (float)slo);
}

// OpenCV's stop tests are written in double on float operands; both have exact FP32 equivalents except inside a
// 2e-4 relative band around epsilon^2, where the double path decides (warp-uniform branch):
//   delta.ddot(delta) <= eps2            dx*dx + dy*dy, products exact in double, one rounding
//   std::abs(dx + prevDelta.x) < 0.01    f < 0.01 (double)  <=>  f <= 0.01f, because 0.01f < 0.01 < nextafter(0.01f)
__device__ __forceinline__ bool klt_eps_reached(float dx, float dy, double eps2, float eps2_lo, float eps2_hi)
{
    const float s = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
    if (s < eps2_lo) return true;
    if (s > eps2_hi) return false;
    return __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= eps2;
}

__device__ __forceinline__ void bilinear_weights(float a, float b, int &iw00, int &iw01, int &iw10, int &iw11)
{
    const float oma = __fsub_rn(1.f, a), omb = __fsub_rn(1.f, b);
    iw00 = __float2int_rn(__fmul_rn(__fmul_rn(oma, omb), (float)(1 << W_BITS)));
    iw01 = __float2int_rn(__fmul_rn(__fmul_rn(a, omb), (float)(1 << W_BITS)));
    iw10 = __float2int_rn(__fmul_rn(__fmul_rn(oma, b), (float)(1 << W_BITS)));
    iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
}

// Result write-back + fused FeatureTracker post-filters (one lane per feature).
__device__ __forceinline__ void klt_epilogue(const KltArgs &a, size_t gi, const SlotDesc &S0, float2 stored, int status, float errv)
{
    {
        a.pts1[gi] = stored;
        if (!status) errv = 0.f;
        if (a.status) a.status[gi] = (uint8_t)status;
        if (a.err) a.err[gi] = errv;
        // ---- fused FeatureTracker post-filters
        const KltPost &P = a.post;
        if (P.mode == 1) {          // track(): feature_tracker.cpp:33-34
            P.mask[gi] = P.mask[gi] && status && errv <= P.thres_err;
        } else if (P.mode == 2) {   // trackWithPrior(): feature_tracker.cpp:191-197
            const float w = (float)S0.lv[0].w, h = (float)S0.lv[0].h;
            P.mask[gi] = P.mask[gi] && status && stored.x > 0.f && stored.x < w && stored.y > 0.f && stored.y < h &&
                         errv <= P.thres_err;
        } else if (P.mode == 4) {   // backward pass of trackBidirection(+WithPrior): :74-83 / :130-149
            const float w = (float)S0.lv[0].w, h = (float)S0.lv[0].h;
            const float2 ref = reinterpret_cast<const float2 *>(P.ref_pts)[gi];
            const float2 fw = reinterpret_cast<const float2 *>(P.fwd_pts)[gi];
            const float ddx = __fsub_rn(stored.x, ref.x), ddy = __fsub_rn(stored.y, ref.y);
            const float dist2 = __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));
            const float bd = (float)P.border;
            const bool inimg = fw.x > bd && fw.x < w - bd && fw.y > bd && fw.y < h - bd;
            const bool fs = P.fwd_status[gi] != 0;
            P.mask[gi] = P.mask[gi] && inimg && fs && status && P.fwd_err[gi] <= P.thres_err && errv <= P.thres_err &&
                         dist2 <= P.thres_bi2;
        }
    }
}

template <int NPX>
__global__ void __launch_bounds__(128)
k_klt(const KltArgs a, const KltBatchIds ids)
{
    const int lane = threadIdx.x & 31;
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int pair = blockIdx.y;
    if (f >= a.n) return;
    const size_t gi = (size_t)pair * a.n + f;
    if (a.skip_mask && !a.skip_mask[gi]) return;
    const SlotDesc &S0 = a.slots[ids.s0.id[pair]];
    const SlotDesc &S1 = a.slots[ids.s1.id[pair]];
    const int win = a.win;
    const int npix = win * win;
    const float halfWin = (float)(win - 1) * 0.5f;

    // per-lane pixel coordinates inside the window (level independent)
    int px[NPX], py[NPX];
#pragma unroll
    for (int k = 0; k < NPX; ++k) {
        const int idx = lane + 32 * k;
        const int y = idx / win;
        py[k] = y;
        px[k] = idx - y * win;
    }

    const float2 p0 = a.pts0[gi];
    float2 stored = (a.flags & VO_KLT_USE_INITIAL_FLOW) ? a.pts1[gi] : p0;  // == nextPts[ptidx]
    int status = 1;
    float errv = 0.f;

    for (int level = a.top_level; level >= 0; --level) {
        const LevelDesc I = S0.lv[level];
        const LevelDesc J = S1.lv[level];
        const float sc = 1.f / (float)(1 << level);     // exact power of two
        float prevx = __fmul_rn(p0.x, sc), prevy = __fmul_rn(p0.y, sc);
        if (level == a.top_level) {
            if (a.flags & VO_KLT_USE_INITIAL_FLOW) { stored.x = __fmul_rn(stored.x, sc); stored.y = __fmul_rn(stored.y, sc); }
            else { stored.x = prevx; stored.y = prevy; }
        } else {
            stored.x = __fmul_rn(stored.x, 2.f); stored.y = __fmul_rn(stored.y, 2.f);
        }
        float nextx = stored.x, nexty = stored.y;

        prevx = __fsub_rn(prevx, halfWin); prevy = __fsub_rn(prevy, halfWin);
        const int ipx = __float2int_rd(prevx), ipy = __float2int_rd(prevy);
        if (ipx < -win || ipx >= I.w || ipy < -win || ipy >= I.h) {
            if (level == 0) { status = 0; errv = 0.f; }
            continue;
        }
        int iw00, iw01, iw10, iw11;
        bilinear_weights(__fsub_rn(prevx, (float)ipx), __fsub_rn(prevy, (float)ipy), iw00, iw01, iw10, iw11);

        // ---- template: I (5 frac bits), Ix, Iy as int16-range ints in registers; exact A sums
        int Iv[NPX], Ix[NPX], Iy[NPX];
        int sA11 = 0, sA12 = 0, sA22 = 0;
        {
            const uint8_t *ib = I.img + (ptrdiff_t)ipy * I.pitch + ipx;
            const short2 *db = I.deriv + (ptrdiff_t)ipy * I.pitch + ipx;
#pragma unroll
            for (int k = 0; k < NPX; ++k) {
                const bool ok = (lane + 32 * k) < npix;
                const int o = ok ? py[k] * I.pitch + px[k] : 0;
                const int s00 = __ldg(ib + o), s01 = __ldg(ib + o + 1);
                const int s10 = __ldg(ib + o + I.pitch), s11 = __ldg(ib + o + I.pitch + 1);
                const short2 d00 = __ldg(db + o), d01 = __ldg(db + o + 1);
                const short2 d10 = __ldg(db + o + I.pitch), d11 = __ldg(db + o + I.pitch + 1);
                int iv = (s00 * iw00 + s01 * iw01 + s10 * iw10 + s11 * iw11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
                int ix = (d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
                int iy = (d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
                if (!ok) { iv = 0; ix = 0; iy = 0; }
                Iv[k] = iv; Ix[k] = ix; Iy[k] = iy;
                sA11 += ix * ix; sA12 += ix * iy; sA22 += iy * iy;
            }
        }
        const float FLT_SCALE = 1.f / (float)(1 << 20);
        const float A11 = __fmul_rn(__ll2float_rn(warp_sum_exact(sA11)), FLT_SCALE);
        const float A12 = __fmul_rn(__ll2float_rn(warp_sum_exact(sA12)), FLT_SCALE);
        const float A22 = __fmul_rn(__ll2float_rn(warp_sum_exact(sA22)), FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dif = __fsub_rn(A11, A22);
        const float minEig = __fdiv_rn(
            __fsub_rn(__fadd_rn(A22, A11),
                      __fsqrt_rn(__fadd_rn(__fmul_rn(dif, dif), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
            (float)(2 * win * win));
        if (a.counters && lane == 0) atomicAdd(a.counters + 2 * level, 1ull);
        if (minEig < a.min_eig || D < FLT_EPSILON) {
            if (level == 0) status = 0;
            continue;
        }
        D = __fdiv_rn(1.f, D);

        nextx = __fsub_rn(nextx, halfWin); nexty = __fsub_rn(nexty, halfWin);
        float pdx = 0.f, pdy = 0.f;
        int j = 0;
        for (; j < a.max_count; ++j) {
            const int inx = __float2int_rd(nextx), iny = __float2int_rd(nexty);
            if (inx < -win || inx >= J.w || iny < -win || iny >= J.h) {
                if (level == 0) status = 0;
                break;
            }
            bilinear_weights(__fsub_rn(nextx, (float)inx), __fsub_rn(nexty, (float)iny), iw00, iw01, iw10, iw11);
            const uint8_t *jb = J.img + (ptrdiff_t)iny * J.pitch + inx;
            int sb1 = 0, sb2 = 0;
#pragma unroll
            for (int k = 0; k < NPX; ++k) {
                const int o = (lane + 32 * k) < npix ? py[k] * J.pitch + px[k] : 0;
                const int s00 = __ldg(jb + o), s01 = __ldg(jb + o + 1);
                const int s10 = __ldg(jb + o + J.pitch), s11 = __ldg(jb + o + J.pitch + 1);
                const int diff = ((s00 * iw00 + s01 * iw01 + s10 * iw10 + s11 * iw11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5)) - Iv[k];
                sb1 += diff * Ix[k];   // Ix = Iy = 0 on padding lanes
                sb2 += diff * Iy[k];
            }
            const float b1 = __fmul_rn(__ll2float_rn(warp_sum_exact(sb1)), FLT_SCALE);
            const float b2 = __fmul_rn(__ll2float_rn(warp_sum_exact(sb2)), FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nextx = __fadd_rn(nextx, dx); nexty = __fadd_rn(nexty, dy);
            stored.x = __fadd_rn(nextx, halfWin); stored.y = __fadd_rn(nexty, halfWin);
            if (__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= a.eps2) { ++j; break; }
            if (j > 0 && fabs((double)__fadd_rn(dx, pdx)) < 0.01 && fabs((double)__fadd_rn(dy, pdy)) < 0.01) {
                stored.x = __fsub_rn(stored.x, __fmul_rn(dx, 0.5f));
                stored.y = __fsub_rn(stored.y, __fmul_rn(dy, 0.5f));
                ++j;
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (a.counters && lane == 0) atomicAdd(a.counters + 2 * level + 1, (unsigned long long)j);

        if (status && level == 0) {
            const float npx = __fsub_rn(stored.x, halfWin), npy = __fsub_rn(stored.y, halfWin);
            const int inx = __float2int_rd(npx), iny = __float2int_rd(npy);
            if (inx < -win || inx >= J.w || iny < -win || iny >= J.h) {
                status = 0;
            } else {
                bilinear_weights(__fsub_rn(npx, (float)inx), __fsub_rn(npy, (float)iny), iw00, iw01, iw10, iw11);
                const uint8_t *jb = J.img + (ptrdiff_t)iny * J.pitch + inx;
                int sabs = 0;
#pragma unroll
                for (int k = 0; k < NPX; ++k) {
                    const bool ok = (lane + 32 * k) < npix;
                    const int o = ok ? py[k] * J.pitch + px[k] : 0;
                    const int s00 = __ldg(jb + o), s01 = __ldg(jb + o + 1);
                    const int s10 = __ldg(jb + o + J.pitch), s11 = __ldg(jb + o + J.pitch + 1);
                    const int diff = ((s00 * iw00 + s01 * iw01 + s10 * iw10 + s11 * iw11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5)) - Iv[k];
                    sabs += ok ? abs(diff) : 0;
                }
                const int tot = __reduce_add_sync(0xffffffffu, sabs);   // <= 961*8160 < 2^31
                errv = __fdiv_rn(__fmul_rn((float)tot, 1.f), (float)(32 * win * win));
            }
        }
    }

    if (lane == 0) klt_epilogue(a, gi, S0, stored, status, errv);
}


// =======================================================================================
// v3: TMA-staged windows + dp2a sampling (every odd window 9 ... 31).
//
// ncu on v1 (profiles/r1_v1_*) showed the kernel bound by L1 wavefronts: 229 M scattered byte gathers per launch.
// v2 (round 1) staged the template patch, its Scharr patch and a search region of the next image in per-warp shared
// memory with LDG.128 -> STS through registers; profiles/r1_v4_* showed it instruction-issue bound, with a third of the
// per-(feature, level) set-up being the staging loop's address arithmetic (172 IMAD/IADD3/LEA + 68 STS + 23 LDG).
// v3 hands the staging to the TMA unit: three cp.async.bulk.tensor.2d boxes per (feature, level) -- template patch
// (u8), Scharr patch (short2 as 32-bit elements), search region (u8) -- issued by one lane against per-(slot, level)
// tensor maps, landing on two per-warp mbarriers:
//   * the staging needs no address arithmetic, no registers and no STS: three box coordinates per request;
//   * the NEXT level's template / Scharr boxes are requested as soon as the current level's template registers are
//     built, the current level's search region right at the level start: both land while the warp computes.
// A TMA box must start on a 16-byte boundary of its innermost dimension (measured on this B200 / driver 13.0 with
// tools/probe/tma_probe3.cu: the programming guide's own example faults with "illegal instruction" at x = 65 int32
// elements and runs at x = 64), so box origins are rounded DOWN to 16 pixels (u8 planes) / 4 elements (short2 plane)
// and the window sits at byte offset ipx & 15 inside the staged rows -- the funnel-shift realignment of load_run absorbs it.
// Box row pitches (48 / 80 B image rows, 28 / 36-element Scharr rows) are the dense TMA layout; they are chosen so that
// r * pitch mod 32 banks has period 8, which keeps the row-segment reads of the iterations at most 2-way conflicted.
// Each lane owns RPL horizontal runs of RL pixels; a run reads 3 words from each of 2 rows, realigns them with funnel
// shifts and evaluates the fixed-point bilinear sample of a pixel with two dp2a (s16 weight pair x u8 pixel pair) --
// bit-identical to the scalar formula.  The search region is re-requested only if the window drifts outside its margin.
// =======================================================================================
#define MJ 5    // minimum search-region margin (pixels) above / below the window

template <int WIN> struct Klt3Cfg {
    static constexpr int W1 = WIN + 1;
    static constexpr int SEG = (WIN > 24) ? 4 : (WIN > 16) ? 3 : 2;       // runs per window row
    static constexpr int RL = (WIN + SEG - 1) / SEG;                     // pixels per run (<= 8)
    static constexpr int NRUN = WIN * SEG;
    static constexpr int RPL = (NRUN + 31) / 32;                         // runs per lane
    static constexpr bool FULL = (WIN % SEG) == 0;                       // every run has RL pixels
    static constexpr int MISS = SEG * RL - WIN;                          // pixels the LAST run of a row is short of RL
    static constexpr int NPX = RPL * RL;                                 // template registers per lane
    // TMA boxes (dense rows in shared memory); origins are 16-byte aligned, so a row holds up to 15 (u8) / 3 (short2) lead-in
    // elements before the window
    static constexpr int JBW = (W1 + MJ + 15 <= 48) ? 48 : 80;           // search region: bytes per row (12 / 20 words)
    static constexpr int IBW = JBW;                                      // template patch: fetched with the search-region box
    static constexpr int DBW = (SEG * RL + 4 <= 28) ? 28 : 36;           // Scharr patch: short2 elements per row
    static constexpr int JR = W1 + 2 * MJ;                               // search region rows
    static constexpr int IPW = IBW / 4, DPW = DBW, JPW = JBW / 4;        // pitches in 32-bit words
    static constexpr int I_BYTES = JR * IBW;                             // the template patch is fetched with the search-region box
    static constexpr int D_BYTES = W1 * DBW * 4, J_BYTES = JR * JBW;
    static constexpr int OFF_I = 0;
    static constexpr int OFF_D = (I_BYTES + 127) & ~127;
    static constexpr int OFF_J = OFF_D + ((D_BYTES + 127) & ~127);
    static constexpr int WARP_BYTES = OFF_J + ((J_BYTES + 16 + 127) & ~127);   // +16: load_run may touch one word past a row
    static constexpr int MINB = NPX <= 14 ? 5 : (NPX <= 24 ? 3 : 2);     // resident CTAs per SM the register budget is sized for
    static_assert(RL <= 8, "load_run covers at most 8 pixels");
    static_assert(W1 + MJ + 15 <= JBW && SEG * RL + 4 <= DBW && IBW == JBW, "boxes too small for this window");
    static_assert(((15 + (SEG - 1) * RL) >> 2) + 3 <= IPW, "template run reads leave the staged row");
};

__device__ __forceinline__ int dp2a_lo_su(int a, unsigned b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(int a, unsigned b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// TMA descriptors of one launch, passed BY VALUE as a __grid_constant__ kernel parameter (the canonical path: no
// descriptor fetch from global memory, no proxy fence).  All slots of a context live in one allocation with a uniform
// stride, so level l of every slot is ONE rank-3 tensor (x, y, slot): [l] image plane, box JBW x JR x 1 (search region; the
// template patch uses the same box and ignores the extra rows), [l] Scharr plane, box DBW x W1 x 1.
struct KltMaps {
    CUtensorMap img[VO_MAX_LEVELS];
    CUtensorMap der[VO_MAX_LEVELS];
};

// ---- TMA / mbarrier plumbing (one lane issues, the warp waits)
struct WarpPipe {
    uint32_t bar_id, bar_j;     // shared-window addresses of the two mbarriers (template + Scharr boxes / search region)
    uint32_t ph_id, ph_j;       // phase parity to wait for next
};
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t phase)
{
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
    return ok != 0;
}
// All lanes wait for the phase; a transfer that never completes (a broken tensor map) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t &phase)
{
    int spins = 0;
    while (!mbar_try_wait(bar, phase)) { if (++spins > (1 << 24)) __trap(); }
    phase ^= 1u;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((unsigned long long)tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// 8-bit sample stream of one run: E holds bytes e0.., O the same shifted by one byte.  A run of RL pixels needs bytes
// 0 .. RL: with RL <= 7 the third word only feeds the funnel shift of E[1] and O[1] is a plain shift.
template <int RL>
__device__ __forceinline__ void load_run(const uint32_t *row, int byte0, uint32_t E[3], uint32_t O[3])
{
    const uint32_t *p = row + (byte0 >> 2);
    const int sh = (byte0 & 3) * 8;
    const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
    E[0] = __funnelshift_r(w0, w1, sh);
    E[1] = __funnelshift_r(w1, w2, sh);
    O[0] = __funnelshift_r(E[0], E[1], 8);
    if (RL <= 7) {
        E[2] = 0; O[1] = E[1] >> 8; O[2] = 0;
    } else {
        E[2] = w2 >> sh;
        O[1] = __funnelshift_r(E[1], E[2], 8);
        O[2] = E[2] >> 8;
    }
}

// fixed-point bilinear sample of pixel k of a run, accumulated onto `acc`:
// rows A (weights wA = {iw00,iw01}) and B (wB = {iw10,iw11})
#define KLT2_SAMPLE(k, EA, OA, EB, OB, wA, wB, acc)                                                   \
    (((k) & 3) == 0 ? dp2a_lo_su(wB, EB[(k) >> 2], dp2a_lo_su(wA, EA[(k) >> 2], acc))                 \
   : ((k) & 3) == 1 ? dp2a_lo_su(wB, OB[(k) >> 2], dp2a_lo_su(wA, OA[(k) >> 2], acc))                 \
   : ((k) & 3) == 2 ? dp2a_hi_su(wB, EB[(k) >> 2], dp2a_hi_su(wA, EA[(k) >> 2], acc))                 \
                    : dp2a_hi_su(wB, OB[(k) >> 2], dp2a_hi_su(wA, OA[(k) >> 2], acc)))

#define KLT2_RND (1 << (W_BITS - 5 - 1))
#define KLT2_SH (W_BITS - 5)

#ifndef KLT2_WPB
#define KLT2_WPB 4          // warps (features) per CTA
#endif

// Template-window origin of a level (depends on pts0 only): false when the window misses the image (level skipped).
__device__ __forceinline__ bool klt_tpl_origin(const float2 p0, const int level, const float halfWin, const int win, const int w, const int h,
                                               float &prevx, float &prevy, int &ipx, int &ipy)
{
    const float sc = 1.f / (float)(1 << level);
    prevx = __fsub_rn(__fmul_rn(p0.x, sc), halfWin); prevy = __fsub_rn(__fmul_rn(p0.y, sc), halfWin);
    ipx = __float2int_rd(prevx); ipy = __float2int_rd(prevy);
    return !(ipx < -win || ipx >= w || ipy < -win || ipy >= h);
}

// One (pair, feature) on one warp: every level, every iteration, the fused post-filter epilogue.
// wbuf: this warp's 128-byte aligned staging area (Klt3Cfg<WIN>::WARP_BYTES); pp: its two mbarriers.
template <int WIN>
__device__ __forceinline__ void klt3_feature(const KltArgs &a, const KltMaps &maps, const int sid0, const int sid1, const int pair, const int f,
                                             uint8_t *wbuf, WarpPipe &pp, const int lane)
{
    using C = Klt3Cfg<WIN>;
    const size_t gi = (size_t)pair * a.n + f;
    if (a.skip_mask && !a.skip_mask[gi]) return;
    uint32_t *Ibuf = reinterpret_cast<uint32_t *>(wbuf + C::OFF_I);
    uint32_t *Dbuf = reinterpret_cast<uint32_t *>(wbuf + C::OFF_D);
    uint32_t *Jbuf = reinterpret_cast<uint32_t *>(wbuf + C::OFF_J);
    const uint32_t sI = smem_addr(Ibuf), sD = smem_addr(Dbuf), sJ = smem_addr(Jbuf);
    const SlotDesc &S0 = a.slots[sid0];
    const SlotDesc &S1 = a.slots[sid1];
    const float halfWin = (float)(WIN - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    const float eps2_lo = (float)a.eps2 * 0.9999f, eps2_hi = (float)a.eps2 * 1.0001f;

    // run geometry of this lane (level independent). Idle slots alias an existing run; their
    // contributions are dropped per RUN (run_ok), never per pixel.
    int run_row[C::RPL], run_x0[C::RPL], run_len[C::RPL];
    bool run_ok[C::RPL];
#pragma unroll
    for (int q = 0; q < C::RPL; ++q) {
        const int id = lane + 32 * q;
        run_ok[q] = id < C::NRUN;
        const int idm = run_ok[q] ? id : id - C::NRUN;
        const int r = idm / C::SEG, sgm = idm - r * C::SEG;
        run_row[q] = r;
        run_x0[q] = sgm * C::RL;
        run_len[q] = min(C::RL, WIN - sgm * C::RL);               // pixels of this run inside the window (RL - MISS for a row's last run)
    }

    const float2 p0 = a.pts0[gi];
    float2 stored = (a.flags & VO_KLT_USE_INITIAL_FLOW) ? a.pts1[gi] : p0;
    int status = 1;
    float errv = 0.f;

    // request the top level's template + Scharr boxes
    {
        float px_, py_; int ix_, iy_;
        const LevelDesc I = S0.lv[a.top_level];
        if (klt_tpl_origin(p0, a.top_level, halfWin, WIN, I.w, I.h, px_, py_, ix_, iy_) && lane == 0) {
            mbar_expect_tx(pp.bar_id, C::I_BYTES + C::D_BYTES);
            tma_load_3d(sI, &maps.img[a.top_level], pp.bar_id, (ix_ & ~15) + VO_PAD, iy_ + VO_PAD, sid0);
            tma_load_3d(sD, &maps.der[a.top_level], pp.bar_id, (ix_ & ~3) + VO_PAD, iy_ + VO_PAD, sid0);
        }
    }

    for (int level = a.top_level; level >= 0; --level) {
        const LevelDesc I = S0.lv[level];
        const LevelDesc J = S1.lv[level];
        const float sc = 1.f / (float)(1 << level);
        if (level == a.top_level) {
            if (a.flags & VO_KLT_USE_INITIAL_FLOW) { stored.x = __fmul_rn(stored.x, sc); stored.y = __fmul_rn(stored.y, sc); }
            else { stored.x = __fmul_rn(p0.x, sc); stored.y = __fmul_rn(p0.y, sc); }
        } else {
            stored.x = __fmul_rn(stored.x, 2.f); stored.y = __fmul_rn(stored.y, 2.f);
        }
        float nextx = stored.x, nexty = stored.y;
        float prevx, prevy;
        int ipx, ipy;
        if (!klt_tpl_origin(p0, level, halfWin, WIN, I.w, I.h, prevx, prevy, ipx, ipy)) {
            // nothing was requested for this level
            if (level == 0) { status = 0; errv = 0.f; }
            continue;
        }
        int iw00, iw01, iw10, iw11;
        bilinear_weights(__fsub_rn(prevx, (float)ipx), __fsub_rn(prevy, (float)ipy), iw00, iw01, iw10, iw11);

        // ---- request the search region around the start position (lands while the template is built)
        nextx = __fsub_rn(nextx, halfWin); nexty = __fsub_rn(nexty, halfWin);
        int jx0, jy0;   // origin (pixel coords) of the staged search region; jx0 is 16-aligned
        {
            int inx = __float2int_rd(nextx), iny = __float2int_rd(nexty);
            // clamp the staging origin so that it stays near the padded plane even for a wild start
            inx = max(-WIN, min(inx, J.w - 1)); iny = max(-WIN, min(iny, J.h - 1));
            jx0 = (inx - MJ) & ~15; jy0 = iny - MJ;
            if (lane == 0) {
                mbar_expect_tx(pp.bar_j, C::J_BYTES);
                tma_load_3d(sJ, &maps.img[level], pp.bar_j, jx0 + VO_PAD, jy0 + VO_PAD, sid1);
            }
        }
        mbar_wait(pp.bar_id, pp.ph_id);        // template + Scharr boxes of this level (requested one level earlier)

        // ---- template. Registers per pixel: Cn = RND - (I << SH) (the dp2a accumulator seed, so that an
        // iteration's diff is one shift after the two dp2a), Ix, Iy. Exact A sums.
        int Cn[C::NPX], Ix[C::NPX], Iy[C::NPX];
        int sA11 = 0, sA12 = 0, sA22 = 0;
        {
            const int wA = (iw00 & 0xffff) | (iw01 << 16), wB = (iw10 & 0xffff) | (iw11 << 16);
            const int offI = ipx & 15, offD = ipx & 3;                    // the boxes start at ipx rounded down to 16 bytes
#pragma unroll
            for (int q = 0; q < C::RPL; ++q) {
                uint32_t EA[3], OA[3], EB[3], OB[3];
                const int r = run_row[q];
                load_run<C::RL>(Ibuf + r * C::IPW, offI + run_x0[q], EA, OA);
                load_run<C::RL>(Ibuf + (r + 1) * C::IPW, offI + run_x0[q], EB, OB);
                const uint32_t *d0 = Dbuf + r * C::DPW + offD + run_x0[q];
                const uint32_t *d1 = d0 + C::DPW;
                uint32_t da = d0[0], db = d1[0];
                int a11 = 0, a12 = 0, a22 = 0;
#pragma unroll
                for (int k = 0; k < C::RL; ++k) {
                    const uint32_t da1 = d0[k + 1], db1 = d1[k + 1];
                    const int iv = KLT2_SAMPLE(k, EA, OA, EB, OB, wA, wB, KLT2_RND) >> KLT2_SH;
                    int ix = ((int)(short)(da & 0xffff) * iw00 + (int)(short)(da1 & 0xffff) * iw01 + (int)(short)(db & 0xffff) * iw10 +
                              (int)(short)(db1 & 0xffff) * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
                    int iy = (((int)da >> 16) * iw00 + ((int)da1 >> 16) * iw01 + ((int)db >> 16) * iw10 + ((int)db1 >> 16) * iw11 +
                              (1 << (W_BITS - 1))) >> W_BITS;
                    if (!C::FULL && k >= C::RL - C::MISS && k >= run_len[q]) { ix = 0; iy = 0; }
                    Cn[q * C::RL + k] = KLT2_RND - (iv << KLT2_SH);
                    Ix[q * C::RL + k] = ix; Iy[q * C::RL + k] = iy;
                    a11 += ix * ix; a12 += ix * iy; a22 += iy * iy;
                    da = da1; db = db1;
                }
                sA11 += run_ok[q] ? a11 : 0; sA12 += run_ok[q] ? a12 : 0; sA22 += run_ok[q] ? a22 : 0;
            }
        }
        // ---- the template / Scharr buffers are free again: request the NEXT level's boxes now
        __syncwarp();
        if (level > 0) {
            float px_, py_; int ix_, iy_;
            const LevelDesc In = S0.lv[level - 1];
            if (klt_tpl_origin(p0, level - 1, halfWin, WIN, In.w, In.h, px_, py_, ix_, iy_) && lane == 0) {
                mbar_expect_tx(pp.bar_id, C::I_BYTES + C::D_BYTES);
                tma_load_3d(sI, &maps.img[level - 1], pp.bar_id, (ix_ & ~15) + VO_PAD, iy_ + VO_PAD, sid0);
                tma_load_3d(sD, &maps.der[level - 1], pp.bar_id, (ix_ & ~3) + VO_PAD, iy_ + VO_PAD, sid0);
            }
        }
        const float A11 = __fmul_rn(warp_sum_exact_f(sA11), FLT_SCALE);
        const float A12 = __fmul_rn(warp_sum_exact_f(sA12), FLT_SCALE);
        const float A22 = __fmul_rn(warp_sum_exact_f(sA22), FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dif = __fsub_rn(A11, A22);
        const float minEig = __fdiv_rn(
            __fsub_rn(__fadd_rn(A22, A11),
                      __fsqrt_rn(__fadd_rn(__fmul_rn(dif, dif), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
            (float)(2 * WIN * WIN));
        if (a.counters && lane == 0) atomicAdd(a.counters + 2 * level, 1ull);
        mbar_wait(pp.bar_j, pp.ph_j);          // search region of this level
        if (minEig < a.min_eig || D < FLT_EPSILON) {
            if (level == 0) status = 0;
            __syncwarp();
            continue;
        }
        // b = sum * 2^-20 feeds only delta = (A12 b2 - A22 b1) * D: a power-of-two factor commutes with every rounding on the way
        // (no under/overflow at these magnitudes), so it is folded into D once per level
        D = __fmul_rn(__fdiv_rn(1.f, D), FLT_SCALE);

        float pdx = 0.f, pdy = 0.f;
        int j = 0;
        for (; j < a.max_count; ++j) {
            const int inx = __float2int_rd(nextx), iny = __float2int_rd(nexty);
            if (inx < -WIN || inx >= J.w || iny < -WIN || iny >= J.h) {
                if (level == 0) status = 0;
                break;
            }
            int offx = inx - jx0, offy = iny - jy0;
            if (offx < 0 || offx + C::W1 > C::JBW || offy < 0 || offy + C::W1 > C::JR) {
                __syncwarp();
                jx0 = (inx - MJ) & ~15; jy0 = iny - MJ;
                if (lane == 0) {
                    mbar_expect_tx(pp.bar_j, C::J_BYTES);
                    tma_load_3d(sJ, &maps.img[level], pp.bar_j, jx0 + VO_PAD, jy0 + VO_PAD, sid1);
                }
                mbar_wait(pp.bar_j, pp.ph_j);
                offx = inx - jx0; offy = iny - jy0;
            }
            bilinear_weights(__fsub_rn(nextx, (float)inx), __fsub_rn(nexty, (float)iny), iw00, iw01, iw10, iw11);
            const int wA = (iw00 & 0xffff) | (iw01 << 16), wB = (iw10 & 0xffff) | (iw11 << 16);
            int sb1 = 0, sb2 = 0;
#pragma unroll
            for (int q = 0; q < C::RPL; ++q) {
                uint32_t EA[3], OA[3], EB[3], OB[3];
                const uint32_t *rowA = Jbuf + (offy + run_row[q]) * C::JPW;
                load_run<C::RL>(rowA, offx + run_x0[q], EA, OA);
                load_run<C::RL>(rowA + C::JPW, offx + run_x0[q], EB, OB);
                int b1q = 0, b2q = 0;
#pragma unroll
                for (int k = 0; k < C::RL; ++k) {
                    const int diff = KLT2_SAMPLE(k, EA, OA, EB, OB, wA, wB, Cn[q * C::RL + k]) >> KLT2_SH;
                    b1q += diff * Ix[q * C::RL + k];     // Ix = Iy = 0 on the pixels a short run does not have
                    b2q += diff * Iy[q * C::RL + k];
                }
                sb1 += run_ok[q] ? b1q : 0; sb2 += run_ok[q] ? b2q : 0;
            }
            const float b1 = warp_sum_exact_f(sb1);        // x 2^20 (the scale lives in D)
            const float b2 = warp_sum_exact_f(sb2);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nextx = __fadd_rn(nextx, dx); nexty = __fadd_rn(nexty, dy);
            stored.x = __fadd_rn(nextx, halfWin); stored.y = __fadd_rn(nexty, halfWin);
            if (klt_eps_reached(dx, dy, a.eps2, eps2_lo, eps2_hi)) { ++j; break; }
            if (j > 0 && fabsf(__fadd_rn(dx, pdx)) <= 0.01f && fabsf(__fadd_rn(dy, pdy)) <= 0.01f) {
                stored.x = __fsub_rn(stored.x, __fmul_rn(dx, 0.5f));
                stored.y = __fsub_rn(stored.y, __fmul_rn(dy, 0.5f));
                ++j;
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (a.counters && lane == 0) atomicAdd(a.counters + 2 * level + 1, (unsigned long long)j);

        if (status && level == 0) {
            const float npx = __fsub_rn(stored.x, halfWin), npy = __fsub_rn(stored.y, halfWin);
            const int inx = __float2int_rd(npx), iny = __float2int_rd(npy);
            if (inx < -WIN || inx >= J.w || iny < -WIN || iny >= J.h) {
                status = 0;
            } else {
                int offx = inx - jx0, offy = iny - jy0;
                if (offx < 0 || offx + C::W1 > C::JBW || offy < 0 || offy + C::W1 > C::JR) {
                    __syncwarp();
                    jx0 = (inx - MJ) & ~15; jy0 = iny - MJ;
                    if (lane == 0) {
                        mbar_expect_tx(pp.bar_j, C::J_BYTES);
                        tma_load_3d(sJ, &maps.img[level], pp.bar_j, jx0 + VO_PAD, jy0 + VO_PAD, sid1);
                    }
                    mbar_wait(pp.bar_j, pp.ph_j);
                    offx = inx - jx0; offy = iny - jy0;
                }
                bilinear_weights(__fsub_rn(npx, (float)inx), __fsub_rn(npy, (float)iny), iw00, iw01, iw10, iw11);
                const int wA = (iw00 & 0xffff) | (iw01 << 16), wB = (iw10 & 0xffff) | (iw11 << 16);
                int sabs = 0;
#pragma unroll
                for (int q = 0; q < C::RPL; ++q) {
                    uint32_t EA[3], OA[3], EB[3], OB[3];
                    const uint32_t *rowA = Jbuf + (offy + run_row[q]) * C::JPW;
                    load_run<C::RL>(rowA, offx + run_x0[q], EA, OA);
                    load_run<C::RL>(rowA + C::JPW, offx + run_x0[q], EB, OB);
                    int sq = 0;
#pragma unroll
                    for (int k = 0; k < C::RL; ++k) {
                        const int diff = KLT2_SAMPLE(k, EA, OA, EB, OB, wA, wB, Cn[q * C::RL + k]) >> KLT2_SH;
                        sq += (!C::FULL && k >= C::RL - C::MISS && k >= run_len[q]) ? 0 : abs(diff);
                    }
                    sabs += run_ok[q] ? sq : 0;
                }
                const int tot = __reduce_add_sync(0xffffffffu, sabs);
                errv = __fdiv_rn(__fmul_rn((float)tot, 1.f), (float)(32 * WIN * WIN));
            }
        }
        __syncwarp();   // the next level's search-region request overwrites the buffer
    }

    if (lane == 0) klt_epilogue(a, gi, S0, stored, status, errv);
}

// per-warp staging area + mbarriers of a CTA of KLT2_WPB warps
template <int WIN>
__device__ __forceinline__ uint8_t *klt3_warp_setup(uint8_t *smem_raw, unsigned long long *bars, const int wib, const int lane, WarpPipe &pp)
{
    using C = Klt3Cfg<WIN>;
    uint8_t *base = smem_raw + ((128u - (smem_addr(smem_raw) & 127u)) & 127u);     // TMA destinations: 128-byte aligned
    pp.bar_id = smem_addr(bars + 2 * wib);
    pp.bar_j = smem_addr(bars + 2 * wib + 1);
    pp.ph_id = 0; pp.ph_j = 0;
    if (lane == 0) {
        mbar_init(pp.bar_id, 1);
        mbar_init(pp.bar_j, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    return base + (size_t)wib * C::WARP_BYTES;
}
#define KLT3_SMEM(WIN) ((size_t)KLT2_WPB * Klt3Cfg<WIN>::WARP_BYTES + 128)

template <int WIN>
__global__ void __launch_bounds__(32 * KLT2_WPB, Klt3Cfg<WIN>::MINB * 4 / KLT2_WPB)
k_klt3(const KltArgs a, const KltBatchIds ids, const __grid_constant__ KltMaps maps)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(8) unsigned long long bars[2 * KLT2_WPB];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (f >= a.n) return;
    WarpPipe pp;
    uint8_t *wbuf = klt3_warp_setup<WIN>(smem_raw, bars, wib, lane, pp);
    klt3_feature<WIN>(a, maps, ids.s0.id[blockIdx.y], ids.s1.id[blockIdx.y], blockIdx.y, f, wbuf, pp, lane);
}

// Fused tracking chain of the stereo frame step (stereo_vo.cpp:533-571): trackWithPrior(l0 -> l1), trackWithScale,
// trackWithPrior(l1 -> r1) for ONE feature on ONE warp, back to back.  The three stages of a feature depend on each
// other, but features do not depend on each other: as three launches every stage waits for its slowest feature (a
// non-converging feature runs 30 iterations x 4 levels ~ 70 us while the median one needs a few), as one launch only
// the sum over one feature's stages matters.  Results are identical to the three-launch sequence.
#define VO_CHAIN_BACK 1     // second LK pass (the backward pass of a bidirectional track)
#define VO_CHAIN_SCALE 2    // trackWithScale
#define VO_CHAIN_NEXT 4     // third LK pass (l1 -> r1)
template <int WIN>
__global__ void __launch_bounds__(32 * KLT2_WPB, (Klt3Cfg<WIN>::MINB > 4 ? 4 : Klt3Cfg<WIN>::MINB) * 4 / KLT2_WPB)
k_track_chain(const KltArgs a1, const KltArgs a2, const KltScaleArgs sc, const KltArgs a3, const int stages, const __grid_constant__ KltMaps maps)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(8) unsigned long long bars[2 * KLT2_WPB];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (f >= a1.n) return;
    WarpPipe pp;
    uint8_t *wbuf = klt3_warp_setup<WIN>(smem_raw, bars, wib, lane, pp);
    klt3_feature<WIN>(a1, maps, a1.sid0, a1.sid1, 0, f, wbuf, pp, lane);
    __syncwarp();                          // lane 0's point / status / mask stores are visible to the whole warp
    if (stages & VO_CHAIN_BACK) { klt3_feature<WIN>(a2, maps, a2.sid0, a2.sid1, 0, f, wbuf, pp, lane); __syncwarp(); }
    if (stages & VO_CHAIN_SCALE) { klt_scale_feature(sc, f, lane); __syncwarp(); }
    if (stages & VO_CHAIN_NEXT) klt3_feature<WIN>(a3, maps, a3.sid0, a3.sid1, 0, f, wbuf, pp, lane);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// every odd window 9 ... 31 has a k_klt3 / k_track_chain instantiation; other sizes (even, < 9) run the v1 kernel
static bool klt3_window(int win) { return win >= 9 && win <= 31 && (win & 1); }
#define KLT3_FOR_EACH_WIN(X) X(9) X(11) X(13) X(15) X(17) X(19) X(21) X(23) X(25) X(27) X(29) X(31)

// run-time copy of the Klt3Cfg box geometry (static_asserts below keep the two in step)
struct Klt3Boxes { int w1, ibw, dbw, jbw, jr; };
static Klt3Boxes klt3_boxes(int win)
{
    const int seg = win > 24 ? 4 : (win > 16 ? 3 : 2), rl = (win + seg - 1) / seg;
    Klt3Boxes b;
    b.w1 = win + 1; b.jbw = (win + 1 + MJ + 15 <= 48) ? 48 : 80; b.ibw = b.jbw; b.dbw = (seg * rl + 4 <= 28) ? 28 : 36; b.jr = win + 1 + 2 * MJ;
    return b;
}
#define KLT3_CHECK_BOXES(W) static_assert(Klt3Cfg<W>::JBW == ((W + 1 + MJ + 15 <= 48) ? 48 : 80) && Klt3Cfg<W>::JR == W + 1 + 2 * MJ && \
                                          Klt3Cfg<W>::DBW == ((((W > 24 ? 4 : (W > 16 ? 3 : 2)) * Klt3Cfg<W>::RL + 4) <= 28) ? 28 : 36), "klt3_boxes out of step");
KLT3_FOR_EACH_WIN(KLT3_CHECK_BOXES)

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled tmap_encoder()
{
    static PFN_tmapEncodeTiled fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }();
    return fn;
}

// TMA descriptor set for images of w x h in this context and window size `win`, encoded once and cached.  Per level:
// image planes of ALL slots as one rank-3 u8 tensor {pitch, h + 2 PAD, n_slots} (strides pitch, slot stride), box JBW x JR x 1;
// Scharr planes as a rank-3 tensor of 32-bit elements (short2), box DBW x W1 x 1.  The tensors cover the whole padded
// planes, so any window origin the kernel can produce is a valid (possibly partly out-of-bounds = zero-filled, never
// used) box coordinate.
struct KltMapKey {
    int w, h, win;
    bool operator<(const KltMapKey &o) const { return w != o.w ? w < o.w : (h != o.h ? h < o.h : win < o.win); }
};
typedef std::map<KltMapKey, KltMaps> KltMapCache;

void vo_klt_maps_free(vo_ctx *ctx)
{
    delete static_cast<KltMapCache *>(ctx->klt_maps);
    ctx->klt_maps = nullptr;
}

static int get_klt_maps(vo_ctx *ctx, int slot, int win, const KltMaps **out)
{
    const Slot &S = ctx->slots[slot];
    if (!ctx->klt_maps) ctx->klt_maps = new KltMapCache();
    KltMapCache &cache = *static_cast<KltMapCache *>(ctx->klt_maps);
    const KltMapKey key{S.w, S.h, win};
    auto it = cache.find(key);
    if (it != cache.end()) { *out = &it->second; return VO_OK; }
    PFN_tmapEncodeTiled enc = tmap_encoder();
    VO_REQUIRE(enc != nullptr, VO_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const Klt3Boxes B = klt3_boxes(win);
    KltMaps m;
    memset(&m, 0, sizeof(m));
    for (int l = 0; l < ctx->max_levels; ++l) {
        const LevelDesc &L = S.desc.lv[l];
        if (!L.img) continue;
        // plane start of this level inside slot 0 (every slot of this geometry has the same layout at the same offsets)
        uint8_t *img_plane0 = ctx->slot_pool + ((L.img - (size_t)VO_PAD * L.pitch - VO_PAD) - S.base);
        uint8_t *der_plane0 = ctx->slot_pool + (reinterpret_cast<uint8_t *>(L.deriv - (size_t)VO_PAD * L.pitch - VO_PAD) - S.base);
        const cuuint64_t rows = (cuuint64_t)L.h + 2 * VO_PAD;
        const cuuint32_t estr[3] = {1, 1, 1};
        const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, rows, (cuuint64_t)ctx->n_slots};
        const cuuint64_t str8[2] = {(cuuint64_t)L.pitch, (cuuint64_t)ctx->slot_stride};
        const cuuint64_t str32[2] = {(cuuint64_t)L.pitch * 4, (cuuint64_t)ctx->slot_stride};
        const cuuint32_t boxJ[3] = {(cuuint32_t)B.jbw, (cuuint32_t)B.jr, 1}, boxD[3] = {(cuuint32_t)B.dbw, (cuuint32_t)B.w1, 1};
        const CUresult r0 = enc(&m.img[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, img_plane0, dims, str8, boxJ, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const CUresult r1 = enc(&m.der[l], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, der_plane0, dims, str32, boxD, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r0 != CUDA_SUCCESS || r1 != CUDA_SUCCESS) {
            ctx->last_error = "cuTensorMapEncodeTiled failed (level " + std::to_string(l) + ", codes " + std::to_string((int)r0) + " " +
                              std::to_string((int)r1) + ")";
            return VO_ERR_CUDA;
        }
    }
    *out = &cache.emplace(key, m).first->second;
    return VO_OK;
}

template <int WIN>
static cudaError_t launch_klt3(const KltArgs &a, const KltBatchIds &ids, const KltMaps &maps, dim3 grd, cudaStream_t st)
{
    grd.x = (grd.x * 4 + KLT2_WPB - 1) / KLT2_WPB;
    k_klt3<WIN><<<grd, 32 * KLT2_WPB, KLT3_SMEM(WIN), st>>>(a, ids, maps);
    return cudaGetLastError();
}
template <int WIN>
static cudaError_t launch_chain(const KltArgs &a1, const KltArgs &a2, const KltScaleArgs &sc, const KltArgs &a3, int stages, int n,
                                const KltMaps &maps, cudaStream_t st)
{
    k_track_chain<WIN><<<vo_div_up(n, KLT2_WPB), 32 * KLT2_WPB, KLT3_SMEM(WIN), st>>>(a1, a2, sc, a3, stages, maps);
    return cudaGetLastError();
}

int vo_klt_launch(vo_ctx *ctx, int n_pairs, const int *slots0, const int *slots1, const float *pts0_d,
                  int n, int win, int max_level, int flags, float *pts1_d, uint8_t *status_d,
                  float *err_d, long long *counters_d, const KltPost *post)
{
    VO_REQUIRE(n >= 0 && n_pairs >= 0, VO_ERR_INVALID_ARG, "negative size");
    if (n == 0 || n_pairs == 0) return VO_OK;
    VO_REQUIRE(win >= 3 && win <= VO_MAX_WIN, VO_ERR_INVALID_ARG, "window size must be in [3, 31]");
    VO_REQUIRE(max_level >= 0, VO_ERR_INVALID_ARG, "max_level < 0");
    for (int i = 0; i < n_pairs; ++i) {
        VO_REQUIRE(slots0[i] >= 0 && slots0[i] < ctx->n_slots && slots1[i] >= 0 && slots1[i] < ctx->n_slots,
                   VO_ERR_INVALID_ARG, "slot id out of range");
        VO_REQUIRE(ctx->slots[slots0[i]].w > 0 && ctx->slots[slots1[i]].w > 0, VO_ERR_INVALID_ARG, "slot has no image");
        VO_REQUIRE(ctx->slots[slots0[i]].w == ctx->slots[slots1[i]].w && ctx->slots[slots0[i]].h == ctx->slots[slots1[i]].h,
                   VO_ERR_SIZE_MISMATCH, "image pair sizes differ");
    }
    const Slot &A = ctx->slots[slots0[0]];
    int eff = vo_effective_max_level(A.w, A.h, win, max_level);
    if (eff > ctx->max_levels - 1) eff = ctx->max_levels - 1;
    int rc = vo_ensure_pyramids(ctx, slots0, n_pairs, eff + 1, 1);
    if (rc) return rc;
    rc = vo_ensure_pyramids(ctx, slots1, n_pairs, eff + 1, 0);
    if (rc) return rc;
    const KltMaps *maps = nullptr;
    if (klt3_window(win)) {
        for (int i = 0; i < n_pairs; ++i)
            VO_REQUIRE(ctx->slots[slots0[i]].w == A.w && ctx->slots[slots0[i]].h == A.h, VO_ERR_SIZE_MISMATCH, "batched pairs must share one image size");
        rc = get_klt_maps(ctx, slots0[0], win, &maps);
        if (rc) return rc;
    }

    const int npx = vo_div_up(win * win, 32);
    for (int c0 = 0; c0 < n_pairs; c0 += VO_IDLIST_MAX) {
        const int nb = n_pairs - c0 < VO_IDLIST_MAX ? n_pairs - c0 : VO_IDLIST_MAX;
        KltArgs a;
        KltBatchIds ids;
        a.slots = ctx->d_slots;
        for (int i = 0; i < nb; ++i) { ids.s0.id[i] = slots0[c0 + i]; ids.s1.id[i] = slots1[c0 + i]; }
        a.sid0 = slots0[c0]; a.sid1 = slots1[c0];
        const size_t off = (size_t)c0 * n;
        a.pts0 = reinterpret_cast<const float2 *>(pts0_d) + off;
        a.pts1 = reinterpret_cast<float2 *>(pts1_d) + off;
        a.status = status_d ? status_d + off : nullptr;
        a.err = err_d ? err_d + off : nullptr;
        a.counters = reinterpret_cast<unsigned long long *>(counters_d);
        a.skip_mask = (post && post->skip_masked && post->mask) ? post->mask + off : nullptr;
        a.n = n; a.win = win; a.top_level = eff; a.flags = flags;
        a.max_count = 30; a.min_eig = 1e-4f; a.eps2 = 0.01 * 0.01;
        if (post) {
            a.post = *post;
            if (a.post.ref_pts) a.post.ref_pts += 2 * off;
            if (a.post.fwd_pts) a.post.fwd_pts += 2 * off;
            if (a.post.fwd_status) a.post.fwd_status += off;
            if (a.post.fwd_err) a.post.fwd_err += off;
            if (a.post.mask) a.post.mask += off;
        } else {
            a.post = KltPost{};
        }
        dim3 grd(vo_div_up(n, 4), nb);
        if (klt3_window(win)) {
            cudaError_t e = cudaErrorInvalidValue;
            switch (win) {
#define KLT3_CASE(W) case W: e = launch_klt3<W>(a, ids, *maps, grd, ctx->stream); break;
                KLT3_FOR_EACH_WIN(KLT3_CASE)
#undef KLT3_CASE
            }
            if (e != cudaSuccess) { ctx->last_error = std::string("k_klt3: ") + cudaGetErrorString(e); return VO_ERR_CUDA; }
        }
        else if (npx <= 6) k_klt<6><<<grd, 128, 0, ctx->stream>>>(a, ids);
        else if (npx <= 8) k_klt<8><<<grd, 128, 0, ctx->stream>>>(a, ids);
        else if (npx <= 10) k_klt<10><<<grd, 128, 0, ctx->stream>>>(a, ids);
        else if (npx <= 14) k_klt<14><<<grd, 128, 0, ctx->stream>>>(a, ids);
        else if (npx <= 20) k_klt<20><<<grd, 128, 0, ctx->stream>>>(a, ids);
        else k_klt<31><<<grd, 128, 0, ctx->stream>>>(a, ids);
        ctx->launches++;
    }
    VO_CUDA(cudaGetLastError());
    return VO_OK;
}

// Fused per-feature chains (one pair, device pointers, asynchronous).  They fall back to separate launches for window
// sizes without a k_klt2 instantiation or when VO_CHAIN_UNFUSED is set (A/B switch).
static void fill_klt_args(vo_ctx *ctx, KltArgs &a, int s0, int s1, const float *pts0_d, float *pts1_d, uint8_t *status_d, float *err_d, int n,
                          int win, int eff, int flags, const KltPost &post)
{
    a.slots = ctx->d_slots;
    a.sid0 = s0; a.sid1 = s1;
    a.pts0 = reinterpret_cast<const float2 *>(pts0_d); a.pts1 = reinterpret_cast<float2 *>(pts1_d);
    a.status = status_d; a.err = err_d; a.counters = nullptr;
    a.skip_mask = (post.skip_masked && post.mask) ? post.mask : nullptr;
    a.n = n; a.win = win; a.top_level = eff; a.flags = flags;
    a.max_count = 30; a.min_eig = 1e-4f; a.eps2 = 0.01 * 0.01;
    a.post = post;
}
static bool chain_unfused(int win) { return !klt3_window(win); }   // even / tiny windows: separate v1 launches
static int launch_chain_win(vo_ctx *ctx, int win, const KltArgs &a1, const KltArgs &a2, const KltScaleArgs &sc, const KltArgs &a3, int stages, int n)
{
    const KltMaps *maps = nullptr;
    int rc = get_klt_maps(ctx, a1.sid0, win, &maps);
    if (rc) return rc;
    cudaError_t e = cudaErrorInvalidValue;
    switch (win) {
#define KLT3_CASE(W) case W: e = launch_chain<W>(a1, a2, sc, a3, stages, n, *maps, ctx->stream); break;
        KLT3_FOR_EACH_WIN(KLT3_CASE)
#undef KLT3_CASE
    }
    ctx->launches++;
    if (e != cudaSuccess) { ctx->last_error = std::string("k_track_chain: ") + cudaGetErrorString(e); return VO_ERR_CUDA; }
    return VO_OK;
}
static int clamp_level(vo_ctx *ctx, int slot, int win, int max_level)
{
    const Slot &A = ctx->slots[slot];
    int eff = vo_effective_max_level(A.w, A.h, win, max_level < 0 ? 0 : max_level);
    return eff > ctx->max_levels - 1 ? ctx->max_levels - 1 : eff;
}

// stereo_vo.cpp:533-571: trackWithPrior(l0 -> l1), trackWithScale, trackWithPrior(l1 -> r1)
int vo_track_chain_launch_d(vo_ctx *ctx, int slot_l0, int slot_l1, int slot_r1, const float *pts_l0_d, float *pts_l1_d, float *pts_r1_d,
                            const float *scale_d, uint8_t *mask_d, int *nan_flag_d, int n, int win, int max_level, float thres_err,
                            int do_scale)
{
    if (n <= 0) return VO_OK;
    const int eff = clamp_level(ctx, slot_l0, win, max_level);
    const int sl[3] = {slot_l0, slot_l1, slot_r1};
    int rc = vo_ensure_pyramids(ctx, sl, 2, eff + 1, 1);          // templates come from l0 and l1
    if (rc) return rc;
    rc = vo_ensure_pyramids(ctx, sl + 2, 1, eff + 1, 0);
    if (rc) return rc;
    KltPost post{};
    post.mode = 2; post.thres_err = thres_err; post.mask = mask_d; post.skip_masked = 1;
    if (chain_unfused(win)) {
        rc = vo_klt_launch(ctx, 1, &slot_l0, &slot_l1, pts_l0_d, n, win, max_level, VO_KLT_USE_INITIAL_FLOW, pts_l1_d, nullptr, nullptr, nullptr, &post);
        if (rc) return rc;
        if (do_scale) { rc = vo_klt_scale_launch_d(ctx, slot_l0, slot_l1, pts_l0_d, scale_d, n, pts_l1_d, mask_d, nan_flag_d); if (rc) return rc; }
        return vo_klt_launch(ctx, 1, &slot_l1, &slot_r1, pts_l1_d, n, win, max_level, VO_KLT_USE_INITIAL_FLOW, pts_r1_d, nullptr, nullptr, nullptr, &post);
    }
    KltArgs a1, a3;
    fill_klt_args(ctx, a1, slot_l0, slot_l1, pts_l0_d, pts_l1_d, nullptr, nullptr, n, win, eff, VO_KLT_USE_INITIAL_FLOW, post);
    fill_klt_args(ctx, a3, slot_l1, slot_r1, pts_l1_d, pts_r1_d, nullptr, nullptr, n, win, eff, VO_KLT_USE_INITIAL_FLOW, post);
    KltScaleArgs sc;
    sc.slots = ctx->d_slots; sc.slot0 = slot_l0; sc.slot1 = slot_l1;
    sc.pts0 = reinterpret_cast<const float2 *>(pts_l0_d); sc.scale = scale_d;
    sc.pts_track = reinterpret_cast<float2 *>(pts_l1_d); sc.mask = mask_d; sc.nan_flag = nan_flag_d; sc.iters = nullptr; sc.n = n;
    if (do_scale && ctx->scale_faithful) {
        // reference-faithful trackWithScale borders: the second pass (k_klt_scale_fixup) must see every feature's scale stage
        // before any feature goes on to l1 -> r1, so the third stage is a launch of its own in this mode
        rc = vo_klt_scale_prepare(ctx, sc);
        if (rc) return rc;
        rc = launch_chain_win(ctx, win, a1, a1, sc, a3, VO_CHAIN_SCALE, n);
        if (rc) return rc;
        rc = vo_klt_scale_fixup_launch(ctx, sc);
        if (rc) return rc;
        return vo_klt_launch(ctx, 1, &slot_l1, &slot_r1, pts_l1_d, n, win, max_level, VO_KLT_USE_INITIAL_FLOW, pts_r1_d, nullptr, nullptr, nullptr, &post);
    }
    return launch_chain_win(ctx, win, a1, a1, sc, a3, (do_scale ? VO_CHAIN_SCALE : 0) | VO_CHAIN_NEXT, n);
}

// Bidirectional track of one pair (feature_tracker.cpp:39-169): forward pass, backward pass seeded with pts0 whose
// epilogue fuses the validity test, optionally followed by trackWithScale (the mono step, mono_vo.cpp:768-783).
// with_prior: both passes start from the given prior and run at max_level (trackBidirectionWithPrior, 5x gate,
// 0-px border); otherwise the forward pass starts at pts0 and the backward pass runs at max_level - 1 (3-px border).
// back_d / st_d / stb_d / err_d / errb_d: scratch [n].  scale_d == nullptr skips the scale stage.
int vo_bidir_chain_launch_d(vo_ctx *ctx, int slot0, int slot1, const float *pts0_d, float *pts1_d, float *back_d, uint8_t *st_d, uint8_t *stb_d,
                            float *err_d, float *errb_d, uint8_t *mask_d, int skip_masked, int n, int win, int max_level, float thres_err,
                            float thres_bi, int with_prior, const float *scale_d, int *nan_flag_d)
{
    if (n <= 0) return VO_OK;
    const int sl[2] = {slot0, slot1};
    const int eff_f = clamp_level(ctx, slot0, win, max_level);
    const int back_lvl = with_prior ? max_level : (max_level - 1 < 0 ? 0 : max_level - 1);
    const int eff_b = clamp_level(ctx, slot0, win, back_lvl);
    int rc = vo_ensure_pyramids(ctx, sl, 2, eff_f + 1, 1);         // both images serve as template
    if (rc) return rc;
    VO_CUDA(cudaMemcpyAsync(back_d, pts0_d, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));   // backward pass starts at pts0
    KltPost pf{}, pb{};
    pf.mode = 3; pf.thres_err = thres_err; pf.mask = mask_d; pf.skip_masked = skip_masked;
    pb = pf;
    pb.mode = 4; pb.border = with_prior ? 0 : 3;
    pb.thres_bi2 = with_prior ? (thres_bi * thres_bi) * 5 : thres_bi * thres_bi;
    pb.ref_pts = pts0_d; pb.fwd_pts = pts1_d; pb.fwd_status = st_d; pb.fwd_err = err_d;
    const int flags_f = with_prior ? VO_KLT_USE_INITIAL_FLOW : 0;
    if (chain_unfused(win)) {
        rc = vo_klt_launch(ctx, 1, &slot0, &slot1, pts0_d, n, win, max_level, flags_f, pts1_d, st_d, err_d, nullptr, &pf);
        if (rc) return rc;
        rc = vo_klt_launch(ctx, 1, &slot1, &slot0, pts1_d, n, win, back_lvl, VO_KLT_USE_INITIAL_FLOW, back_d, stb_d, errb_d, nullptr, &pb);
        if (rc) return rc;
        if (scale_d) return vo_klt_scale_launch_d(ctx, slot0, slot1, pts0_d, scale_d, n, pts1_d, mask_d, nan_flag_d);
        return VO_OK;
    }
    KltArgs a1, a2;
    fill_klt_args(ctx, a1, slot0, slot1, pts0_d, pts1_d, st_d, err_d, n, win, eff_f, flags_f, pf);
    fill_klt_args(ctx, a2, slot1, slot0, pts1_d, back_d, stb_d, errb_d, n, win, eff_b, VO_KLT_USE_INITIAL_FLOW, pb);
    KltScaleArgs sc;
    sc.slots = ctx->d_slots; sc.slot0 = slot0; sc.slot1 = slot1;
    sc.pts0 = reinterpret_cast<const float2 *>(pts0_d); sc.scale = scale_d;
    sc.pts_track = reinterpret_cast<float2 *>(pts1_d); sc.mask = mask_d; sc.nan_flag = nan_flag_d; sc.iters = nullptr; sc.n = n;
    if (scale_d) { rc = vo_klt_scale_prepare(ctx, sc); if (rc) return rc; }
    rc = launch_chain_win(ctx, win, a1, a2, sc, a1, VO_CHAIN_BACK | (scale_d ? VO_CHAIN_SCALE : 0), n);
    if (rc) return rc;
    return scale_d ? vo_klt_scale_fixup_launch(ctx, sc) : VO_OK;        // the scale stage is the last one of this chain
}
