// orb.cu -- K-orb: the keypoints of cv::ORB::detect, bucketed as the reference does.
//
// The reference's extractor is cv::ORB (core/visual_odometry/feature_extractor.cpp:26-60: maxFeatures 10000, scaleFactor 1.2,
// 8 levels, edgeThreshold 31, HARRIS_SCORE, patchSize 31, fastThreshold from the yaml) followed by the WeightBin bucketing of
// extractORBwithBinning_fast (:211-282).  OpenCV is third-party and not under /root/reference; this file restates the
// published algorithm of cv::ORB::detect stage by stage.  Its numpy twin (test infrastructure) is pinned bit-exact against
// the cv2 4.13 wheel: INTER_LINEAR_EXACT pyramid, FAST-9/16 keypoints + scores, and the final keypoint set (octave, point,
// Harris response) of cv2.ORB.detect.
//
//   k_orb_resize   level l from level l-1: cv::resize(INTER_LINEAR_EXACT), 8.8 fixed-point weights from the double-precision
//                  source coordinate, 16.16 vertical pass, round-half-up to u8                       (7 dependent launches)
//   k_orb_fast     all levels in one launch: FAST-9/16 segment test + cornerScore (largest threshold that still passes, -1)
//                  for every pixel that can be, or can suppress, a keypoint inside the 31-px border
//   k_orb_nms      3x3 non-maximum suppression (strict), border filter, per-level 256-bin score histogram
//   k_orb_cut      retainBest(2 * quota) on the FAST score: cutoff score per level from the histogram (ties kept)
//   k_orb_emit     surviving pixels -> per-level keypoint lists
//   k_orb_harris   Harris response of every listed keypoint (orb.cpp HarrisResponses: 7x7 block of integer Sobel-like
//                  gradients, float32 response), one warp per keypoint
//   k_orb_select   retainBest(quota) on the Harris response: exact n-th largest float per level by a 4-pass radix select
//   k_orb_bucket   kept keypoints, scaled to level-0 pixels -> per-bin 64-bit atomicMax of (response, first keypoint)
//   k_orb_out      ordered scan over the bins -> new points in bin order; k_orb_gather: the whole keypoint list (tests)
// Ties: OpenCV's order inside a level is whatever std::nth_element leaves; the reference's strict '<' then keeps the first
// of two bit-equal responses in one bin.  Here the tie goes to the lower (level, y, x).
// Compiled with -fmad=false (float32 Harris response in OpenCV's operation order).
#include "vo_internal.cuh"

#include <cmath>
#include <cstring>

#define ORB_LEVELS 8

namespace {

struct OrbLevel {
    const uint8_t *img;
    uint8_t *wimg;           // writable alias for levels >= 1 (null for level 0)
    uint8_t *score, *nms;    // dense w x h planes
    unsigned *xy;            // keypoint list: y << 16 | x
    float *resp;
    int w, h, pitch, quota, cap, active;
    float scale;
};

struct OrbArgs {
    OrbLevel lv[ORB_LEVELS];
    int n_levels, thr, edge;
    int *hist;               // [levels][256]
    int *cut;                // [levels] FAST score cutoff
    int *count;              // [levels]
    unsigned *cutkey;        // [levels] Harris cutoff (sortable key)
    int *overflow;
    int n_bins_u, n_bins_v, u_step, v_step;
    int *weight;
    unsigned long long *best;
    const float2 *occ;
    const int *n_occ_d;
    int n_occ;
    float2 *out;
    uint8_t *out_mask;
    int *n_out;
    int max_out;
    // full keypoint list (vo_orb_detect)
    float2 *all_pt;
    float *all_resp;
    int *all_octave, *n_all;
    int max_all;
};

__global__ void __launch_bounds__(256) k_orb_reset(const OrbArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ORB_LEVELS * 256) a.hist[i] = 0;
    if (i < ORB_LEVELS) { a.count[i] = 0; a.cut[i] = 256; a.cutkey[i] = 0xFFFFFFFFu; }
    if (i == 0) { *a.overflow = 0; if (a.n_all) *a.n_all = 0; }
    if (i < a.n_bins_u * a.n_bins_v) { a.weight[i] = 1; a.best[i] = 0ull; }
}

// WeightBin::update (feature_extractor.h:119-131)
__global__ void __launch_bounds__(256) k_orb_mark(const OrbArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = a.n_occ_d ? *a.n_occ_d : a.n_occ;
    if (i >= n) return;
    const float2 p = a.occ[i];
    const long long ui = (long long)floorf(__fdiv_rn(p.x, (float)a.u_step));
    const long long vi = (long long)floorf(__fdiv_rn(p.y, (float)a.v_step));
    const long long b = vi * a.n_bins_u + ui;
    if (b >= 0 && b < (long long)a.n_bins_u * a.n_bins_v) a.weight[b] = 0;
}

// interpolationLinear<uchar>::getCoeffs (imgproc resize.cpp): source offset(s) and the 8.8 weight of the second tap
__device__ __forceinline__ void orb_coeff(int d, int src, int dst, int &o0, int &o1, int &c1)
{
    const double scale = 1.0 / ((double)dst / (double)src);
    const double fval = scale * ((double)d + 0.5) - 0.5;
    const int ival = (int)floor(fval);
    if (ival >= 0 && src > 1) {
        if (ival < src - 1) { o0 = ival; o1 = ival + 1; c1 = __double2int_rn((fval - (double)ival) * 256.0); }
        else { o0 = o1 = src - 1; c1 = 0; }
    } else { o0 = o1 = 0; c1 = 0; }
}

__global__ void __launch_bounds__(256) k_orb_resize(const OrbLevel s, const OrbLevel d)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= d.w || y >= d.h) return;
    int xo, xo1, xc, yo, yo1, yc;
    orb_coeff(x, s.w, d.w, xo, xo1, xc);
    orb_coeff(y, s.h, d.h, yo, yo1, yc);
    const uint8_t *r0 = s.img + (size_t)yo * s.pitch, *r1 = s.img + (size_t)yo1 * s.pitch;
    const unsigned h0 = min((unsigned)r0[xo] * (unsigned)(256 - xc) + (unsigned)r0[xo1] * (unsigned)xc, 65535u);      // ufixedpoint16
    const unsigned h1 = min((unsigned)r1[xo] * (unsigned)(256 - xc) + (unsigned)r1[xo1] * (unsigned)xc, 65535u);
    const unsigned v = h0 * (unsigned)(256 - yc) + h1 * (unsigned)yc;                                                  // ufixedpoint32
    d.wimg[(size_t)y * d.pitch + x] = (uint8_t)min((v + 32768u) >> 16, 255u);
}

// FAST-9/16: d[k] = centre - ring[k]; a corner has 9 contiguous ring pixels all darker (d > t) or all brighter (d < -t);
// cornerScore<16> = max(t, best arc minimum of d, best arc minimum of -d) - 1
__global__ void __launch_bounds__(256) k_orb_fast(const OrbArgs a)
{
    const OrbLevel L = a.lv[blockIdx.z];
    if (!L.active) return;
    const int x0 = a.edge - 1, y0 = a.edge - 1;
    const int x = x0 + blockIdx.x * 32 + (threadIdx.x & 31), y = y0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x > L.w - a.edge || y > L.h - a.edge) return;
    const uint8_t *p = L.img + (size_t)y * L.pitch + x;
    const int P = L.pitch, t = a.thr;
    const int c = p[0];
    // any arc of 9 contains at least two of the four compass pixels
    const int d0 = c - p[3 * P], d4 = c - p[3], d8 = c - p[-3 * P], d12 = c - p[-3];
    const int dark = (d0 > t) + (d4 > t) + (d8 > t) + (d12 > t), bright = (d0 < -t) + (d4 < -t) + (d8 < -t) + (d12 < -t);
    int score = 0;
    if (dark >= 2 || bright >= 2) {
        int d[16];
        d[0] = d0; d[1] = c - p[3 * P + 1]; d[2] = c - p[2 * P + 2]; d[3] = c - p[P + 3];
        d[4] = d4; d[5] = c - p[-P + 3]; d[6] = c - p[-2 * P + 2]; d[7] = c - p[-3 * P + 1];
        d[8] = d8; d[9] = c - p[-3 * P - 1]; d[10] = c - p[-2 * P - 2]; d[11] = c - p[-P - 3];
        d[12] = d12; d[13] = c - p[P - 3]; d[14] = c - p[2 * P - 2]; d[15] = c - p[3 * P - 1];
        // best arc minimum of d (ring darker than the centre) and of -d (ring brighter), both as running minima.  (Written
        // as min chains on purpose: tracking min(max(d)) and negating it at the end was mis-compiled by nvcc 12.9 for sm_100a
        // -- the negation got lost in the fused 3-input min/max -- and cost an afternoon of bisecting.)
        int dark_best = -1000, bright_best = -1000;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            int lo = d[k], lo_n = -d[k];
#pragma unroll
            for (int j = 1; j < 9; ++j) { const int v = d[(k + j) & 15]; lo = min(lo, v); lo_n = min(lo_n, -v); }
            dark_best = max(dark_best, lo); bright_best = max(bright_best, lo_n);
        }
        const int best = max(dark_best, bright_best);
        if (best > t) score = best - 1;                 // cornerScore: max(t, best) - 1 for a pixel that passes the test
    }
    L.score[(size_t)y * L.w + x] = (uint8_t)score;
}

__global__ void __launch_bounds__(256) k_orb_nms(const OrbArgs a)
{
    __shared__ int s_hist[256];
    const OrbLevel L = a.lv[blockIdx.z];
    if (!L.active) return;
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int x = a.edge + blockIdx.x * 32 + (threadIdx.x & 31), y = a.edge + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x < L.w - a.edge && y < L.h - a.edge) {                   // KeyPointsFilter::runByImageBorder
        const uint8_t *s = L.score + (size_t)y * L.w + x;
        const int v = s[0];
        int keep = 0;
        if (v > 0) {
            const int W = L.w;
            const int m = max(max(max((int)s[-W - 1], (int)s[-W]), max((int)s[-W + 1], (int)s[-1])),
                              max(max((int)s[1], (int)s[W - 1]), max((int)s[W], (int)s[W + 1])));
            keep = v > m;
        }
        L.nms[(size_t)y * L.w + x] = keep ? (uint8_t)v : 0;
        if (keep) atomicAdd(&s_hist[v], 1);
    }
    __syncthreads();
    const int hv = s_hist[threadIdx.x];
    if (hv) atomicAdd(a.hist + blockIdx.z * 256 + threadIdx.x, hv);
}

// KeyPointsFilter::retainBest(keypoints, 2 * featuresNum) on the FAST score: everything >= the n-th largest.
// One CTA per level: histogram into shared memory, one thread scans it.
__global__ void __launch_bounds__(256) k_orb_cut(const OrbArgs a)
{
    __shared__ int s_h[256];
    const int lv = blockIdx.x;
    if (!a.lv[lv].active) return;
    s_h[threadIdx.x] = a.hist[lv * 256 + threadIdx.x];
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int n = 2 * a.lv[lv].quota;
    int total = 0;
    for (int s = 1; s < 256; ++s) total += s_h[s];
    int cut = 1;
    if (n == 0) cut = 256;
    else if (total > n) {
        int acc = 0;
        for (int s = 255; s >= 1; --s) { acc += s_h[s]; if (acc >= n) { cut = s; break; } }
    }
    a.cut[lv] = cut;
}

// orb.cpp HarrisResponses (blockSize 7, HARRIS_K 0.04): one warp per keypoint, lanes over the 49 block positions
__device__ __forceinline__ float orb_harris_warp(const uint8_t *img, int pitch, int x, int y, int lane)
{
    int sa = 0, sb = 0, sc = 0;
    for (int k = lane; k < 49; k += 32) {
        const int dy = k / 7 - 3, dx = k % 7 - 3;
        const uint8_t *q = img + (size_t)(y + dy) * pitch + x + dx;
        const int Ix = ((int)q[1] - (int)q[-1]) * 2 + ((int)q[-pitch + 1] - (int)q[-pitch - 1]) + ((int)q[pitch + 1] - (int)q[pitch - 1]);
        const int Iy = ((int)q[pitch] - (int)q[-pitch]) * 2 + ((int)q[pitch - 1] - (int)q[-pitch - 1]) + ((int)q[pitch + 1] - (int)q[-pitch + 1]);
        sa += Ix * Ix; sb += Iy * Iy; sc += Ix * Iy;
    }
    sa = __reduce_add_sync(0xffffffffu, sa); sb = __reduce_add_sync(0xffffffffu, sb); sc = __reduce_add_sync(0xffffffffu, sc);
    const float scale = 1.f / ((float)((1 << 2) * 7) * 255.f);
    const float s4 = scale * scale * scale * scale;
    const float fa = (float)sa, fb = (float)sb, fc = (float)sc;
    return ((fa * fb - fc * fc) - (0.04f * (fa + fb)) * (fa + fb)) * s4;
}

__global__ void __launch_bounds__(256) k_orb_emit(const OrbArgs a)
{
    const OrbLevel L = a.lv[blockIdx.z];
    if (!L.active) return;
    const int x = a.edge + blockIdx.x * 32 + (threadIdx.x & 31), y = a.edge + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= L.w - a.edge || y >= L.h - a.edge) return;
    const int v = L.nms[(size_t)y * L.w + x];
    if (v == 0 || v < a.cut[blockIdx.z]) return;
    const int idx = atomicAdd(a.count + blockIdx.z, 1);
    if (idx >= L.cap) { *a.overflow = 1; return; }
    L.xy[idx] = ((unsigned)y << 16) | (unsigned)x;
}

__global__ void __launch_bounds__(256) k_orb_harris(const OrbArgs a)
{
    const int lv = blockIdx.y, lane = threadIdx.x & 31;
    const OrbLevel L = a.lv[lv];
    if (!L.active) return;
    const int n = min(a.count[lv], L.cap);
    for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) {
        const unsigned xy = L.xy[i];
        const float r = orb_harris_warp(L.img, L.pitch, (int)(xy & 0xFFFFu), (int)(xy >> 16), lane);
        if (lane == 0) L.resp[i] = r;
    }
}

__device__ __forceinline__ unsigned orb_key(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// KeyPointsFilter::retainBest(keypoints, featuresNum) on the Harris response: exact n-th largest key (ties kept)
__global__ void __launch_bounds__(1024) k_orb_select(const OrbArgs a)
{
    __shared__ int s_hist[256];
    __shared__ unsigned s_prefix, s_mask;
    __shared__ int s_want;
    const int lv = blockIdx.x, tid = threadIdx.x;
    const OrbLevel L = a.lv[lv];
    if (!L.active) return;
    const int n = min(a.count[lv], L.cap), q = L.quota;
    if (n <= q) { if (tid == 0) a.cutkey[lv] = 0u; return; }               // keep everything
    if (q == 0) { if (tid == 0) a.cutkey[lv] = 0xFFFFFFFFu; return; }      // keypoints.clear() (a key never reaches it: no NaNs)
    if (tid == 0) { s_prefix = 0u; s_mask = 0u; s_want = q; }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        const unsigned prefix = s_prefix, mask = s_mask;
        for (int i = tid; i < n; i += 1024) {
            const unsigned k = orb_key(L.resp[i]);
            if ((k & mask) == prefix) atomicAdd(&s_hist[(k >> shift) & 255u], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0, want = s_want, digit = 0;
            for (int dgt = 255; dgt >= 0; --dgt) {
                if (acc + s_hist[dgt] >= want) { digit = dgt; want -= acc; break; }
                acc += s_hist[dgt];
            }
            s_want = want;
            s_prefix = prefix | ((unsigned)digit << shift);
            s_mask = mask | (0xFFu << shift);
        }
        __syncthreads();
    }
    if (tid == 0) a.cutkey[lv] = s_prefix;
}

// extractORBwithBinning_fast (:246-273): bin of the level-0 point, weight gate, strictly larger response wins
__global__ void __launch_bounds__(256) k_orb_bucket(const OrbArgs a)
{
    const int lv = blockIdx.y;
    const OrbLevel L = a.lv[lv];
    if (!L.active) return;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= min(a.count[lv], L.cap)) return;
    const float r = L.resp[i];
    const unsigned key = orb_key(r);
    if (key < a.cutkey[lv]) return;
    const unsigned xy = L.xy[i];
    const int x = (int)(xy & 0xFFFFu), y = (int)(xy >> 16);
    const float px = __fmul_rn((float)x, L.scale), py = __fmul_rn((float)y, L.scale);          // keypoint.pt *= scale
    if (a.all_pt) {
        const int j = atomicAdd(a.n_all, 1);
        if (j < a.max_all) { a.all_pt[j] = make_float2(px, py); a.all_resp[j] = r; a.all_octave[j] = lv; }
    }
    if (a.n_bins_u <= 0) return;
    const float inv_u = __fdiv_rn(1.0f, (float)a.u_step), inv_v = __fdiv_rn(1.0f, (float)a.v_step);
    const int u = (int)floorf(__fmul_rn(px, inv_u)), v = (int)floorf(__fmul_rn(py, inv_v));
    if (u < 0 || u >= a.n_bins_u || v < 0 || v >= a.n_bins_v) return;
    const int b = v * a.n_bins_u + u;
    if (a.weight[b] == 0 || !(r > -1.0f)) return;                                              // max_score_ starts at -1
    const unsigned tb = ((unsigned)lv << 28) | ((unsigned)y << 14) | (unsigned)x;
    atomicMax(a.best + b, ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - tb));
}

__global__ void __launch_bounds__(1024) k_orb_out(const OrbArgs a)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    const int nb = a.n_bins_u * a.n_bins_v;
    for (int c0 = 0; c0 < nb; c0 += 1024) {
        const int b = c0 + tid;
        const unsigned long long e = b < nb ? a.best[b] : 0ull;
        const bool keep = b < nb && e != 0ull && a.weight[b] > 0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const int within = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        if (wid == 0) {
            int v = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            s_warp[lane] = v;
        }
        __syncthreads();
        const int pos = s_base + (wid ? s_warp[wid - 1] : 0) + within;
        if (keep && pos < a.max_out) {
            const unsigned tb = 0xFFFFFFFFu - (unsigned)(e & 0xFFFFFFFFull);
            const int lv = (int)(tb >> 28), y = (int)((tb >> 14) & 0x3FFFu), x = (int)(tb & 0x3FFFu);
            a.out[pos] = make_float2(__fmul_rn((float)x, a.lv[lv].scale), __fmul_rn((float)y, a.lv[lv].scale));
        }
        __syncthreads();
        if (tid == 0) s_base += s_warp[31];
        __syncthreads();
    }
    const int n = s_base < a.max_out ? s_base : a.max_out;
    if (tid == 0) *a.n_out = n;
    if (a.out_mask)
        for (int i = tid; i < a.max_out; i += 1024) a.out_mask[i] = i < n ? 1 : 0;
}

inline size_t al(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

// Device-pointer launcher (same contract as vo_detect_launch_d).  all_* (nullable): the whole cv::ORB keypoint list.
int vo_orb_launch_d(vo_ctx *ctx, int slot, const float *occ_d, const int *n_occ_d, int n_occ, int n_bins_u, int n_bins_v, int edge,
                    float *out_d, uint8_t *out_mask_d, int *n_out_d, int max_out, float *all_pt_d, float *all_resp_d, int *all_octave_d,
                    int *n_all_d, int max_all)
{
    VO_REQUIRE(slot >= 0 && slot < ctx->n_slots && ctx->slots[slot].w > 0, VO_ERR_INVALID_ARG, "slot has no image");
    VO_REQUIRE(edge >= 4 && edge <= 64, VO_ERR_INVALID_ARG, "bad detector arguments");
    VO_REQUIRE(ctx->orb_fast_threshold >= 1 && ctx->orb_fast_threshold <= 254, VO_ERR_INVALID_ARG, "FAST threshold out of range");
    const Slot &S = ctx->slots[slot];
    VO_REQUIRE(S.w < 16384 && S.h < 16384, VO_ERR_INVALID_ARG, "image too large for the keypoint packing");
    const bool bucket = n_bins_u > 0 && n_bins_v > 0;
    VO_REQUIRE(!bucket || (S.w / n_bins_u >= 1 && S.h / n_bins_v >= 1), VO_ERR_INVALID_ARG, "more bins than pixels");
    int rc = vo_ensure_pyramids(ctx, &slot, 1, 1, 0);
    if (rc) return rc;
    // ---- level geometry and quotas (orb.cpp detectAndCompute / computeKeyPoints), float arithmetic as OpenCV's
    const int nfeatures = 10000, n_levels = ORB_LEVELS;
    const double scale_factor = 1.2;
    OrbArgs a;
    memset(&a, 0, sizeof(a));
    a.n_levels = n_levels; a.thr = ctx->orb_fast_threshold; a.edge = edge;
    {
        const float factor = (float)(1.0 / scale_factor);
        float nd = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)n_levels));
        int sum = 0;
        for (int l = 0; l < n_levels - 1; ++l) { a.lv[l].quota = (int)lrintf(nd); sum += a.lv[l].quota; nd *= factor; }
        a.lv[n_levels - 1].quota = nfeatures - sum > 0 ? nfeatures - sum : 0;
    }
    size_t off = 0, o_img[ORB_LEVELS], o_sc[ORB_LEVELS], o_nms[ORB_LEVELS], o_xy[ORB_LEVELS], o_rs[ORB_LEVELS];
    for (int l = 0; l < n_levels; ++l) {
        OrbLevel &L = a.lv[l];
        // size: cvRound(dim / scale) with the DOUBLE scale (129 / 1.2 = 107.5 -> 108; the float 1.2f would give 107);
        // keypoint coordinates: multiplied by layerScale = (float)pow(scaleFactor, level)
        const double sd = std::pow(scale_factor, (double)l);
        L.scale = (float)sd;
        L.w = l ? (int)lrint((double)S.w / sd) : S.w;
        L.h = l ? (int)lrint((double)S.h / sd) : S.h;
        L.pitch = l ? L.w : S.desc.lv[0].pitch;
        L.active = (L.w > 2 * edge && L.h > 2 * edge) ? 1 : 0;
        const size_t px = (size_t)L.w * L.h;
        L.cap = ((L.w + 1) / 2) * ((L.h + 1) / 2);      // 3x3 strict non-maximum suppression: no two keypoints are 8-neighbours
        o_img[l] = off; off += l ? al(px) : 0;
        o_sc[l] = off; off += al(px);
        o_nms[l] = off; off += al(px);
        o_xy[l] = off; off += al((size_t)L.cap * 4);
        o_rs[l] = off; off += al((size_t)L.cap * 4);
    }
    const int nb = bucket ? n_bins_u * n_bins_v : 0;
    const size_t o_hist = off; off += al(ORB_LEVELS * 256 * 4);
    const size_t o_small = off; off += 256;
    const size_t o_w = off; off += al((size_t)nb * 4 + 4);
    const size_t o_best = off; off += al((size_t)nb * 8 + 8);
    if (off > ctx->orb_bytes) {
        VO_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_orb) cudaFree(ctx->d_orb);
        ctx->d_orb = nullptr; ctx->orb_bytes = 0;
        VO_CUDA(cudaMalloc(&ctx->d_orb, off));
        ctx->orb_bytes = off;
    }
    uint8_t *base = (uint8_t *)ctx->d_orb;
    for (int l = 0; l < n_levels; ++l) {
        OrbLevel &L = a.lv[l];
        if (l) { L.wimg = base + o_img[l]; L.img = L.wimg; } else { L.img = S.desc.lv[0].img; L.wimg = nullptr; }
        L.score = base + o_sc[l]; L.nms = base + o_nms[l]; L.xy = (unsigned *)(base + o_xy[l]); L.resp = (float *)(base + o_rs[l]);
    }
    for (int l = 0; l < n_levels; ++l) ctx->orb_dbg[l] = {a.lv[l].img, a.lv[l].score, a.lv[l].nms, a.lv[l].w, a.lv[l].h, a.lv[l].pitch};
    a.hist = (int *)(base + o_hist);
    int *small = (int *)(base + o_small);
    a.cut = small; a.count = small + 8; a.cutkey = (unsigned *)(small + 16); a.overflow = small + 24;
    a.n_bins_u = bucket ? n_bins_u : 0; a.n_bins_v = bucket ? n_bins_v : 0;
    a.u_step = bucket ? S.w / n_bins_u : 1; a.v_step = bucket ? S.h / n_bins_v : 1;
    a.weight = (int *)(base + o_w); a.best = (unsigned long long *)(base + o_best);
    a.occ = (const float2 *)occ_d; a.n_occ_d = n_occ_d; a.n_occ = n_occ;
    a.out = (float2 *)out_d; a.out_mask = out_mask_d; a.n_out = n_out_d; a.max_out = max_out;
    a.all_pt = (float2 *)all_pt_d; a.all_resp = all_resp_d; a.all_octave = all_octave_d; a.n_all = n_all_d; a.max_all = max_all;

    const int n_reset = nb > ORB_LEVELS * 256 ? nb : ORB_LEVELS * 256;
    k_orb_reset<<<vo_div_up(n_reset, 256), 256, 0, ctx->stream>>>(a);
    if (bucket && n_occ > 0) { k_orb_mark<<<vo_div_up(n_occ, 256), 256, 0, ctx->stream>>>(a); ctx->launches++; }
    for (int l = 1; l < n_levels; ++l) {
        dim3 g(vo_div_up(a.lv[l].w, 32), vo_div_up(a.lv[l].h, 8));
        k_orb_resize<<<g, 256, 0, ctx->stream>>>(a.lv[l - 1], a.lv[l]);
    }
    const int cw = S.w - 2 * edge + 2, ch = S.h - 2 * edge + 2;
    if (cw > 2 && ch > 2) {
        dim3 g(vo_div_up(cw, 32), vo_div_up(ch, 8), n_levels);
        k_orb_fast<<<g, 256, 0, ctx->stream>>>(a);
        k_orb_nms<<<g, 256, 0, ctx->stream>>>(a);
        k_orb_cut<<<n_levels, 256, 0, ctx->stream>>>(a);
        k_orb_emit<<<g, 256, 0, ctx->stream>>>(a);
        k_orb_harris<<<dim3(148, n_levels), 256, 0, ctx->stream>>>(a);
        k_orb_select<<<n_levels, 1024, 0, ctx->stream>>>(a);
        dim3 gb(vo_div_up(a.lv[0].cap, 256), n_levels);
        k_orb_bucket<<<gb, 256, 0, ctx->stream>>>(a);
        ctx->launches += 7;
    }
    if (bucket) { k_orb_out<<<1, 1024, 0, ctx->stream>>>(a); ctx->launches++; }
    ctx->launches += 1 + (n_levels - 1);
    VO_CUDA(cudaGetLastError());
    return VO_OK;
}

extern "C" int vo_set_detector(vo_ctx *ctx, int kind, int fast_threshold)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(kind == VO_DETECTOR_HARRIS_SCHARR || kind == VO_DETECTOR_ORB, VO_ERR_INVALID_ARG, "unknown detector");
    VO_REQUIRE(kind != VO_DETECTOR_ORB || (fast_threshold >= 1 && fast_threshold <= 254), VO_ERR_INVALID_ARG, "FAST threshold out of range");
    ctx->detector = kind;
    if (kind == VO_DETECTOR_ORB) ctx->orb_fast_threshold = fast_threshold;
    return VO_OK;
}

extern "C" int vo_orb_detect(vo_ctx *ctx, int slot, int fast_threshold, int edge, int max_keypoints, float *pts, float *response,
                             int *octave, int *n_out)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(max_keypoints > 0 && pts && response && octave && n_out, VO_ERR_INVALID_ARG, "bad arguments");
    VO_REQUIRE(fast_threshold >= 1 && fast_threshold <= 254, VO_ERR_INVALID_ARG, "FAST threshold out of range");
    VO_CUDA(cudaSetDevice(ctx->device));
    const size_t M = (size_t)max_keypoints;
    const size_t o_pt = 0, o_r = M * 8, o_o = o_r + M * 4, o_n = o_o + M * 4, total = o_n + 64;
    int rc = vo_stage_reserve(ctx, total);
    if (rc) return rc;
    uint8_t *h = ctx->h_stage, *d = ctx->d_stage;
    const int saved = ctx->orb_fast_threshold;
    ctx->orb_fast_threshold = fast_threshold;
    rc = vo_orb_launch_d(ctx, slot, nullptr, nullptr, 0, 0, 0, edge, nullptr, nullptr, nullptr, 0, (float *)(d + o_pt), (float *)(d + o_r),
                         (int *)(d + o_o), (int *)(d + o_n), max_keypoints);
    ctx->orb_fast_threshold = saved;
    if (rc) return rc;
    VO_CUDA(cudaMemcpyAsync(h, d, total, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    const int n = *(const int *)(h + o_n);
    VO_REQUIRE(n <= max_keypoints, VO_ERR_INVALID_ARG, "max_keypoints too small for this image");
    *n_out = n;
    memcpy(pts, h + o_pt, (size_t)n * 8); memcpy(response, h + o_r, (size_t)n * 4); memcpy(octave, h + o_o, (size_t)n * 4);
    return VO_OK;
}

// Test hook: a plane of the last K-orb run.  plane 0 = pyramid level, 1 = FAST score, 2 = score after non-maximum suppression
// (the last two are only defined inside the border band the detector evaluates).
extern "C" int vo_orb_read_level(vo_ctx *ctx, int level, int plane, uint8_t *dst, int *w, int *h)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(level >= 0 && level < ORB_LEVELS && plane >= 0 && plane <= 2 && ctx->orb_dbg[level].img, VO_ERR_INVALID_ARG, "no such plane");
    VO_CUDA(cudaSetDevice(ctx->device));
    const auto &D = ctx->orb_dbg[level];
    if (w) *w = D.w;
    if (h) *h = D.h;
    if (!dst) return VO_OK;
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint8_t *src = plane == 0 ? D.img : (plane == 1 ? D.score : D.nms);
    const size_t pitch = plane == 0 ? (size_t)D.pitch : (size_t)D.w;
    VO_CUDA(cudaMemcpy2D(dst, D.w, src, pitch, D.w, D.h, cudaMemcpyDeviceToHost));
    return VO_OK;
}
