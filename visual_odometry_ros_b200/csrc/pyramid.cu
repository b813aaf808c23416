// pyramid.cu -- K-pyr: image pyramid (pyrDown chain) + Scharr derivative pyramid, sm_100a.
//
// Replaces the cv::buildOpticalFlowPyramid + calcScharrDeriv work that
// cv::calcOpticalFlowPyrLK redoes on every call of the reference
// (core/visual_odometry/feature_tracker.cpp:29,60,69,108,117,186): here each image's
// pyramid is built once, stays resident in HBM, and is bit-exact with OpenCV:
//   pyrDown : 5x5 [1 4 6 4 1]^2 at even coordinates, reflect-101, (sum+128)>>8
//   Scharr  : [3 10 3] x [-1 0 1], int16 interleaved (Ix, Iy), reflect-101 at the image edge
// All kernels are HBM/L2-bound byte work: threads own 4 adjacent output pixels so every
// global access is an aligned 32/64/128-bit word and stores are full words.
#include "vo_internal.cuh"

__device__ __forceinline__ int reflect101(int p, int len)
{
    // valid for overshoot < len (VO_PAD <= smallest level size is enforced on the host)
    p = p < 0 ? -p : p;
    p = p >= len ? 2 * (len - 1) - p : p;
    return p;
}

// ---------------------------------------------------------------------------------------
// ingest: dense raw upload (w x h, pitch w) -> interior of the padded level-0 plane.
// 16 output bytes per thread; the raw side is read with byte-granular alignment handling
// because w (1241) is odd, the padded side is written as aligned 128-bit words.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_ingest(const SlotDesc *__restrict__ slots, const IdList ids)
{
    const SlotDesc &S = slots[ids.id[blockIdx.z]];
    const LevelDesc L = S.lv[0];
    const int y = blockIdx.y;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (x0 >= L.w) return;
    const uint8_t *src = S.raw + (size_t)y * L.w + x0;
    uint8_t *dst = L.img + (size_t)y * L.pitch + x0;
    if (x0 + 16 <= L.w) {
        // gather 16 bytes from an arbitrarily aligned address: 5 aligned words + funnel shifts
        const uintptr_t ad = reinterpret_cast<uintptr_t>(src);
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(ad & ~(uintptr_t)3);
        const int sh = (int)(ad & 3) * 8;
        const uint32_t a = __ldg(wp), b = __ldg(wp + 1), c = __ldg(wp + 2), d = __ldg(wp + 3);
        const uint32_t e = sh ? __ldg(wp + 4) : 0u;
        uint4 v;
        v.x = __funnelshift_r(a, b, sh); v.y = __funnelshift_r(b, c, sh);
        v.z = __funnelshift_r(c, d, sh); v.w = __funnelshift_r(d, e, sh);
        *reinterpret_cast<uint4 *>(dst) = v;
    } else {
        for (int k = 0; x0 + k < L.w; ++k) dst[k] = src[k];
    }
}

// ---------------------------------------------------------------------------------------
// pyrDown: one thread -> 4 output columns x 2 output rows.
// ---------------------------------------------------------------------------------------
#define PD_RPT 2
__device__ __forceinline__ void pd_hrow(const uint8_t *__restrict__ row, int t, int w, bool fast, int h4[4])
{
    int b[11];
    if (fast) {
        const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t *>(row + 8 * t - 4));
        const uint2 w12 = __ldg(reinterpret_cast<const uint2 *>(row + 8 * t));
        const uint32_t w3 = __ldg(reinterpret_cast<const uint32_t *>(row + 8 * t + 8));
        b[0] = (w0 >> 16) & 0xff; b[1] = w0 >> 24;
        b[2] = w12.x & 0xff; b[3] = (w12.x >> 8) & 0xff; b[4] = (w12.x >> 16) & 0xff; b[5] = w12.x >> 24;
        b[6] = w12.y & 0xff; b[7] = (w12.y >> 8) & 0xff; b[8] = (w12.y >> 16) & 0xff; b[9] = w12.y >> 24;
        b[10] = w3 & 0xff;
    } else {
#pragma unroll
        for (int i = 0; i < 11; ++i) b[i] = row[reflect101(8 * t - 2 + i, w)];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        h4[j] = b[2 * j] + 4 * b[2 * j + 1] + 6 * b[2 * j + 2] + 4 * b[2 * j + 3] + b[2 * j + 4];
}

__global__ void __launch_bounds__(256)
k_pyrdown(const SlotDesc *__restrict__ slots, const IdList ids, int src_level)
{
    const SlotDesc &S = slots[ids.id[blockIdx.z]];
    const LevelDesc src = S.lv[src_level];
    const LevelDesc dst = S.lv[src_level + 1];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;    // 4-column group
    const int yo = (blockIdx.y * blockDim.y + threadIdx.y) * PD_RPT;
    if (4 * t >= dst.w || yo >= dst.h) return;
    const bool fast = (8 * t - 2 >= 0) && (8 * t + 8 <= src.w - 1);

    int hr[2 * PD_RPT + 3][4];
#pragma unroll
    for (int r = 0; r < 2 * PD_RPT + 3; ++r) {
        const int sy = reflect101(2 * yo - 2 + r, src.h);
        pd_hrow(src.img + (size_t)sy * src.pitch, t, src.w, fast, hr[r]);
    }
#pragma unroll
    for (int k = 0; k < PD_RPT; ++k) {
        const int y = yo + k;
        if (y >= dst.h) break;
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = (hr[2 * k][j] + 4 * hr[2 * k + 1][j] + 6 * hr[2 * k + 2][j] +
                           4 * hr[2 * k + 3][j] + hr[2 * k + 4][j] + 128) >> 8;
            packed |= (uint32_t)v << (8 * j);
        }
        uint8_t *o = dst.img + (size_t)y * dst.pitch + 4 * t;
        if (4 * t + 3 < dst.w) {
            *reinterpret_cast<uint32_t *>(o) = packed;
        } else {
            for (int j = 0; j < 4 && 4 * t + j < dst.w; ++j) o[j] = (uint8_t)(packed >> (8 * j));
        }
    }
}

// ---------------------------------------------------------------------------------------
// finish kernel: reflect-101 border ring of every level + Scharr derivative of every level
// in ONE launch (block roles by blockIdx.x range; the two roles are independent because
// the Scharr role reflects at the image edge itself).
// ---------------------------------------------------------------------------------------
struct FinishPlan {
    int n_levels;
    int do_border_from;            // first level whose border must be filled
    int do_deriv_from;             // first level whose derivative must be computed (n_levels = none)
    int blk_start[2 * VO_MAX_LEVELS + 1];   // [0..L) border roles, [L..2L) scharr roles
};

__device__ __forceinline__ void border_role(const LevelDesc L, int i)
{
    // word-granular enumeration of the ring: top band, bottom band, left band, right band
    const int wp = L.w + 2 * VO_PAD;
    const int wpw = (wp + 3) >> 2;                // words per full padded row
    const int n_band = VO_PAD * wpw;
    int x0, y;
    if (i < 2 * n_band) {
        const int j = i < n_band ? i : i - n_band;
        const int r = j / wpw;
        x0 = (j - r * wpw) * 4 - VO_PAD;
        y = i < n_band ? r - VO_PAD : L.h + r;
    } else {
        int j = i - 2 * n_band;
        const int per_row = 2 * (VO_PAD / 4);
        if (j >= L.h * per_row) return;
        y = j / per_row;
        j -= y * per_row;
        x0 = j < VO_PAD / 4 ? j * 4 - VO_PAD : L.w + (j - VO_PAD / 4) * 4;
        // right band starts at w which may be unaligned: handled bytewise below
    }
    const int sy = reflect101(y, L.h);
    const uint8_t *srow = L.img + (size_t)sy * L.pitch;
    uint8_t *drow = L.img + (size_t)y * L.pitch;
    const bool in_rows = (y >= 0 && y < L.h);
    if ((((size_t)(drow + x0)) & 3) == 0 && !(in_rows && x0 + 3 >= L.w + VO_PAD) && !(in_rows && x0 < L.w && x0 + 3 >= 0)) {
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int x = x0 + k;
            x = x >= L.w + VO_PAD ? L.w + VO_PAD - 1 : x;
            packed |= (uint32_t)srow[reflect101(x, L.w)] << (8 * k);
        }
        *reinterpret_cast<uint32_t *>(drow + x0) = packed;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = x0 + k;
            if (x >= L.w + VO_PAD) break;
            if (in_rows && x >= 0 && x < L.w) continue;   // never touch interior pixels
            drow[x] = srow[reflect101(x, L.w)];
        }
    }
}

__device__ __forceinline__ void scharr_role(const LevelDesc L, int i)
{
    const int gpr = (L.w + 3) >> 2;    // 4-px groups per row
    const int y = i / gpr;
    if (y >= L.h) return;
    const int t = i - y * gpr;
    const int x0 = 4 * t;
    const bool fast = (x0 >= 4) && (x0 + 4 <= L.w - 1);
    int p[3][6];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int sy = reflect101(y - 1 + r, L.h);
        const uint8_t *row = L.img + (size_t)sy * L.pitch;
        if (fast) {
            const uint32_t a = __ldg(reinterpret_cast<const uint32_t *>(row + x0 - 4));
            const uint32_t b = __ldg(reinterpret_cast<const uint32_t *>(row + x0));
            const uint32_t c = __ldg(reinterpret_cast<const uint32_t *>(row + x0 + 4));
            p[r][0] = a >> 24;
            p[r][1] = b & 0xff; p[r][2] = (b >> 8) & 0xff; p[r][3] = (b >> 16) & 0xff; p[r][4] = b >> 24;
            p[r][5] = c & 0xff;
        } else {
#pragma unroll
            for (int k = 0; k < 6; ++k) p[r][k] = row[reflect101(x0 - 1 + k, L.w)];
        }
    }
    short2 o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int dx = 3 * (p[0][k + 2] - p[0][k]) + 10 * (p[1][k + 2] - p[1][k]) + 3 * (p[2][k + 2] - p[2][k]);
        const int dy = 3 * (p[2][k] - p[0][k]) + 10 * (p[2][k + 1] - p[0][k + 1]) + 3 * (p[2][k + 2] - p[0][k + 2]);
        o[k] = make_short2((short)dx, (short)dy);
    }
    short2 *d = L.deriv + (size_t)y * L.pitch + x0;
    if (x0 + 3 < L.w) {
        uint4 v;
        v.x = *reinterpret_cast<uint32_t *>(&o[0]); v.y = *reinterpret_cast<uint32_t *>(&o[1]);
        v.z = *reinterpret_cast<uint32_t *>(&o[2]); v.w = *reinterpret_cast<uint32_t *>(&o[3]);
        *reinterpret_cast<uint4 *>(d) = v;
    } else {
        for (int k = 0; k < 4 && x0 + k < L.w; ++k) d[k] = o[k];   // keep the zero ring intact
    }
}

__global__ void __launch_bounds__(256)
k_pyr_finish(const SlotDesc *__restrict__ slots, const IdList ids, FinishPlan plan)
{
    const SlotDesc &S = slots[ids.id[blockIdx.y]];
    const int b = blockIdx.x;
    // locate role (<= 16 entries, warp-uniform)
    int role = 0;
    while (role + 1 < 2 * plan.n_levels && b >= plan.blk_start[role + 1]) ++role;
    const int i = (b - plan.blk_start[role]) * blockDim.x + threadIdx.x;
    if (role < plan.n_levels) {
        const LevelDesc L = S.lv[role];
        const int wpw = (L.w + 2 * VO_PAD + 3) >> 2;
        const int total = 2 * VO_PAD * wpw + L.h * 2 * (VO_PAD / 4);
        if (i < total) border_role(L, i);
    } else {
        scharr_role(S.lv[role - plan.n_levels], i);
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
int vo_ensure_pyramids(vo_ctx *ctx, const int *slot_ids, int n, int n_levels, int with_deriv)
{
    VO_REQUIRE(n_levels >= 1 && n_levels <= ctx->max_levels, VO_ERR_INVALID_ARG, "n_levels out of range");
    // Which slots need work? Group them (a batch normally shares one state).
    std::vector<int> todo;
    int min_levels = n_levels, min_deriv = with_deriv ? n_levels : 0;
    bool need_border0 = false;
    int w = -1, h = -1;
    for (int i = 0; i < n; ++i) {
        const int s = slot_ids[i];
        VO_REQUIRE(s >= 0 && s < ctx->n_slots, VO_ERR_INVALID_ARG, "slot id out of range");
        Slot &S = ctx->slots[s];
        VO_REQUIRE(S.w > 0, VO_ERR_INVALID_ARG, "slot has no image");
        const bool stale = S.levels_built < n_levels || (with_deriv && S.deriv_built < n_levels) || !S.border0;
        if (!stale) continue;
        if (w < 0) { w = S.w; h = S.h; }
        VO_REQUIRE(S.w == w && S.h == h, VO_ERR_INVALID_ARG, "batched slots must share one image size");
        bool dup = false;
        for (int t : todo) dup |= (t == s);
        if (dup) continue;
        todo.push_back(s);
        min_levels = S.levels_built < min_levels ? S.levels_built : min_levels;
        if (with_deriv) min_deriv = S.deriv_built < min_deriv ? S.deriv_built : min_deriv;
        need_border0 |= !S.border0;
    }
    if (todo.empty()) return VO_OK;
    const int nb_total = (int)todo.size();
    {   // slots whose pixels are still in the raw staging area
        std::vector<int> pend;
        for (int s : todo) if (ctx->slots[s].raw_pending) pend.push_back(s);
        for (size_t c0 = 0; c0 < pend.size(); c0 += VO_IDLIST_MAX) {
            const int nb = (int)(pend.size() - c0 < VO_IDLIST_MAX ? pend.size() - c0 : VO_IDLIST_MAX);
            IdList ids;
            for (int i = 0; i < nb; ++i) ids.id[i] = pend[c0 + i];
            k_ingest<<<dim3(vo_div_up(vo_div_up(w, 16), 128), h, nb), 128, 0, ctx->stream>>>(ctx->d_slots, ids);
            ctx->launches++;
        }
        for (int s : pend) ctx->slots[s].raw_pending = false;
    }
    const Slot &S0 = ctx->slots[todo[0]];
    if (min_levels < 1) min_levels = 1;
    for (int c0 = 0; c0 < nb_total; c0 += VO_IDLIST_MAX) {
        const int nb = nb_total - c0 < VO_IDLIST_MAX ? nb_total - c0 : VO_IDLIST_MAX;
        IdList ids;
        for (int i = 0; i < nb; ++i) ids.id[i] = todo[c0 + i];
        for (int l = min_levels - 1; l + 1 < n_levels; ++l) {
            const LevelDesc &D = S0.desc.lv[l + 1];
            dim3 blk(32, 8);
            dim3 grd(vo_div_up(vo_div_up(D.w, 4), 32), vo_div_up(vo_div_up(D.h, PD_RPT), 8), nb);
            k_pyrdown<<<grd, blk, 0, ctx->stream>>>(ctx->d_slots, ids, l);
            ctx->launches++;
        }
    }
    FinishPlan plan;
    plan.n_levels = n_levels;
    plan.do_border_from = need_border0 ? 0 : min_levels;
    plan.do_deriv_from = with_deriv ? min_deriv : n_levels;
    int blocks = 0;
    for (int l = 0; l < n_levels; ++l) {
        plan.blk_start[l] = blocks;
        if (l >= plan.do_border_from) {
            const LevelDesc &L = S0.desc.lv[l];
            const int wpw = (L.w + 2 * VO_PAD + 3) >> 2;
            blocks += vo_div_up(2 * VO_PAD * wpw + L.h * 2 * (VO_PAD / 4), 256);
        }
    }
    for (int l = 0; l < n_levels; ++l) {
        plan.blk_start[n_levels + l] = blocks;
        if (l >= plan.do_deriv_from) {
            const LevelDesc &L = S0.desc.lv[l];
            blocks += vo_div_up(vo_div_up(L.w, 4) * L.h, 256);
        }
    }
    plan.blk_start[2 * n_levels] = blocks;
    for (int c0 = 0; c0 < nb_total && blocks > 0; c0 += VO_IDLIST_MAX) {
        const int nb = nb_total - c0 < VO_IDLIST_MAX ? nb_total - c0 : VO_IDLIST_MAX;
        IdList ids;
        for (int i = 0; i < nb; ++i) ids.id[i] = todo[c0 + i];
        k_pyr_finish<<<dim3(blocks, nb), 256, 0, ctx->stream>>>(ctx->d_slots, ids, plan);
        ctx->launches++;
    }
    VO_CUDA(cudaGetLastError());
    for (int s : todo) {
        Slot &S = ctx->slots[s];
        S.levels_built = n_levels > S.levels_built ? n_levels : S.levels_built;
        if (with_deriv) S.deriv_built = n_levels > S.deriv_built ? n_levels : S.deriv_built;
        S.border0 = true;
    }
    return VO_OK;
}
