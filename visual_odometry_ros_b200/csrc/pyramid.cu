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

#include <cstdlib>

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
// Fused pyramid: ONE launch builds, for a batch of images, level 0 (ingest from the dense raw upload or
// re-use of the resident plane), the pyrDown chain up to 3 more levels, the reflect-101 ring of every level
// and the Scharr plane of every level.  One CTA owns a 128 x 64 level-0 tile and the matching
// (128>>l) x (64>>l) tile of every level; the level-0 tile plus the halo the chain needs is staged ONCE in
// shared memory and every coarser level is computed from shared memory, so HBM sees each pixel once on the
// way in and each output byte once on the way out (the per-level kernels above re-read every level 2-3 times
// and need 5 launches per image batch).
//   halo rows   : HY[top] = 1 (Scharr), HY[l] = 2 HY[l+1] + 2 (5-tap decimation)  ->  4 levels: 22, 10, 4, 1
//   halo columns: the same rounded up to a multiple of 4 below the top level (28, 12, 4, 1) so that both the
//                 tile origin and the owned origin are word aligned in shared memory.
//   Out-of-image halo positions hold the reflect-101 value, so the arithmetic is branch-free and
//   bit-identical to cv::pyrDown / cv::Scharr with BORDER_REFLECT_101.
//   The byte arithmetic runs on dp4a: a 5-tap row sum is two dp4a on aligned words (weights 1 4 6 4 | 1 and
//   0 0 1 4 | 6 4 1 0 for the two column parities); the Scharr taps are dp4a on byte-transposed columns.
// ---------------------------------------------------------------------------------------
#define PF_TW 128
#define PF_TH 64
template <int NL> struct PFCfg {
    __host__ __device__ static constexpr int HY(int l) { return l >= NL - 1 ? 1 : 2 * HY(l + 1) + 2; }
    __host__ __device__ static constexpr int HX(int l) { return l >= NL - 1 ? 1 : (2 * HX(l + 1) + 2 + 3) / 4 * 4; }
    __host__ __device__ static constexpr int P(int l) { return (HX(l) + (PF_TW >> l) + HX(l) + 3) / 4 * 4 + (l >= NL - 1 ? 4 : 0); }   // smem pitch (bytes)
    __host__ __device__ static constexpr int R(int l) { return (PF_TH >> l) + 2 * HY(l); }                // smem rows
    __host__ __device__ static constexpr int OFF(int l) { return l == 0 ? 0 : OFF(l - 1) + P(l - 1) * R(l - 1); }
    // smem column of level coordinate x is x - ax + CX(l): word aligned owned origin on every level
    __host__ __device__ static constexpr int CX(int l) { return (HX(l) + 3) / 4 * 4; }
    static constexpr int BYTES = OFF(NL - 1) + P(NL - 1) * R(NL - 1) + 16;
};

__device__ __forceinline__ int dp4a_uu(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// 5-tap [1 4 6 4 1] sums of 4 decimated outputs from 4 aligned words; the first tap of output j is byte OFFB + 2j.
template <int OFFB>
__device__ __forceinline__ void pf_hsum4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int h[4])
{
    constexpr uint32_t K_A = 0x04060401u;   // bytes (1,4,6,4): taps 0..3 on an aligned word
    constexpr uint32_t K_B = 0x00000001u;   // tap 4 = first byte of the next word
    constexpr uint32_t K_C = 0x04010000u;   // bytes (0,0,1,4): taps 0..1 on bytes 2,3
    constexpr uint32_t K_D = 0x00010406u;   // bytes (6,4,1,0): taps 2..4 on the next word
    if (OFFB == 0) {
        h[0] = dp4a_uu(w1, K_B, dp4a_uu(w0, K_A, 0));
        h[1] = dp4a_uu(w1, K_D, dp4a_uu(w0, K_C, 0));
        h[2] = dp4a_uu(w2, K_B, dp4a_uu(w1, K_A, 0));
        h[3] = dp4a_uu(w2, K_D, dp4a_uu(w1, K_C, 0));
    } else {            // OFFB == 2
        h[0] = dp4a_uu(w1, K_D, dp4a_uu(w0, K_C, 0));
        h[1] = dp4a_uu(w2, K_B, dp4a_uu(w1, K_A, 0));
        h[2] = dp4a_uu(w2, K_D, dp4a_uu(w1, K_C, 0));
        h[3] = dp4a_uu(w3, K_B, dp4a_uu(w2, K_A, 0));
    }
}

// Write one level from its shared-memory tile: interior (optional), ring copies of the owned pixels, Scharr plane.
// s points at the tile, (CXl, HYl) is the tile position of the owned origin (ax, ay).  Requires w, h >= VO_PAD + 2.
template <int LOGWQ>      // log2 of the words per owned row of a full tile: (128 >> l) / 4
__device__ __forceinline__ void pf_emit(const LevelDesc L, const uint8_t *__restrict__ s, int P, int CXl, int HYl, int ax, int bx, int ay,
                                        int by, bool write_interior, bool write_ring, bool with_deriv)
{
    const int ow = bx - ax, oh = by - ay;
    if (ow <= 0 || oh <= 0) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int wq = (ow + 3) >> 2;
    constexpr int WQF = 1 << LOGWQ;
    const int mx = L.w - 1, my = L.h - 1;
    // ---- interior words + the vertical ring copies of the same words (rows 1..32 -> -1..-32, rows h-33..h-2 -> h..h+31)
    if (write_interior || write_ring) {
        for (int i = threadIdx.x; i < (oh << LOGWQ); i += 256) {
            const int r = i >> LOGWQ, c = i & (WQF - 1);
            if (c >= wq) continue;
            const int y = ay + r;
            // up to three destination rows: the row itself, its copy above the image, its copy below the image
            const int yt0 = write_interior ? y : 0x7fffffff;
            const int yt1 = (write_ring && y >= 1 && y <= VO_PAD) ? -y : 0x7fffffff;
            const int yt2 = (write_ring && my - y >= 1 && my - y <= VO_PAD) ? 2 * my - y : 0x7fffffff;
            const int x = ax + 4 * c;
            const uint32_t v = *reinterpret_cast<const uint32_t *>(s + (r + HYl) * P + CXl + 4 * c);
            const bool full = x + 3 < L.w;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int yt = q == 0 ? yt0 : (q == 1 ? yt1 : yt2);
                if (yt == 0x7fffffff) continue;
                uint8_t *o = L.img + (ptrdiff_t)yt * L.pitch + x;
                if (full) *reinterpret_cast<uint32_t *>(o) = v;
                else for (int k = 0; x + k < L.w; ++k) o[k] = (uint8_t)(v >> (8 * k));
            }
        }
    }
    // ---- horizontal ring copies (columns 1..32 -> -1..-32, columns w-33..w-2 -> w..w+31) incl. the corners
    if (write_ring && (ax <= VO_PAD || bx >= L.w - VO_PAD - 1)) {
        for (int r = ty; r < oh; r += 8) {
            const int y = ay + r;
            const int ytA = (y >= 1 && y <= VO_PAD) ? -y : 0x7fffffff;
            const int ytB = (my - y >= 1 && my - y <= VO_PAD) ? 2 * my - y : 0x7fffffff;
            const uint8_t *srow = s + (r + HYl) * P + CXl;
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const int k = tx + 1;                               // 1..32
                const int x = side == 0 ? k : mx - k;               // source column
                const int xt = side == 0 ? -k : mx + k;             // ring column
                if (x < ax || x >= bx || x < 1 || x > mx - 1) continue;
                const uint8_t v = srow[x - ax];
                L.img[(ptrdiff_t)y * L.pitch + xt] = v;
                if (ytA != 0x7fffffff) L.img[(ptrdiff_t)ytA * L.pitch + xt] = v;
                if (ytB != 0x7fffffff) L.img[(ptrdiff_t)ytB * L.pitch + xt] = v;
            }
        }
    }
    // ---- Scharr plane: 4 pixels per thread, dp4a on byte-transposed columns (a_j, b_j, c_j, 0)
    if (with_deriv) {
        for (int i = threadIdx.x; i < (oh << LOGWQ); i += 256) {
            const int r = i >> LOGWQ, c = i & (WQF - 1);
            if (c >= wq) continue;
            const int y = ay + r;
            const uint8_t *r0 = s + (r + HYl - 1) * P + CXl, *r1 = r0 + P, *r2 = r1 + P;
            {
                const int x0 = ax + 4 * c;
                const uint32_t a = *reinterpret_cast<const uint32_t *>(r0 + 4 * c);
                const uint32_t b = *reinterpret_cast<const uint32_t *>(r1 + 4 * c);
                const uint32_t cc = *reinterpret_cast<const uint32_t *>(r2 + 4 * c);
                uint32_t col[6];
                col[0] = (uint32_t)r0[4 * c - 1] | ((uint32_t)r1[4 * c - 1] << 8) | ((uint32_t)r2[4 * c - 1] << 16);
                col[5] = (uint32_t)r0[4 * c + 4] | ((uint32_t)r1[4 * c + 4] << 8) | ((uint32_t)r2[4 * c + 4] << 16);
                const uint32_t ab01 = __byte_perm(a, b, 0x5140), ab23 = __byte_perm(a, b, 0x7362);   // (a0,b0,a1,b1), (a2,b2,a3,b3)
                col[1] = __byte_perm(ab01, cc, 0x7410); col[2] = __byte_perm(ab01, cc, 0x7532);      // (a,b,c,*) -> top byte is masked by a 0 weight
                col[3] = __byte_perm(ab23, cc, 0x7610); col[4] = __byte_perm(ab23, cc, 0x7732);
                int sm[6], df[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    sm[j] = dp4a_uu(col[j], 0x00030a03u, 0);            // 3a + 10b + 3c
                    df[j] = dp4a_us(col[j], 0x000100ff, 0);             // c - a  (weights -1, 0, +1, 0)
                }
                uint32_t o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int dx = sm[k + 2] - sm[k];
                    const int dy = 3 * (df[k] + df[k + 2]) + 10 * df[k + 1];
                    o[k] = __byte_perm((uint32_t)dx, (uint32_t)dy, 0x5410);
                }
                short2 *d = L.deriv + (size_t)y * L.pitch + x0;
                if (x0 + 3 < L.w) *reinterpret_cast<uint4 *>(d) = make_uint4(o[0], o[1], o[2], o[3]);
                else for (int k = 0; k < 4 && x0 + k < L.w; ++k) reinterpret_cast<uint32_t *>(d)[k] = o[k];   // keep the zero ring intact
            }
        }
    }
}

// 5x5 decimation of the source tile into the destination tile (all staged positions; the ones outside the image
// are overwritten by the reflection pass that follows).  Work item = 4 columns x 2 rows.
template <int OFFB, int GW, int GH>
__device__ __forceinline__ void pf_down(const uint8_t *__restrict__ ss, int Ps, uint8_t *__restrict__ sd, int Pd, int dcol0, bool byte_store)
{
    // ss points at the source byte of tap 0 of destination (row 0, col 0) rounded down to a word; sd at destination (0, 0)
    for (int i = threadIdx.x; i < GW * GH; i += 256) {
        const int g = i % GW, rp = i / GW;
        int hr[7][4];
#pragma unroll
        for (int rr = 0; rr < 7; ++rr) {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(ss + (4 * rp + rr) * Ps + 8 * g);
            pf_hsum4<OFFB>(q[0], q[1], q[2], q[3], hr[rr]);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            uint32_t packed = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int v = (hr[2 * k][j] + 4 * hr[2 * k + 1][j] + 6 * hr[2 * k + 2][j] + 4 * hr[2 * k + 3][j] + hr[2 * k + 4][j] + 128) >> 8;
                packed |= (uint32_t)v << (8 * j);
            }
            uint8_t *o = sd + (2 * rp + k) * Pd + dcol0 + 4 * g;
            if (!byte_store) *reinterpret_cast<uint32_t *>(o) = packed;
            else { o[0] = (uint8_t)packed; o[1] = (uint8_t)(packed >> 8); o[2] = (uint8_t)(packed >> 16); o[3] = (uint8_t)(packed >> 24); }
        }
    }
}

template <int NL, int l>
__device__ __forceinline__ void pf_level(const SlotDesc &S, uint8_t *smem, int X0, int Y0, int from_raw, int with_deriv, int write_ring0)
{
    using C = PFCfg<NL>;
    const int tid = threadIdx.x;
    const LevelDesc L = S.lv[l];
    const int ax = X0 >> l, ay = Y0 >> l;
    const int bx = min(L.w, (X0 + PF_TW) >> l), by = min(L.h, (Y0 + PF_TH) >> l);
    const uint8_t *sl = smem + C::OFF(l);
    pf_emit<5 - l>(L, sl, C::P(l), C::CX(l), C::HY(l), ax, bx, ay, by, l > 0 || from_raw, l > 0 || write_ring0, with_deriv != 0);
    if constexpr (l + 1 < NL) {
        const LevelDesc D = S.lv[l + 1];
        uint8_t *sd = smem + C::OFF(l + 1);
        constexpr int Ps = C::P(l), Pd = C::P(l + 1);
        constexpr int HXd = C::HX(l + 1), HYd = C::HY(l + 1), CXd = C::CX(l + 1);
        constexpr int cw = (PF_TW >> (l + 1)) + 2 * HXd, chh = (PF_TH >> (l + 1)) + 2 * HYd;   // staged destination extent
        constexpr int GW = (cw + 3) / 4, GH = (chh + 1) / 2;
        // tap 0 of destination column (dax - HXd) is source column ax - 2 HXd - 2  ->  tile column CX(l) - 2 HXd - 2
        constexpr int scol = C::CX(l) - 2 * HXd - 2;
        static_assert(scol >= 0 && (scol % 4 == 0 || scol % 4 == 2), "decimation source alignment");
        static_assert(C::HY(l) - 2 * HYd - 2 >= 0, "decimation source rows");
        static_assert((scol & ~3) + 8 * (GW - 1) + 16 <= Ps + 16, "decimation source row overrun");
        const uint8_t *ss = sl + (C::HY(l) - 2 * HYd - 2) * Ps + (scol & ~3);
        pf_down<(scol & 3), GW, GH>(ss, Ps, sd, Pd, CXd - HXd, ((CXd - HXd) & 3) != 0);
        __syncthreads();
        // ---- reflect-101 fill of the staged positions outside the image (edge tiles only)
        const int dax = X0 >> (l + 1), day = Y0 >> (l + 1);
        if (dax - HXd < 0 || day - HYd < 0 || dax + (PF_TW >> (l + 1)) + HXd > D.w || day + (PF_TH >> (l + 1)) + HYd > D.h) {
            for (int i = tid; i < cw * chh; i += 256) {
                const int r = i / cw, c = i - r * cw;
                const int x = dax - HXd + c, y = day - HYd + r;
                if (x >= 0 && x < D.w && y >= 0 && y < D.h) continue;
                const int rx = reflect101(x > D.w - 1 + VO_PAD ? D.w - 1 + VO_PAD : x, D.w), ry = reflect101(y > D.h - 1 + VO_PAD ? D.h - 1 + VO_PAD : y, D.h);
                const int cc = rx - dax + HXd, rr = ry - day + HYd;
                uint8_t v = 0;
                if (cc >= 0 && cc < cw && rr >= 0 && rr < chh) v = sd[rr * Pd + (cc - HXd + CXd)];
                sd[r * Pd + (c - HXd + CXd)] = v;
            }
            __syncthreads();
        }
        pf_level<NL, l + 1>(S, smem, X0, Y0, from_raw, with_deriv, write_ring0);
    }
}

template <int NL>
__global__ void __launch_bounds__(256)
k_pyr_fused(const SlotDesc *__restrict__ slots, const IdList ids, int from_raw, int with_deriv, int write_ring0)
{
    using C = PFCfg<NL>;
    __shared__ __align__(16) uint8_t smem[C::BYTES];
    const SlotDesc &S = slots[ids.id[blockIdx.z]];
    const int X0 = blockIdx.x * PF_TW, Y0 = blockIdx.y * PF_TH;
    const int tid = threadIdx.x;

    // ---- stage the level-0 tile + halo (reflect-101 outside the image): columns [X0 - CX0, X0 + 128 + HX0), all rows
    {
        const LevelDesc L = S.lv[0];
        const uint8_t *src = from_raw ? S.raw : L.img;
        const int sp = from_raw ? L.w : L.pitch;
        constexpr int HY0 = C::HY(0), CX0 = C::CX(0), P0 = C::P(0), R0 = C::R(0);
        constexpr int WQ = P0 / 4;
        uint8_t *s0 = smem;
        for (int i = tid; i < WQ * R0; i += 256) {
            const int r = i / WQ, c = i - r * WQ;
            const int gy = Y0 - HY0 + r, gx = X0 - CX0 + 4 * c;
            const int sy = reflect101(gy > L.h - 1 + VO_PAD ? L.h - 1 + VO_PAD : gy, L.h);
            const uint8_t *row = src + (size_t)sy * sp;
            uint32_t v;
            if (gx >= 0 && gx + 3 < L.w) {
                const uintptr_t ad = reinterpret_cast<uintptr_t>(row + gx);
                const uint32_t *wp = reinterpret_cast<const uint32_t *>(ad & ~(uintptr_t)3);
                const int sh = (int)(ad & 3) * 8;
                const uint32_t a = __ldg(wp);
                const uint32_t b = sh ? __ldg(wp + 1) : 0u;
                v = __funnelshift_r(a, b, sh);
            } else {
                v = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    int x = gx + k;
                    x = x > L.w - 1 + VO_PAD ? L.w - 1 + VO_PAD : x;
                    v |= (uint32_t)row[reflect101(x, L.w)] << (8 * k);
                }
            }
            *reinterpret_cast<uint32_t *>(s0 + r * P0 + 4 * c) = v;
        }
    }
    __syncthreads();
    pf_level<NL, 0>(S, smem, X0, Y0, from_raw, with_deriv, write_ring0);
}

template <int NL>
static void launch_pyr_fused(vo_ctx *ctx, const IdList &ids, int nb, int w, int h, int from_raw, int with_deriv, int write_ring0)
{
    dim3 grd(vo_div_up(w, PF_TW), vo_div_up(h, PF_TH), nb);
    k_pyr_fused<NL><<<grd, 256, 0, ctx->stream>>>(ctx->d_slots, ids, from_raw, with_deriv, write_ring0);
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
    // ---- fused path: every stale slot is built from its level-0 pixels in ONE launch per <= 64 images
    bool all_fresh = w >= VO_PAD + 2 && h >= VO_PAD + 2;
    for (int s : todo) all_fresh &= ctx->slots[s].levels_built <= 1 && ctx->slots[s].deriv_built == 0;
    if (all_fresh) {
        // fused levels: at most 4, and only levels at least VO_PAD + 2 px wide/high (single-bounce ring reflection)
        int nl_f = 0;
        for (int lw = w, lh = h; nl_f < n_levels && nl_f < 4 && lw >= VO_PAD + 2 && lh >= VO_PAD + 2; ++nl_f) { lw = (lw + 1) / 2; lh = (lh + 1) / 2; }
        for (int pass = 0; pass < 2; ++pass) {          // pass 0: pixels still in the raw staging area, pass 1: resident
            std::vector<int> grp;
            for (int s : todo) if ((ctx->slots[s].raw_pending ? 0 : 1) == pass) grp.push_back(s);
            // the level-0 ring must be (re)written unless every slot of the group still has it
            bool ring0 = pass == 0;
            for (int s : grp) ring0 |= !ctx->slots[s].border0;
            for (size_t c0 = 0; c0 < grp.size(); c0 += VO_IDLIST_MAX) {
                const int nb = (int)(grp.size() - c0 < VO_IDLIST_MAX ? grp.size() - c0 : VO_IDLIST_MAX);
                IdList ids;
                for (int i = 0; i < nb; ++i) ids.id[i] = grp[c0 + i];
                const int fr = pass == 0 ? 1 : 0, wd = with_deriv ? 1 : 0, r0 = ring0 ? 1 : 0;
                if (nl_f == 1) launch_pyr_fused<1>(ctx, ids, nb, w, h, fr, wd, r0);
                else if (nl_f == 2) launch_pyr_fused<2>(ctx, ids, nb, w, h, fr, wd, r0);
                else if (nl_f == 3) launch_pyr_fused<3>(ctx, ids, nb, w, h, fr, wd, r0);
                else launch_pyr_fused<4>(ctx, ids, nb, w, h, fr, wd, r0);
                ctx->launches++;
            }
        }
        VO_CUDA(cudaGetLastError());
        for (int s : todo) {
            Slot &S = ctx->slots[s];
            S.raw_pending = false;
            S.levels_built = nl_f;
            if (with_deriv) S.deriv_built = nl_f;
            S.border0 = true;
        }
        if (nl_f == n_levels) return VO_OK;
        // deeper levels (rare: maxLevel > 3) continue on the per-level kernels below
        min_levels = nl_f;
        min_deriv = with_deriv ? nl_f : 0;
        need_border0 = false;
    }
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
