// lba.cu -- K-lba-*: sliding-window local bundle adjustment (Schur-complement LM), FP64.
//
// Replaces SparseBundleAdjustmentSolver::solveForFiniteIterations
// (core/visual_odometry/ba_solver/sparse_bundle_adjustment.cpp:150-768) on the flat problem that
// SparseBAParameters::setPosesAndPoints packs (ba_solver/sparse_ba_parameters.h:292-465).
//
// Per LM iteration two launches (deterministic: no floating-point atomics anywhere):
//   k_lba_build  one CTA per TILE of landmarks, one warp per landmark, one lane per observation.
//                Applies the previous iteration's point update, evaluates residuals / Jacobians,
//                C_i, b_i (warp shuffles), A_j, a_j (per-warp shared accumulators, summed in a fixed
//                order), the last-writer cross blocks B[j][i], damps and inverts C_i (3x3 pivoted
//                LDLT), forms BCinv and stages the tile as two dense shared-memory panels
//                P = [BCinv_i], Q = [B_i | b_i]; the tile's Schur contribution P^T Q -- a dense
//                (6 N_opt) x (6 N_opt + 1) x (3 TL) contraction -- is then formed with every thread
//                owning fixed output entries.  The reference instead memsets and walks four dense
//                N_opt x M block arrays (23 MB) per iteration (:844-885).
//   k_lba_solve  one CTA: fixed-order reduction of the tile partials, damping, reduced camera system
//                S = blkdiag(A) - BCinvBt, pivoted LDLT (the algorithm of Eigen::LDLT) in shared
//                memory, pose retraction T <- Exp(Log(Exp(x) Exp(Log T))), error bookkeeping.
// Reference defects kept: B assigned not accumulated (last observation
// of (landmark, keyframe) wins), right-camera Q fed to the Q(0,1)=Q(1,0)=0 shortcut, strict '>' Huber
// gate, SE3Log's w=0 snap, transposed diagonal blocks of BCinvBt.  Compiled with -fmad=false.
#include "vo_internal.cuh"

#include <cstdio>
#include <mutex>
#include <cstdlib>

#include <cstring>
#include <chrono>
#include <vector>

#define LBA_WARPS 8
#define LBA_THREADS (LBA_WARPS * 32)
#define LBA_BWARPS 16           // k_lba_build: warps per tile (latency-bound per-landmark chains: more warps in flight)
#define LBA_BTHREADS (LBA_BWARPS * 32)
#define LBA_MAX_OPT 16
#define LBA_NA 27            // 21 upper entries of Q^T W Q + 6 entries of Q^T W r

struct LbaDev {
    int n_frames, n_opt, n_points, n_obs, n_tiles, tile, n6;
    double *poses;           // [n_frames][16]
    const int *opt_index;    // [n_frames]
    int *opt_frame;          // [n_opt]
    double *points;          // [n_points][3]
    const int *obs_ptr;
    const int *obs_frame;
    const uint8_t *obs_right;
    const double *obs_px;
    double *bcinv;           // [n_obs][18]   BCinv[j][i] of left-opt observations (row-major 6x3)
    double *cinv_b;          // [n_points][3]
    double *s_part;          // [n_tiles][n6*(n6+1)]
    double *a_part;          // [n_tiles][n_opt][27]
    double *err_part;        // [n_tiles]
    double *red;             // [n6*(n6+1) + n_opt*27 + 1] tile partials summed in tile order (k_lba_reduce)
    double *x;               // [n6]
    double *avg_err;         // [max_iter]
    int *nan_flag;
    long long *dbg;          // nullable: clock64 stamps of k_lba_solve (VO_LBA_TRACE)
    double K_l[4], K_r[4];
    double R_rl[9], t_rl[3];
    double huber, lambda;
};

// ------------------------------------------------------------------ small FP64 helpers
__device__ void ldlt3_inverse(const double *Cin, double *Cinv)
{
    // Eigen::LDLT<Matrix3d>::solve(Identity): unblocked, diagonal pivoting, lower storage
    double m[9];
    int tr[3];
    for (int i = 0; i < 9; ++i) m[i] = Cin[i];
#define M3(i, j) m[(i) * 3 + (j)]
    for (int k = 0; k < 3; ++k) {
        int big = k;
        double bv = fabs(M3(k, k));
        for (int i = k + 1; i < 3; ++i) if (fabs(M3(i, i)) > bv) { bv = fabs(M3(i, i)); big = i; }
        tr[k] = big;
        if (big != k) {
            const int s = 3 - big - 1;
            for (int j = 0; j < k; ++j) { const double t = M3(k, j); M3(k, j) = M3(big, j); M3(big, j) = t; }
            for (int i = 0; i < s; ++i) { const double t = M3(big + 1 + i, k); M3(big + 1 + i, k) = M3(big + 1 + i, big); M3(big + 1 + i, big) = t; }
            { const double t = M3(k, k); M3(k, k) = M3(big, big); M3(big, big) = t; }
            for (int i = k + 1; i < big; ++i) { const double t = M3(i, k); M3(i, k) = M3(big, i); M3(big, i) = t; }
        }
        const int rs = 3 - k - 1;
        if (k > 0) {
            double temp[3];
            for (int j = 0; j < k; ++j) temp[j] = M3(j, j) * M3(k, j);
            double s = 0;
            for (int j = 0; j < k; ++j) s += M3(k, j) * temp[j];
            M3(k, k) -= s;
            for (int i = 0; i < rs; ++i) {
                double s2 = 0;
                for (int j = 0; j < k; ++j) s2 += M3(k + 1 + i, j) * temp[j];
                M3(k + 1 + i, k) -= s2;
            }
        }
        const double akk = M3(k, k);
        if (rs > 0 && fabs(akk) > 0.0) for (int i = 0; i < rs; ++i) M3(k + 1 + i, k) /= akk;
    }
    const double tol = 1.0 / 1.7976931348623157e308;
    for (int c = 0; c < 3; ++c) {
        double y[3] = {c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0};
        for (int k = 0; k < 3; ++k) if (tr[k] != k) { const double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
        for (int i = 0; i < 3; ++i) { double s = y[i]; for (int j = 0; j < i; ++j) s -= M3(i, j) * y[j]; y[i] = s; }
        for (int i = 0; i < 3; ++i) y[i] = fabs(M3(i, i)) > tol ? y[i] / M3(i, i) : 0.0;
        for (int i = 2; i >= 0; --i) { double s = y[i]; for (int j = i + 1; j < 3; ++j) s -= M3(j, i) * y[j]; y[i] = s; }
        for (int k = 2; k >= 0; --k) if (tr[k] != k) { const double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
        for (int i = 0; i < 3; ++i) Cinv[i * 3 + c] = y[i];
    }
#undef M3
}

__device__ void se3exp_d(const double *xi, double *T)
{
    const double v[3] = {xi[0], xi[1], xi[2]}, w[3] = {xi[3], xi[4], xi[5]};
    const double theta = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const double wx[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double wxwx[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += wx[i * 3 + k] * wx[k * 3 + j];
            wxwx[i * 3 + j] = s;
        }
    double a, b, c;
    if (theta < 1e-9) { a = 1.0; b = 0.5; c = 0.33333333333333333333333333; }
    else {
        const double invtheta2 = 1.0 / (theta * theta);
        a = sin(theta) / theta;
        b = (1 - cos(theta)) * invtheta2;
        c = (theta - sin(theta)) / (theta * theta * theta);
    }
    double V[9];
    for (int i = 0; i < 16; ++i) T[i] = 0.0;
    for (int i = 0; i < 9; ++i) {
        const double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        const int r = i / 3, cc = i - 3 * r;
        T[r * 4 + cc] = (id + a * wx[i]) + b * wxwx[i];
        V[i] = (id + b * wx[i]) + c * wxwx[i];
    }
    for (int i = 0; i < 3; ++i) T[i * 4 + 3] = (V[i * 3] * v[0] + V[i * 3 + 1] * v[1]) + V[i * 3 + 2] * v[2];
    T[15] = 1.0;
}

__device__ void se3log_d(const double *T, double *xi)
{
    double R[9], t[3];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R[i * 3 + j] = T[i * 4 + j]; t[i] = T[i * 4 + 3]; }
    const double inCos = ((R[0] + R[4] + R[8]) - 1.0) * 0.5;
    double w[3] = {0, 0, 0};
    double Vin[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (!(inCos >= 0.999999999)) {
        const double theta = acos(inCos);
        const double invTheta = 1.0 / theta, invTheta2 = invTheta * invTheta;
        const double k = theta / (2.0 * sin(theta));
        double lnR[9];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) lnR[i * 3 + j] = k * (R[i * 3 + j] - R[j * 3 + i]);
        w[0] = -lnR[5]; w[1] = lnR[2]; w[2] = -lnR[1];
        const double wx[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
        double wxwx[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int kk = 0; kk < 3; ++kk) s += wx[i * 3 + kk] * wx[kk * 3 + j];
                wxwx[i * 3 + j] = s;
            }
        const double A = sin(theta) * invTheta;
        const double B = (1.0 - cos(theta)) * invTheta2;
        const double cc = invTheta2 * (1.0 - A / (2.0 * B));
        for (int i = 0; i < 9; ++i) {
            const double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
            Vin[i] = (id - 0.5 * wx[i]) + cc * wxwx[i];
        }
    }
    for (int i = 0; i < 3; ++i) xi[i] = (Vin[i * 3] * t[0] + Vin[i * 3 + 1] * t[1]) + Vin[i * 3 + 2] * t[2];
    xi[3] = w[0]; xi[4] = w[1]; xi[5] = w[2];
}

__device__ void pose_retract(double *T, const double *x)
{
    double xi[6], Tjw[16], dT[16], Tn[16];
    se3log_d(T, xi);
    se3exp_d(xi, Tjw);
    se3exp_d(x, dT);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += dT[i * 4 + k] * Tjw[k * 4 + j];
            Tn[i * 4 + j] = s;
        }
    se3log_d(Tn, xi);
    se3exp_d(xi, T);
}

__device__ __forceinline__ double wsum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// frame of observation o, clamped into the window: the host checks the range while the kernels already run
__device__ __forceinline__ int lba_frame(const LbaDev &d, const int o)
{
    return min(max(d.obs_frame[o], 0), d.n_frames - 1);
}

// ------------------------------------------------------------------ k_lba_build
// dynamic shared memory layout (doubles):
//   P     [3*TL][n6]        P[3*il+m][6j+r] = BCinv[j][i](r, m)
//   Q     [3*TL][n6+1]      Q[3*il+m][6k+c] = B[k][i](c, m) ; Q[3*il+m][n6] = b_i(m)
//   Aacc  [LBA_BWARPS][n_opt][27]
//   Btmp  [LBA_BWARPS][NG][n_opt][18]
//   errw  [LBA_BWARPS]
// A warp works on NG landmarks at a time, one per group of GS = 32 / NG lanes (lane of the group = observation).  The
// per-landmark work is one long dependent chain (observation tables -> projection -> divisions -> group sums -> 3 x 3
// LDLT inverse -> B Cinv); a landmark of the benchmark window has 5 observations on average, so a whole warp per landmark
// left 27 lanes idle and every warp walked 3+ landmarks one after the other (issue slots 36 % busy, profiles/r2_lba_build_*).
// With four groups the 16 warps of a tile hold 64 landmarks in flight: one round for a tile of the 7.5 k-landmark window.
// Only the accumulation into the warp's A_j blocks is serialised over the groups (fixed order: deterministic sums).
template <int GS>
__device__ __forceinline__ double gsum(double v)
{
#pragma unroll
    for (int o = GS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NG>
__global__ void __launch_bounds__(LBA_BTHREADS, 1)
k_lba_build(const LbaDev d, int apply_update)
{
    constexpr int GS = 32 / NG;
    extern __shared__ __align__(16) double smem[];
    const int n6 = d.n6, No = d.n_opt, TL = d.tile;
    double *P = smem;
    double *Q = P + (size_t)3 * TL * n6;
    double *Aacc = Q + (size_t)3 * TL * (n6 + 1);
    double *Btmp = Aacc + (size_t)LBA_BWARPS * No * LBA_NA;
    double *errw = Btmp + (size_t)LBA_BWARPS * NG * No * 18;
    double *s_x = errw + LBA_BWARPS;                       // [n6]        the pose update of the previous iteration
    double *s_T = s_x + n6;                                // [N][16]     keyframe poses
    int *s_opt = reinterpret_cast<int *>(s_T + (size_t)16 * d.n_frames);   // [N]  frame -> optimisable index
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n_smem = (int)(errw + LBA_BWARPS - smem);
    {   // zero fill, 16 bytes per store (the panels are dense: keyframes a landmark does not see stay zero)
        double2 *z = reinterpret_cast<double2 *>(smem);
        for (int i = tid; i < n_smem / 2; i += LBA_BTHREADS) z[i] = make_double2(0.0, 0.0);
        if (tid == 0 && (n_smem & 1)) smem[n_smem - 1] = 0.0;
    }
    // the per-frame tables every observation looks up: once per tile from global memory instead of a dependent L2 round trip
    // (frame -> index -> pose) in each of a landmark's three passes
    for (int i = tid; i < n6; i += LBA_BTHREADS) s_x[i] = apply_update ? d.x[i] : 0.0;
    for (int i = tid; i < 16 * d.n_frames; i += LBA_BTHREADS) s_T[i] = d.poses[i];
    for (int i = tid; i < d.n_frames; i += LBA_BTHREADS) s_opt[i] = d.opt_index[i];
    __syncthreads();

    const int tile_base = blockIdx.x * TL;
    const int grp = lane / GS, gl = lane % GS;
    double *myA = Aacc + (size_t)wid * No * LBA_NA;
    double *myB = Btmp + ((size_t)wid * NG + grp) * No * 18;
    double err_w = 0.0;

    for (int il0 = wid * NG; il0 < TL; il0 += LBA_BWARPS * NG) {
        if (tile_base + il0 >= d.n_points) break;                    // warp-uniform: every group is past the end
        const int il = il0 + grp;
        const int i = tile_base + il;
        const bool valid = il < TL && i < d.n_points;                // this group has a landmark in this round
        const int o_beg = valid ? d.obs_ptr[i] : 0, o_end = valid ? d.obs_ptr[i + 1] : 0;
        const int n_chunks = __reduce_max_sync(0xffffffffu, (o_end - o_beg + GS - 1) / GS);      // warp-uniform trip count
        double Xi[3] = {0.0, 0.0, 1.0};
        if (valid) { Xi[0] = d.points[3 * i]; Xi[1] = d.points[3 * i + 1]; Xi[2] = d.points[3 * i + 2]; }

        // ---- apply the previous iteration's landmark update: X_i += Cinv_b_i - sum_j CinvBt[i][j] x_j
        if (apply_update) {
            double cb[3] = {0, 0, 0};
            for (int o = o_beg + gl; o < o_end; o += GS) {
                if (d.obs_right[o]) continue;
                const int j = s_opt[lba_frame(d, o)];
                if (j < 0) continue;
                const double *BC = d.bcinv + (size_t)o * 18;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    double s = 0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) s += BC[c * 3 + r] * s_x[6 * j + c];
                    cb[r] += s;
                }
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                cb[r] = gsum<GS>(cb[r]);
                if (valid) Xi[r] += d.cinv_b[3 * i + r] - cb[r];
            }
            __syncwarp();
            if (gl == 0 && valid) { d.points[3 * i] = Xi[0]; d.points[3 * i + 1] = Xi[1]; d.points[3 * i + 2] = Xi[2]; }
        }

        // ---- observations: lane per observation (chunks of 32)
        double Cs[6] = {0, 0, 0, 0, 0, 0}, bs[3] = {0, 0, 0};
        for (int ch = 0; ch < n_chunks; ++ch) {
            const int o = o_beg + ch * GS + gl;
            const bool act = o < o_end;
            int j = -1, right = 0;
            double Rij[6], rij[2] = {0, 0}, weight = 1.0, Qm[12];
            if (act) {
                const int f = lba_frame(d, o);
                right = d.obs_right[o];
                j = s_opt[f];
                const double *T = s_T + 16 * f;
                double R_jw[9], t_jw[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) { R_jw[r * 3] = T[r * 4]; R_jw[r * 3 + 1] = T[r * 4 + 1]; R_jw[r * 3 + 2] = T[r * 4 + 2]; t_jw[r] = T[r * 4 + 3]; }
                double Xij[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) Xij[r] = ((R_jw[r * 3] * Xi[0] + R_jw[r * 3 + 1] * Xi[1]) + R_jw[r * 3 + 2] * Xi[2]) + t_jw[r];
                double Rm[9], Xc[3], fx, fy, cx, cy;
                if (right) {
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            Rm[r * 3 + c] = (d.R_rl[r * 3] * R_jw[c] + d.R_rl[r * 3 + 1] * R_jw[3 + c]) + d.R_rl[r * 3 + 2] * R_jw[6 + c];
#pragma unroll
                    for (int r = 0; r < 3; ++r) Xc[r] = ((d.R_rl[r * 3] * Xij[0] + d.R_rl[r * 3 + 1] * Xij[1]) + d.R_rl[r * 3 + 2] * Xij[2]) + d.t_rl[r];
                    fx = d.K_r[0]; fy = d.K_r[1]; cx = d.K_r[2]; cy = d.K_r[3];
                } else {
#pragma unroll
                    for (int k = 0; k < 9; ++k) Rm[k] = R_jw[k];
                    Xc[0] = Xij[0]; Xc[1] = Xij[1]; Xc[2] = Xij[2];
                    fx = d.K_l[0]; fy = d.K_l[1]; cx = d.K_l[2]; cy = d.K_l[3];
                }
                const double invz = 1.0 / Xc[2];
                const double fxinvz = fx * invz, fyinvz = fy * invz, xinvz = Xc[0] * invz, yinvz = Xc[1] * invz;
                const double fx_xinvz2 = fxinvz * xinvz, fy_yinvz2 = fyinvz * yinvz, xinvz_yinvz = xinvz * yinvz;
                rij[0] = (fx * xinvz + cx) - d.obs_px[2 * o];
                rij[1] = (fy * yinvz + cy) - d.obs_px[2 * o + 1];
                const double absrxry = fabs(rij[0]) + fabs(rij[1]);
                if (absrxry > d.huber) weight = d.huber / absrxry;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    Rij[c] = fxinvz * Rm[c] - fx_xinvz2 * Rm[6 + c];
                    Rij[3 + c] = fyinvz * Rm[3 + c] - fy_yinvz2 * Rm[6 + c];
                }
                Cs[0] += weight * (Rij[0] * Rij[0] + Rij[3] * Rij[3]);
                Cs[1] += weight * (Rij[0] * Rij[1] + Rij[3] * Rij[4]);
                Cs[2] += weight * (Rij[0] * Rij[2] + Rij[3] * Rij[5]);
                Cs[3] += weight * (Rij[1] * Rij[1] + Rij[4] * Rij[4]);
                Cs[4] += weight * (Rij[1] * Rij[2] + Rij[4] * Rij[5]);
                Cs[5] += weight * (Rij[2] * Rij[2] + Rij[5] * Rij[5]);
#pragma unroll
                for (int c = 0; c < 3; ++c) bs[c] += -(weight * (Rij[c] * rij[0] + Rij[3 + c] * rij[1]));
                err_w += rij[0] * rij[0] + rij[1] * rij[1];
                if (j >= 0) {
                    if (right) {
                        const double dp[6] = {fxinvz, 0, -fx_xinvz2, 0, fyinvz, -fy_yinvz2};
                        double dpR[6];
#pragma unroll
                        for (int r = 0; r < 2; ++r)
#pragma unroll
                            for (int c = 0; c < 3; ++c)
                                dpR[r * 3 + c] = (dp[r * 3] * d.R_rl[c] + dp[r * 3 + 1] * d.R_rl[3 + c]) + dp[r * 3 + 2] * d.R_rl[6 + c];
                        const double sk[9] = {0, -Xij[2], Xij[1], Xij[2], 0, -Xij[0], -Xij[1], Xij[0], 0};
#pragma unroll
                        for (int r = 0; r < 2; ++r)
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                Qm[r * 6 + c] = dpR[r * 3 + c];
                                Qm[r * 6 + 3 + c] = ((-dpR[r * 3]) * sk[c] + (-dpR[r * 3 + 1]) * sk[3 + c]) + (-dpR[r * 3 + 2]) * sk[6 + c];
                            }
                    } else {
                        Qm[0] = fxinvz; Qm[1] = 0; Qm[2] = -fx_xinvz2; Qm[3] = -fx * xinvz_yinvz; Qm[4] = fx * (1.0 + xinvz * xinvz); Qm[5] = -fx * yinvz;
                        Qm[6] = 0; Qm[7] = fyinvz; Qm[8] = -fy_yinvz2; Qm[9] = -fy * (1.0 + yinvz * yinvz); Qm[10] = fy * xinvz_yinvz; Qm[11] = fy * xinvz;
                    }
                }
            }
            // A_j / a_j: left-camera lanes first, then right-camera lanes -- within one pass a keyframe
            // occurs at most once per landmark, so the per-warp accumulators see no write conflict
            // and the summation order is fixed.
#pragma unroll 1
            for (int gp = 0; gp < 2 * NG; ++gp) {
                const int pass = gp & 1;
                if (grp == (gp >> 1) && act && j >= 0 && right == pass) {
                    double wa[12];
#pragma unroll
                    for (int k = 0; k < 12; ++k) wa[k] = weight * Qm[k];
                    double *Aj = myA + (size_t)j * LBA_NA;
                    // upper triangle in row order (21 entries), shortcut of calc_Qij_t_Qij_weight
                    int idx = 0;
#pragma unroll
                    for (int r = 0; r < 6; ++r)
#pragma unroll
                        for (int c = r; c < 6; ++c, ++idx) {
                            double v;
                            if (r == 0) v = (c == 1) ? 0.0 : wa[0] * Qm[c];
                            else if (r == 1) v = wa[7] * Qm[6 + c];
                            else v = wa[r] * Qm[c] + wa[6 + r] * Qm[6 + c];
                            Aj[idx] += v;
                        }
#pragma unroll
                    for (int r = 0; r < 6; ++r) Aj[21 + r] += -(weight * (Qm[r] * rij[0] + Qm[6 + r] * rij[1]));
                }
                __syncwarp();
            }
            // B[j][i]: last writer (highest observation index) of each optimisable keyframe
            {
                const int key = (act && j >= 0) ? grp * LBA_MAX_OPT + j : -1 - lane;
                const unsigned peers = __match_any_sync(0xffffffffu, key);
                const bool last = act && j >= 0 && (lane == 31 - __clz(peers));
                if (last) {
                    double *Bj = myB + (size_t)j * 18;
#pragma unroll
                    for (int r = 0; r < 6; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c) Bj[r * 3 + c] = weight * (Qm[r] * Rij[c] + Qm[6 + r] * Rij[3 + c]);
                }
                __syncwarp();
            }
        }
        // ---- C_i, b_i across lanes; damping; inverse; Cinv_b
        double C[9], b[3];
        {
            double c6[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) c6[k] = gsum<GS>(Cs[k]);
#pragma unroll
            for (int k = 0; k < 3; ++k) b[k] = gsum<GS>(bs[k]);
            C[0] = c6[0]; C[1] = c6[1]; C[2] = c6[2]; C[3] = c6[1]; C[4] = c6[3]; C[5] = c6[4]; C[6] = c6[2]; C[7] = c6[4]; C[8] = c6[5];
        }
        C[0] += d.lambda * C[0]; C[4] += d.lambda * C[4]; C[8] += d.lambda * C[8];
        double Cinv[9];
        ldlt3_inverse(C, Cinv);
        double cib[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) cib[r] = (Cinv[r * 3] * b[0] + Cinv[r * 3 + 1] * b[1]) + Cinv[r * 3 + 2] * b[2];
        if (gl == 0 && valid) { d.cinv_b[3 * i] = cib[0]; d.cinv_b[3 * i + 1] = cib[1]; d.cinv_b[3 * i + 2] = cib[2]; }
        if (gl < 3 && valid) Q[(size_t)(3 * il + gl) * (n6 + 1) + n6] = b[gl];
        // ---- BCinv for the left-camera observations of optimisable keyframes; stage P and Q
        for (int o = o_beg + gl; o < o_end; o += GS) {
            if (!d.obs_right[o]) {
                const int j = s_opt[lba_frame(d, o)];
                if (j >= 0) {
                    const double *Bj = myB + (size_t)j * 18;
                    double BC[18];
#pragma unroll
                    for (int r = 0; r < 6; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            BC[r * 3 + c] = (Bj[r * 3] * Cinv[c] + Bj[r * 3 + 1] * Cinv[3 + c]) + Bj[r * 3 + 2] * Cinv[6 + c];
                    double *g = d.bcinv + (size_t)o * 18;
#pragma unroll
                    for (int k = 0; k < 18; ++k) g[k] = BC[k];
#pragma unroll
                    for (int m = 0; m < 3; ++m)
#pragma unroll
                        for (int r = 0; r < 6; ++r) {
                            P[(size_t)(3 * il + m) * n6 + 6 * j + r] = BC[r * 3 + m];
                            Q[(size_t)(3 * il + m) * (n6 + 1) + 6 * j + r] = Bj[r * 3 + m];
                        }
                }
            }
        }
        __syncwarp();
        // clear the last-writer scratch for the next landmark of this warp
        for (int k = gl; k < No * 18; k += GS) myB[k] = 0.0;
        __syncwarp();
    }
    err_w = wsum(err_w);
    if (lane == 0) errw[wid] = err_w;
    __syncthreads();

    // ---- tile contribution to the reduced camera system: out[r][c] = sum_k P[k][r] * Q[k][c]
    const int ncol = n6 + 1;
    double *out = d.s_part + (size_t)blockIdx.x * n6 * ncol;
    const int K3 = 3 * TL;
    // Register-tiled: a work item is a 2 x 3 patch of one 6 x 6 block (jb <= kb: the upper block triangle) or a 6 x 1 slice
    // of the right-hand-side column; per k it loads 2 + 3 (6 + 1) panel entries for 6 multiply-adds on six independent
    // accumulators (the first version ran one dependent 3 TL-long dot product per entry, two loads per multiply-add:
    // 41 % of the kernel's instructions, profiles/r2_lba_build_*).  Every entry is still summed over k in ascending order.
    {
        const int n_pairs = No * (No + 1) / 2;
        const int n_items = n_pairs * 6 + No;
        for (int e = tid; e < n6 * ncol; e += LBA_BTHREADS) {      // entries outside the upper block triangle: zero
            const int r = e / ncol, c = e - r * ncol;
            if (c != n6 && (c / 6) < (r / 6)) out[e] = 0.0;
        }
        for (int it = tid; it < n_items; it += LBA_BTHREADS) {
            if (it < n_pairs * 6) {
                const int pair = it / 6, sub = it - pair * 6;       // sub: 3 row pairs x 2 column triples
                int jb = 0, rem = pair;
                while (rem >= No - jb) { rem -= No - jb; ++jb; }
                const int kb = jb + rem;
                const int r0 = 6 * jb + 2 * (sub >> 1), c0 = 6 * kb + 3 * (sub & 1);
                double a00 = 0, a01 = 0, a02 = 0, a10 = 0, a11 = 0, a12 = 0;
                const double *pp = P + r0, *qq = Q + c0;
#pragma unroll 4
                for (int k = 0; k < K3; ++k) {
                    const double p0 = pp[0], p1 = pp[1], q0 = qq[0], q1 = qq[1], q2 = qq[2];
                    a00 += p0 * q0; a01 += p0 * q1; a02 += p0 * q2;
                    a10 += p1 * q0; a11 += p1 * q1; a12 += p1 * q2;
                    pp += n6; qq += ncol;
                }
                double *o = out + (size_t)r0 * ncol + c0;
                o[0] = a00; o[1] = a01; o[2] = a02;
                o[ncol] = a10; o[ncol + 1] = a11; o[ncol + 2] = a12;
            } else {
                const int jb = it - n_pairs * 6;
                double a[6] = {0, 0, 0, 0, 0, 0};
                const double *pp = P + 6 * jb, *qq = Q + n6;
#pragma unroll 4
                for (int k = 0; k < K3; ++k) {
                    const double q = qq[0];
#pragma unroll
                    for (int r = 0; r < 6; ++r) a[r] += pp[r] * q;
                    pp += n6; qq += ncol;
                }
#pragma unroll
                for (int r = 0; r < 6; ++r) out[(size_t)(6 * jb + r) * ncol + n6] = a[r];
            }
        }
    }
    for (int e = tid; e < No * LBA_NA; e += LBA_BTHREADS) {
        double acc = 0.0;
        for (int w = 0; w < LBA_BWARPS; ++w) acc += Aacc[(size_t)w * No * LBA_NA + e];
        d.a_part[(size_t)blockIdx.x * No * LBA_NA + e] = acc;
    }
    if (tid == 0) {
        double acc = 0.0;
        for (int w = 0; w < LBA_BWARPS; ++w) acc += errw[w];
        d.err_part[blockIdx.x] = acc;
    }
}

// ------------------------------------------------------------------ k_lba_reduce
// Sums of the per-tile partial systems in a FIXED order: eight lanes per entry, lane g adds the tiles g, g + 8, ... in
// sequence (independent loads in flight), then the eight partials are added in lane order.  Coalesced across entries.
// (Inside the single solve CTA this reduction was a 100-microsecond chain of dependent L2 loads per iteration; one thread
// per entry over 133 tiles still took 17 us.)
#define LBA_RED_G 8
__global__ void __launch_bounds__(256) k_lba_reduce(const LbaDev d)
{
    const int n = d.n6, ncol = n + 1, nS = n * ncol, nA = d.n_opt * LBA_NA;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    // entries vary fastest across the 32 / LBA_RED_G = 4 ... keep an entry's lanes in one warp: e = warp-local mapping
    const int warp = gid >> 5, lane = gid & 31;
    const int e = warp * (32 / LBA_RED_G) + (lane % (32 / LBA_RED_G)), g = lane / (32 / LBA_RED_G);
    const bool act = e < nS + nA + 1;
    const double *src = d.err_part; size_t stride = 1;
    if (act) {
        if (e < nS) { src = d.s_part + e; stride = (size_t)nS; }
        else if (e < nS + nA) { src = d.a_part + (e - nS); stride = (size_t)nA; }
    }
    double acc = 0.0;
    if (act) {
        int t = g;
        for (; t + 3 * LBA_RED_G < d.n_tiles; t += 4 * LBA_RED_G) {
            const double v0 = src[(size_t)t * stride], v1 = src[(size_t)(t + LBA_RED_G) * stride], v2 = src[(size_t)(t + 2 * LBA_RED_G) * stride],
                         v3 = src[(size_t)(t + 3 * LBA_RED_G) * stride];
            acc += v0; acc += v1; acc += v2; acc += v3;
        }
        for (; t < d.n_tiles; t += LBA_RED_G) acc += src[(size_t)t * stride];
    }
    // lane order g = 0, 1, ..., 7 (lanes of one entry are 4 apart)
    double tot = __shfl_sync(0xffffffffu, acc, lane % (32 / LBA_RED_G));
#pragma unroll
    for (int k = 1; k < LBA_RED_G; ++k) tot += __shfl_sync(0xffffffffu, acc, (lane % (32 / LBA_RED_G)) + k * (32 / LBA_RED_G));
    if (act && g == 0) d.red[e] = tot;
    // observation count of this shard, summed over the ranks with everything else in the distributed variant
    if (gid == 0) d.red[nS + nA + 1] = (double)d.n_obs;
}

// ------------------------------------------------------------------ k_lba_solve (one CTA)
// Reduced camera system S = blkdiag(A damped) - BCinvBt (with the reference's mirror quirk), pivoted LDLT, solve, pose
// retraction.  Round 1 factorised S in shared memory: 48 steps of {pivot search, swap, scale, trailing update}, four
// block barriers and a shared-memory read-modify-write of the whole trailing triangle per step -- 98 k cycles per LM
// iteration (VO_LBA_TRACE), the longest stage of the local BA.  Here the matrix lives in REGISTERS: the CTA is a 16 x 16
// thread grid, thread (ty, tx) owns the B x B block of rows ty*B.. and columns tx*B.. of the FULL symmetric matrix (B = 3
// for n <= 48, i.e. the reference's window of 8 optimised keyframes; B = 6 up to n = 96).  Nothing is ever swapped:
// diagonal pivoting only chooses the ORDER of elimination, so step k eliminates original index p_k in place.  A step:
//   1. every warp holds the whole diagonal and a position table in registers (lane = index mod 32) and finds Eigen's pivot
//      (largest remaining |diagonal|, FIRST POSITION on ties -- the table mirrors the transpositions Eigen would have made)
//      with one redux on the leading 32 bits; only if two lanes tie there do the second redux and the position redux run.
//      The lane-local candidate's reciprocal is computed in the shadow of the reduction and shuffled with the pivot.
//   2. the owners of column p publish l = L(:,p) and t = D_p l (double-buffered by step parity) and park the factor column;
//      the pivot's row and column are set to ZERO in the registers, so every later update is an unconditional
//      A -= u * v with no liveness masks (the in-block column index is a warp-uniform switch: registers cannot be indexed).
//   3. ONE named barrier, then every thread updates its own registers (A(i,j) -= L(i,p) * (D_p L(j,p)), the products of
//      Eigen's unblocked LDLT, fused) and its copy of the diagonal.
// History (VO_LBA_TRACE cycles of the factorisation at n = 48): shared-memory version 98 k; 6 x 6 register blocks on two
// warps with a shared diagonal / position table and two barriers per step 72 k; register-resident search, one barrier 56 k;
// zeroed retired rows instead of masks, fused updates 51 k; 3 x 3 blocks on eight warps + single-redux search 51 k.
// What a step costs now (ablation with the pieces switched off, tools/probe notes in DESIGN.md): ~550 cycles of dependent
// register / shuffle work (select, redux, ballot, shuffle, reciprocal, scale: ~7 cycles per dependent ALU operation with one
// or two warps per scheduler), ~200 for the barrier, ~110 for the column switch, ~170 for the shared-memory round trip.
// The triangular solves then run on one warp over the factor parked in shared memory, in original index order (no
// permutation of the right-hand side).
#define SOLVE_G 16

// S, rhs, Aj in shared memory from the reduced tile sums (d.red), exactly as sparse_bundle_adjustment.cpp:456-531 leaves them
__device__ __forceinline__ void lba_assemble(const LbaDev &d, double *S, const int ld, double *rhs, double *Aj, double *s_err, const int tid)
{
    const int n = d.n6, No = d.n_opt, ncol = n + 1;
#define SM(i, j) S[(size_t)(i) * ld + (j)]
    for (int e = tid; e < n * ncol; e += LBA_THREADS) {
        const double acc = d.red[e];
        const int r = e / ncol, c = e - r * ncol;
        if (c == n) rhs[r] = acc; else SM(r, c) = acc;
    }
    for (int e = tid; e < No * LBA_NA; e += LBA_THREADS) Aj[e] = d.red[n * ncol + e];
    if (tid == 0) *s_err = d.red[n * ncol + No * LBA_NA];
    __syncthreads();
    // S <- blkdiag(A damped) - BCinvBt with the reference's mirror (upper blocks -> lower, diagonal
    // blocks transposed, sparse_bundle_adjustment.cpp:501-512); rhs <- a - BCinv_b
    for (int e = tid; e < n * n; e += LBA_THREADS) {
        const int r = e / n, c = e - r * n;
        const int jb = r / 6, kb = c / 6;
        if (jb > kb) continue;                      // handle each upper-block entry once
        const double up = SM(r, c);                 // BCinvBt[jb][kb](r%6, c%6), untouched so far
        if (jb < kb) SM(c, r) = -up;                // lower block = transpose
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += LBA_THREADS) {
        const int r = e / n, c = e - r * n;
        const int jb = r / 6, kb = c / 6;
        if (jb < kb) SM(r, c) = -SM(r, c);
    }
    __syncthreads();
    // diagonal blocks: A_j(r,c) - BCinvBt[j][j]^T(r,c)
    for (int e = tid; e < No * 36; e += LBA_THREADS) {
        const int jb = e / 36, rr = (e % 36) / 6, cc = e % 6;
        if (rr > cc) continue;                      // one thread handles the (rr,cc)/(cc,rr) pair
        const int lo = rr, hi = cc;
        const int idx = lo * 6 - lo * (lo - 1) / 2 + (hi - lo);      // upper index of (lo,hi) in the 21-entry packing
        double a_v = Aj[jb * LBA_NA + idx];
        if (lo == hi) a_v += d.lambda * a_v;        // damping A(k,k) += lambda*A(k,k)
        const double b_rc = SM(6 * jb + rr, 6 * jb + cc), b_cr = SM(6 * jb + cc, 6 * jb + rr);
        SM(6 * jb + rr, 6 * jb + cc) = a_v - b_cr;  // transposed diagonal block
        if (rr != cc) SM(6 * jb + cc, 6 * jb + rr) = a_v - b_rc;
    }
    for (int r = tid; r < n; r += LBA_THREADS) rhs[r] = Aj[(r / 6) * LBA_NA + 21 + (r % 6)] - rhs[r];
    __syncthreads();
#undef SM
}

__device__ __forceinline__ unsigned redux_max_u32(unsigned v)
{
    unsigned r;
    asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ int redux_min_s32(int v)
{
    int r;
    asm volatile("redux.sync.min.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}

// One elimination step's column work with the pivot's in-block column PA known at compile time (register arrays cannot be
// indexed dynamically; the caller switches on the warp-uniform value).  Owners of column p (tx == pb) publish
// L(:,p) = A(:,p) * rk and D_p L(:,p), park the factor column, and retire the column; owners of row p (ty == pb) retire the
// row.  Retired rows / columns are ZERO in the registers from then on (0 - 0 * t stays 0), so no liveness masks are needed.
template <int B, int PA>
__device__ __forceinline__ void ldlt_column_step(double (&A)[B][B], bool col_owner, bool row_owner, double rk, double akk,
                                                 double *l_rows, double *t_rows, double *S_col, int LD)
{
    if (col_owner) {
        double l[B], t[B];
#pragma unroll
        for (int a = 0; a < B; ++a) {
            const double c = (a == PA && row_owner) ? 0.0 : A[a][PA];      // the pivot's own row takes no update
            l[a] = c * rk;
            t[a] = akk * l[a];
            A[a][PA] = 0.0;
        }
        if constexpr (B % 2 == 0) {
#pragma unroll
            for (int h = 0; h < B / 2; ++h) {
                reinterpret_cast<double2 *>(l_rows)[h] = make_double2(l[2 * h], l[2 * h + 1]);
                reinterpret_cast<double2 *>(t_rows)[h] = make_double2(t[2 * h], t[2 * h + 1]);
            }
        } else {
#pragma unroll
            for (int a = 0; a < B; ++a) { l_rows[a] = l[a]; t_rows[a] = t[a]; }
        }
#pragma unroll
        for (int a = 0; a < B; ++a) S_col[(size_t)a * LD] = l[a];           // rows eliminated earlier store a zero nobody reads
    }
    if (row_owner) {
#pragma unroll
        for (int b = 0; b < B; ++b) A[PA][b] = 0.0;
    }
}

template <int G, int B>      // G x G threads own B x B register blocks of the (G*B)-padded system
__global__ void __launch_bounds__(LBA_THREADS, 1)
k_lba_solve(const LbaDev d, int iter)
{
    constexpr int NP = G * B;            // padded system size
    constexpr int LD = NP + 1;           // shared-memory row pitch (odd: conflict-free column walks)
    extern __shared__ double smem[];
    const int n = d.n6, No = d.n_opt;
    double *S = smem;                    // [NP][LD]: assembly, then (column by column) the factor L
    double *rhs = S + (size_t)NP * LD;   // [NP]
    double *Aj = rhs + NP;               // [No][27]
    __shared__ __align__(16) double s_lt[4 * NP];      // L(:,p) and D_p L(:,p) of the current step, double-buffered by step parity: [l 0 | l 1 | t 0 | t 1]
    __shared__ double s_D[NP];
    __shared__ int s_piv[NP];            // step -> original index
    __shared__ int s_step[NP];           // original index -> step
    __shared__ double s_err;
    const int tid = threadIdx.x, lane = tid & 31;
    const bool act = tid < G * G;        // threads of the factorisation grid (the other warps wait at the barrier behind the loop)
    const int ty = tid / G, tx = tid % G;
#define SM(i, j) S[(size_t)(i) * LD + (j)]
#define STAMP(k) do { if (d.dbg && tid == 0) d.dbg[k] = clock64(); } while (0)
    STAMP(0);
    lba_assemble(d, S, LD, rhs, Aj, &s_err, tid);
    STAMP(1);

    // ---- registers <- the LOWER triangle (what Eigen::LDLT<.., Lower> reads), mirrored to a full symmetric matrix
    double A[B][B];
    // Every active warp also keeps, redundantly and in REGISTERS, the whole diagonal (lane owns the original indices lane,
    // lane + 32, ...) and the POSITION each index would have after the transpositions Eigen applies (its tie rule is "first
    // position"); -1 = eliminated or padding.  The pivot search then needs no shared memory and no second barrier.
    constexpr int NPS = (NP + 31) / 32;
    double dg[NPS];
    int pos[NPS];
    if (act) {
#pragma unroll
        for (int a = 0; a < B; ++a) {
            const int i = ty * B + a;
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const int j = tx * B + b;
                A[a][b] = (i < n && j < n) ? (i >= j ? SM(i, j) : SM(j, i)) : 0.0;
            }
        }
#pragma unroll
        for (int q = 0; q < NPS; ++q) {
            const int i = lane + 32 * q;
            dg[q] = i < n ? SM(i, i) : 0.0;
            pos[q] = i < n ? i : -1;
        }
    }
    if (tid < NP) { s_step[tid] = -1; s_D[tid] = 0.0; }
    // Which of {l, t = D l} a thread multiplies is fixed by its block position: element (i, j) takes L(i,p) * (D_p L(j,p)) with
    // i >= j (the lower-triangle product) and its mirror image otherwise; inside a diagonal block the split is a >= b.
    const int oUA = (ty >= tx) ? 0 : 2 * NP, oVA = (ty >= tx) ? 2 * NP : 0;         // l or t: offsets into s_lt relative to the step's l buffer
    const int oUB = (ty > tx) ? 0 : 2 * NP, oVB = (ty > tx) ? 2 * NP : 0;
    __syncthreads();                     // every register copy is made before the factor overwrites S

    // One barrier per step.  After the barrier of step k every thread holds L(:,p_k) and D L(:,p_k); the register diagonal is
    // updated first (one fused multiply-subtract per owned index), the search for step k+1 runs on it, and the 36 independent
    // fused multiply-subtracts of the trailing update are issued in the shadow of the search's reductions and of the pivot
    // reciprocal.  (No clock stamps inside the loop: they pin the instruction order and undo exactly that overlap.)
    if (act) {
        for (int k = 0; k < n; ++k) {
            // ---- pivot of step k: largest |diagonal| among the live indices, first POSITION on ties (every active warp).
            // Keys are the bit patterns of |d| (monotonic for non-negative doubles); eliminated / padding / NaN entries and
            // exact zeros carry key 0 and never win -- if nothing wins the pivot is the index at position k, as in Eigen.
            // lane-local candidate (value key, position, which of the lane's entries) ...
            unsigned long long bk = 0ull;
            int bp = 0x7fffffff, bq = 0;
            double bd = dg[0];
#pragma unroll
            for (int q = 0; q < NPS; ++q) {
                const double v = fabs(dg[q]);
                const unsigned long long key = (pos[q] >= 0 && v == v) ? (unsigned long long)__double_as_longlong(v) : 0ull;
                const bool take = key > bk || (key == bk && key != 0ull && pos[q] < bp);
                bk = take ? key : bk;
                bp = take ? pos[q] : bp;
                bq = take ? q : bq;
                bd = take ? dg[q] : bd;
            }
            // ... and its reciprocal, started now so that it runs in the shadow of the warp reductions: two Newton steps and a
            // final correction on the hardware approximation (within one ulp of the quotient Eigen forms; no slow-path call)
            double br;
            {
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(br) : "d"(bd));
                double e = __fma_rn(-bd, br, 1.0);
                br = __fma_rn(br, e, br);
                e = __fma_rn(-bd, br, 1.0);
                br = __fma_rn(br, e, br);
                e = __fma_rn(-bd, br, 1.0);
                br = __fma_rn(br, e, br);
            }
            const unsigned hi = (unsigned)(bk >> 32);
            const unsigned mhi = redux_max_u32(hi);
            unsigned owners = __ballot_sync(0xffffffffu, bk != 0ull && hi == mhi);
            if (__popc(owners) != 1) {
                // rare: several lanes share the leading 32 bits, or nothing is comparable (all zero / NaN: Eigen keeps position k)
                const unsigned lo = (bk != 0ull && hi == mhi) ? (unsigned)bk : 0u;
                const unsigned mlo = redux_max_u32(lo);
                const bool win = bk != 0ull && hi == mhi && (unsigned)bk == mlo;
                int big = redux_min_s32(win ? bp : 0x7fffffff);
                big = big == 0x7fffffff ? k : big;
                int my_q = -1;
#pragma unroll
                for (int q = 0; q < NPS; ++q) my_q = (pos[q] == big) ? q : my_q;
                owners = __ballot_sync(0xffffffffu, my_q >= 0);
                if (my_q >= 0 && (my_q != bq || bp != big)) {          // the owner's entry is not its lane-local candidate
                    bq = my_q; bp = big;
                    bd = dg[0];
#pragma unroll
                    for (int q = 1; q < NPS; ++q) bd = (my_q == q) ? dg[q] : bd;
                    br = 1.0 / bd;
                }
            }
            const int src = __ffs(owners) - 1;
            const int p = __shfl_sync(0xffffffffu, lane + 32 * bq, src);
            const int big_pos = __shfl_sync(0xffffffffu, bp, src);
            const double akk = __shfl_sync(0xffffffffu, bd, src);
            const double rcp_akk = __shfl_sync(0xffffffffu, br, src);
            // the transposition: the pivot leaves the table, the index that sat at position k takes the pivot's old position
#pragma unroll
            for (int q = 0; q < NPS; ++q) {
                const int moved = (pos[q] == k) ? big_pos : pos[q];
                pos[q] = (lane == src && q == bq) ? -1 : moved;
            }
            const int pb = p / B, pa = p - pb * B;
            // L(:,p) = A(:,p) * (1 / D_p).  Eigen leaves the column unscaled at the last step and for an exactly zero pivot: factor 1
            const double rk = ((k < n - 1) && fabs(akk) > 0.0) ? rcp_akk : 1.0;
            double *bl = s_lt + (k & 1) * NP, *bt = bl + 2 * NP;
            {
                const bool co = tx == pb, ro = ty == pb;
                double *lr = bl + ty * B, *tr = bt + ty * B, *sc = &SM(ty * B, p);
                switch (pa) {
                case 0: ldlt_column_step<B, 0>(A, co, ro, rk, akk, lr, tr, sc, LD); break;
                case 1: ldlt_column_step<B, 1>(A, co, ro, rk, akk, lr, tr, sc, LD); break;
                case 2: ldlt_column_step<B, 2>(A, co, ro, rk, akk, lr, tr, sc, LD); break;
                case 3: ldlt_column_step<B, (3 < B ? 3 : 0)>(A, co, ro, rk, akk, lr, tr, sc, LD); break;
                case 4: ldlt_column_step<B, (4 < B ? 4 : 0)>(A, co, ro, rk, akk, lr, tr, sc, LD); break;
                default: ldlt_column_step<B, (5 < B ? 5 : 0)>(A, co, ro, rk, akk, lr, tr, sc, LD); break;
                }
            }
            if (tid == 0) { s_piv[k] = p; s_step[p] = k; s_D[p] = akk; }
            asm volatile("bar.sync 1, %0;" ::"n"(G * G) : "memory");    // column visible (the idle warps are not involved)
            // ---- the diagonal first (all the next search needs), then the trailing update
#pragma unroll
            for (int q = 0; q < NPS; ++q) {
                const int i = lane + 32 * q;
                if (i < NP) dg[q] = __fma_rn(-bl[i], bt[i], dg[q]);
            }
            double uA[B], vA[B], uB[B], vB[B];
#pragma unroll
            for (int a = 0; a < B; ++a) {
                uA[a] = bl[oUA + ty * B + a]; uB[a] = bl[oUB + ty * B + a];
                vA[a] = bl[oVA + tx * B + a]; vB[a] = bl[oVB + tx * B + a];
            }
#pragma unroll
            for (int a = 0; a < B; ++a)
#pragma unroll
                for (int b = 0; b < B; ++b) A[a][b] = (a >= b) ? __fma_rn(-uA[a], vA[b], A[a][b]) : __fma_rn(-uB[a], vB[b], A[a][b]);
        }
    }
    __syncthreads();
    STAMP(2);
    const double tol = 1.0 / 1.7976931348623157e308;
    if (tid < 32) {
        // The triangular solves on one warp with the right-hand side in REGISTERS (lane owns indices lane, lane + 32, ...): a
        // step is one shuffle of the finished component plus one multiply-subtract per owned index; the factor entries and
        // the pivot list are plain shared-memory loads off the dependent chain (the first version kept the right-hand side
        // in shared memory: 400 cycles of load -> multiply -> store -> barrier per step, 38 k cycles per LM iteration).
        constexpr int NPL = (NP + 31) / 32;
        double r[NPL];
        int st[NPL];
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
            const int i = lane + 32 * q;
            r[q] = i < n ? rhs[i] : 0.0;
            st[q] = i < n ? s_step[i] : -1;
        }
        // forward: for k ascending, y(p_k) is final; every index eliminated later takes its update
#pragma unroll 4
        for (int k = 0; k < n; ++k) {
            const int p = s_piv[k];
            double v = r[0];
#pragma unroll
            for (int q = 1; q < NPL; ++q) v = (p >> 5) == q ? r[q] : v;
            const double yk = __shfl_sync(0xffffffffu, v, p & 31);
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                const int i = lane + 32 * q;
                if (st[q] > k) r[q] -= SM(i, p) * yk;
            }
        }
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
            const int i = lane + 32 * q;
            if (i < n) { const double D = s_D[i]; r[q] = fabs(D) > tol ? r[q] / D : 0.0; }
        }
        // backward: L^T x = y, column-oriented: once x(p_k) is final, every index eliminated EARLIER takes its update
#pragma unroll 4
        for (int k = n - 1; k > 0; --k) {
            const int p = s_piv[k];
            double v = r[0];
#pragma unroll
            for (int q = 1; q < NPL; ++q) v = (p >> 5) == q ? r[q] : v;
            const double xk = __shfl_sync(0xffffffffu, v, p & 31);
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                const int i = lane + 32 * q;
                if (st[q] >= 0 && st[q] < k) r[q] -= SM(p, i) * xk;
            }
        }
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
            const int i = lane + 32 * q;
            if (i < n) rhs[i] = r[q];
        }
    }
    __syncthreads();
    STAMP(3);
    for (int i = tid; i < n; i += LBA_THREADS) d.x[i] = rhs[i];
    // ---- pose retraction of the optimisable keyframes + error bookkeeping
    if (tid < No) {
        double T[16], xj[6];
        double *Tg = d.poses + 16 * d.opt_frame[tid];
        for (int k = 0; k < 16; ++k) T[k] = Tg[k];
        for (int k = 0; k < 6; ++k) xj[k] = rhs[6 * tid + k];
        pose_retract(T, xj);
        for (int k = 0; k < 16; ++k) Tg[k] = T[k];
    }
    STAMP(4);
    if (tid == 0) {
        const double n_obs_all = d.red[n * (n + 1) + No * LBA_NA + 1];     // == d.n_obs on one GPU; the all-reduced count otherwise
        d.avg_err[iter] = sqrt(s_err / n_obs_all);
        if (isnan(s_err)) atomicExch(d.nan_flag, 1);
    }
#undef SM
#undef STAMP
}

// final landmark update after the last iteration
__global__ void __launch_bounds__(256) k_lba_update_points(const LbaDev d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n_points) return;
    double cb[3] = {0, 0, 0};
    for (int o = d.obs_ptr[i]; o < d.obs_ptr[i + 1]; ++o) {
        if (d.obs_right[o]) continue;
        const int j = d.opt_index[lba_frame(d, o)];
        if (j < 0) continue;
        const double *BC = d.bcinv + (size_t)o * 18;
        for (int r = 0; r < 3; ++r) {
            double s = 0;
            for (int c = 0; c < 6; ++c) s += BC[c * 3 + r] * d.x[6 * j + c];
            cb[r] += s;
        }
    }
    for (int r = 0; r < 3; ++r) d.points[3 * i + r] += d.cinv_b[3 * i + r] - cb[r];
}

// ------------------------------------------------------------------ host
static void inv_se3_host_d(const double *T, double *O)
{
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) O[i * 4 + j] = T[j * 4 + i];
        double s = 0;
        for (int k = 0; k < 3; ++k) s += T[k * 4 + i] * T[k * 4 + 3];
        O[i * 4 + 3] = -s;
    }
    O[12] = O[13] = O[14] = 0.0; O[15] = 1.0;
}

static size_t a256(size_t v) { return (v + 255) / 256 * 256; }

// ------------------------------------------------------------------ NCCL (loaded at run time; only the distributed entry needs it)
// Oversize windows (north_star / SURVEY 8e): the LANDMARKS are partitioned over the ranks, the keyframe poses are replicated.
// Each rank builds the partial reduced camera system of its landmarks; ONE ncclAllReduce(sum, double) of
// (6 N_opt)(6 N_opt + 1) + 27 N_opt + 2 values per LM iteration -- enqueued on the LBA stream right after k_lba_reduce, so it
// rides NVLink between the build of this iteration and the redundant dense solve -- makes the system complete on every rank;
// the solve, the pose retraction and the back-substitution of the rank's own landmarks stay local.
#include <dlfcn.h>
typedef struct ncclComm *vo_ncclComm_t;
typedef struct { char internal[128]; } vo_ncclUniqueId;
struct NcclApi {
    int (*GetUniqueId)(vo_ncclUniqueId *) = nullptr;
    int (*CommInitRank)(vo_ncclComm_t *, int, vo_ncclUniqueId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, vo_ncclComm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(vo_ncclComm_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
static const NcclApi &nccl_api()
{
    static NcclApi api = [] {
        NcclApi a;
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);      // the copy torch already loaded, else the system one
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return a;
        a.GetUniqueId = (int (*)(vo_ncclUniqueId *))dlsym(h, "ncclGetUniqueId");
        a.CommInitRank = (int (*)(vo_ncclComm_t *, int, vo_ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
        a.AllReduce = (int (*)(const void *, void *, size_t, int, int, vo_ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
        a.CommDestroy = (int (*)(vo_ncclComm_t))dlsym(h, "ncclCommDestroy");
        a.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy;
        return a;
    }();
    return api;
}
#define VO_NCCL_DOUBLE 8     /* ncclFloat64 */
#define VO_NCCL_SUM 0        /* ncclSum */

extern "C" int vo_dist_unique_id(void *id_out)
{
    if (!id_out || !nccl_api().ok) return VO_ERR_INVALID_ARG;
    vo_ncclUniqueId id;
    if (nccl_api().GetUniqueId(&id) != 0) return VO_ERR_CUDA;
    memcpy(id_out, &id, sizeof(id));
    return VO_OK;
}

extern "C" int vo_dist_init(vo_ctx *ctx, int rank, int world, const void *unique_id)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(unique_id && world >= 1 && rank >= 0 && rank < world, VO_ERR_INVALID_ARG, "vo_dist_init: bad rank / world / id");
    VO_REQUIRE(nccl_api().ok, VO_ERR_CUDA, "libnccl.so.2 could not be loaded");
    VO_REQUIRE(!ctx->nccl_comm, VO_ERR_INVALID_ARG, "vo_dist_init: communicator already initialised");
    VO_CUDA(cudaSetDevice(ctx->device));
    vo_ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    vo_ncclComm_t comm = nullptr;
    const int rc = nccl_api().CommInitRank(&comm, world, id, rank);
    if (rc != 0) {
        ctx->last_error = std::string("ncclCommInitRank: ") + (nccl_api().GetErrorString ? nccl_api().GetErrorString(rc) : "error");
        return VO_ERR_CUDA;
    }
    ctx->nccl_comm = comm; ctx->dist_rank = rank; ctx->dist_world = world;
    return VO_OK;
}

extern "C" int vo_dist_finalize(vo_ctx *ctx)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    if (ctx->nccl_comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        nccl_api().CommDestroy((vo_ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr; ctx->dist_world = 1; ctx->dist_rank = 0;
    }
    return VO_OK;
}

static int lba_solve_impl(vo_ctx *ctx, const vo_lba_problem *p, double *poses_out, double *points_out, double *avg_err_out, int *success,
                          bool dist);

extern "C" int vo_lba_reserve(vo_ctx *ctx, size_t bytes)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_CUDA(cudaSetDevice(ctx->device));
    if (bytes <= ctx->lba_bytes) return VO_OK;
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_lba) cudaFree(ctx->d_lba);
    ctx->d_lba = nullptr; ctx->lba_bytes = 0;
    VO_CUDA(cudaMalloc(&ctx->d_lba, bytes));
    ctx->lba_bytes = bytes;
    // one tiny well-posed window through the whole path: function attributes, first launches of the three kernels and the
    // staging buffers are then out of the way before the first real keyframe (it was 0.9 ms slower than the next ones)
    {
        const int N = 3, M = 8;
        double poses[N * 16], pts[M * 3], px[M * N * 2], poses_o[N * 16], pts_o[M * 3], avg[2];
        int opt_index[N] = {-1, -1, 0}, obs_ptr[M + 1], obs_frame[M * N];
        uint8_t obs_right[M * N];
        for (int f = 0; f < N; ++f)
            for (int k = 0; k < 16; ++k) poses[f * 16 + k] = (k % 5 == 0) ? 1.0 : (k == 3 ? -0.1 * f : 0.0);     // T_jw: camera f sits at x = 0.1 f
        for (int i = 0; i < M; ++i) {
            pts[3 * i] = 0.2 * (i % 3) - 0.2; pts[3 * i + 1] = 0.15 * (i / 3) - 0.15; pts[3 * i + 2] = 1.0 + 0.1 * i;
            obs_ptr[i] = i * N;
            for (int f = 0; f < N; ++f) {
                const int o = i * N + f;
                obs_frame[o] = f; obs_right[o] = 0;
                px[2 * o] = 700.0 * (pts[3 * i] - 0.1 * f) / pts[3 * i + 2] + 600.0 + 0.3 * ((i + f) % 3 - 1);
                px[2 * o + 1] = 700.0 * pts[3 * i + 1] / pts[3 * i + 2] + 180.0;
            }
        }
        obs_ptr[M] = M * N;
        vo_lba_problem w;
        memset(&w, 0, sizeof(w));
        w.n_frames = N; w.n_opt = 1; w.n_points = M; w.n_obs = M * N;
        w.poses = poses; w.opt_index = opt_index; w.points = pts; w.obs_ptr = obs_ptr; w.obs_frame = obs_frame; w.obs_right = obs_right; w.obs_px = px;
        w.K_l[0] = w.K_r[0] = 700; w.K_l[1] = w.K_r[1] = 700; w.K_l[2] = w.K_r[2] = 600; w.K_l[3] = w.K_r[3] = 180;
        for (int k = 0; k < 16; ++k) w.T_lr[k] = (k % 5 == 0) ? 1.0 : 0.0;
        w.is_stereo = 0; w.huber = 0.5; w.lambda = 1e-5; w.max_iter = 2;
        int ok = 0;
        const int rc = lba_solve_impl(ctx, &w, poses_o, pts_o, avg, &ok, false);
        if (rc != VO_OK && rc != VO_ERR_NAN) return rc;       // the warm-up's numbers do not matter, a CUDA failure does
    }
    return VO_OK;
}

extern "C" int vo_lba_solve(vo_ctx *ctx, const vo_lba_problem *p, double *poses_out, double *points_out,
                            double *avg_err_out, int *success)
{
    return lba_solve_impl(ctx, p, poses_out, points_out, avg_err_out, success, false);
}

extern "C" int vo_lba_solve_dist(vo_ctx *ctx, const vo_lba_problem *local_part, double *poses_out, double *points_out,
                                 double *avg_err_out, int *success)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(ctx->nccl_comm != nullptr, VO_ERR_INVALID_ARG, "vo_lba_solve_dist: call vo_dist_init first");
    return lba_solve_impl(ctx, local_part, poses_out, points_out, avg_err_out, success, true);
}

// Deferred part of the problem validation (see lba_solve_impl): nullptr if the observation lists are well formed.
static const char *lba_check_observations(const vo_lba_problem *p)
{
    const int N = p->n_frames, M = p->n_points;
    for (int i = 0; i < M; ++i) {
        int last_left_opt = -1;
        unsigned seen_l = 0, seen_r = 0;
        for (int o = p->obs_ptr[i]; o < p->obs_ptr[i + 1]; ++o) {
            const int f = p->obs_frame[o];
            if (f < 0 || f >= N) return "obs_frame out of range";
            const int j = p->opt_index[f];
            if (j < 0) continue;
            unsigned &seen = p->obs_right[o] ? seen_r : seen_l;
            if (seen & (1u << j)) return "duplicate (landmark, keyframe, camera) observation";
            seen |= 1u << j;
            if (!p->obs_right[o]) {
                // the reference accumulates BCinvBt[j][k] for observation order jj <= kk and then mirrors
                // the upper block triangle: only chronological (ascending opt index) lists are well defined
                if (j <= last_left_opt) return "left observations must be in ascending keyframe order";
                last_left_opt = j;
            }
        }
    }
    return nullptr;
}

static int lba_solve_impl(vo_ctx *ctx, const vo_lba_problem *p, double *poses_out, double *points_out, double *avg_err_out, int *success,
                          bool dist)
{
    if (!ctx || !p) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(p->n_frames > 0 && p->n_points >= 0 && p->n_obs >= 0 && p->max_iter >= 1, VO_ERR_INVALID_ARG, "bad sizes");
    VO_REQUIRE(p->n_opt >= 1 && p->n_opt <= LBA_MAX_OPT, VO_ERR_INVALID_ARG, "n_opt must be in [1, 16]");
    VO_REQUIRE(p->poses && p->opt_index && p->points && p->obs_ptr && p->obs_frame && p->obs_right && p->obs_px && poses_out && points_out,
               VO_ERR_INVALID_ARG, "null pointer");
    const int N = p->n_frames, No = p->n_opt, M = p->n_points, n_obs = p->n_obs, n6 = 6 * No;
    static const bool trace = getenv("VO_LBA_TRACE") != nullptr;
    const auto h_t0 = std::chrono::steady_clock::now();
    auto h_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h_t0).count(); };
    double h_valid = 0, h_stage = 0, h_enq = 0;
    // host-side validation of the index structure (cheap, O(n_obs)); the kernels rely on it
    std::vector<int> opt_frame(No, -1);
    for (int f = 0; f < N; ++f) {
        const int j = p->opt_index[f];
        VO_REQUIRE(j >= -1 && j < No, VO_ERR_INVALID_ARG, "opt_index out of range");
        if (j >= 0) { VO_REQUIRE(opt_frame[j] < 0, VO_ERR_INVALID_ARG, "duplicate opt_index"); opt_frame[j] = f; }
    }
    for (int j = 0; j < No; ++j) VO_REQUIRE(opt_frame[j] >= 0, VO_ERR_INVALID_ARG, "opt_index does not cover [0, n_opt)");
    VO_REQUIRE(p->obs_ptr[0] == 0 && p->obs_ptr[M] == n_obs, VO_ERR_SIZE_MISMATCH, "obs_ptr inconsistent with n_obs");
    for (int i = 0; i < M; ++i) VO_REQUIRE(p->obs_ptr[i + 1] >= p->obs_ptr[i], VO_ERR_INVALID_ARG, "obs_ptr not monotone");
    // The O(n_obs) part of the validation (frame range, duplicates, chronological left observations: 2.4 ns per observation,
    // 0.11 - 0.17 ms at the benchmark windows) runs AFTER the work is enqueued, while the GPU iterates; the kernels clamp
    // the frame index they read, so a malformed list cannot make them touch memory they do not own, and a violation found
    // here discards their output.
    h_valid = h_ms();
    VO_CUDA(cudaSetDevice(ctx->device));

    // Tile size and landmarks per warp.  About one tile per SM (more tiles only cost a longer k_lba_reduce); inside a tile
    // the 16 warps take NG landmarks each per round, so NG is the smallest of {1, 2, 4} that covers the tile in one round --
    // as far as the shared-memory budget allows (the last-writer scratch grows with NG, the panels with the tile).
    const size_t budget = 227 * 1024;          // opt-in maximum of dynamic shared memory per CTA on sm_100
    const size_t per_lm = (size_t)3 * (2 * n6 + 1) * 8;
    auto fixed_of = [&](int ng) {
        return ((size_t)LBA_BWARPS * No * LBA_NA + (size_t)LBA_BWARPS * ng * No * 18 + LBA_BWARPS + n6 + (size_t)16 * N) * 8 + (size_t)N * 4 + 8;
    };
    VO_REQUIRE(fixed_of(1) + per_lm <= budget, VO_ERR_INVALID_ARG, "window too large for the shared-memory tile");
    static int n_sm = 0;
    if (!n_sm) { cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device); if (n_sm <= 0) n_sm = 148; }
    int want = vo_div_up(M > 0 ? M : 1, n_sm);
    if (want > 64) want = 64;
    int NG = 1, TL = 1;
    {
        long best_rounds = -1;
        for (int ng = 1; ng <= 4; ng *= 2) {
            if (fixed_of(ng) + per_lm > budget) break;
            int tl = (int)((budget - fixed_of(ng)) / per_lm);
            if (tl > want) tl = want;
            // cost model: waves of tiles over the SMs x rounds per tile (a round is one landmark latency)
            const long rounds = (long)vo_div_up(vo_div_up(M > 0 ? M : 1, tl), n_sm) * vo_div_up(tl, LBA_BWARPS * ng);
            if (best_rounds < 0 || rounds < best_rounds) { best_rounds = rounds; NG = ng; TL = tl; }
        }
    }
    const size_t fixed = fixed_of(NG);
    const int n_tiles = M > 0 ? vo_div_up(M, TL) : 1;
    const size_t smem_build = fixed + per_lm * TL;
    const int solve_B = n6 <= 48 ? 3 : 6;      // k_lba_solve<16, B>: 16 x 16 threads own B x B register blocks
    const int solve_NP = 16 * solve_B;
    const size_t smem_solve = ((size_t)solve_NP * (solve_NP + 1) + solve_NP + (size_t)No * LBA_NA) * 8;

    // device scratch (one allocation, grow-only)
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += a256(bytes); return o; };
    const size_t o_poses = take((size_t)N * 128), o_opt = take((size_t)N * 4), o_optf = take((size_t)No * 4);
    const size_t o_pts = take((size_t)M * 24), o_ptr = take((size_t)(M + 1) * 4), o_of = take((size_t)n_obs * 4);
    const size_t o_or = take((size_t)n_obs), o_px = take((size_t)n_obs * 16);
    const size_t in_bytes = off;
    const size_t o_bc = take((size_t)n_obs * 144), o_cb = take((size_t)M * 24);
    const size_t o_sp = take((size_t)n_tiles * n6 * (n6 + 1) * 8), o_ap = take((size_t)n_tiles * No * LBA_NA * 8);
    const size_t o_ep = take((size_t)n_tiles * 8), o_x = take((size_t)n6 * 8);
    const size_t n_red = (size_t)n6 * (n6 + 1) + (size_t)No * LBA_NA + 2;      // reduced system + A blocks + error + observation count
    const size_t o_red = take(n_red * 8);
    const size_t o_ae = take((size_t)p->max_iter * 8), o_nan = take(16), o_dbg = take(128);
    const size_t total = off;
    if (total > ctx->lba_bytes) {
        VO_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_lba) cudaFree(ctx->d_lba);
        ctx->d_lba = nullptr; ctx->lba_bytes = 0;
        // grow geometrically: window problems grow by a few landmarks per keyframe and a cudaFree + cudaMalloc
        // pair costs milliseconds
        // (and cudaFree synchronises the whole device, stalling every other context's stream): start at 32 MB,
        // enough for ~10 k landmarks with 10 stereo observations each
        size_t want = total + total / 2;
        if (want < ((size_t)32 << 20)) want = (size_t)32 << 20;
        VO_CUDA(cudaMalloc(&ctx->d_lba, want));
        ctx->lba_bytes = want;
    }
    int rc = vo_stage_reserve(ctx, in_bytes > (size_t)N * 128 + (size_t)M * 24 + 4096 ? in_bytes : (size_t)N * 128 + (size_t)M * 24 + 4096);
    if (rc) return rc;
    uint8_t *h = ctx->h_stage, *dv = (uint8_t *)ctx->d_lba;
    memcpy(h + o_poses, p->poses, (size_t)N * 128);
    memcpy(h + o_opt, p->opt_index, (size_t)N * 4);
    memcpy(h + o_optf, opt_frame.data(), (size_t)No * 4);
    memcpy(h + o_pts, p->points, (size_t)M * 24);
    memcpy(h + o_ptr, p->obs_ptr, (size_t)(M + 1) * 4);
    memcpy(h + o_of, p->obs_frame, (size_t)n_obs * 4);
    memcpy(h + o_or, p->obs_right, (size_t)n_obs);
    memcpy(h + o_px, p->obs_px, (size_t)n_obs * 16);
    h_stage = h_ms();
    VO_CUDA(cudaMemcpyAsync(dv, h, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    VO_CUDA(cudaMemsetAsync(dv + o_bc, 0, total - o_bc, ctx->stream));

    LbaDev d;
    d.n_frames = N; d.n_opt = No; d.n_points = M; d.n_obs = n_obs; d.n_tiles = n_tiles; d.tile = TL; d.n6 = n6;
    d.poses = (double *)(dv + o_poses); d.opt_index = (const int *)(dv + o_opt); d.opt_frame = (int *)(dv + o_optf);
    d.points = (double *)(dv + o_pts); d.obs_ptr = (const int *)(dv + o_ptr); d.obs_frame = (const int *)(dv + o_of);
    d.obs_right = dv + o_or; d.obs_px = (const double *)(dv + o_px);
    d.bcinv = (double *)(dv + o_bc); d.cinv_b = (double *)(dv + o_cb); d.s_part = (double *)(dv + o_sp);
    d.a_part = (double *)(dv + o_ap); d.err_part = (double *)(dv + o_ep); d.x = (double *)(dv + o_x); d.red = (double *)(dv + o_red);
    d.avg_err = (double *)(dv + o_ae); d.nan_flag = (int *)(dv + o_nan);
    d.dbg = getenv("VO_LBA_TRACE") ? (long long *)(dv + o_dbg) : nullptr;
    memcpy(d.K_l, p->K_l, 32); memcpy(d.K_r, p->K_r, 32);
    double T_rl[16];
    inv_se3_host_d(p->T_lr, T_rl);   // geometry::inverseSE3(T_lr), sparse_bundle_adjustment.cpp:174
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) d.R_rl[i * 3 + j] = T_rl[i * 4 + j]; d.t_rl[i] = T_rl[i * 4 + 3]; }
    d.huber = p->huber; d.lambda = p->lambda;

    {   // Function attributes are process-wide: set them ONCE to the largest value any problem can need (several
        // contexts may solve concurrently from different host threads; a per-call value would race).  One carve-out
        // for the three kernels, which alternate 21 times per solve.
        static std::once_flag once;
        static cudaError_t attr_err = cudaSuccess;
        std::call_once(once, [&]() {
            const size_t max_solve = ((size_t)(SOLVE_G * 6) * (SOLVE_G * 6 + 1) + SOLVE_G * 6 + (size_t)LBA_MAX_OPT * LBA_NA) * 8;
            cudaError_t e = cudaFuncSetAttribute(k_lba_build<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lba_build<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lba_build<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lba_solve<16, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_solve);
            cudaFuncSetAttribute(k_lba_build<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_lba_build<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_lba_build<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_lba_solve<16, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_lba_solve<16, 6>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_lba_update_points, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            attr_err = e;
        });
        if (attr_err != cudaSuccess) { ctx->last_error = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(attr_err); return VO_ERR_CUDA; }
    }
    std::vector<cudaEvent_t> evs;
    if (trace) { evs.resize(2 * p->max_iter + 1); for (auto &e : evs) cudaEventCreate(&e); cudaEventRecord(evs[0], ctx->stream); }
    for (int it = 0; it < p->max_iter; ++it) {
        if (NG == 4) k_lba_build<4><<<n_tiles, LBA_BTHREADS, smem_build, ctx->stream>>>(d, it > 0 ? 1 : 0);
        else if (NG == 2) k_lba_build<2><<<n_tiles, LBA_BTHREADS, smem_build, ctx->stream>>>(d, it > 0 ? 1 : 0);
        else k_lba_build<1><<<n_tiles, LBA_BTHREADS, smem_build, ctx->stream>>>(d, it > 0 ? 1 : 0);
        k_lba_reduce<<<vo_div_up((n6 * (n6 + 1) + No * LBA_NA + 1) * LBA_RED_G, 256), 256, 0, ctx->stream>>>(d);
        if (dist) {
            // the one exchange step of the path: partial reduced systems of the ranks' landmark shards -> complete system everywhere
            const int nrc = nccl_api().AllReduce(d.red, d.red, n_red, VO_NCCL_DOUBLE, VO_NCCL_SUM, (vo_ncclComm_t)ctx->nccl_comm, ctx->stream);
            if (nrc != 0) { ctx->last_error = std::string("ncclAllReduce: ") + (nccl_api().GetErrorString ? nccl_api().GetErrorString(nrc) : "error"); return VO_ERR_CUDA; }
        }
        if (trace) cudaEventRecord(evs[2 * it + 1], ctx->stream);
        if (solve_B == 3) k_lba_solve<16, 3><<<1, LBA_THREADS, smem_solve, ctx->stream>>>(d, it);
        else k_lba_solve<16, 6><<<1, LBA_THREADS, smem_solve, ctx->stream>>>(d, it);
        if (trace) cudaEventRecord(evs[2 * it + 2], ctx->stream);
        ctx->launches += 3;
    }
    if (trace) {
        cudaStreamSynchronize(ctx->stream);
        for (int it = 0; it < p->max_iter; ++it) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, evs[2 * it], evs[2 * it + 1]); cudaEventElapsedTime(&b, evs[2 * it + 1], evs[2 * it + 2]);
            fprintf(stderr, "lba it %d: build %.1f us, solve %.1f us (tiles %d x %d landmarks, %d per warp, smem %zu / %zu)\n", it, a * 1e3f, b * 1e3f, n_tiles, TL, NG, smem_build, smem_solve);
        }
        for (auto &e : evs) cudaEventDestroy(e);
        long long st[8];
        cudaMemcpy(st, dv + o_dbg, 40, cudaMemcpyDeviceToHost);
        fprintf(stderr, "k_lba_solve cycles: assemble %lld, ldlt %lld, solves %lld, retract %lld\n", st[1] - st[0], st[2] - st[1], st[3] - st[2], st[4] - st[3]);
    }
    if (M > 0) { k_lba_update_points<<<vo_div_up(M, 256), 256, 0, ctx->stream>>>(d); ctx->launches++; }
    VO_CUDA(cudaGetLastError());
    // results: poses + points (+ avg_err, nan flag) through the pinned staging
    const size_t r_poses = 0, r_pts = a256((size_t)N * 128), r_ae = r_pts + a256((size_t)M * 24), r_nan = r_ae + a256((size_t)p->max_iter * 8);
    rc = vo_stage_reserve(ctx, r_nan + 256);
    if (rc) return rc;
    h = ctx->h_stage;
    VO_CUDA(cudaMemcpyAsync(h + r_poses, dv + o_poses, (size_t)N * 128, cudaMemcpyDeviceToHost, ctx->stream));
    if (M > 0) VO_CUDA(cudaMemcpyAsync(h + r_pts, dv + o_pts, (size_t)M * 24, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaMemcpyAsync(h + r_ae, dv + o_ae, (size_t)p->max_iter * 8, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaMemcpyAsync(h + r_nan, dv + o_nan, 4, cudaMemcpyDeviceToHost, ctx->stream));
    h_enq = h_ms();
    const char *bad = lba_check_observations(p);
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (bad) { ctx->last_error = bad; return VO_ERR_INVALID_ARG; }
    if (trace) fprintf(stderr, "vo_lba_solve host ms: validated at %.3f, staged at %.3f, enqueued at %.3f, results at %.3f\n", h_valid, h_stage, h_enq, h_ms());
    memcpy(poses_out, h + r_poses, (size_t)N * 128);
    if (M > 0) memcpy(points_out, h + r_pts, (size_t)M * 24);
    const double *ae = (const double *)(h + r_ae);
    if (avg_err_out) memcpy(avg_err_out, ae, (size_t)p->max_iter * 8);
    const bool nan = *(int *)(h + r_nan) != 0;
    if (success) *success = (!nan && ae[p->max_iter - 1] <= 1.0) ? 1 : 0;
    if (nan) { ctx->last_error = "Local BA NAN!"; return VO_ERR_NAN; }   // sparse_bundle_adjustment.cpp:624,761
    return VO_OK;
}
