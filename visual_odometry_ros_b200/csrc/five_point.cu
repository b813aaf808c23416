// five_point.cu -- M1: MotionEstimator::calcPose5PointsAlgorithm (core/visual_odometry/motion_estimator.cpp:21-123) and
// findCorrectRT (:205-263), device resident.
//
// The reference calls cv::findEssentialMat(pts0, pts1, K, cv::RANSAC, 0.999, thres_5p) (:41; OpenCV is third-party and not
// under /root/reference), decomposes E with an SVD (:70-96) and keeps the (R, t) of the four candidates that puts most
// DLT-triangulated points in front of both cameras (:205-263).  Here:
//   k_5pt_norm    normalised coordinates of all correspondences (FP64, as OpenCV converts them)
//   k_5pt_solve   one thread per hypothesis: 5 distinct correspondences from a counter-based hash of (seed, hypothesis),
//                 Nister's minimal solver -- null space of the 5x9 epipolar system, the ten cubic constraints
//                 det(E) = 0 and 2 E E^T E - tr(E E^T) E = 0 as a 10x20 matrix, Gauss-Jordan, the 3x3 polynomial matrix in z,
//                 its degree-10 determinant, real roots by monotone-interval safeguarded Newton on [-1, 1] (polynomial and its
//                 reversal) -- up to 10 essential matrices
//   k_5pt_score   one block per hypothesis: the error OpenCV's RANSAC uses, (x1^T E x0)^2 / (|E x0|_xy^2 + |E^T x1|_xy^2)
//                 on normalised coordinates against (thres / mean focal)^2, for every candidate over ALL correspondences;
//                 the best (count, first hypothesis, first candidate) is kept with one 64-bit atomicMax
//   k_5pt_decomp  SVD decomposition of the best E into the four (R, t) candidates
//   k_5pt_cheir   one thread per correspondence: inlier bit of the best E and the four cheirality tests
//                 (mapping::triangulateDLT, core/util/triangulate_3d.cpp:5-50), counted with block + global atomics
//   k_5pt_final   the winning candidate: R10 / t10 / X0 / mask
// All hypotheses are evaluated in parallel (a fixed number, default 1024) instead of OpenCV's sequential, adaptively
// terminated loop: same model class, same error, same acceptance rule; the random samples differ, so parity with the
// reference is statistical -- tests/test_five_point_gpu.py states the criteria.
// Compiled with -fmad=false: the FP32 triangulation follows the reference's operation order; the FP64 solver uses fma().
#include "vo_internal.cuh"
#include "tri_device.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

struct FpDev {
    const float2 *p0, *p1;
    int n;
    const int *n_d;          // optional device-resident count (overrides n)
    double fx, fy, cx, cy;
    double thr2;
    int H;
    unsigned seed;
    double4 *q;              // [n] (x0, y0, x1, y1) normalised
    double *Es;              // [H][10][9]
    int *nsol;               // [H]
    unsigned long long *best;
    float *R10, *t10, *X0, *E_out;
    uint8_t *mask;
    int *info;               // 0 RANSAC inliers, 1 cheirality inliers of the chosen candidate, 2 ok
    float K[4];
    struct FpPose *pose;     // decomposition of the best model + counters
    uint8_t *bits;           // [n] bit 0 RANSAC inlier, bits 1..4 cheirality of the four candidates
    float *X0c;              // [n][4][3] triangulated points of the four candidates
};

struct FpPose {
    double E[9];
    float R[2][9], t[3];
    int valid;
    int cnt[5];              // RANSAC inliers, cheirality counts of the four candidates
};

__device__ __forceinline__ int fp_count(const FpDev &d) { return d.n_d ? min(*d.n_d, d.n) : d.n; }

__global__ void __launch_bounds__(256) k_5pt_norm(const FpDev d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *d.best = 0ull;
    if (i >= fp_count(d)) return;
    const float2 a = d.p0[i], b = d.p1[i];
    d.q[i] = make_double4(((double)a.x - d.cx) / d.fx, ((double)a.y - d.cy) / d.fy, ((double)b.x - d.cx) / d.fx, ((double)b.y - d.cy) / d.fy);
}

// ------------------------------------------------------------------------------ polynomial helpers
// degree-1: {x, y, z, 1}; degree-2: {xx, yy, zz, xy, xz, yz, x, y, z, 1};
// degree-3 (Nister's elimination order): {x3, y3, x2y, xy2, x2z, x2, y2z, y2, xyz, xy | xz2, xz, x, yz2, yz, y, z3, z2, z, 1}
__constant__ signed char c_M2[4][4] = {{0, 3, 4, 6}, {3, 1, 5, 7}, {4, 5, 2, 8}, {6, 7, 8, 9}};
__constant__ signed char c_M3[10][4] = {{0, 2, 4, 5}, {3, 1, 6, 7}, {10, 13, 16, 17}, {2, 3, 8, 9}, {4, 8, 10, 11},
                                        {8, 6, 13, 14}, {5, 9, 11, 12}, {9, 7, 14, 15}, {11, 14, 17, 18}, {12, 15, 18, 19}};

__device__ __forceinline__ void p1p1_acc(const double *a, const double *b, double *o, double s)
{
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) o[c_M2[i][j]] = fma(s * a[i], b[j], o[c_M2[i][j]]);
}
__device__ __forceinline__ void p2p1_acc(const double *a, const double *b, double *o, double s)
{
    for (int i = 0; i < 10; ++i)
        for (int j = 0; j < 4; ++j) o[c_M3[i][j]] = fma(s * a[i], b[j], o[c_M3[i][j]]);
}
__device__ __forceinline__ double horner(const double *c, int m, double x)
{
    double r = c[m];
    for (int i = m - 1; i >= 0; --i) r = fma(r, x, c[i]);
    return r;
}
// o (deg da + db) = a * b, polynomials in z, ascending coefficients
__device__ __forceinline__ void zmul(const double *a, int da, const double *b, int db, double *o)
{
    for (int i = 0; i <= da + db; ++i) o[i] = 0.0;
    for (int i = 0; i <= da; ++i)
        for (int j = 0; j <= db; ++j) o[i + j] = fma(a[i], b[j], o[i + j]);
}

// (i + k)! / i!: coefficient i of the k-th derivative is c[i + k] * c_dfact[k][i]
__constant__ double c_dfact[11][11] = {
    {1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0},
    {1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.0, 10.0, 0.0},
    {2.0, 6.0, 12.0, 20.0, 30.0, 42.0, 56.0, 72.0, 90.0, 0.0, 0.0},
    {6.0, 24.0, 60.0, 120.0, 210.0, 336.0, 504.0, 720.0, 0.0, 0.0, 0.0},
    {24.0, 120.0, 360.0, 840.0, 1680.0, 3024.0, 5040.0, 0.0, 0.0, 0.0, 0.0},
    {120.0, 720.0, 2520.0, 6720.0, 15120.0, 30240.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {720.0, 5040.0, 20160.0, 60480.0, 151200.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {5040.0, 40320.0, 181440.0, 604800.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {40320.0, 362880.0, 1814400.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {362880.0, 3628800.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {3628800.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0}};

// p(x) and p'(x) of a degree-10 polynomial held in registers (coefficients above the actual degree are zero)
__device__ __forceinline__ void horner11(const double (&d)[11], double x, double &p, double &dp)
{
    p = d[10]; dp = 0.0;
#pragma unroll
    for (int i = 9; i >= 0; --i) { dp = fma(dp, x, p); p = fma(p, x, d[i]); }
}

// Real roots of a polynomial of degree <= 10 inside [-1, 1]: the roots of the k-th derivative split the interval into pieces
// on which the (k-1)-th derivative is monotone, so every sign change brackets exactly one root (safeguarded Newton).
// 32 hypotheses run in lockstep in a warp: the level's coefficients sit in registers and every evaluation is the same
// fully unrolled 10-step Horner pair, so lanes with different root counts share one instruction stream.
__device__ int roots_unit(const double *c, double *roots, bool closed)
{
    double ra[10], rb[10];
    int na = 0;
    for (int k = 9; k >= 0; --k) {
        double d[11];
#pragma unroll
        for (int i = 0; i <= 10; ++i) d[i] = (i + k <= 10) ? c[min(i + k, 10)] * c_dfact[k][i] : 0.0;
        int nb = 0;
        double lo = -1.0, flo, tmp;
        horner11(d, lo, flo, tmp);
        if (flo == 0.0 && (closed || k > 0)) rb[nb++] = lo;
        for (int j = 0; j <= na; ++j) {
            const double hi = j < na ? ra[j] : 1.0;
            double fhi;
            horner11(d, hi, fhi, tmp);
            if (fhi == 0.0) {
                if (nb < 10 && (j < na || closed || k > 0)) rb[nb++] = hi;
            } else if (flo != 0.0 && ((flo < 0.0) != (fhi < 0.0))) {
                // safeguarded Newton inside the bracket (bisection whenever the Newton step leaves it or stalls)
                double a = lo, b = hi;
                const bool neg_a = flo < 0.0;
                double x = 0.5 * (a + b), dx_old = b - a, dx = dx_old;
                for (int it = 0; it < 100; ++it) {
                    double p, dp;
                    horner11(d, x, p, dp);
                    if (p == 0.0) break;
                    if ((p < 0.0) == neg_a) a = x; else b = x;
                    const double tol = 4.0e-16;                                        // |x| <= 1
                    const double step = dp != 0.0 ? p / dp : 0.0;
                    if ((dp != 0.0 && fabs(step) <= tol) || b - a <= tol) break;       // converged (x is one end of the bracket)
                    const double xn = x - step;
                    const bool newton_ok = dp != 0.0 && xn > a && xn < b && fabs(2.0 * p) <= fabs(dx_old * dp);
                    dx_old = dx;
                    const double x_new = newton_ok ? xn : 0.5 * (a + b);
                    dx = x_new - x;
                    if (x_new == x) break;
                    x = x_new;
                }
                if (nb < 10) rb[nb++] = x;
            }
            lo = hi; flo = fhi;
        }
        na = nb;
        for (int j = 0; j < nb; ++j) ra[j] = rb[j];
    }
    for (int j = 0; j < na; ++j) roots[j] = ra[j];
    return na;
}

// All real roots of a degree-10 polynomial: those in [-1, 1] directly, the others as reciprocals of the roots of the
// reversed polynomial in (-1, 1).  Both searches live on a bounded interval: no Cauchy-bound brackets to bisect through.
__device__ int real_roots(const double *c_in, int deg, double *roots)
{
    double c[11], cr[11];
    double cmax = 0.0;
    for (int i = 0; i <= deg; ++i) cmax = fmax(cmax, fabs(c_in[i]));
    if (!(cmax > 0.0) || !isfinite(cmax)) return 0;
    for (int i = 0; i <= 10; ++i) c[i] = i <= deg ? c_in[i] / cmax : 0.0;
    for (int i = 0; i <= 10; ++i) cr[i] = c[10 - i];
    int n = roots_unit(c, roots, true);
    double w[10];
    const int nw = roots_unit(cr, w, false);
    for (int j = 0; j < nw && n < 10; ++j)
        if (fabs(w[j]) > 1e-12 && fabs(w[j]) < 1.0) roots[n++] = 1.0 / w[j];           // w = 0 <=> a root at infinity (degree < 10)
    return n;
}

// Nister's five-point minimal solver.  q[i] = (x0, y0, x1, y1) with x1^T E x0 = 0.  E_out[s][9] row-major, Frobenius
// norm 1; returns the number of real solutions (<= 10).
__device__ int solve5(const double4 *q, double *E_out, long long *trace = nullptr)
{
#define FP_STAMP(k) do { if (trace) trace[k] = clock64(); } while (0)
    FP_STAMP(0);
    // ---- null space of the 5x9 system: Gauss-Jordan with full pivoting
    double Q[5][9];
    for (int i = 0; i < 5; ++i) {
        const double x0 = q[i].x, y0 = q[i].y, x1 = q[i].z, y1 = q[i].w;
        Q[i][0] = x1 * x0; Q[i][1] = x1 * y0; Q[i][2] = x1;
        Q[i][3] = y1 * x0; Q[i][4] = y1 * y0; Q[i][5] = y1;
        Q[i][6] = x0; Q[i][7] = y0; Q[i][8] = 1.0;
    }
    int pc[5];
    unsigned used = 0;
    for (int r = 0; r < 5; ++r) {
        int bi = r, bj = -1;
        double bv = 0.0;
        for (int i = r; i < 5; ++i)
            for (int j = 0; j < 9; ++j)
                if (!((used >> j) & 1) && fabs(Q[i][j]) > bv) { bv = fabs(Q[i][j]); bi = i; bj = j; }
        if (bj < 0 || bv < 1e-14) return 0;                          // degenerate sample
        for (int j = 0; j < 9; ++j) { const double t = Q[r][j]; Q[r][j] = Q[bi][j]; Q[bi][j] = t; }
        pc[r] = bj; used |= 1u << bj;
        const double inv = 1.0 / Q[r][bj];
        for (int j = 0; j < 9; ++j) Q[r][j] *= inv;
        for (int i = 0; i < 5; ++i) {
            if (i == r) continue;
            const double f = Q[i][bj];
            if (f != 0.0) for (int j = 0; j < 9; ++j) Q[i][j] = fma(-f, Q[r][j], Q[i][j]);
        }
    }
    FP_STAMP(1);
    double Ep[9][4];                                                 // entry -> {X, Y, Z, W} coefficients
    {
        int k = 0;
        for (int f = 0; f < 9; ++f) {
            if ((used >> f) & 1) continue;
            double nrm = 1.0;
            for (int r = 0; r < 5; ++r) nrm = fma(Q[r][f], Q[r][f], nrm);
            const double s = rsqrt(nrm);
            for (int e = 0; e < 9; ++e) Ep[e][k] = 0.0;
            Ep[f][k] = s;
            for (int r = 0; r < 5; ++r) Ep[pc[r]][k] = -Q[r][f] * s;
            ++k;
        }
    }
    // ---- the ten cubic constraints
    double A[10][20];
    for (int r = 0; r < 10; ++r)
        for (int c = 0; c < 20; ++c) A[r][c] = 0.0;
    {
        double t2[10];
        // det(E)
        for (int c = 0; c < 3; ++c) {
            const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;            // cofactor expansion along row 0 (cyclic, sign +)
            for (int i = 0; i < 10; ++i) t2[i] = 0.0;
            p1p1_acc(Ep[3 + c1], Ep[6 + c2], t2, 1.0);
            p1p1_acc(Ep[3 + c2], Ep[6 + c1], t2, -1.0);
            p2p1_acc(t2, Ep[c], A[0], 1.0);
        }
        // L = E E^T - 0.5 tr(E E^T) I ; rows 1..9 = L E
        double L[6][10];                                             // 00 01 02 11 12 22
        const int li[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
        for (int a = 0; a < 6; ++a)
            for (int i = 0; i < 10; ++i) L[a][i] = 0.0;
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j)
                for (int k = 0; k < 3; ++k) p1p1_acc(Ep[3 * i + k], Ep[3 * j + k], L[li[i][j]], 1.0);
        for (int i = 0; i < 10; ++i) {
            const double tr = 0.5 * (L[0][i] + L[3][i] + L[5][i]);
            L[0][i] -= tr; L[3][i] -= tr; L[5][i] -= tr;
        }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                for (int k = 0; k < 3; ++k) p2p1_acc(L[li[i][k]], Ep[3 * k + j], A[1 + 3 * i + j], 1.0);
    }
    FP_STAMP(2);
    // ---- Gauss-Jordan on the first ten columns (partial pivoting)
    for (int c = 0; c < 10; ++c) {
        int br = c;
        double bv = fabs(A[c][c]);
        for (int r = c + 1; r < 10; ++r)
            if (fabs(A[r][c]) > bv) { bv = fabs(A[r][c]); br = r; }
        if (bv < 1e-300) return 0;
        if (br != c)
            for (int j = c; j < 20; ++j) { const double t = A[c][j]; A[c][j] = A[br][j]; A[br][j] = t; }
        const double inv = 1.0 / A[c][c];
        for (int j = c; j < 20; ++j) A[c][j] *= inv;
        for (int r = 0; r < 10; ++r) {
            if (r == c) continue;
            const double f = A[r][c];
            if (f != 0.0) for (int j = c; j < 20; ++j) A[r][j] = fma(-f, A[c][j], A[r][j]);
        }
    }
    FP_STAMP(3);
    // ---- B(z): rows <k> = <e> - z<f>, <l> = <g> - z<h>, <m> = <i> - z<j>; columns x (deg 3), y (deg 3), 1 (deg 4)
    double B[3][3][5];
    for (int r = 0; r < 3; ++r) {
        const double *p = A[4 + 2 * r], *s = A[5 + 2 * r];
        for (int v = 0; v < 2; ++v) {                                // x: columns 10..12, y: columns 13..15
            const int o = 10 + 3 * v;
            B[r][v][0] = p[o + 2];
            B[r][v][1] = p[o + 1] - s[o + 2];
            B[r][v][2] = p[o] - s[o + 1];
            B[r][v][3] = -s[o];
            B[r][v][4] = 0.0;
        }
        B[r][2][0] = p[19];
        B[r][2][1] = p[18] - s[19];
        B[r][2][2] = p[17] - s[18];
        B[r][2][3] = p[16] - s[17];
        B[r][2][4] = -s[16];
    }
    double poly[11];
    for (int i = 0; i < 11; ++i) poly[i] = 0.0;
    {
        double t7[8], u7[8], t6[7], u6[7], pr[11];
        // kx (ly m1 - l1 my) - ky (lx m1 - l1 mx) + k1 (lx my - ly mx)
        zmul(B[1][1], 3, B[2][2], 4, t7); zmul(B[1][2], 4, B[2][1], 3, u7);
        for (int i = 0; i < 8; ++i) t7[i] -= u7[i];
        zmul(B[0][0], 3, t7, 7, pr);
        for (int i = 0; i < 11; ++i) poly[i] += pr[i];
        zmul(B[1][0], 3, B[2][2], 4, t7); zmul(B[1][2], 4, B[2][0], 3, u7);
        for (int i = 0; i < 8; ++i) t7[i] -= u7[i];
        zmul(B[0][1], 3, t7, 7, pr);
        for (int i = 0; i < 11; ++i) poly[i] -= pr[i];
        zmul(B[1][0], 3, B[2][1], 3, t6); zmul(B[1][1], 3, B[2][0], 3, u6);
        for (int i = 0; i < 7; ++i) t6[i] -= u6[i];
        zmul(B[0][2], 4, t6, 6, pr);
        for (int i = 0; i < 11; ++i) poly[i] += pr[i];
    }
    FP_STAMP(4);
    double zs[10];
    const int nz = real_roots(poly, 10, zs);
    FP_STAMP(5);
    int ns = 0;
    for (int s = 0; s < nz; ++s) {
        const double z = zs[s];
        double b[3][3];
        for (int r = 0; r < 3; ++r) {
            b[r][0] = horner(B[r][0], 3, z);
            b[r][1] = horner(B[r][1], 3, z);
            b[r][2] = horner(B[r][2], 4, z);
        }
        // (x, y) from the two rows whose 2x2 system is best conditioned
        int r0 = 0, r1 = 1;
        double bd = 0.0;
        for (int a = 0; a < 3; ++a)
            for (int c = a + 1; c < 3; ++c) {
                const double na2 = (b[a][0] * b[a][0] + b[a][1] * b[a][1]) * (b[c][0] * b[c][0] + b[c][1] * b[c][1]);
                const double dd = b[a][0] * b[c][1] - b[a][1] * b[c][0];
                const double sc = na2 > 0.0 ? dd * dd / na2 : 0.0;
                if (sc > bd) { bd = sc; r0 = a; r1 = c; }
            }
        const double D = b[r0][0] * b[r1][1] - b[r0][1] * b[r1][0];
        if (!(bd > 0.0) || D == 0.0) continue;
        const double x = (b[r0][1] * b[r1][2] - b[r0][2] * b[r1][1]) / D;
        const double y = (b[r0][2] * b[r1][0] - b[r0][0] * b[r1][2]) / D;
        double E[9], nrm = 0.0;
        for (int e = 0; e < 9; ++e) {
            E[e] = fma(x, Ep[e][0], fma(y, Ep[e][1], fma(z, Ep[e][2], Ep[e][3])));
            nrm = fma(E[e], E[e], nrm);
        }
        if (!(nrm > 0.0) || !isfinite(nrm)) continue;
        const double inv = rsqrt(nrm);
        for (int e = 0; e < 9; ++e) E_out[ns * 9 + e] = E[e] * inv;
        ++ns;
    }
    FP_STAMP(6);
    return ns;
#undef FP_STAMP
}

__device__ __forceinline__ unsigned fp_hash(unsigned a, unsigned b, unsigned c)
{
    unsigned long long z = (((unsigned long long)a << 32) ^ (unsigned long long)b) + 0x9E3779B97F4A7C15ull * ((unsigned long long)c + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (unsigned)((z ^ (z >> 31)) >> 16);
}

__global__ void __launch_bounds__(32) k_5pt_solve(const FpDev d)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= d.H) return;
    const int n = fp_count(d);
    int ns = 0;
    if (n >= 5) {
        int id[5];
        unsigned ctr = 0;
        for (int k = 0; k < 5; ++k) {
            for (;;) {
                const int c = (int)(fp_hash(d.seed, (unsigned)h, ctr++) % (unsigned)n);
                bool dup = false;
                for (int j = 0; j < k; ++j) dup |= id[j] == c;
                if (!dup) { id[k] = c; break; }
            }
        }
        double4 q[5];
        for (int k = 0; k < 5; ++k) q[k] = d.q[id[k]];
        ns = solve5(q, d.Es + (size_t)h * 90);
    }
    d.nsol[h] = ns;
}

// thread per correspondence set: the minimal solver alone (parity tests against the action-matrix oracle)
__global__ void __launch_bounds__(64) k_5pt_minimal(const double4 *q, int n_sets, double *Es, int *nsol, long long *trace)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sets) return;
    nsol[s] = solve5(q + 5 * (size_t)s, Es + (size_t)s * 90, (trace && s == 0) ? trace : nullptr);
}

__device__ __forceinline__ bool fp_inlier(const double *E, const double4 q, double thr2)
{
    const double a0 = fma(E[0], q.x, fma(E[1], q.y, E[2]));
    const double a1 = fma(E[3], q.x, fma(E[4], q.y, E[5]));
    const double a2 = fma(E[6], q.x, fma(E[7], q.y, E[8]));
    const double b0 = fma(E[0], q.z, fma(E[3], q.w, E[6]));
    const double b1 = fma(E[1], q.z, fma(E[4], q.w, E[7]));
    const double num = fma(q.z, a0, fma(q.w, a1, a2));
    const double den = fma(a0, a0, fma(a1, a1, fma(b0, b0, b1 * b1)));
    return num * num <= thr2 * den;
}

__global__ void __launch_bounds__(128) k_5pt_score(const FpDev d)
{
    __shared__ double sE[90];
    __shared__ int sc[10];
    const int h = blockIdx.x, tid = threadIdx.x;
    const int ns = d.nsol[h];
    if (ns == 0) return;
    for (int i = tid; i < ns * 9; i += 128) sE[i] = d.Es[(size_t)h * 90 + i];
    if (tid < 10) sc[tid] = 0;
    __syncthreads();
    const int n = fp_count(d);
    int cnt[10];
#pragma unroll
    for (int c = 0; c < 10; ++c) cnt[c] = 0;
    for (int i = tid; i < n; i += 128) {
        const double4 q = d.q[i];
#pragma unroll
        for (int c = 0; c < 10; ++c)
            if (c < ns) cnt[c] += fp_inlier(sE + 9 * c, q, d.thr2) ? 1 : 0;
    }
#pragma unroll
    for (int c = 0; c < 10; ++c) {
        if (c < ns) {
            const int w = __reduce_add_sync(0xffffffffu, cnt[c]);
            if ((tid & 31) == 0) atomicAdd(&sc[c], w);
        }
    }
    __syncthreads();
    if (tid == 0) {
        int bc = 0, bn = sc[0];
        for (int c = 1; c < ns; ++c)
            if (sc[c] > bn) { bn = sc[c]; bc = c; }
        // OpenCV keeps the first model whose count is strictly larger: highest count, then lowest (hypothesis, candidate)
        const unsigned long long key = ((unsigned long long)(unsigned)bn << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(h * 16 + bc));
        atomicMax(d.best, key);
    }
}

// eigen-decomposition of a symmetric 3x3 (cyclic Jacobi, FP64): A = V diag(w) V^T
__device__ void jacobi_eig3(double A[3][3], double V[3][3], double w[3])
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

// decomposition of the best model (one thread): E = U diag(s, s, 0) V^T (motion_estimator.cpp:70-96)
__global__ void k_5pt_decomp(const FpDev d)
{
    if (threadIdx.x != 0) return;
    FpPose &P = *d.pose;
    for (int i = 0; i < 5; ++i) P.cnt[i] = 0;
    const unsigned long long key = *d.best;
    const int count = (int)(key >> 32);
    P.valid = count >= 5 ? 1 : 0;                     // no model: cv::findEssentialMat returns an empty matrix
    if (!P.valid) return;
    const unsigned id = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
    const int h = (int)(id >> 4), c = (int)(id & 15u);
    double sE[9];
    for (int i = 0; i < 9; ++i) { sE[i] = d.Es[(size_t)h * 90 + c * 9 + i]; P.E[i] = sE[i]; }
    // V from the eigenvectors of E^T E, u_i = E v_i / s_i, u_3 = u_1 x u_2 (det U = +1, what the reference's sign fix
    // produces), det V forced to +1 through v_3
    double A[3][3], V[3][3], w[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = sE[0 + i] * sE[0 + j] + sE[3 + i] * sE[3 + j] + sE[6 + i] * sE[6 + j];
    jacobi_eig3(A, V, w);
    int o[3] = {0, 1, 2};
    for (int a = 0; a < 2; ++a)
        for (int b = a + 1; b < 3; ++b)
            if (w[o[b]] > w[o[a]]) { const int t = o[a]; o[a] = o[b]; o[b] = t; }
    double v[3][3], u[3][3];                          // v[k] = k-th right singular vector
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < 3; ++i) v[k][i] = V[i][o[k]];
    for (int k = 0; k < 2; ++k) {
        double nn = 0.0;
        for (int i = 0; i < 3; ++i) u[k][i] = sE[3 * i] * v[k][0] + sE[3 * i + 1] * v[k][1] + sE[3 * i + 2] * v[k][2];
        if (k == 1) {
            const double dp = u[1][0] * u[0][0] + u[1][1] * u[0][1] + u[1][2] * u[0][2];
            for (int i = 0; i < 3; ++i) u[1][i] -= dp * u[0][i];
        }
        for (int i = 0; i < 3; ++i) nn += u[k][i] * u[k][i];
        const double inv = 1.0 / sqrt(nn);
        for (int i = 0; i < 3; ++i) u[k][i] *= inv;
    }
    u[2][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1];
    u[2][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2];
    u[2][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
    const double cx0 = v[0][1] * v[1][2] - v[0][2] * v[1][1], cx1 = v[0][2] * v[1][0] - v[0][0] * v[1][2],
                 cx2 = v[0][0] * v[1][1] - v[0][1] * v[1][0];
    if (cx0 * v[2][0] + cx1 * v[2][1] + cx2 * v[2][2] < 0.0)
        for (int i = 0; i < 3; ++i) v[2][i] = -v[2][i];
    // W = [0 -1 0; 1 0 0; 0 0 1]: the columns of U W are (u2, -u1, u3)
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            P.R[0][3 * i + j] = (float)(u[1][i] * v[0][j] - u[0][i] * v[1][j] + u[2][i] * v[2][j]);   // U W V^T
            P.R[1][3 * i + j] = (float)(-u[1][i] * v[0][j] + u[0][i] * v[1][j] + u[2][i] * v[2][j]);  // U W^T V^T
        }
    for (int i = 0; i < 3; ++i) P.t[i] = (float)u[2][i];
    if (d.E_out) for (int i = 0; i < 9; ++i) d.E_out[i] = (float)sE[i];
}

// one thread per correspondence: RANSAC inlier bit of the best model (what cv::findEssentialMat returns) and the
// cheirality of the four candidates (findCorrectRT :205-263 triangulates ALL correspondences for every candidate)
__global__ void __launch_bounds__(128) k_5pt_cheir(const FpDev d)
{
    __shared__ int s_cnt[5];
    const FpPose &P = *d.pose;
    if (!P.valid) return;
    const int tid = threadIdx.x, i = blockIdx.x * 128 + tid;
    if (tid < 5) s_cnt[tid] = 0;
    __syncthreads();
    const int n = fp_count(d);
    int bits = 0;
    if (i < n) {
        if (fp_inlier(P.E, d.q[i], d.thr2)) bits |= 1;
        const float2 a = d.p0[i], b = d.p1[i];
#pragma unroll 1
        for (int cand = 0; cand < 4; ++cand) {
            const float sg = (cand & 1) ? -1.f : 1.f;
            const float t[3] = {sg * P.t[0], sg * P.t[1], sg * P.t[2]};
            float X0[3], X1[3];
            tri_point(a, b, P.R[cand >> 1], t, d.K, d.K, X0, X1);
            if (X0[2] > 0.f && X1[2] > 0.f) bits |= 2 << cand;
            float *o = d.X0c + ((size_t)i * 4 + cand) * 3;
            o[0] = X0[0]; o[1] = X0[1]; o[2] = X0[2];
        }
        d.bits[i] = (uint8_t)bits;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int w = __reduce_add_sync(0xffffffffu, (bits >> k) & 1);
        if ((tid & 31) == 0 && w) atomicAdd(&s_cnt[k], w);
    }
    __syncthreads();
    if (tid < 5 && s_cnt[tid]) atomicAdd(&d.pose->cnt[tid], s_cnt[tid]);
}

// the candidate with strictly most points in front of both cameras, first one on ties (:244-252); outputs
__global__ void __launch_bounds__(128) k_5pt_final(const FpDev d)
{
    const FpPose &P = *d.pose;
    const int n = fp_count(d);
    const int i = blockIdx.x * 128 + threadIdx.x;
    int mx = 0, b = -1;
    if (P.valid)
        for (int cand = 0; cand < 4; ++cand)
            if (P.cnt[1 + cand] > mx) { mx = P.cnt[1 + cand]; b = cand; }
    if (i == 0) {
        d.info[0] = P.valid ? P.cnt[0] : 0; d.info[1] = mx; d.info[2] = b >= 0 ? 1 : 0;
        if (b >= 0) {
            const float sg = (b & 1) ? -1.f : 1.f;
            for (int k = 0; k < 9; ++k) d.R10[k] = P.R[b >> 1][k];
            for (int k = 0; k < 3; ++k) d.t10[k] = sg * P.t[k];
        }
    }
    if (i >= n) return;
    if (b < 0) { d.mask[i] = 0; return; }             // the reference leaves R10_true / t10_true unset here; reported as failure
    const int bits = d.bits[i];
    d.mask[i] = ((bits & 1) && ((bits >> (1 + b)) & 1)) ? 1 : 0;                  // mask_inlier = verify && 5p (:115)
    if (d.X0) {
        const float *o = d.X0c + ((size_t)i * 4 + b) * 3;
        d.X0[3 * i] = o[0]; d.X0[3 * i + 1] = o[1]; d.X0[3 * i + 2] = o[2];
    }
}

int fp_scratch(vo_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->fp_bytes) return VO_OK;
    if (ctx->d_fp) { VO_CUDA(cudaStreamSynchronize(ctx->stream)); VO_CUDA(cudaFree(ctx->d_fp)); ctx->d_fp = nullptr; ctx->fp_bytes = 0; }
    size_t want = ctx->fp_bytes ? ctx->fp_bytes * 2 : (size_t)1 << 20;
    if (want < bytes) want = bytes;
    VO_CUDA(cudaMalloc(&ctx->d_fp, want));
    ctx->fp_bytes = want;
    return VO_OK;
}

}  // namespace

// Asynchronous, device pointers.  n: capacity of the point arrays; n_d (nullable): device-resident count <= n.
// R10_d[9], t10_d[3], mask_d[n], info_d[3]; X0_d [n][3] and E_d[9] nullable.
int vo_5pt_launch_d(vo_ctx *ctx, const float *pts0_d, const float *pts1_d, int n, const int *n_d, const float *K, float thres_px,
                    int n_hyp, unsigned seed, float *R10_d, float *t10_d, float *X0_d, uint8_t *mask_d, float *E_d, int *info_d)
{
    if (n_hyp <= 0) n_hyp = 1024;
    VO_REQUIRE(n_hyp <= (1 << 20), VO_ERR_INVALID_ARG, "too many hypotheses");
    const size_t a = 256;
    const int n_cap = n > 8192 ? n : 8192;      // scratch sized for at least 8192 correspondences: no regrowth inside a sequence
    const size_t o_q = 0, o_E = o_q + ((size_t)n_cap * 32 + a - 1) / a * a, o_ns = o_E + (size_t)n_hyp * 720, o_b = o_ns + ((size_t)n_hyp * 4 + a - 1) / a * a;
    const size_t o_P = o_b + a, o_bits = o_P + (sizeof(FpPose) + a - 1) / a * a, o_Xc = o_bits + ((size_t)n_cap + a - 1) / a * a;
    const int rc = fp_scratch(ctx, o_Xc + (size_t)n_cap * 48);
    if (rc) return rc;
    uint8_t *s = (uint8_t *)ctx->d_fp;
    FpDev d;
    memset(&d, 0, sizeof(d));
    d.p0 = (const float2 *)pts0_d; d.p1 = (const float2 *)pts1_d; d.n = n; d.n_d = n_d;
    d.fx = K[0]; d.fy = K[1]; d.cx = K[2]; d.cy = K[3];
    const double thr = (double)thres_px / (((double)K[0] + (double)K[1]) / 2.0);     // cv::findEssentialMat: threshold /= (fx + fy) / 2
    d.thr2 = thr * thr;
    d.H = n_hyp; d.seed = seed;
    d.q = (double4 *)(s + o_q); d.Es = (double *)(s + o_E); d.nsol = (int *)(s + o_ns); d.best = (unsigned long long *)(s + o_b);
    d.R10 = R10_d; d.t10 = t10_d; d.X0 = X0_d; d.E_out = E_d; d.mask = mask_d; d.info = info_d;
    d.pose = (FpPose *)(s + o_P); d.bits = s + o_bits; d.X0c = (float *)(s + o_Xc);
    memcpy(d.K, K, 16);
    k_5pt_norm<<<vo_div_up(n > 0 ? n : 1, 256), 256, 0, ctx->stream>>>(d);
    k_5pt_solve<<<vo_div_up(n_hyp, 32), 32, 0, ctx->stream>>>(d);
    k_5pt_score<<<n_hyp, 128, 0, ctx->stream>>>(d);
    k_5pt_decomp<<<1, 32, 0, ctx->stream>>>(d);
    k_5pt_cheir<<<vo_div_up(n > 0 ? n : 1, 128), 128, 0, ctx->stream>>>(d);
    k_5pt_final<<<vo_div_up(n > 0 ? n : 1, 128), 128, 0, ctx->stream>>>(d);
    ctx->launches += 6;
    VO_CUDA(cudaGetLastError());
    return VO_OK;
}

extern "C" int vo_pose_5point(vo_ctx *ctx, const float *pts0, const float *pts1, int n, const float *K, float thres_px, int n_hypotheses,
                              unsigned seed, float *R10, float *t10, float *X0, uint8_t *mask, float *E, int *info)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(pts0 && pts1 && K && R10 && t10 && mask, VO_ERR_INVALID_ARG, "null pointer");
    // motion_estimator.cpp:26-34: both size checks throw; fewer than five correspondences have no model
    VO_REQUIRE(n > 0, VO_ERR_SIZE_MISMATCH, "calcPose5PointsAlgorithm(): pts0.size() == pts1.size() == 0");
    VO_REQUIRE(n >= 5, VO_ERR_INVALID_ARG, "calcPose5PointsAlgorithm(): fewer than five correspondences");
    VO_CUDA(cudaSetDevice(ctx->device));
    const size_t N = (size_t)n;
    const size_t o_p0 = 0, o_p1 = N * 8, in_bytes = N * 16;
    const size_t o_res = (in_bytes + 255) / 256 * 256;
    const size_t o_R = o_res, o_t = o_R + 48, o_E = o_t + 16, o_i = o_E + 48, o_m = o_i + 16, o_X = o_m + (N + 15) / 16 * 16, total = o_X + N * 12;
    int rc = vo_stage_reserve(ctx, total);
    if (rc) return rc;
    uint8_t *hs = ctx->h_stage, *dv = ctx->d_stage;
    memcpy(hs + o_p0, pts0, N * 8); memcpy(hs + o_p1, pts1, N * 8);
    VO_CUDA(cudaMemcpyAsync(dv, hs, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    VO_CUDA(cudaMemsetAsync(dv + o_res, 0, o_m - o_res, ctx->stream));
    rc = vo_5pt_launch_d(ctx, (const float *)(dv + o_p0), (const float *)(dv + o_p1), n, nullptr, K, thres_px, n_hypotheses, seed,
                         (float *)(dv + o_R), (float *)(dv + o_t), (float *)(dv + o_X), dv + o_m, (float *)(dv + o_E), (int *)(dv + o_i));
    if (rc) return rc;
    VO_CUDA(cudaMemcpyAsync(hs + o_res, dv + o_res, total - o_res, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    const int *hi = (const int *)(hs + o_i);
    if (info) { info[0] = hi[0]; info[1] = hi[1]; info[2] = hi[2]; }
    memcpy(mask, hs + o_m, N);
    if (!hi[2]) { ctx->last_error = "calcPose5PointsAlgorithm() is failed."; return VO_ERR_MODE; }     // mono_vo.cpp:590
    memcpy(R10, hs + o_R, 36); memcpy(t10, hs + o_t, 12);
    if (E) memcpy(E, hs + o_E, 36);
    if (X0) memcpy(X0, hs + o_X, N * 12);
    return VO_OK;
}

extern "C" int vo_five_point_minimal(vo_ctx *ctx, const double *q, int n_sets, double *E, int *n_solutions)
{
    if (!ctx) return VO_ERR_INVALID_ARG;
    VO_REQUIRE(q && E && n_solutions && n_sets > 0, VO_ERR_INVALID_ARG, "null pointer");
    VO_CUDA(cudaSetDevice(ctx->device));
    const size_t S = (size_t)n_sets;
    const size_t o_q = 0, o_E = S * 160, o_n = o_E + S * 720, total = o_n + S * 4;
    const int rc = vo_stage_reserve(ctx, total);
    if (rc) return rc;
    uint8_t *hs = ctx->h_stage, *dv = ctx->d_stage;
    memcpy(hs + o_q, q, S * 160);
    VO_CUDA(cudaMemcpyAsync(dv, hs, S * 160, cudaMemcpyHostToDevice, ctx->stream));
    VO_CUDA(cudaMemsetAsync(dv + o_E, 0, total - o_E, ctx->stream));
    static const bool trace = getenv("VO_5PT_TRACE") != nullptr;      // phase clocks of set 0 (profiling aid)
    long long *trace_d = nullptr;
    if (trace) { VO_CUDA(cudaMalloc(&trace_d, 64)); }
    k_5pt_minimal<<<vo_div_up(n_sets, 64), 64, 0, ctx->stream>>>((const double4 *)(dv + o_q), n_sets, (double *)(dv + o_E), (int *)(dv + o_n), trace_d);
    ctx->launches++;
    VO_CUDA(cudaGetLastError());
    VO_CUDA(cudaMemcpyAsync(hs + o_E, dv + o_E, total - o_E, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (trace) {
        long long t[8];
        VO_CUDA(cudaMemcpy(t, trace_d, 56, cudaMemcpyDeviceToHost));
        fprintf(stderr, "[5pt trace] null %lld | constraints %lld | gauss-jordan %lld | B(z)+det %lld | roots %lld | back-subst %lld cycles\n",
                t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5]);
        cudaFree(trace_d);
    }
    memcpy(E, hs + o_E, S * 720);
    memcpy(n_solutions, hs + o_n, S * 4);
    return VO_OK;
}
