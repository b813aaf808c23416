/*
 * oracle/lba_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Eigen-free CPU restatement, FP64 like the reference (define_ba_type.h:9), of
 *   SparseBundleAdjustmentSolver::solveForFiniteIterations
 *        core/visual_odometry/ba_solver/sparse_bundle_adjustment.cpp:150-768
 *   calc_Rij_t_Rij_weight / calc_Qij_t_Qij_weight          :913-934, :1022-1109
 *   geometry::se3Exp / SE3Log / addFrontse3 / inverseSE3   core/util/geometry_library.cpp:336-384, 442-495, 546-552, 561-567
 *   Eigen::LDLT (3x3 and dynamic 6N_opt) -- third-party Eigen3, unpinned, absent here: restated
 *        from the published algorithm (unblocked, diagonal pivoting, lower storage).
 * working on the flat problem SparseBAParameters::setPosesAndPoints packs
 * (ba_solver/sparse_ba_parameters.h:292-465): left-keyframe poses T_jw in the reference keyframe's
 * frame with translations / points divided by pose_scale (=10), observations per landmark in
 * chronological order, left then right per keyframe (keyframes.cpp:177-215).
 *
 * Reference quirks reproduced: cross block B[j][i] ASSIGNED not accumulated
 * (#1, :307,:411 -- last observation of (landmark i, keyframe j) wins, normally the right-camera one),
 * right-camera Q built as [dp_dX*R_rl, -dp_dX*R_rl*[Xij]x] and fed to the shortcut QtQ that assumes
 * Q(0,1)=Q(1,0)=0 (#3), Huber gate strict '>' (#6), SE3Log returning w=0 when (trR-1)/2 >= 0.999999999,
 * fixed lambda, fixed iteration count, no step rejection, diagonal blocks of BCinvBt transposed by the
 * mirror loop (:501-503).
 * PARITY UNPINNED by the reference (no golden vectors, cannot be compiled here).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ geometry (double) */
void orc_se3exp_d(const double *xi, double *T)
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
    double a, b, bV, c;
    if (theta < 1e-9) { a = 1.0; b = 0.5; bV = 0.5; c = 0.33333333333333333333333333; }
    else {
        const double invtheta2 = 1.0 / (theta * theta);
        a = sin(theta) / theta;
        b = (1 - cos(theta)) * invtheta2;
        bV = b;
        c = (theta - sin(theta)) / (theta * theta * theta);
    }
    double V[9];
    memset(T, 0, sizeof(double) * 16);
    for (int i = 0; i < 9; ++i) {
        const double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        const int r = i / 3, cc = i % 3;
        T[r * 4 + cc] = (id + a * wx[i]) + b * wxwx[i];
        V[i] = (id + bV * wx[i]) + c * wxwx[i];
    }
    for (int i = 0; i < 3; ++i) T[i * 4 + 3] = (V[i * 3] * v[0] + V[i * 3 + 1] * v[1]) + V[i * 3 + 2] * v[2];
    T[15] = 1.0;
}

void orc_se3log_d(const double *T, double *xi)
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
        w[0] = -lnR[1 * 3 + 2]; w[1] = lnR[0 * 3 + 2]; w[2] = -lnR[0 * 3 + 1];
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

static void mat4_mul_d(const double *A, const double *B, double *C)
{
    double T[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j];
            T[i * 4 + j] = s;
        }
    memcpy(C, T, sizeof(T));
}

void orc_inverse_se3_d(const double *T, double *Ti)
{
    double O[16] = {0};
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) O[i * 4 + j] = T[j * 4 + i];
        double s = 0;
        for (int k = 0; k < 3; ++k) s += T[k * 4 + i] * T[k * 4 + 3];
        O[i * 4 + 3] = -s;
    }
    O[15] = 1.0;
    memcpy(Ti, O, sizeof(O));
}

/* T <- Exp(Log(Exp(x) * Exp(Log(T))))   (sparse_bundle_adjustment.cpp:584-596) */
void orc_pose_retract_d(double *T, const double *x)
{
    double xi[6], Tjw[16], dT[16];
    orc_se3log_d(T, xi);
    orc_se3exp_d(xi, Tjw);
    orc_se3exp_d(x, dT);
    mat4_mul_d(dT, Tjw, Tjw);
    orc_se3log_d(Tjw, xi);
    orc_se3exp_d(xi, T);
}

/* Eigen::LDLT (lower, unblocked, diagonal pivoting) solve of A x = b, n x n row-major, nrhs columns. */
void orc_ldlt_solve_d(const double *A_in, int n, const double *B, int nrhs, double *X)
{
    double *m = (double *)malloc(sizeof(double) * n * n);
    int *tr = (int *)malloc(sizeof(int) * n);
    double *temp = (double *)malloc(sizeof(double) * n);
    memcpy(m, A_in, sizeof(double) * n * n);
#define M(i, j) m[(size_t)(i) * n + (j)]
    for (int k = 0; k < n; ++k) {
        int big = k;
        double bv = fabs(M(k, k));
        for (int i = k + 1; i < n; ++i) if (fabs(M(i, i)) > bv) { bv = fabs(M(i, i)); big = i; }
        tr[k] = big;
        if (big != k) {
            const int s = n - big - 1;
            for (int j = 0; j < k; ++j) { double t = M(k, j); M(k, j) = M(big, j); M(big, j) = t; }
            for (int i = 0; i < s; ++i) { double t = M(big + 1 + i, k); M(big + 1 + i, k) = M(big + 1 + i, big); M(big + 1 + i, big) = t; }
            { double t = M(k, k); M(k, k) = M(big, big); M(big, big) = t; }
            for (int i = k + 1; i < big; ++i) { double t = M(i, k); M(i, k) = M(big, i); M(big, i) = t; }
        }
        const int rs = n - k - 1;
        if (k > 0) {
            for (int j = 0; j < k; ++j) temp[j] = M(j, j) * M(k, j);
            double s = 0;
            for (int j = 0; j < k; ++j) s += M(k, j) * temp[j];
            M(k, k) -= s;
            for (int i = 0; i < rs; ++i) {
                double s2 = 0;
                for (int j = 0; j < k; ++j) s2 += M(k + 1 + i, j) * temp[j];
                M(k + 1 + i, k) -= s2;
            }
        }
        const double akk = M(k, k);
        if (rs > 0 && fabs(akk) > 0.0)
            for (int i = 0; i < rs; ++i) M(k + 1 + i, k) /= akk;
    }
    double *y = (double *)malloc(sizeof(double) * n);
    const double tol = 1.0 / 1.7976931348623157e308;
    for (int c = 0; c < nrhs; ++c) {
        for (int i = 0; i < n; ++i) y[i] = B[(size_t)i * nrhs + c];
        for (int k = 0; k < n; ++k) if (tr[k] != k) { double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
        for (int i = 0; i < n; ++i) { double s = y[i]; for (int j = 0; j < i; ++j) s -= M(i, j) * y[j]; y[i] = s; }
        for (int i = 0; i < n; ++i) y[i] = fabs(M(i, i)) > tol ? y[i] / M(i, i) : 0.0;
        for (int i = n - 1; i >= 0; --i) { double s = y[i]; for (int j = i + 1; j < n; ++j) s -= M(j, i) * y[j]; y[i] = s; }
        for (int k = n - 1; k >= 0; --k) if (tr[k] != k) { double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
        for (int i = 0; i < n; ++i) X[(size_t)i * nrhs + c] = y[i];
    }
#undef M
    free(m); free(tr); free(temp); free(y);
}

/* ------------------------------------------------------------------ the solver */
typedef struct {
    int n_frames, n_opt, n_points, n_obs;
    const double *poses;
    const int *opt_index;
    const double *points;
    const int *obs_ptr;
    const int *obs_frame;
    const uint8_t *obs_right;
    const double *obs_px;
    double K_l[4], K_r[4];
    double T_lr[16];
    int is_stereo;
    double huber;
    double lambda;
    int max_iter;
} orc_lba_problem;

/* fix_b_accumulate: 0 = reference behaviour (B assigned, last writer wins); 1 = "fixed" variant that
 * accumulates B (for reporting only). Returns 0, or -4 on NaN. */
int orc_lba_solve(const orc_lba_problem *P, double *poses_out, double *points_out, double *avg_err_out,
                  int *success, int fix_b_accumulate)
{
    const int N = P->n_frames, No = P->n_opt, Mp = P->n_points;
    double *poses = (double *)malloc(sizeof(double) * 16 * N);
    double *X = (double *)malloc(sizeof(double) * 3 * Mp);
    memcpy(poses, P->poses, sizeof(double) * 16 * N);
    memcpy(X, P->points, sizeof(double) * 3 * Mp);
    /* opt slot -> frame */
    int *opt_frame = (int *)malloc(sizeof(int) * (No > 0 ? No : 1));
    for (int f = 0; f < N; ++f) if (P->opt_index[f] >= 0) opt_frame[P->opt_index[f]] = f;

    double T_rl[16], R_rl[9], t_rl[3];
    orc_inverse_se3_d(P->T_lr, T_rl);
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R_rl[i * 3 + j] = T_rl[i * 4 + j]; t_rl[i] = T_rl[i * 4 + 3]; }

    double *A = (double *)calloc((size_t)No * 36, sizeof(double)), *a = (double *)calloc((size_t)No * 6, sizeof(double));
    double *B = (double *)calloc((size_t)No * Mp * 18, sizeof(double));      /* [j][i] 6x3 */
    double *BCinv = (double *)calloc((size_t)No * Mp * 18, sizeof(double));
    double *C = (double *)calloc((size_t)Mp * 9, sizeof(double)), *b = (double *)calloc((size_t)Mp * 3, sizeof(double));
    double *Cinv = (double *)calloc((size_t)Mp * 9, sizeof(double)), *Cinv_b = (double *)calloc((size_t)Mp * 3, sizeof(double));
    const int n6 = 6 * No;
    double *BCinvBt = (double *)calloc((size_t)n6 * n6, sizeof(double)), *BCinv_b = (double *)calloc((size_t)n6, sizeof(double));
    double *S = (double *)calloc((size_t)n6 * n6, sizeof(double)), *rhs = (double *)calloc((size_t)n6, sizeof(double));
    double *x = (double *)calloc((size_t)n6, sizeof(double));
    int rc = 0, flag_success = 1;
    const double huber = P->huber, lambda = P->lambda;

    for (int iter = 0; iter < P->max_iter; ++iter) {
        memset(A, 0, sizeof(double) * No * 36); memset(a, 0, sizeof(double) * No * 6);
        memset(B, 0, sizeof(double) * (size_t)No * Mp * 18); memset(BCinv, 0, sizeof(double) * (size_t)No * Mp * 18);
        memset(C, 0, sizeof(double) * Mp * 9); memset(b, 0, sizeof(double) * Mp * 3);
        memset(BCinvBt, 0, sizeof(double) * n6 * n6); memset(BCinv_b, 0, sizeof(double) * n6);
        double err = 0.0;
        for (int i = 0; i < Mp; ++i) {
            const double *Xi = X + 3 * i;
            for (int o = P->obs_ptr[i]; o < P->obs_ptr[i + 1]; ++o) {
                const int f = P->obs_frame[o];
                const int right = P->obs_right[o];
                const int j = P->opt_index[f];
                const double *T = poses + 16 * f;
                double R_jw[9], t_jw[3];
                for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R_jw[r * 3 + c] = T[r * 4 + c]; t_jw[r] = T[r * 4 + 3]; }
                double Xij[3];
                for (int r = 0; r < 3; ++r) Xij[r] = ((R_jw[r * 3] * Xi[0] + R_jw[r * 3 + 1] * Xi[1]) + R_jw[r * 3 + 2] * Xi[2]) + t_jw[r];
                double Rm[9], Xc[3], fx, fy, cx, cy;   /* rotation used in Rij, point in the observing camera */
                if (right) {
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 3; ++c)
                            Rm[r * 3 + c] = (R_rl[r * 3] * R_jw[c] + R_rl[r * 3 + 1] * R_jw[3 + c]) + R_rl[r * 3 + 2] * R_jw[6 + c];
                    for (int r = 0; r < 3; ++r) Xc[r] = ((R_rl[r * 3] * Xij[0] + R_rl[r * 3 + 1] * Xij[1]) + R_rl[r * 3 + 2] * Xij[2]) + t_rl[r];
                    fx = P->K_r[0]; fy = P->K_r[1]; cx = P->K_r[2]; cy = P->K_r[3];
                } else {
                    memcpy(Rm, R_jw, sizeof(Rm)); memcpy(Xc, Xij, sizeof(Xc));
                    fx = P->K_l[0]; fy = P->K_l[1]; cx = P->K_l[2]; cy = P->K_l[3];
                }
                const double invz = 1.0 / Xc[2];
                const double fxinvz = fx * invz, fyinvz = fy * invz, xinvz = Xc[0] * invz, yinvz = Xc[1] * invz;
                const double fx_xinvz2 = fxinvz * xinvz, fy_yinvz2 = fyinvz * yinvz, xinvz_yinvz = xinvz * yinvz;
                const double rij[2] = {(fx * xinvz + cx) - P->obs_px[2 * o], (fy * yinvz + cy) - P->obs_px[2 * o + 1]};
                const double absrxry = fabs(rij[0]) + fabs(rij[1]);
                double weight = 1.0;
                if (absrxry > huber) weight = huber / absrxry;
                double Rij[6];
                for (int c = 0; c < 3; ++c) {
                    Rij[c] = fxinvz * Rm[c] - fx_xinvz2 * Rm[6 + c];
                    Rij[3 + c] = fyinvz * Rm[3 + c] - fy_yinvz2 * Rm[6 + c];
                }
                /* C_i += w RtR (upper then mirrored), b_i -= w Rt r */
                double RtR[9];
                RtR[0] = weight * (Rij[0] * Rij[0] + Rij[3] * Rij[3]);
                RtR[1] = weight * (Rij[0] * Rij[1] + Rij[3] * Rij[4]);
                RtR[2] = weight * (Rij[0] * Rij[2] + Rij[3] * Rij[5]);
                RtR[4] = weight * (Rij[1] * Rij[1] + Rij[4] * Rij[4]);
                RtR[5] = weight * (Rij[1] * Rij[2] + Rij[4] * Rij[5]);
                RtR[8] = weight * (Rij[2] * Rij[2] + Rij[5] * Rij[5]);
                RtR[3] = RtR[1]; RtR[6] = RtR[2]; RtR[7] = RtR[5];
                for (int k = 0; k < 9; ++k) C[9 * i + k] += RtR[k];
                for (int c = 0; c < 3; ++c) b[3 * i + c] += -(weight * (Rij[c] * rij[0] + Rij[3 + c] * rij[1]));
                if (j >= 0) {
                    double Q[12]; /* 2x6 */
                    if (right) {
                        const double dp[6] = {fxinvz, 0, -fx_xinvz2, 0, fyinvz, -fy_yinvz2};
                        double dpR[6];
                        for (int r = 0; r < 2; ++r)
                            for (int c = 0; c < 3; ++c)
                                dpR[r * 3 + c] = (dp[r * 3] * R_rl[c] + dp[r * 3 + 1] * R_rl[3 + c]) + dp[r * 3 + 2] * R_rl[6 + c];
                        const double sk[9] = {0, -Xij[2], Xij[1], Xij[2], 0, -Xij[0], -Xij[1], Xij[0], 0};
                        for (int r = 0; r < 2; ++r)
                            for (int c = 0; c < 3; ++c) {
                                Q[r * 6 + c] = dpR[r * 3 + c];
                                /* (-dp_dX*R_rl) * skew : unary minus applied to the product first */
                                Q[r * 6 + 3 + c] = ((-dpR[r * 3]) * sk[c] + (-dpR[r * 3 + 1]) * sk[3 + c]) + (-dpR[r * 3 + 2]) * sk[6 + c];
                            }
                    } else {
                        Q[0] = fxinvz; Q[1] = 0; Q[2] = -fx_xinvz2; Q[3] = -fx * xinvz_yinvz; Q[4] = fx * (1.0 + xinvz * xinvz); Q[5] = -fx * yinvz;
                        Q[6] = 0; Q[7] = fyinvz; Q[8] = -fy_yinvz2; Q[9] = -fy * (1.0 + yinvz * yinvz); Q[10] = fy * xinvz_yinvz; Q[11] = fy * xinvz;
                    }
                    double wa[12];
                    for (int k = 0; k < 12; ++k) wa[k] = weight * Q[k];
                    double QtQ[36] = {0};
                    QtQ[0 * 6 + 0] = wa[0] * Q[0];
                    for (int c = 2; c < 6; ++c) QtQ[0 * 6 + c] = wa[0] * Q[c];
                    for (int c = 1; c < 6; ++c) QtQ[1 * 6 + c] = wa[6 + 1] * Q[6 + c];
                    for (int r = 2; r < 6; ++r) for (int c = r; c < 6; ++c) QtQ[r * 6 + c] = wa[r] * Q[c] + wa[6 + r] * Q[6 + c];
                    for (int r = 0; r < 6; ++r) for (int c = 0; c < r; ++c) QtQ[r * 6 + c] = QtQ[c * 6 + r];
                    QtQ[0 * 6 + 1] = 0; QtQ[1 * 6 + 0] = 0;
                    for (int k = 0; k < 36; ++k) A[36 * j + k] += QtQ[k];
                    double QtR[18];
                    for (int r = 0; r < 6; ++r)
                        for (int c = 0; c < 3; ++c) QtR[r * 3 + c] = weight * (Q[r] * Rij[c] + Q[6 + r] * Rij[3 + c]);
                    double *Bji = B + ((size_t)j * Mp + i) * 18;
                    if (fix_b_accumulate) for (int k = 0; k < 18; ++k) Bji[k] += QtR[k];
                    else memcpy(Bji, QtR, sizeof(QtR));
                    for (int r = 0; r < 6; ++r) a[6 * j + r] += -(weight * (Q[r] * rij[0] + Q[6 + r] * rij[1]));
                }
                err += rij[0] * rij[0] + rij[1] * rij[1];
            }
        }
        for (int j = 0; j < No; ++j) for (int k = 0; k < 6; ++k) A[36 * j + 7 * k] += lambda * A[36 * j + 7 * k];
        const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        for (int i = 0; i < Mp; ++i) {
            for (int k = 0; k < 3; ++k) C[9 * i + 4 * k] += lambda * C[9 * i + 4 * k];
            orc_ldlt_solve_d(C + 9 * i, 3, I3, 3, Cinv + 9 * i);
            for (int r = 0; r < 3; ++r)
                Cinv_b[3 * i + r] = (Cinv[9 * i + r * 3] * b[3 * i] + Cinv[9 * i + r * 3 + 1] * b[3 * i + 1]) + Cinv[9 * i + r * 3 + 2] * b[3 * i + 2];
        }
        for (int i = 0; i < Mp; ++i) {
            for (int o = P->obs_ptr[i]; o < P->obs_ptr[i + 1]; ++o) {
                if (P->obs_right[o]) continue;
                const int j = P->opt_index[P->obs_frame[o]];
                if (j < 0) continue;
                const double *Bji = B + ((size_t)j * Mp + i) * 18;
                double *BCi = BCinv + ((size_t)j * Mp + i) * 18;
                for (int r = 0; r < 6; ++r)
                    for (int c = 0; c < 3; ++c)
                        BCi[r * 3 + c] = (Bji[r * 3] * Cinv[9 * i + c] + Bji[r * 3 + 1] * Cinv[9 * i + 3 + c]) + Bji[r * 3 + 2] * Cinv[9 * i + 6 + c];
                for (int r = 0; r < 6; ++r)
                    BCinv_b[6 * j + r] += (BCi[r * 3] * b[3 * i] + BCi[r * 3 + 1] * b[3 * i + 1]) + BCi[r * 3 + 2] * b[3 * i + 2];
                for (int o2 = o; o2 < P->obs_ptr[i + 1]; ++o2) {
                    if (P->obs_right[o2]) continue;
                    const int k = P->opt_index[P->obs_frame[o2]];
                    if (k < 0) continue;
                    const double *Bki = B + ((size_t)k * Mp + i) * 18; /* Bt[i][k] = B[k][i]^T */
                    for (int r = 0; r < 6; ++r)
                        for (int c = 0; c < 6; ++c)
                            BCinvBt[(size_t)(6 * j + r) * n6 + 6 * k + c] += (BCi[r * 3] * Bki[c * 3] + BCi[r * 3 + 1] * Bki[c * 3 + 1]) + BCi[r * 3 + 2] * Bki[c * 3 + 2];
                }
            }
        }
        /* mirror (also transposes the diagonal blocks), :501-503 */
        for (int j = 0; j < No; ++j)
            for (int u = j; u < No; ++u) {
                double blk[36];
                for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) blk[r * 6 + c] = BCinvBt[(size_t)(6 * j + c) * n6 + 6 * u + r];
                for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) BCinvBt[(size_t)(6 * u + r) * n6 + 6 * j + c] = blk[r * 6 + c];
            }
        for (int r = 0; r < n6; ++r) for (int c = 0; c < n6; ++c) S[(size_t)r * n6 + c] = -BCinvBt[(size_t)r * n6 + c];
        for (int j = 0; j < No; ++j)
            for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c)
                S[(size_t)(6 * j + r) * n6 + 6 * j + c] = A[36 * j + r * 6 + c] - BCinvBt[(size_t)(6 * j + r) * n6 + 6 * j + c];
        for (int r = 0; r < n6; ++r) rhs[r] = a[r] - BCinv_b[r];
        if (n6 > 0) orc_ldlt_solve_d(S, n6, rhs, 1, x);
        /* y_i and updates */
        for (int i = 0; i < Mp; ++i) {
            double cbx[3] = {0, 0, 0};
            for (int o = P->obs_ptr[i]; o < P->obs_ptr[i + 1]; ++o) {
                if (P->obs_right[o]) continue;
                const int j = P->opt_index[P->obs_frame[o]];
                if (j < 0) continue;
                const double *BCi = BCinv + ((size_t)j * Mp + i) * 18; /* CinvBt[i][j] = BCinv[j][i]^T */
                for (int r = 0; r < 3; ++r) {
                    double s = 0;
                    for (int c = 0; c < 6; ++c) s += BCi[c * 3 + r] * x[6 * j + c];
                    cbx[r] += s;
                }
            }
            for (int r = 0; r < 3; ++r) X[3 * i + r] += Cinv_b[3 * i + r] - cbx[r];
        }
        for (int j = 0; j < No; ++j) orc_pose_retract_d(poses + 16 * opt_frame[j], x + 6 * j);
        const double average_error = sqrt(err / (double)P->n_obs);
        if (avg_err_out) avg_err_out[iter] = average_error;
        if (isnan(err)) { rc = -4; break; }
        flag_success = (average_error <= 1.0);
    }
    memcpy(poses_out, poses, sizeof(double) * 16 * N);
    memcpy(points_out, X, sizeof(double) * 3 * Mp);
    if (success) *success = (rc == 0) ? flag_success : 0;
    free(poses); free(X); free(opt_frame); free(A); free(a); free(B); free(BCinv); free(C); free(b); free(Cinv); free(Cinv_b);
    free(BCinvBt); free(BCinv_b); free(S); free(rhs); free(x);
    return rc;
}
