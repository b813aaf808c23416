"""Mono tracking step oracle -- TEST INFRASTRUCTURE ONLY.

Composes the per-stage oracles in the order of the steady-state branch of MonoVO::trackImage
(core/visual_odometry/mono_vo/mono_vo.cpp:724-1010): constant-velocity prior + patch scale for bundled landmarks
(:739-761), trackBidirectionWithPrior (cv2, :768), trackWithScale (C restatement, :783), landmark selection for the
pose-only BA (:799-827), mono poseOnlyBundleAdjustment (C restatement, :864), Sampson gate
(motion_estimator.cpp:539-568, mono_vo.cpp:957-962), bucketed extraction of new features on the current image and their
back-tracking into the previous one with trackBidirection (:985-992).  The 5-point fallback (:909-949) and the
initialisation step of the second image (:562-659) take ``five_point`` = a callable (pts0, pts1) -> (ok, R10, t10, mask):
by default the reference's own library call (oracle/five_point.py: cv2.findEssentialMat + restated decomposition).
"""
import numpy as np

from . import detect as odet
from . import five_point as ofp
from . import klt as oklt
from . import pose as opose
from . import step as ostep

f32 = np.float32


def inv3_f32(M):
    """Eigen::Matrix3f::inverse() (camera.cpp:42, third-party, unpinned) restated: cofactors times 1/det."""
    m = np.asarray(M, f32)
    c = np.zeros((3, 3), f32)

    def cof(a, b, cc, d):   # m[a]*m[b] - m[cc]*m[d] on flattened indices
        mf = m.reshape(9)
        return f32(f32(mf[a] * mf[b]) - f32(mf[cc] * mf[d]))
    c[0, 0] = cof(4, 8, 5, 7); c[0, 1] = cof(2, 7, 1, 8); c[0, 2] = cof(1, 5, 2, 4)
    c[1, 0] = cof(5, 6, 3, 8); c[1, 1] = cof(0, 8, 2, 6); c[1, 2] = cof(2, 3, 0, 5)
    c[2, 0] = cof(3, 7, 4, 6); c[2, 1] = cof(1, 6, 0, 7); c[2, 2] = cof(0, 4, 1, 3)
    det = f32(f32(f32(m[0, 0] * c[0, 0]) + f32(m[0, 1] * c[1, 0])) + f32(m[0, 2] * c[2, 0]))
    return (c * f32(f32(1.0) / det)).astype(f32)


def mul3_f32(A, B):
    A, B = np.asarray(A, f32), np.asarray(B, f32)
    C = np.zeros((3, 3), f32)
    for i in range(3):
        for j in range(3):
            s = f32(0)
            for k in range(3):
                s = f32(s + f32(A[i, k] * B[k, j]))
            C[i, j] = s
    return C


def fundamental(K4, R10, t10):
    """E10 = skew(t10) R10, F10 = Kinv^T E10 Kinv (motion_estimator.cpp:549-551)."""
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], f32)
    Kinv = inv3_f32(K)
    t = np.asarray(t10, f32)
    S = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]], f32)
    E = mul3_f32(S, R10)
    return mul3_f32(mul3_f32(Kinv.T.copy(), E), Kinv)


def sampson(pts0, pts1, F):
    """motion_estimator.cpp:553-568, float32 in the reference's operation order."""
    p0, p1 = np.asarray(pts0, f32).reshape(-1, 2), np.asarray(pts1, f32).reshape(-1, 2)
    F = np.asarray(F, f32)
    one = f32(1.0)
    a = [((F[r, 0] * p0[:, 0] + F[r, 1] * p0[:, 1]) + F[r, 2] * one).astype(f32) for r in range(3)]      # F10 p0
    b = [((F[0, r] * p1[:, 0] + F[1, r] * p1[:, 1]) + F[2, r] * one).astype(f32) for r in range(3)]      # F10^T p1
    num = ((p1[:, 0] * a[0] + p1[:, 1] * a[1]) + one * a[2]).astype(f32)
    num = (num * num).astype(f32)
    den = (((a[0] * a[0] + a[1] * a[1]) + b[0] * b[0]) + b[1] * b[1]).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (num / den).astype(f32)


def mono_frame_step(I0, I1, pts0, Xw, triangulated, bundled, T_wc_prev, dT01_prior, K4, win, max_level, thres_err, thres_bi,
                    thres_sampson, thres_poseba, use_bundled_only, n_bins_u=0, n_bins_v=0, det_edge=31, det_min_score=0,
                    do_scale_refine=True, lk=oklt.lk_cv2, five_point=None, thres_5p=0.0, detect_fn=None, faithful_scale=False):
    h, w = I0.shape
    pts0 = np.asarray(pts0, f32).reshape(-1, 2)
    Xw = np.asarray(Xw, f32).reshape(-1, 3)
    tri, bun = np.asarray(triangulated).astype(bool), np.asarray(bundled).astype(bool)
    n = len(pts0)
    K4 = np.asarray(K4, f32)
    Twc_prev = np.asarray(T_wc_prev, f32)
    Tcw_prev = opose.inverse_se3_f(Twc_prev)
    dT01_prior = np.asarray(dT01_prior, f32)
    Twc_prior = ostep.mul4_f32(Twc_prev, dT01_prior)
    Tcw_prior = opose.inverse_se3_f(Twc_prior)
    # prior + scale (:739-761)
    Xp = ostep._xform(Tcw_prev, Xw)
    Xc = ostep._xform(Tcw_prior, Xw)
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = np.where(bun, Xp[:, 2] / Xc[:, 2], f32(1)).astype(f32)
        invz = f32(1) / Xc[:, 2]
        proj = np.stack([K4[0] * Xc[:, 0] * invz + K4[2], K4[1] * Xc[:, 1] * invz + K4[3]], 1).astype(f32)
    use = bun & (Xc[:, 2] > 0)
    prior = np.where(use[:, None], proj, pts0).astype(f32)
    idx = np.arange(n)
    counts = []
    # K4 (:768)
    p1, m = oklt.track_bidirection_with_prior(lk, I0, I1, pts0, prior, win, max_level, thres_err, thres_bi)
    idx, p0c, p1c, sc = idx[m], pts0[m], p1[m], scale[m]
    counts.append(len(idx))
    # K7 (:783)
    if do_scale_refine:
        p1c, m = oklt.track_with_scale(I0, I1, p0c, sc, p1c, faithful=faithful_scale)
        idx, p0c, p1c = idx[m], p0c[m], p1c[m]
    counts.append(len(idx))
    # selection (:799-827)
    flag = bun[idx] if use_bundled_only else tri[idx]
    sel = flag & (Xp[idx, 2].astype(np.float64) > 0.1)
    counts.append(int(sel.sum()))
    ok, iters = False, 0
    if sel.sum() > 10:
        ok, R01, t01, mask_ba, iters = opose.pose_gn_mono(Xp[idx[sel]], p1c[sel], K4, int(thres_poseba), dT01_prior[:3, :3], dT01_prior[:3, 3])
    mask_motion = np.ones(len(idx), bool)
    used_5point = False
    if ok:
        mask_motion[np.flatnonzero(sel)] = mask_ba
        dT01 = np.eye(4, dtype=f32)
        dT01[:3, :3], dT01[:3, 3] = R01, t01
        dT10 = opose.inverse_se3_f(dT01)
    else:
        # :909-949: five-point on all K7 survivors, translation scaled to the previous motion's length
        if five_point is None and not thres_5p > 0:
            raise RuntimeError("insufficient points / pose-only BA failed: 5-point fallback required")
        fp = five_point or (lambda a, b: _five_point_cv(a, b, K4, thres_5p))
        ok5, R10, t10, m5 = fp(p0c, p1c)
        if not ok5:
            raise RuntimeError("'calcPose5PointsAlgorithm()' is failed. Terminate the algorithm.")
        t10 = np.asarray(t10, f32)
        scale5 = f32(np.sqrt(f32(f32(f32(dT01_prior[0, 3] * dT01_prior[0, 3]) + f32(dT01_prior[1, 3] * dT01_prior[1, 3])) + f32(dT01_prior[2, 3] * dT01_prior[2, 3]))))
        nrm = f32(np.sqrt(f32(f32(f32(t10[0] * t10[0]) + f32(t10[1] * t10[1])) + f32(t10[2] * t10[2]))))
        dT10 = np.eye(4, dtype=f32)
        dT10[:3, :3] = np.asarray(R10, f32)
        dT10[:3, 3] = (f32(scale5 / nrm) * t10).astype(f32)
        dT01 = opose.inverse_se3_f(dT10)
        mask_motion = np.asarray(m5, bool)
        used_5point = True
    T_wc = ostep.mul4_f32(Twc_prev, dT01)
    idx, p0c, p1c = idx[mask_motion], p0c[mask_motion], p1c[mask_motion]
    counts.append(len(idx))
    # Sampson gate (:957-962)
    F = fundamental(K4, dT10[:3, :3], dT10[:3, 3])
    d = sampson(p0c, p1c, F)
    with np.errstate(invalid="ignore"):
        keep = d < f32(thres_sampson)
    idx, p0c, p1c = idx[keep], p0c[keep], p1c[keep]
    counts.append(len(idx))
    out = dict(T_wc=T_wc, dT01=dT01, dT10=dT10, index=idx.astype(np.int32), pts1=p1c, counts=counts, gn_iters=iters, F10=F, used_5point=used_5point)
    # new features (:981-992): extraction on I1 with the survivors as occupancy, back-tracking I1 -> I0
    if n_bins_u * n_bins_v > 0:
        pts_new = detect_fn(I1, p1c) if detect_fn else odet.detect_bucketed(I1, p1c, n_bins_u, n_bins_v, det_edge, det_min_score)
        if len(pts_new):
            p0_new, m = oklt.track_bidirection(lk, I1, I0, pts_new, win, max_level, thres_err, thres_bi)
            out.update(new_p1=pts_new[m], new_p0=p0_new[m])
        else:
            out.update(new_p1=np.zeros((0, 2), f32), new_p0=np.zeros((0, 2), f32))
        out["n_detected"] = len(pts_new)
    return out


def _five_point_cv(p0, p1, K4, thres_5p):
    ok, R, t, X0, m, E = ofp.calc_pose_5point(p0, p1, K4, thres_5p)
    return ok, R, t, m


def mono_init_step(I0, I1, pts0, T_wc_prev, K4, win, max_level, thres_err, thres_bi, thres_sampson, thres_5p, n_bins_u=0, n_bins_v=0,
                   det_edge=31, det_min_score=0, lk=oklt.lk_cv2, five_point=None, detect_fn=None):
    """The second image of a sequence (mono_vo.cpp:562-659): track (:573), calcPose5PointsAlgorithm (:589), Sampson gate
    (:594-597), |t10| = 1 (:606-609), pose (:611), new features back-tracked with trackBidirection (:623-636)."""
    pts0 = np.asarray(pts0, f32).reshape(-1, 2)
    K4 = np.asarray(K4, f32)
    Twc_prev = np.asarray(T_wc_prev, f32)
    idx = np.arange(len(pts0))
    p1, m = oklt.track(lk, I0, I1, pts0, win, max_level, thres_err)
    idx, p0c, p1c = idx[m], pts0[m], p1[m]
    counts = [len(idx)]
    fp = five_point or (lambda a, b: _five_point_cv(a, b, K4, thres_5p))
    ok5, R10, t10, m5 = fp(p0c, p1c)
    if not ok5:
        raise RuntimeError("calcPose5PointsAlgorithm() is failed.")
    t10 = np.asarray(t10, f32)
    nrm = f32(np.sqrt(f32(f32(f32(t10[0] * t10[0]) + f32(t10[1] * t10[1])) + f32(t10[2] * t10[2]))))
    dT10 = np.eye(4, dtype=f32)
    dT10[:3, :3] = np.asarray(R10, f32)
    dT10[:3, 3] = ((t10 / nrm).astype(f32) * f32(1.0)).astype(f32)
    dT01 = opose.inverse_se3_f(dT10)
    T_wc = ostep.mul4_f32(Twc_prev, dT01)
    F = fundamental(K4, dT10[:3, :3], dT10[:3, 3])
    d = sampson(p0c, p1c, F)
    with np.errstate(invalid="ignore"):
        keep = np.asarray(m5, bool) & (d < f32(thres_sampson))
    counts.append(int(np.asarray(m5, bool).sum()))
    idx, p0c, p1c = idx[keep], p0c[keep], p1c[keep]
    counts.append(len(idx))
    out = dict(T_wc=T_wc, dT01=dT01, dT10=dT10, index=idx.astype(np.int32), pts1=p1c, counts=counts, F10=F, used_5point=True)
    if n_bins_u * n_bins_v > 0:
        pts_new = detect_fn(I1, p1c) if detect_fn else odet.detect_bucketed(I1, p1c, n_bins_u, n_bins_v, det_edge, det_min_score)
        if len(pts_new):
            p0_new, m = oklt.track_bidirection(lk, I1, I0, pts_new, win, max_level, thres_err, thres_bi)
            out.update(new_p1=pts_new[m], new_p0=p0_new[m])
        else:
            out.update(new_p1=np.zeros((0, 2), f32), new_p0=np.zeros((0, 2), f32))
        out["n_detected"] = len(pts_new)
    return out


def symmetric_epipolar(pts0, pts1, F):
    """motion_estimator.cpp:638-652, float32 in the reference's operation order."""
    p0, p1 = np.asarray(pts0, f32).reshape(-1, 2), np.asarray(pts1, f32).reshape(-1, 2)
    F = np.asarray(F, f32)
    one = f32(1.0)
    a = [((F[r, 0] * p0[:, 0] + F[r, 1] * p0[:, 1]) + F[r, 2] * one).astype(f32) for r in range(3)]
    b = [((F[0, r] * p1[:, 0] + F[1, r] * p1[:, 1]) + F[2, r] * one).astype(f32) for r in range(2)]
    num = np.abs(((p1[:, 0] * a[0] + p1[:, 1] * a[1]) + one * a[2]).astype(f32))
    with np.errstate(divide="ignore", invalid="ignore"):
        den = (one / np.sqrt((a[0] * a[0] + a[1] * a[1]).astype(f32)).astype(f32) +
               one / np.sqrt((b[0] * b[0] + b[1] * b[1]).astype(f32)).astype(f32)).astype(f32)
        return (num * den).astype(f32)


def inliers_1point_histogram(pts0, pts1, K4, thres_1p):
    """MotionEstimator::findInliers1PointHistogram (motion_estimator.cpp:471-537) with histogram::makeHistogram
    (core/util/histogram.h:11-35) and medianHistogram (histogram.cpp:4-28: the centre of the fullest bin; std::sort leaves
    ties unspecified -- the first such bin here).  Returns (theta_opt, mask, theta, counts, R10, t10)."""
    p0, p1 = np.asarray(pts0, f32).reshape(-1, 2), np.asarray(pts1, f32).reshape(-1, 2)
    K4 = np.asarray(K4, f32)
    ifx, ify = f32(f32(1.0) / K4[0]), f32(f32(1.0) / K4[1])
    x0, y0 = ((p0[:, 0] - K4[2]) * ifx).astype(f32), ((p0[:, 1] - K4[3]) * ify).astype(f32)
    x1, y1 = ((p1[:, 0] - K4[2]) * ifx).astype(f32), ((p1[:, 1] - K4[3]) * ify).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        val = ((x0 * y1 - y0 * x1).astype(f32) / (y0 * f32(1) + f32(1) * y1).astype(f32)).astype(f32)
    theta = (np.float64(-2.0) * np.arctan(val).astype(f32).astype(np.float64)).astype(f32)
    nb = 400
    hmin, hmax = f32(-0.5), f32(0.5)
    step = f32(f32(hmax - hmin) / f32(nb))
    centers = np.zeros(nb, f32)
    centers[0] = hmin
    for i in range(1, nb - 1):
        centers[i] = f32(centers[i - 1] + step)
    centers[nb - 1] = hmax
    with np.errstate(invalid="ignore"):
        q = np.floor(((theta - hmin).astype(f32) / step).astype(f32))
        ok = (q >= 0) & (q < nb)
    counts = np.bincount(q[ok].astype(np.int64), minlength=nb)
    th = centers[int(np.argmax(counts))]
    c, s_ = f32(np.cos(th)), f32(np.sin(th))
    R10 = np.array([[c, 0, s_], [0, 1, 0], [-s_, 0, c]], f32)
    t10 = np.array([f32(np.sin(f32(th * f32(0.5)))), 0, f32(np.cos(f32(th * f32(0.5))))], f32)
    d = symmetric_epipolar(p0, p1, fundamental(K4, R10, t10))
    with np.errstate(invalid="ignore"):
        mask = d <= f32(f32(thres_1p) * f32(thres_1p))
    return float(th), mask, theta, counts, R10, t10
