"""Stereo tracking step oracle -- TEST INFRASTRUCTURE ONLY.

Composes the per-stage oracles in the order of StereoVO::trackStereoImages
(core/visual_odometry/stereo_vo/stereo_vo.cpp:475-670): constant-velocity prior, trackWithPrior
(cv2), trackWithScale (C restatement), trackWithPrior (cv2), stereo pose-only GN (C restatement),
the y > 660 stub, with the stable compactions of StereoLandmarkTracking(src, mask) after each gate.
"""
import numpy as np

from . import klt as oklt
from . import pose as opose

f32 = np.float32


def mul4_f32(A, B):
    A = np.asarray(A, f32)
    B = np.asarray(B, f32)
    C = np.zeros((4, 4), f32)
    for i in range(4):
        for j in range(4):
            s = f32(0)
            for k in range(4):
                s = f32(s + f32(A[i, k] * B[k, j]))
            C[i, j] = s
    return C


def _xform(T, X):
    T = np.asarray(T, f32)
    return np.stack([((T[r, 0] * X[:, 0] + T[r, 1] * X[:, 1]) + T[r, 2] * X[:, 2]) + T[r, 3] for r in range(3)], 1).astype(f32)


def stereo_track_step(I0l, I1l, I1r, pts_l0, pts_r0, Xw, tri, T_wp, dT_prev, K_l, K_r, T_lr, win, max_level, thres_err,
                      thres_poseba, do_scale_refine=True, sampson_y=660.0, lk=oklt.lk_cv2, faithful_scale=False):
    h, w = I0l.shape
    pts_l0 = np.asarray(pts_l0, f32).reshape(-1, 2)
    pts_r0 = np.asarray(pts_r0, f32).reshape(-1, 2)
    Xw = np.asarray(Xw, f32).reshape(-1, 3)
    tri = np.asarray(tri).astype(bool)
    n = len(pts_l0)
    K_l = np.asarray(K_l, f32)
    K_r = np.asarray(K_r, f32)
    T_wc_prior = mul4_f32(T_wp, dT_prev)                      # :478
    T_cw_prior = opose.inverse_se3_f(T_wc_prior)              # :479
    T_pw = opose.inverse_se3_f(np.asarray(T_wp, f32))
    T_rl = opose.inverse_se3_f(np.asarray(T_lr, f32))
    # [3] priors (:485-522)
    Xl1 = _xform(T_cw_prior, Xw)
    Xr1 = _xform(T_rl, Xl1)
    Xl0 = _xform(T_pw, Xw)
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = np.where(tri, Xl0[:, 2] / Xl1[:, 2], f32(1)).astype(f32)
        izl, izr = f32(1) / Xl1[:, 2], f32(1) / Xr1[:, 2]
        pl = np.stack([K_l[0] * Xl1[:, 0] * izl + K_l[2], K_l[1] * Xl1[:, 1] * izl + K_l[3]], 1).astype(f32)
        pr = np.stack([K_r[0] * Xr1[:, 0] * izr + K_r[2], K_r[1] * Xr1[:, 1] * izr + K_r[3]], 1).astype(f32)
    off = f32(3.0)

    def in_image(p):
        return ~((p[:, 0] < off) | (p[:, 1] < off) | (p[:, 0] >= f32(w) - off) | (p[:, 1] >= f32(h) - off))
    use = tri & in_image(pl) & in_image(pr) & ~(Xl1[:, 2].astype(np.float64) < 0.1) & ~(Xr1[:, 2].astype(np.float64) < 0.1)
    prior_l1 = np.where(use[:, None], pl, pts_l0).astype(f32)
    prior_r1 = np.where(use[:, None], pr, pts_r0).astype(f32)
    idx = np.arange(n)
    counts = []
    # [4] l0 -> l1, compaction
    p_l1, m = oklt.track_with_prior(lk, I0l, I1l, pts_l0, prior_l1, win, max_level, thres_err)
    idx, l0, l1, r1p, sc = idx[m], pts_l0[m], p_l1[m], prior_r1[m], scale[m]
    counts.append(len(idx))
    # [4-1] scale refinement
    if do_scale_refine:
        l1, m = oklt.track_with_scale(I0l, I1l, l0, sc, l1, faithful=faithful_scale)
        idx, l0, l1, r1p = idx[m], l0[m], l1[m], r1p[m]
    counts.append(len(idx))
    # [5] l1 -> r1
    r1, m = oklt.track_with_prior(lk, I1l, I1r, l1, r1p, win, max_level, thres_err)
    idx, l1, r1 = idx[m], l1[m], r1[m]
    counts.append(len(idx))
    # [6] pose-only GN on triangulated survivors
    sel = tri[idx]
    Xp = _xform(T_pw, Xw[idx[sel]])
    counts.append(int(sel.sum()))
    ok, dT, mask_po, iters = opose.pose_gn_stereo(Xp, l1[sel], r1[sel], K_l, K_r, T_lr, thres_poseba, dT_prev)
    if not ok:
        raise RuntimeError("PoseOnlyStereoBA is failed!")
    mask_motion = np.ones(len(idx), bool)
    mask_motion[np.flatnonzero(sel)] = mask_po
    T_wc = mul4_f32(T_wp, dT)
    idx, l1, r1 = idx[mask_motion], l1[mask_motion], r1[mask_motion]
    # [7] stub
    keep = ~(l1[:, 1] > f32(sampson_y))
    idx, l1, r1 = idx[keep], l1[keep], r1[keep]
    counts.append(len(idx))
    return dict(T_wc=T_wc, dT_pc=dT, index=idx.astype(np.int32), pts_l1=l1, pts_r1=r1, counts=counts, gn_iters=iters)
