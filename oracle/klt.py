"""KLT oracle -- TEST INFRASTRUCTURE ONLY.

* ``lk_cv2``   : ``cv2.calcOpticalFlowPyrLK`` exactly as the reference calls it.
* ``lk_c``     : the scalar C restatement (oracle/klt_oracle.c), pinned against cv2.
* ``track*``   : the four ``FeatureTracker`` front-ends
  (``core/visual_odometry/feature_tracker.cpp:13-206``) restated line for line on top
  of a pluggable LK callable, so the same post-filters can be evaluated with either
  LK implementation.
"""
import ctypes

import numpy as np

from . import lib

OPTFLOW_USE_INITIAL_FLOW = 4

_u8p = ctypes.POINTER(ctypes.c_uint8)
_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)
_i16p = ctypes.POINTER(ctypes.c_int16)


def _p(a, t):
    return a.ctypes.data_as(t)


def effective_max_level(w, h, win, max_level):
    return lib().orc_effective_max_level(int(w), int(h), int(win), int(max_level))


def pyrdown(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().orc_pyrdown_u8(_p(img, _u8p), w, h, w, _p(out, _u8p), out.shape[1])
    return out


def scharr(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    out = np.empty((h, w, 2), np.int16)
    lib().orc_scharr_s16(_p(img, _u8p), w, h, w, _p(out, _i16p), 2 * w)
    return out


def build_pyramid(img, win, max_level):
    """Returns ([levels u8], [derivs int16 HxWx2])."""
    L = lib()
    L.orc_pyramid_build.restype = ctypes.c_void_p
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    P = ctypes.c_void_p(L.orc_pyramid_build(_p(img, _u8p), w, h, w, int(win), int(max_level), 1))
    n = L.orc_pyramid_nlevels(P)
    levels, derivs = [], []
    for l in range(n):
        lw, lh = ctypes.c_int(), ctypes.c_int()
        L.orc_pyramid_level_size(P, l, ctypes.byref(lw), ctypes.byref(lh))
        a = np.empty((lh.value, lw.value), np.uint8)
        d = np.empty((lh.value, lw.value, 2), np.int16)
        L.orc_pyramid_get_level(P, l, _p(a, _u8p))
        L.orc_pyramid_get_deriv(P, l, _p(d, _i16p))
        levels.append(a)
        derivs.append(d)
    L.orc_pyramid_free(P)
    return levels, derivs


def lk_c(img0, img1, pts0, win, max_level, flags=0, prior=None, return_iters=False):
    """C restatement of cv::calcOpticalFlowPyrLK with default criteria."""
    img0 = np.ascontiguousarray(img0, np.uint8)
    img1 = np.ascontiguousarray(img1, np.uint8)
    h, w = img0.shape
    pts0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2)
    n = pts0.shape[0]
    nxt = (np.ascontiguousarray(prior, np.float32).reshape(-1, 2).copy()
           if (flags & OPTFLOW_USE_INITIAL_FLOW) else pts0.copy())
    status = np.zeros(n, np.uint8)
    err = np.zeros(n, np.float32)
    iters = np.zeros((16, max(n, 1)), np.int32)
    nl = lib().orc_calc_optical_flow_pyr_lk(
        _p(img0, _u8p), _p(img1, _u8p), w, h, img0.strides[0], img1.strides[0],
        _p(pts0, _f32p), _p(nxt, _f32p), n, int(win), int(max_level), int(flags),
        _p(status, _u8p), _p(err, _f32p), _p(iters, _i32p))
    if return_iters:
        return nxt, status, err, iters[:nl, :n]
    return nxt, status, err


def lk_cv2(img0, img1, pts0, win, max_level, flags=0, prior=None):
    """The reference's own library call (feature_tracker.cpp:29 etc.)."""
    import cv2
    pts0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 1, 2)
    if pts0.shape[0] == 0:
        return np.zeros((0, 2), np.float32), np.zeros(0, np.uint8), np.zeros(0, np.float32)
    if flags & OPTFLOW_USE_INITIAL_FLOW:
        nxt = np.ascontiguousarray(prior, np.float32).reshape(-1, 1, 2).copy()
        p1, st, err = cv2.calcOpticalFlowPyrLK(img0, img1, pts0, nxt, winSize=(win, win),
                                               maxLevel=max_level, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    else:
        p1, st, err = cv2.calcOpticalFlowPyrLK(img0, img1, pts0, None, winSize=(win, win),
                                               maxLevel=max_level)
    return p1.reshape(-1, 2), st.reshape(-1), err.reshape(-1)


# --------------------------------------------------------------------------
# FeatureTracker front-ends (feature_tracker.cpp:13-206), LK pluggable.
# mask_in mirrors `mask_valid.resize(n, true)` keeping pre-existing entries
# (SURVEY Appendix B #7): pass None for a fresh all-true mask.
# --------------------------------------------------------------------------
def _mask0(mask_in, n):
    if mask_in is None:
        return np.ones(n, bool)
    m = np.ones(n, bool)
    k = min(n, len(mask_in))
    m[:k] = np.asarray(mask_in, bool)[:k]
    return m


def track(lk, img0, img1, pts0, win, max_lvl, thres_err, mask_in=None):
    """feature_tracker.cpp:13-37"""
    pts0 = np.asarray(pts0, np.float32).reshape(-1, 2)
    m = _mask0(mask_in, len(pts0))
    p1, st, err = lk(img0, img1, pts0, win, max_lvl)
    with np.errstate(invalid="ignore"):
        m &= (st > 0) & (np.where(st > 0, err, 0) <= np.float32(thres_err))
    return p1, m


def track_with_prior(lk, img0, img1, pts0, prior, win, max_lvl, thres_err, mask_in=None):
    """feature_tracker.cpp:171-206"""
    pts0 = np.asarray(pts0, np.float32).reshape(-1, 2)
    h, w = img0.shape
    m = _mask0(mask_in, len(pts0))
    p1, st, err = lk(img0, img1, pts0, win, max_lvl, OPTFLOW_USE_INITIAL_FLOW, prior)
    m &= (st > 0) & (p1[:, 0] > 0) & (p1[:, 0] < w) & (p1[:, 1] > 0) & (p1[:, 1] < h)
    m &= np.where(st > 0, err, 0) <= np.float32(thres_err)
    return p1, m


def track_bidirection(lk, img0, img1, pts0, win, max_lvl, thres_err, thres_bi, mask_in=None):
    """feature_tracker.cpp:39-86 (backward pass at maxLevel-1, INITIAL_FLOW from pts0)."""
    pts0 = np.asarray(pts0, np.float32).reshape(-1, 2)
    h, w = img0.shape
    m = _mask0(mask_in, len(pts0))
    thres_bi2 = np.float32(thres_bi) * np.float32(thres_bi)
    p1, stf, errf = lk(img0, img1, pts0, win, max_lvl)
    pb, stb, errb = lk(img1, img0, p1, win, max_lvl - 1, OPTFLOW_USE_INITIAL_FLOW, pts0)
    dp = pb - pts0
    dist2 = dp[:, 0] * dp[:, 0] + dp[:, 1] * dp[:, 1]
    m &= (p1[:, 0] > 3) & (p1[:, 0] < w - 3) & (p1[:, 1] > 3) & (p1[:, 1] < h - 3)
    ok = (stf > 0) & (stb > 0)
    m &= ok & (np.where(stf > 0, errf, 0) <= np.float32(thres_err)) \
            & (np.where(stb > 0, errb, 0) <= np.float32(thres_err)) & (dist2 <= thres_bi2)
    return p1, m


def track_bidirection_with_prior(lk, img0, img1, pts0, prior, win, max_lvl, thres_err, thres_bi,
                                 mask_in=None):
    """feature_tracker.cpp:88-169 (both passes INITIAL_FLOW at maxLevel; 5x bidirectional gate)."""
    pts0 = np.asarray(pts0, np.float32).reshape(-1, 2)
    h, w = img0.shape
    m = _mask0(mask_in, len(pts0))
    thres_bi2 = np.float32(thres_bi) * np.float32(thres_bi)
    p1, stf, errf = lk(img0, img1, pts0, win, max_lvl, OPTFLOW_USE_INITIAL_FLOW, prior)
    pb, stb, errb = lk(img1, img0, p1, win, max_lvl, OPTFLOW_USE_INITIAL_FLOW, pts0)
    dp = pb - pts0
    dist2 = dp[:, 0] * dp[:, 0] + dp[:, 1] * dp[:, 1]
    m &= (p1[:, 0] > 0) & (p1[:, 0] < w) & (p1[:, 1] > 0) & (p1[:, 1] < h)
    m &= (stf > 0) & (np.where(stf > 0, errf, 0) <= np.float32(thres_err))
    m &= (stb > 0) & (np.where(stb > 0, errb, 0) <= np.float32(thres_err))
    m &= dist2 <= thres_bi2 * np.float32(5)
    return p1, m


# --------------------------------------------------------------------------
# trackWithScale (feature_tracker.cpp:236-504) -- oracle/klt_scale_oracle.c
# --------------------------------------------------------------------------
def sobel3_f32(img):
    """cv::Sobel(img, CV_32F, 1,0 / 0,1, ksize=3, BORDER_DEFAULT) restated (stereo_vo.cpp:551-552)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    du = np.empty((h, w), np.float32)
    dv = np.empty((h, w), np.float32)
    lib().orc_sobel3_f32(_p(img, _u8p), w, h, w, _p(du, _f32p), _p(dv, _f32p))
    return du, dv


def track_with_scale(img0, img1, pts0, scale_est, pts_track, mask_in=None, du0=None, dv0=None, faithful=False,
                     return_iters=False):
    img0 = np.ascontiguousarray(img0, np.uint8)
    img1 = np.ascontiguousarray(img1, np.uint8)
    h, w = img0.shape
    if du0 is None:
        du0, dv0 = sobel3_f32(img0)
    du0 = np.ascontiguousarray(du0, np.float32)
    dv0 = np.ascontiguousarray(dv0, np.float32)
    pts0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2)
    n = len(pts0)
    sc = np.ascontiguousarray(scale_est, np.float32)
    pt = np.ascontiguousarray(pts_track, np.float32).reshape(-1, 2).copy()
    if len(pt) != n:
        raise RuntimeError("pts_track.size() != pts0.size()")
    m = _mask0(mask_in, n).astype(np.uint8)
    iters = np.zeros(max(n, 1), np.int32)
    rc = lib().orc_track_with_scale(_p(img0, _u8p), _p(du0, _f32p), _p(dv0, _f32p), _p(img1, _u8p), w, h, w, w,
                                    _p(pts0, _f32p), _p(sc, _f32p), n, _p(pt, _f32p), _p(m, _u8p),
                                    1 if faithful else 0, _p(iters, _i32p))
    if rc != 0:
        raise RuntimeError("ax ay nan / dtu dtv nan")
    if return_iters:
        return pt, m.astype(bool), iters[:n]
    return pt, m.astype(bool)
