"""Triangulation / depth-filter / prior / compaction oracle (oracle/misc_oracle.c) -- TEST INFRASTRUCTURE ONLY."""
import ctypes

import numpy as np

from . import lib

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_i32p = ctypes.POINTER(ctypes.c_int)


def _p(a, t):
    return a.ctypes.data_as(t)


def svd4_null_f(M):
    M = np.ascontiguousarray(M, np.float32)
    v = np.zeros(4, np.float32)
    lib().orc_svd4_null_f(_p(M, _f32p), _p(v, _f32p))
    return v


def triangulate_dlt(pts0, pts1, R10, t10, K0, K1):
    pts0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2)
    pts1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    n = len(pts0)
    assert len(pts1) == n
    R10 = np.ascontiguousarray(R10, np.float32)
    t10 = np.ascontiguousarray(t10, np.float32)
    K0 = np.ascontiguousarray(K0, np.float32)
    K1 = np.ascontiguousarray(K1, np.float32)
    X0 = np.zeros((n, 3), np.float32)
    X1 = np.zeros((n, 3), np.float32)
    lib().orc_triangulate_dlt(_p(pts0, _f32p), _p(pts1, _f32p), n, _p(R10, _f32p), _p(t10, _f32p), _p(K0, _f32p),
                              _p(K1, _f32p), _p(X0, _f32p), _p(X1, _f32p))
    return X0, X1


def depth_filter_normal(x_prev, cov_prev, x_curr, cov_curr):
    a = [np.ascontiguousarray(v, np.float64) for v in (x_prev, cov_prev, x_curr, cov_curr)]
    n = len(a[0])
    x = np.zeros(n)
    c = np.zeros(n)
    lib().orc_depth_filter_normal(*[_p(v, _f64p) for v in a], n, _p(x, _f64p), _p(c, _f64p))
    return x, c


def depth_filter_student_t(x_prev, cov_prev, a, b, x_min, x_max, x_curr, cov_curr):
    """Returns (x_upd, cov_upd, a', b', x_min', x_max')."""
    xp, cp, xc, cc = [np.ascontiguousarray(v, np.float64) for v in (x_prev, cov_prev, x_curr, cov_curr)]
    a, b, lo, hi = [np.ascontiguousarray(v, np.float64).copy() for v in (a, b, x_min, x_max)]
    n = len(xp)
    x = np.zeros(n)
    c = np.zeros(n)
    lib().orc_depth_filter_student_t(_p(xp, _f64p), _p(cp, _f64p), _p(a, _f64p), _p(b, _f64p), _p(lo, _f64p),
                                     _p(hi, _f64p), _p(xc, _f64p), _p(cc, _f64p), n, _p(x, _f64p), _p(c, _f64p))
    return x, c, a, b, lo, hi


def calc_prior(pts0, Xw, Tw1, K4):
    pts0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2)
    Xw = np.ascontiguousarray(Xw, np.float32).reshape(-1, 3)
    n = len(Xw)
    Tw1 = np.ascontiguousarray(Tw1, np.float32)
    K4 = np.ascontiguousarray(K4, np.float32)
    out = pts0.copy()
    lib().orc_calc_prior(_p(pts0, _f32p), _p(Xw, _f32p), n, _p(Tw1, _f32p), _p(K4, _f32p), _p(out, _f32p))
    return out


def compact(mask):
    m = np.ascontiguousarray(mask).astype(np.uint8)
    idx = np.zeros(len(m), np.int32)
    k = lib().orc_compact(_p(m, _u8p), len(m), _p(idx, _i32p))
    return idx[:k]
