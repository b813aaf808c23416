"""Pose-only Gauss-Newton oracle (oracle/pose_oracle.c) -- TEST INFRASTRUCTURE ONLY."""
import ctypes

import numpy as np

from . import lib

_f32p = ctypes.POINTER(ctypes.c_float)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def _p(a, t):
    return a.ctypes.data_as(t)


def se3exp_f(xi):
    xi = np.ascontiguousarray(xi, np.float32)
    T = np.zeros((4, 4), np.float32)
    lib().orc_se3exp_f(_p(xi, _f32p), _p(T, _f32p))
    return T


def inverse_se3_f(T):
    T = np.ascontiguousarray(T, np.float32)
    Ti = np.zeros((4, 4), np.float32)
    lib().orc_inverse_se3_f(_p(T, _f32p), _p(Ti, _f32p))
    return Ti


def ldlt6_solve_f(A, b):
    A = np.ascontiguousarray(A, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    x = np.zeros(6, np.float32)
    lib().orc_ldlt6_solve_f(_p(A, _f32p), _p(b, _f32p), _p(x, _f32p))
    return x


def pose_gn_stereo(X, pl, pr, Kl, Kr, T_lr, thres, T01_init, want_trace=False, max_iter=100, no_early_stop=False):
    """poseOnlyBundleAdjustment_Stereo. Returns (success, T01, mask, iters[, trace]).
    max_iter / no_early_stop are test-only knobs (fixed-point and iterate-by-iterate comparisons)."""
    X = np.ascontiguousarray(X, np.float32).reshape(-1, 3)
    pl = np.ascontiguousarray(pl, np.float32).reshape(-1, 2)
    pr = np.ascontiguousarray(pr, np.float32).reshape(-1, 2)
    n = len(X)
    assert len(pl) == n and len(pr) == n
    Kl = np.ascontiguousarray(Kl, np.float32)
    Kr = np.ascontiguousarray(Kr, np.float32)
    T_lr = np.ascontiguousarray(T_lr, np.float32)
    T01 = np.ascontiguousarray(T01_init, np.float32).copy()
    mask = np.zeros(n, np.uint8)
    iters = ctypes.c_int(0)
    trace = np.zeros((100, 24), np.float32)
    ok = lib().orc_pose_gn_stereo_ex(_p(X, _f32p), _p(pl, _f32p), _p(pr, _f32p), n, _p(Kl, _f32p), _p(Kr, _f32p),
                                     _p(T_lr, _f32p), ctypes.c_float(thres), _p(T01, _f32p), _p(mask, _u8p),
                                     ctypes.byref(iters), _p(trace, _f32p), int(max_iter), int(bool(no_early_stop)))
    out = (bool(ok), T01, mask.astype(bool), iters.value)
    return out + (trace[:iters.value],) if want_trace else out


def pose_gn_mono(X, p1, K, thres, R01_init, t01_init, variant=0, want_trace=False, max_iter=100, no_early_stop=False):
    """poseOnlyBundleAdjustment (variant 0 = core, 1 = standalone)."""
    X = np.ascontiguousarray(X, np.float32).reshape(-1, 3)
    p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 2)
    n = len(X)
    assert len(p1) == n
    R01 = np.ascontiguousarray(R01_init, np.float32).copy()
    t01 = np.ascontiguousarray(t01_init, np.float32).copy()
    mask = np.zeros(n, np.uint8)
    iters = ctypes.c_int(0)
    trace = np.zeros((100, 24), np.float32)
    f = [ctypes.c_float(float(k)) for k in K]
    ok = lib().orc_pose_gn_mono_ex(_p(X, _f32p), _p(p1, _f32p), n, f[0], f[1], f[2], f[3], int(thres), _p(R01, _f32p),
                                   _p(t01, _f32p), _p(mask, _u8p), int(variant), ctypes.byref(iters), _p(trace, _f32p),
                                   int(max_iter), int(bool(no_early_stop)))
    out = (bool(ok), R01, t01, mask.astype(bool), iters.value)
    return out + (trace[:iters.value],) if want_trace else out
