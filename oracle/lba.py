"""Local bundle adjustment oracle (oracle/lba_oracle.c) -- TEST INFRASTRUCTURE ONLY."""
import ctypes

import numpy as np

from . import lib

_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int)
_u8p = ctypes.POINTER(ctypes.c_uint8)


class _Problem(ctypes.Structure):
    _fields_ = [
        ("n_frames", ctypes.c_int), ("n_opt", ctypes.c_int), ("n_points", ctypes.c_int), ("n_obs", ctypes.c_int),
        ("poses", _f64p), ("opt_index", _i32p), ("points", _f64p), ("obs_ptr", _i32p),
        ("obs_frame", _i32p), ("obs_right", _u8p), ("obs_px", _f64p),
        ("K_l", ctypes.c_double * 4), ("K_r", ctypes.c_double * 4), ("T_lr", ctypes.c_double * 16),
        ("is_stereo", ctypes.c_int), ("huber", ctypes.c_double), ("lam", ctypes.c_double), ("max_iter", ctypes.c_int),
    ]


def pack(p, struct_cls=_Problem):
    """dict (synth.lba_problem layout) -> (ctypes struct, keep-alive list)."""
    keep = {}
    for k, dt in (("poses", np.float64), ("opt_index", np.int32), ("points", np.float64), ("obs_ptr", np.int32),
                  ("obs_frame", np.int32), ("obs_right", np.uint8), ("obs_px", np.float64)):
        keep[k] = np.ascontiguousarray(p[k], dt)
    s = struct_cls()
    s.n_frames, s.n_opt, s.n_points, s.n_obs = int(p["n_frames"]), int(p["n_opt"]), int(p["n_points"]), int(p["n_obs"])
    for k in keep:
        ptr_t = dict(struct_cls._fields_)[k]
        setattr(s, k, keep[k].ctypes.data_as(ptr_t))
    s.K_l = (ctypes.c_double * 4)(*[float(v) for v in p["K_l"]])
    s.K_r = (ctypes.c_double * 4)(*[float(v) for v in p["K_r"]])
    s.T_lr = (ctypes.c_double * 16)(*[float(v) for v in np.asarray(p["T_lr"], np.float64).ravel()])
    s.is_stereo = int(p["is_stereo"])
    s.huber = float(p["huber"])
    setattr(s, dict(struct_cls._fields_).get("lam") and "lam" or "lambda_", float(p["lam"]))
    s.max_iter = int(p["max_iter"])
    return s, keep


def lba_solve(p, fix_b_accumulate=False):
    """Returns (rc, poses [n_frames,4,4], points [M,3], avg_err [max_iter], success)."""
    s, keep = pack(p)
    poses = np.zeros((s.n_frames, 4, 4))
    points = np.zeros((s.n_points, 3))
    avg = np.zeros(s.max_iter)
    ok = ctypes.c_int(0)
    rc = lib().orc_lba_solve(ctypes.byref(s), poses.ctypes.data_as(_f64p), points.ctypes.data_as(_f64p),
                             avg.ctypes.data_as(_f64p), ctypes.byref(ok), 1 if fix_b_accumulate else 0)
    return rc, poses, points, avg, bool(ok.value)


def se3exp_d(xi):
    xi = np.ascontiguousarray(xi, np.float64)
    T = np.zeros((4, 4))
    lib().orc_se3exp_d(xi.ctypes.data_as(_f64p), T.ctypes.data_as(_f64p))
    return T


def se3log_d(T):
    T = np.ascontiguousarray(T, np.float64)
    xi = np.zeros(6)
    lib().orc_se3log_d(T.ctypes.data_as(_f64p), xi.ctypes.data_as(_f64p))
    return xi
