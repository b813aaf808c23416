"""ctypes binding of the C++ MonoVO class (visual_odometry_ros_b200/host/mono_vo.h) -- the drop-in for
core/visual_odometry/mono_vo/mono_vo.h:235-267 -- for tests and bench.py.  No arithmetic happens here."""
import ctypes
import os

import numpy as np

from . import capi
from .stereo_vo import HOST_LIB_PATH

_host = None
vp = ctypes.c_void_p


class Parameters(ctypes.Structure):
    """Mirror of MonoVO::Parameters."""
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("K", ctypes.c_float * 4), ("thres_error", ctypes.c_float),
                ("thres_bidirection", ctypes.c_float), ("thres_sampson", ctypes.c_float), ("window_size", ctypes.c_int),
                ("max_level", ctypes.c_int), ("thres_parallax_deg", ctypes.c_float), ("n_bins_u", ctypes.c_int), ("n_bins_v", ctypes.c_int),
                ("thres_5p_error", ctypes.c_float), ("thres_poseba_error", ctypes.c_float), ("thres_overlap_ratio", ctypes.c_float),
                ("thres_translation", ctypes.c_float), ("thres_rotation_deg", ctypes.c_float), ("n_max_keyframes_in_window", ctypes.c_int),
                ("do_scale_refine", ctypes.c_int), ("det_edge", ctypes.c_int), ("det_min_score", ctypes.c_longlong), ("device", ctypes.c_int),
                ("n_hypotheses", ctypes.c_int), ("seed", ctypes.c_uint), ("collect_gate_counts", ctypes.c_int),
                ("record_frame_mappoints", ctypes.c_int), ("detector", ctypes.c_int), ("thres_fastscore", ctypes.c_int),
                ("do_undistortion", ctypes.c_int), ("D", ctypes.c_float * 5), ("pose_strict", ctypes.c_int),
                ("scale_faithful_borders", ctypes.c_int)]


class FrameInfo(ctypes.Structure):
    _fields_ = [("frame", ctypes.c_int), ("keyframe", ctypes.c_int), ("n_in", ctypes.c_int), ("n_tracked", ctypes.c_int),
                ("n_detected", ctypes.c_int), ("n_new", ctypes.c_int), ("n_recon", ctypes.c_int), ("used_5point", ctypes.c_int),
                ("lba_points", ctypes.c_int), ("lba_obs", ctypes.c_int), ("lba_ok", ctypes.c_int), ("counts", ctypes.c_int * 5),
                ("ms_step", ctypes.c_float), ("ms_book", ctypes.c_float), ("ms_recon", ctypes.c_float), ("ms_lba_pack", ctypes.c_float),
                ("ms_lba_solve", ctypes.c_float), ("ms_stats", ctypes.c_float), ("ms_total", ctypes.c_float)]


def host_lib():
    global _host
    if _host is not None:
        return _host
    capi.lib()      # libvo_b200.so first (raises VoLibraryMissing if absent: there is no CPU fallback)
    if not os.path.exists(HOST_LIB_PATH):
        raise capi.VoLibraryMissing(f"{HOST_LIB_PATH} not found: run __graft_entry__.build()")
    H = ctypes.CDLL(HOST_LIB_PATH)
    H.vo_mvo_create.argtypes = [ctypes.POINTER(Parameters), ctypes.POINTER(vp)]
    H.vo_mvo_create_from_yaml.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
    H.vo_mvo_destroy.argtypes = [vp]
    H.vo_mvo_destroy.restype = None
    H.vo_mvo_track.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_double]
    H.vo_mvo_pose.argtypes = [vp, vp]
    H.vo_mvo_frame_pose.argtypes = [vp, ctypes.c_int, vp]
    H.vo_mvo_frame_info.argtypes = [vp, ctypes.POINTER(FrameInfo)]
    H.vo_mvo_tracks.argtypes = [vp, ctypes.c_int, vp, vp]
    H.vo_mvo_stats_consistent.argtypes = [vp]
    H.vo_mvo_launch_count.argtypes = [vp]
    H.vo_mvo_launch_count.restype = ctypes.c_longlong
    H.vo_mvo_last_error.restype = ctypes.c_char_p
    _host = H
    return H


def make_parameters(w, h, K, *, window_size=21, max_level=6, thres_error=60.0, thres_bidirection=0.5, thres_sampson=1000.0,
                    thres_parallax_deg=1.0, n_bins_u=30, n_bins_v=12, thres_5p_error=1.0, thres_poseba_error=5.0, thres_overlap_ratio=0.6,
                    thres_translation=4.0, thres_rotation_deg=10.0, n_max_keyframes_in_window=9, do_scale_refine=True, det_edge=31,
                    det_min_score=0, device=0, n_hypotheses=0, seed=0, collect_gate_counts=False, record_frame_mappoints=False,
                    detector="harris", thres_fastscore=20, D=None, pose_strict=False, scale_faithful_borders=False):
    """Defaults = config/mono/kitti_00.yaml.  D = (k1, k2, p1, p2, k3) switches the on-device undistortion on."""
    p = Parameters()
    p.width, p.height = int(w), int(h)
    p.K = (ctypes.c_float * 4)(*[float(v) for v in K])
    p.thres_error, p.thres_bidirection, p.thres_sampson = float(thres_error), float(thres_bidirection), float(thres_sampson)
    p.window_size, p.max_level, p.thres_parallax_deg = int(window_size), int(max_level), float(thres_parallax_deg)
    p.n_bins_u, p.n_bins_v = int(n_bins_u), int(n_bins_v)
    p.thres_5p_error, p.thres_poseba_error = float(thres_5p_error), float(thres_poseba_error)
    p.thres_overlap_ratio, p.thres_translation, p.thres_rotation_deg = float(thres_overlap_ratio), float(thres_translation), float(thres_rotation_deg)
    p.n_max_keyframes_in_window, p.do_scale_refine = int(n_max_keyframes_in_window), int(bool(do_scale_refine))
    p.det_edge, p.det_min_score, p.device = int(det_edge), int(det_min_score), int(device)
    p.n_hypotheses, p.seed = int(n_hypotheses), int(seed)
    p.collect_gate_counts, p.record_frame_mappoints = int(bool(collect_gate_counts)), int(bool(record_frame_mappoints))
    p.detector, p.thres_fastscore = {"harris": 0, "orb": 1}[detector], int(thres_fastscore)
    p.pose_strict = int(bool(pose_strict))
    p.scale_faithful_borders = int(bool(scale_faithful_borders))
    if D is not None:
        p.do_undistortion = 1
        p.D = (ctypes.c_float * 5)(*[float(v) for v in D])
    return p


class MonoVO:
    """trackImage(img, timestamp) like the reference class; poses are 4x4 row-major float32."""

    def __init__(self, params=None, yaml_path=None):
        self.H = host_lib()
        self.h = vp()
        if yaml_path is not None:
            rc = self.H.vo_mvo_create_from_yaml(os.fsencode(yaml_path), ctypes.byref(self.h))
        else:
            rc = self.H.vo_mvo_create(ctypes.byref(params), ctypes.byref(self.h))
        if rc != 0:
            raise capi.VoError(rc, self.H.vo_mvo_last_error().decode())

    def close(self):
        if self.h:
            self.H.vo_mvo_destroy(self.h)
            self.h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trackImage(self, img, timestamp=0.0):
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        h, w = img.shape
        rc = self.H.vo_mvo_track(self.h, img.ctypes.data_as(vp), w, h, img.strides[0], float(timestamp))
        if rc != 0:
            raise capi.VoError(rc, self.H.vo_mvo_last_error().decode())

    def track_ptr(self, ptr, w, h, step, timestamp=0.0):
        rc = self.H.vo_mvo_track(self.h, ptr, w, h, step, float(timestamp))
        if rc != 0:
            raise capi.VoError(rc, self.H.vo_mvo_last_error().decode())

    def pose(self):
        T = np.zeros((4, 4), np.float32)
        self.H.vo_mvo_pose(self.h, T.ctypes.data_as(vp))
        return T

    def frame_pose(self, frame_id):
        T = np.zeros((4, 4), np.float32)
        rc = self.H.vo_mvo_frame_pose(self.h, int(frame_id), T.ctypes.data_as(vp))
        if rc != 0:
            raise capi.VoError(rc, "no such frame")
        return T

    def frame_info(self):
        fi = FrameInfo()
        self.H.vo_mvo_frame_info(self.h, ctypes.byref(fi))
        return {k: (list(getattr(fi, k)) if k == "counts" else getattr(fi, k)) for k, _ in FrameInfo._fields_}

    def tracks(self):
        n = self.H.vo_mvo_tracks(self.h, 0, None, None)
        ids, pts = np.zeros(n, np.int32), np.zeros((n, 2), np.float32)
        self.H.vo_mvo_tracks(self.h, n, ids.ctypes.data_as(vp), pts.ctypes.data_as(vp))
        return ids, pts

    def stats_consistent(self):
        return bool(self.H.vo_mvo_stats_consistent(self.h))

    @property
    def launch_count(self):
        return int(self.H.vo_mvo_launch_count(self.h))
