"""ctypes binding of the C++ StereoVO class (visual_odometry_ros_b200/host/stereo_vo.h) -- the drop-in for
core/visual_odometry/stereo_vo/stereo_vo.h:233-249 -- for tests and bench.py.  No arithmetic happens here."""
import ctypes
import os

import numpy as np

from . import capi

HOST_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libvo_b200_host.so")
_host = None
vp = ctypes.c_void_p


class Parameters(ctypes.Structure):
    """Mirror of StereoVO::Parameters."""
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("K_l", ctypes.c_float * 4), ("K_r", ctypes.c_float * 4),
                ("T_lr", ctypes.c_float * 16), ("thres_error", ctypes.c_float), ("thres_bidirection", ctypes.c_float),
                ("thres_sampson", ctypes.c_float), ("window_size", ctypes.c_int), ("max_level", ctypes.c_int),
                ("n_bins_u", ctypes.c_int), ("n_bins_v", ctypes.c_int), ("thres_poseba_error", ctypes.c_float),
                ("thres_alive_ratio", ctypes.c_float), ("thres_trans", ctypes.c_float), ("thres_rotation_deg", ctypes.c_float),
                ("n_max_keyframes_in_window", ctypes.c_int), ("do_scale_refine", ctypes.c_int), ("det_edge", ctypes.c_int),
                ("det_min_score", ctypes.c_longlong), ("device", ctypes.c_int), ("do_undistortion", ctypes.c_int),
                ("D_l", ctypes.c_float * 5), ("D_r", ctypes.c_float * 5), ("collect_gate_counts", ctypes.c_int),
                ("detector", ctypes.c_int), ("thres_fastscore", ctypes.c_int), ("pose_strict", ctypes.c_int),
                ("scale_faithful_borders", ctypes.c_int)]


class FrameInfo(ctypes.Structure):
    _fields_ = [("frame", ctypes.c_int), ("keyframe", ctypes.c_int), ("n_in", ctypes.c_int), ("n_tracked", ctypes.c_int),
                ("n_detected", ctypes.c_int), ("n_new", ctypes.c_int), ("n_recon", ctypes.c_int), ("lba_points", ctypes.c_int),
                ("lba_obs", ctypes.c_int), ("lba_ok", ctypes.c_int), ("counts", ctypes.c_int * 5),
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
    H.vo_svo_create.argtypes = [ctypes.POINTER(Parameters), ctypes.POINTER(vp)]
    H.vo_svo_create_from_yaml.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
    H.vo_svo_destroy.argtypes = [vp]
    H.vo_svo_destroy.restype = None
    H.vo_svo_track.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_double]
    H.vo_svo_pose.argtypes = [vp, vp]
    H.vo_svo_frame_info.argtypes = [vp, ctypes.POINTER(FrameInfo)]
    H.vo_svo_tracks.argtypes = [vp, ctypes.c_int, vp, vp, vp]
    H.vo_svo_keyframe_poses.argtypes = [vp, ctypes.c_int, vp]
    H.vo_svo_stats_consistent.argtypes = [vp]
    H.vo_svo_launch_count.argtypes = [vp]
    H.vo_svo_launch_count.restype = ctypes.c_longlong
    H.vo_svo_last_error.restype = ctypes.c_char_p
    _host = H
    return H


def make_parameters(w, h, K_l, K_r, T_lr, *, window_size=21, max_level=3, thres_error=80.0, thres_bidirection=0.5, thres_sampson=60.0,
                    n_bins_u=64, n_bins_v=32, thres_poseba_error=3.0, thres_alive_ratio=0.6, thres_trans=10.0, thres_rotation_deg=15.0,
                    n_max_keyframes_in_window=9, do_scale_refine=True, det_edge=31, det_min_score=0, device=0, D_l=None, D_r=None, collect_gate_counts=False,
                    detector="harris", thres_fastscore=20, pose_strict=False, scale_faithful_borders=False):
    p = Parameters()
    p.width, p.height = int(w), int(h)
    p.K_l = (ctypes.c_float * 4)(*[float(v) for v in K_l])
    p.K_r = (ctypes.c_float * 4)(*[float(v) for v in K_r])
    p.T_lr = (ctypes.c_float * 16)(*[float(v) for v in np.asarray(T_lr, np.float32).ravel()])
    p.thres_error, p.thres_bidirection, p.thres_sampson = float(thres_error), float(thres_bidirection), float(thres_sampson)
    p.window_size, p.max_level, p.n_bins_u, p.n_bins_v = int(window_size), int(max_level), int(n_bins_u), int(n_bins_v)
    p.thres_poseba_error, p.thres_alive_ratio, p.thres_trans = float(thres_poseba_error), float(thres_alive_ratio), float(thres_trans)
    p.thres_rotation_deg, p.n_max_keyframes_in_window = float(thres_rotation_deg), int(n_max_keyframes_in_window)
    p.do_scale_refine, p.det_edge, p.det_min_score, p.device = int(bool(do_scale_refine)), int(det_edge), int(det_min_score), int(device)
    p.collect_gate_counts = int(bool(collect_gate_counts))
    p.detector, p.thres_fastscore = {"harris": 0, "orb": 1}[detector], int(thres_fastscore)
    p.pose_strict = int(bool(pose_strict))
    p.scale_faithful_borders = int(bool(scale_faithful_borders))
    if D_l is not None or D_r is not None:
        p.do_undistortion = 1
        p.D_l = (ctypes.c_float * 5)(*[float(v) for v in (D_l if D_l is not None else [0] * 5)])
        p.D_r = (ctypes.c_float * 5)(*[float(v) for v in (D_r if D_r is not None else [0] * 5)])
    return p


class StereoVO:
    """trackStereoImages(img_left, img_right, timestamp) like the reference class; poses are 4x4 row-major float32."""

    def __init__(self, params=None, yaml_path=None):
        self.H = host_lib()
        self.h = vp()
        if yaml_path is not None:
            rc = self.H.vo_svo_create_from_yaml(os.fsencode(yaml_path), ctypes.byref(self.h))
        else:
            rc = self.H.vo_svo_create(ctypes.byref(params), ctypes.byref(self.h))
        if rc != 0:
            raise capi.VoError(rc, self.H.vo_svo_last_error().decode())

    def close(self):
        if self.h:
            self.H.vo_svo_destroy(self.h)
            self.h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trackStereoImages(self, img_left, img_right, timestamp=0.0):
        for im in (img_left, img_right):
            assert im.dtype == np.uint8 and im.ndim == 2 and im.strides[1] == 1
        h, w = img_left.shape
        rc = self.H.vo_svo_track(self.h, img_left.ctypes.data_as(vp), img_right.ctypes.data_as(vp), w, h, img_left.strides[0], float(timestamp))
        if rc != 0:
            raise capi.VoError(rc, self.H.vo_svo_last_error().decode())

    def track_ptr(self, ptr_l, ptr_r, w, h, step, timestamp=0.0):
        rc = self.H.vo_svo_track(self.h, ptr_l, ptr_r, w, h, step, float(timestamp))
        if rc != 0:
            raise capi.VoError(rc, self.H.vo_svo_last_error().decode())

    def pose(self):
        T = np.zeros((4, 4), np.float32)
        self.H.vo_svo_pose(self.h, T.ctypes.data_as(vp))
        return T

    def frame_info(self):
        fi = FrameInfo()
        self.H.vo_svo_frame_info(self.h, ctypes.byref(fi))
        return {k: (list(getattr(fi, k)) if k == "counts" else getattr(fi, k)) for k, _ in FrameInfo._fields_}

    def tracks(self):
        n = self.H.vo_svo_tracks(self.h, 0, None, None, None)
        ids = np.zeros(n, np.int32)
        pl, pr = np.zeros((n, 2), np.float32), np.zeros((n, 2), np.float32)
        self.H.vo_svo_tracks(self.h, n, ids.ctypes.data_as(vp), pl.ctypes.data_as(vp), pr.ctypes.data_as(vp))
        return ids, pl, pr

    def keyframe_poses(self):
        n = self.H.vo_svo_keyframe_poses(self.h, 0, None)
        T = np.zeros((n, 4, 4), np.float32)
        self.H.vo_svo_keyframe_poses(self.h, n, T.ctypes.data_as(vp))
        return T

    def stats_consistent(self):
        return bool(self.H.vo_svo_stats_consistent(self.h))

    @property
    def launch_count(self):
        return int(self.H.vo_svo_launch_count(self.h))
