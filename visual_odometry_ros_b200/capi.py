"""ctypes binding of ``libvo_b200.so`` (the C ABI in ``include/vo_b200.h``).

This is the only way Python reaches the CUDA path; there is no CPU fallback.  If the
shared library is missing, importing the symbols raises ``VoLibraryMissing`` loudly.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvo_b200.so")

VO_OK = 0
VO_ERR_INVALID_ARG = -1
VO_ERR_CUDA = -2
VO_ERR_SIZE_MISMATCH = -3
VO_ERR_NAN = -4
VO_ERR_MODE = -5
VO_ERR_NO_DEVICE = -6
VO_ERR_LARGE_UPDATE = -7
VO_KLT_USE_INITIAL_FLOW = 4
VO_POSE_FAST, VO_POSE_STRICT, VO_POSE_NO_EARLY_STOP = 0, 1, 2
VO_MAX_LEVELS = 8


class VoLibraryMissing(RuntimeError):
    pass


class VoError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"vo_b200 status {status}: {text}")
        self.status = status


_lib = None

c_int_p = ctypes.POINTER(ctypes.c_int)
c_f32_p = ctypes.POINTER(ctypes.c_float)
c_f64_p = ctypes.POINTER(ctypes.c_double)
c_u8_p = ctypes.POINTER(ctypes.c_uint8)
c_i16_p = ctypes.POINTER(ctypes.c_int16)
c_i64_p = ctypes.POINTER(ctypes.c_longlong)
vp = ctypes.c_void_p


class StereoStepParams(ctypes.Structure):
    _fields_ = [("window_size", ctypes.c_int), ("max_level", ctypes.c_int), ("thres_error", ctypes.c_float),
                ("thres_poseba_error", ctypes.c_float), ("K_l", ctypes.c_float * 4), ("K_r", ctypes.c_float * 4),
                ("T_lr", ctypes.c_float * 16), ("do_scale_refine", ctypes.c_int), ("sampson_y", ctypes.c_float)]


class StereoFrameParams(ctypes.Structure):
    _fields_ = [("track", StereoStepParams), ("thres_bidirection", ctypes.c_float), ("n_bins_u", ctypes.c_int),
                ("n_bins_v", ctypes.c_int), ("det_edge", ctypes.c_int), ("det_min_score", ctypes.c_longlong),
                ("new_depth_gate", ctypes.c_int)]


class StereoFrameResult(ctypes.Structure):
    _fields_ = [("T_wc", vp), ("dT_pc", vp), ("n_tracked", ctypes.c_int), ("index", vp), ("pts_l1", vp), ("pts_r1", vp),
                ("counts", vp), ("n_detected", ctypes.c_int), ("n_new", ctypes.c_int), ("new_l1", vp), ("new_r1", vp)]


class MonoFrameParams(ctypes.Structure):
    _fields_ = [("window_size", ctypes.c_int), ("max_level", ctypes.c_int), ("thres_error", ctypes.c_float),
                ("thres_bidirection", ctypes.c_float), ("thres_sampson", ctypes.c_float), ("thres_poseba_error", ctypes.c_float),
                ("K", ctypes.c_float * 4), ("use_bundled_only", ctypes.c_int), ("do_scale_refine", ctypes.c_int),
                ("n_bins_u", ctypes.c_int), ("n_bins_v", ctypes.c_int), ("det_edge", ctypes.c_int), ("det_min_score", ctypes.c_longlong),
                ("thres_5p", ctypes.c_float), ("n_hypotheses", ctypes.c_int), ("seed", ctypes.c_uint), ("init_mode", ctypes.c_int)]


class MonoFrameResult(ctypes.Structure):
    _fields_ = [("T_wc", vp), ("dT01", vp), ("dT10", vp), ("n_tracked", ctypes.c_int), ("index", vp), ("pts1", vp), ("counts", vp),
                ("n_detected", ctypes.c_int), ("n_new", ctypes.c_int), ("new_p1", vp), ("new_p0", vp),
                ("used_5point", ctypes.c_int), ("n_5p_ransac", ctypes.c_int)]


class LbaProblem(ctypes.Structure):
    _fields_ = [
        ("n_frames", ctypes.c_int), ("n_opt", ctypes.c_int), ("n_points", ctypes.c_int), ("n_obs", ctypes.c_int),
        ("poses", c_f64_p), ("opt_index", c_int_p), ("points", c_f64_p), ("obs_ptr", c_int_p),
        ("obs_frame", c_int_p), ("obs_right", c_u8_p), ("obs_px", c_f64_p),
        ("K_l", ctypes.c_double * 4), ("K_r", ctypes.c_double * 4), ("T_lr", ctypes.c_double * 16),
        ("is_stereo", ctypes.c_int), ("huber", ctypes.c_double), ("lambda_", ctypes.c_double),
        ("max_iter", ctypes.c_int),
    ]


def lib():
    """Load libvo_b200.so (built by ``__graft_entry__.build()`` / ``make -C csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VoLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    L.vo_status_string.restype = ctypes.c_char_p
    L.vo_last_error.restype = ctypes.c_char_p
    L.vo_last_error.argtypes = [vp]
    L.vo_build_info.restype = ctypes.c_char_p
    L.vo_ctx_launch_count.restype = ctypes.c_longlong
    L.vo_ctx_launch_count.argtypes = [vp]
    L.vo_ctx_create.argtypes = [ctypes.c_int] * 5 + [vp, ctypes.POINTER(vp)]
    L.vo_ctx_destroy.argtypes = [vp]
    L.vo_ctx_synchronize.argtypes = [vp]
    L.vo_upload_image.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t]
    L.vo_set_image_d.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t]
    L.vo_build_pyramids.argtypes = [vp, c_int_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.vo_read_pyramid_level.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, c_int_p, c_int_p]
    L.vo_klt_track.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                               ctypes.c_int, vp, vp, vp]
    L.vo_klt_track_batch_d.argtypes = [vp, ctypes.c_int, c_int_p, c_int_p, vp, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
    ft = [vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float]
    L.vo_ft_track.argtypes = ft + [vp, vp]
    L.vo_ft_track_with_prior.argtypes = ft + [vp, vp]
    L.vo_ft_track_bidirection.argtypes = ft + [ctypes.c_float, vp, vp]
    L.vo_ft_track_bidirection_with_prior.argtypes = ft + [ctypes.c_float, vp, vp]
    L.vo_invalidate_pyramids.argtypes = [vp, c_int_p, ctypes.c_int]
    L.vo_ft_track_batch.argtypes = [vp, ctypes.c_int, c_int_p, c_int_p, vp, vp, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_size_t, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                    ctypes.c_int, vp, vp]
    f32 = ctypes.c_float
    L.vo_pose_gn_mono.argtypes = [vp, vp, vp, ctypes.c_int, f32, f32, f32, f32, ctypes.c_int, ctypes.c_int, vp, vp, vp,
                                  c_int_p, c_int_p]
    L.vo_pose_gn_stereo.argtypes = [vp, vp, vp, vp, ctypes.c_int, vp, vp, vp, f32, vp, vp, c_int_p, c_int_p]
    L.vo_pose_gn_stereo_batch_d.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp, f32, vp, vp, vp, vp]
    L.vo_pose_gn_mono_ex.argtypes = L.vo_pose_gn_mono.argtypes + [ctypes.c_int, ctypes.c_int, vp]
    L.vo_pose_gn_stereo_ex.argtypes = L.vo_pose_gn_stereo.argtypes + [ctypes.c_int, ctypes.c_int, vp]
    L.vo_pose_gn_stereo_batch_ex_d.argtypes = L.vo_pose_gn_stereo_batch_d.argtypes + [ctypes.c_int, ctypes.c_int, vp]
    L.vo_dist_unique_id.argtypes = [vp]
    L.vo_dist_init.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp]
    L.vo_dist_finalize.argtypes = [vp]
    L.vo_set_pose_mode.argtypes = [vp, ctypes.c_int]
    L.vo_set_scale_mode.argtypes = [vp, ctypes.c_int]
    L.vo_get_scale_mode.argtypes = [vp]
    L.vo_get_pose_mode.argtypes = [vp]
    L.vo_triangulate_dlt.argtypes = [vp, vp, vp, ctypes.c_int, vp, vp, vp, vp, vp, vp]
    L.vo_triangulate_dlt_grouped.argtypes = [vp, vp, vp, ctypes.c_int, vp, ctypes.c_int, vp, vp, vp, vp, vp, vp]
    L.vo_depth_filter_normal.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, vp, vp]
    L.vo_depth_filter_student_t.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_int, vp, vp]
    L.vo_depth_filter_normal_d.argtypes = L.vo_depth_filter_normal.argtypes
    L.vo_depth_filter_student_t_d.argtypes = L.vo_depth_filter_student_t.argtypes
    L.vo_ft_calc_prior.argtypes = [vp, vp, vp, ctypes.c_int, vp, vp, vp]
    L.vo_compact.argtypes = [vp, vp, ctypes.c_int, vp, c_int_p]
    L.vo_ft_track_with_scale.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int, vp, vp]
    L.vo_lba_solve.argtypes = [vp, ctypes.POINTER(LbaProblem), vp, vp, vp, c_int_p]
    L.vo_stereo_track_step.argtypes = [vp, ctypes.POINTER(StereoStepParams), ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp, vp,
                                       c_int_p, vp, vp, vp, vp]
    L.vo_detect_bucketed.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_longlong, vp, ctypes.c_int, c_int_p]
    L.vo_stereo_frame_step.argtypes = [vp, ctypes.POINTER(StereoFrameParams), ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_int, vp, vp, vp, vp, vp, vp,
                                       ctypes.POINTER(StereoFrameResult)]
    L.vo_mono_frame_step.argtypes = [vp, ctypes.POINTER(MonoFrameParams), ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_size_t, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.POINTER(MonoFrameResult)]
    L.vo_rectify_init.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp, vp]
    L.vo_undistort_init.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_int]
    L.vo_read_rectify_maps.argtypes = [vp, ctypes.c_int, vp, vp]
    L.vo_upload_image_rectified.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t]
    L.vo_sampson_distance.argtypes = [vp, vp, vp, ctypes.c_int, vp, vp, vp, vp]
    L.vo_sampson_distance_F.argtypes = [vp, vp, vp, ctypes.c_int, vp, vp]
    L.vo_symmetric_epipolar_distance.argtypes = [vp, vp, vp, ctypes.c_int, vp, vp, vp, vp]
    L.vo_inliers_1point_histogram.argtypes = [vp, vp, vp, ctypes.c_int, vp, f32, vp, vp, vp, vp, vp]
    L.vo_set_detector.argtypes = [vp, ctypes.c_int, ctypes.c_int]
    L.vo_orb_detect.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, c_int_p]
    L.vo_orb_read_level.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, c_int_p, c_int_p]
    L.vo_pose_5point.argtypes = [vp, vp, vp, ctypes.c_int, vp, f32, ctypes.c_int, ctypes.c_uint, vp, vp, vp, vp, vp, vp]
    L.vo_five_point_minimal.argtypes = [vp, vp, ctypes.c_int, vp, vp]
    L.vo_stereo_reconstruct.argtypes = [vp, vp, vp, ctypes.c_int, vp, vp, vp, vp, vp, vp]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(vp) if a is not None else None


def check(ctx_handle, rc):
    if rc != VO_OK:
        L = lib()
        txt = L.vo_status_string(rc).decode()
        if ctx_handle:
            txt += ": " + L.vo_last_error(ctx_handle).decode()
        raise VoError(rc, txt)


def dist_unique_id():
    """128-byte NCCL unique id (rank 0 creates it, the host program ships it to the other ranks)."""
    buf = (ctypes.c_char * 128)()
    rc = lib().vo_dist_unique_id(buf)
    if rc != 0:
        raise VoError(rc, "vo_dist_unique_id failed (libnccl.so.2 not loadable?)")
    return bytes(buf)


class Context:
    """One GPU + one stream + ``n_slots`` device-resident image pyramids (``vo_ctx``)."""

    def __init__(self, device=0, max_w=1241, max_h=376, n_slots=4, max_feat=4096, stream=None):
        L = lib()
        h = vp()
        rc = L.vo_ctx_create(device, max_w, max_h, n_slots, max_feat, vp(stream) if stream else None,
                             ctypes.byref(h))
        if rc != VO_OK:
            raise VoError(rc, L.vo_status_string(rc).decode())
        self.h = h
        self.L = L
        self.n_slots = n_slots

    def close(self):
        if getattr(self, "h", None):
            self.L.vo_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(self.h, self.L.vo_ctx_synchronize(self.h))

    @property
    def launch_count(self):
        return int(self.L.vo_ctx_launch_count(self.h))

    # ---------------------------------------------------------------- images
    def upload_image(self, slot, img):
        img = np.asarray(img)
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        self._keep = img  # keep alive until the async copy is consumed
        check(self.h, self.L.vo_upload_image(self.h, slot, _ptr(img), img.shape[1], img.shape[0], img.strides[0]))

    def set_image_d(self, slot, dev_ptr, w, h, step):
        check(self.h, self.L.vo_set_image_d(self.h, slot, vp(dev_ptr), w, h, step))

    def build_pyramids(self, slots, n_levels, with_deriv=True):
        ids = np.ascontiguousarray(slots, np.int32)
        check(self.h, self.L.vo_build_pyramids(self.h, ids.ctypes.data_as(c_int_p), len(ids), n_levels,
                                               1 if with_deriv else 0))

    def invalidate_pyramids(self, slots):
        ids = np.ascontiguousarray(slots, np.int32)
        check(self.h, self.L.vo_invalidate_pyramids(self.h, ids.ctypes.data_as(c_int_p), len(ids)))

    def ft_track_batch(self, slots0, slots1, img_ptrs0, img_ptrs1, w, h, step, pts0, win, lvl, thres_err,
                       pts_track=None, mask=None, with_prior=False):
        """Batched FeatureTracker::track(WithPrior) with host buffers. img_ptrs*: lists of host addresses
        (ints; 0 keeps the slot's image). pts0: [n_pairs, n, 2] float32. Returns (pts_track, mask)."""
        s0 = np.ascontiguousarray(slots0, np.int32)
        s1 = np.ascontiguousarray(slots1, np.int32)
        npairs = len(s0)
        pts0 = np.ascontiguousarray(pts0, np.float32).reshape(npairs, -1, 2)
        n = pts0.shape[1]
        pt = (np.ascontiguousarray(pts_track, np.float32).reshape(npairs, n, 2) if pts_track is not None
              else np.zeros_like(pts0))
        m = np.ones((npairs, n), np.uint8) if mask is None else np.ascontiguousarray(mask, np.uint8).reshape(npairs, n)
        a0 = (vp * npairs)(*[vp(int(p)) if p else None for p in img_ptrs0]) if img_ptrs0 is not None else None
        a1 = (vp * npairs)(*[vp(int(p)) if p else None for p in img_ptrs1]) if img_ptrs1 is not None else None
        check(self.h, self.L.vo_ft_track_batch(self.h, npairs, s0.ctypes.data_as(c_int_p), s1.ctypes.data_as(c_int_p),
                                               a0, a1, w, h, step, _ptr(pts0), n, win, lvl, thres_err,
                                               1 if with_prior else 0, _ptr(pt), _ptr(m)))
        return pt, m

    def read_pyramid_level(self, slot, level, want_deriv=True):
        w, h = ctypes.c_int(), ctypes.c_int()
        check(self.h, self.L.vo_read_pyramid_level(self.h, slot, level, None, None, ctypes.byref(w), ctypes.byref(h)))
        img = np.empty((h.value, w.value), np.uint8)
        der = np.empty((h.value, w.value, 2), np.int16) if want_deriv else None
        check(self.h, self.L.vo_read_pyramid_level(self.h, slot, level, _ptr(img), _ptr(der), ctypes.byref(w),
                                                   ctypes.byref(h)))
        return img, der

    # ---------------------------------------------------------------- raw LK
    def klt_track(self, slot0, slot1, pts0, win, max_level, flags=0, prior=None):
        pts0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2)
        n = len(pts0)
        p1 = (np.ascontiguousarray(prior, np.float32).reshape(-1, 2).copy()
              if flags & VO_KLT_USE_INITIAL_FLOW else np.zeros_like(pts0))
        st = np.zeros(n, np.uint8)
        err = np.zeros(n, np.float32)
        check(self.h, self.L.vo_klt_track(self.h, slot0, slot1, _ptr(pts0), n, win, max_level, flags, _ptr(p1),
                                          _ptr(st), _ptr(err)))
        return p1, st, err

    def klt_track_batch_d(self, slots0, slots1, pts0_d, n, win, max_level, flags, pts1_d, status_d, err_d,
                          counters_d=None):
        s0 = np.ascontiguousarray(slots0, np.int32)
        s1 = np.ascontiguousarray(slots1, np.int32)
        check(self.h, self.L.vo_klt_track_batch_d(
            self.h, len(s0), s0.ctypes.data_as(c_int_p), s1.ctypes.data_as(c_int_p), vp(pts0_d), n, win, max_level,
            flags, vp(pts1_d), vp(status_d) if status_d else None, vp(err_d) if err_d else None,
            vp(counters_d) if counters_d else None))

    # ---------------------------------------------------------------- FeatureTracker methods
    def _ft(self, fn, slot0, slot1, pts0, win, lvl, thres_err, extra, pts_track, mask):
        pts0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2)
        n = len(pts0)
        pt = (np.ascontiguousarray(pts_track, np.float32).reshape(-1, 2).copy() if pts_track is not None
              else np.zeros_like(pts0))
        m = np.ones(n, np.uint8) if mask is None else np.ascontiguousarray(mask).astype(np.uint8).copy()
        if len(pt) != n or len(m) != n:
            raise VoError(VO_ERR_SIZE_MISMATCH, "pts_track.size() != pts0.size()")
        args = [self.h, slot0, slot1, _ptr(pts0), n, win, lvl, thres_err] + extra + [_ptr(pt), _ptr(m)]
        check(self.h, fn(*args))
        return pt, m.astype(bool)

    def ft_track(self, slot0, slot1, pts0, win, lvl, thres_err, mask=None):
        return self._ft(self.L.vo_ft_track, slot0, slot1, pts0, win, lvl, thres_err, [], None, mask)

    def ft_track_with_prior(self, slot0, slot1, pts0, prior, win, lvl, thres_err, mask=None):
        return self._ft(self.L.vo_ft_track_with_prior, slot0, slot1, pts0, win, lvl, thres_err, [], prior, mask)

    def ft_track_bidirection(self, slot0, slot1, pts0, win, lvl, thres_err, thres_bi, mask=None):
        return self._ft(self.L.vo_ft_track_bidirection, slot0, slot1, pts0, win, lvl, thres_err,
                        [ctypes.c_float(thres_bi)], None, mask)

    def ft_track_bidirection_with_prior(self, slot0, slot1, pts0, prior, win, lvl, thres_err, thres_bi, mask=None):
        return self._ft(self.L.vo_ft_track_bidirection_with_prior, slot0, slot1, pts0, win, lvl, thres_err,
                        [ctypes.c_float(thres_bi)], prior, mask)

    def set_scale_mode(self, faithful_borders):
        """trackWithScale samples outside the image: True = the reference's stale sample buffers reproduced, False = masked out."""
        check(self.h, self.L.vo_set_scale_mode(self.h, int(bool(faithful_borders))))

    # ---------------------------------------------------------------- pose-only Gauss-Newton
    def set_pose_mode(self, flags):
        """VO_POSE_FAST (0) / VO_POSE_STRICT (1): accumulation mode of every non-_ex pose-GN call and of the frame steps."""
        check(self.h, self.L.vo_set_pose_mode(self.h, int(flags)))

    def pose_gn_stereo(self, X, pts_l1, pts_r1, K_l, K_r, T_lr, thres, T01_init, flags=None, max_iter=0, want_trace=False):
        """MotionEstimator::poseOnlyBundleAdjustment_Stereo -> (success, T01, mask, iters[, trace]).
        flags None = the context's mode; else VO_POSE_STRICT | VO_POSE_NO_EARLY_STOP through vo_pose_gn_stereo_ex."""
        X = np.ascontiguousarray(X, np.float32).reshape(-1, 3)
        pl = np.ascontiguousarray(pts_l1, np.float32).reshape(-1, 2)
        pr = np.ascontiguousarray(pts_r1, np.float32).reshape(-1, 2)
        n = len(X)
        if len(pl) != n or len(pr) != n:
            # reference text: motion_estimator.cpp:873
            raise VoError(VO_ERR_SIZE_MISMATCH,
                          "In 'poseOnlyStereoBundleAdjustment()': X.size() != pts_l1.size() || X.size() != pts_r1.size().")
        Kl = np.ascontiguousarray(K_l, np.float32)
        Kr = np.ascontiguousarray(K_r, np.float32)
        Tlr = np.ascontiguousarray(T_lr, np.float32)
        T01 = np.ascontiguousarray(T01_init, np.float32).copy()
        mask = np.ones(n, np.uint8)
        ok, it = ctypes.c_int(0), ctypes.c_int(0)
        if flags is None and not want_trace and not max_iter:
            check(self.h, self.L.vo_pose_gn_stereo(self.h, _ptr(X), _ptr(pl), _ptr(pr), n, _ptr(Kl), _ptr(Kr), _ptr(Tlr),
                                                   thres, _ptr(T01), _ptr(mask), ctypes.byref(ok), ctypes.byref(it)))
            return bool(ok.value), T01, mask.astype(bool), it.value
        mi = max_iter if 0 < max_iter < 100 else 100
        trace = np.zeros((mi, 24), np.float32)
        fl = self.L.vo_get_pose_mode(self.h) if flags is None else int(flags)
        check(self.h, self.L.vo_pose_gn_stereo_ex(self.h, _ptr(X), _ptr(pl), _ptr(pr), n, _ptr(Kl), _ptr(Kr), _ptr(Tlr),
                                                  thres, _ptr(T01), _ptr(mask), ctypes.byref(ok), ctypes.byref(it), fl,
                                                  int(max_iter), _ptr(trace) if want_trace else None))
        out = (bool(ok.value), T01, mask.astype(bool), it.value)
        return out + (trace[:it.value],) if want_trace else out

    def pose_gn_mono(self, X, pts1, K, thres, R01_init, t01_init, standalone_variant=0, flags=None, max_iter=0,
                     want_trace=False):
        """MotionEstimator::poseOnlyBundleAdjustment -> (success, R01, t01, mask, iters[, trace])."""
        X = np.ascontiguousarray(X, np.float32).reshape(-1, 3)
        p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
        n = len(X)
        if len(p1) != n:
            raise VoError(VO_ERR_SIZE_MISMATCH, "In 'poseOnlyBundleAdjustment()': X.size() != pts1.size().")
        R = np.ascontiguousarray(R01_init, np.float32).copy()
        t = np.ascontiguousarray(t01_init, np.float32).copy()
        mask = np.ones(n, np.uint8)
        ok, it = ctypes.c_int(0), ctypes.c_int(0)
        if flags is None and not want_trace and not max_iter:
            check(self.h, self.L.vo_pose_gn_mono(self.h, _ptr(X), _ptr(p1), n, float(K[0]), float(K[1]), float(K[2]),
                                                 float(K[3]), int(thres), int(standalone_variant), _ptr(R), _ptr(t),
                                                 _ptr(mask), ctypes.byref(ok), ctypes.byref(it)))
            return bool(ok.value), R, t, mask.astype(bool), it.value
        mi = max_iter if 0 < max_iter < 100 else 100
        trace = np.zeros((mi, 24), np.float32)
        fl = self.L.vo_get_pose_mode(self.h) if flags is None else int(flags)
        check(self.h, self.L.vo_pose_gn_mono_ex(self.h, _ptr(X), _ptr(p1), n, float(K[0]), float(K[1]), float(K[2]),
                                                float(K[3]), int(thres), int(standalone_variant), _ptr(R), _ptr(t),
                                                _ptr(mask), ctypes.byref(ok), ctypes.byref(it), fl, int(max_iter),
                                                _ptr(trace) if want_trace else None))
        out = (bool(ok.value), R, t, mask.astype(bool), it.value)
        return out + (trace[:it.value],) if want_trace else out

    def pose_gn_stereo_batch_d(self, n_prob, offsets_d, X_d, pl_d, pr_d, K_l, K_r, T_lr, thres, T01_d, mask_d,
                               success_d=None, iters_d=None, flags=None, max_iter=0, trace_d=None):
        Kl = np.ascontiguousarray(K_l, np.float32)
        Kr = np.ascontiguousarray(K_r, np.float32)
        Tlr = np.ascontiguousarray(T_lr, np.float32)
        fl = self.L.vo_get_pose_mode(self.h) if flags is None else int(flags)
        check(self.h, self.L.vo_pose_gn_stereo_batch_ex_d(
            self.h, n_prob, vp(offsets_d), vp(X_d), vp(pl_d), vp(pr_d), _ptr(Kl), _ptr(Kr), _ptr(Tlr), thres,
            vp(T01_d), vp(mask_d), vp(success_d) if success_d else None, vp(iters_d) if iters_d else None, fl,
            int(max_iter), vp(trace_d) if trace_d else None))

    # ---------------------------------------------------------------- elementwise rows
    def triangulate_dlt(self, pts0, pts1, R10, t10, K0, K1=None):
        """mapping::triangulateDLT -> (X0, X1)."""
        p0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2)
        p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
        n = len(p0)
        if len(p1) != n:
            raise VoError(VO_ERR_SIZE_MISMATCH, "pts0.size() != pts1.size()")   # triangulate_3d.cpp:10
        R = np.ascontiguousarray(R10, np.float32)
        t = np.ascontiguousarray(t10, np.float32)
        K0 = np.ascontiguousarray(K0, np.float32)
        K1 = K0 if K1 is None else np.ascontiguousarray(K1, np.float32)
        X0 = np.zeros((n, 3), np.float32)
        X1 = np.zeros((n, 3), np.float32)
        check(self.h, self.L.vo_triangulate_dlt(self.h, _ptr(p0), _ptr(p1), n, _ptr(R), _ptr(t), _ptr(K0), _ptr(K1),
                                                _ptr(X0), _ptr(X1)))
        return X0, X1

    def triangulate_dlt_grouped(self, pts0, pts1, group, R10s, t10s, K0, K1=None):
        """triangulateDLT with one relative pose per group of points -> (X0, X1)."""
        p0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2)
        p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
        g = np.ascontiguousarray(group, np.int32)
        n = len(p0)
        Rs = np.ascontiguousarray(R10s, np.float32).reshape(-1, 9)
        ts = np.ascontiguousarray(t10s, np.float32).reshape(-1, 3)
        K0 = np.ascontiguousarray(K0, np.float32)
        K1 = K0 if K1 is None else np.ascontiguousarray(K1, np.float32)
        X0 = np.zeros((n, 3), np.float32)
        X1 = np.zeros((n, 3), np.float32)
        check(self.h, self.L.vo_triangulate_dlt_grouped(self.h, _ptr(p0), _ptr(p1), n, _ptr(g), len(Rs), _ptr(Rs), _ptr(ts), _ptr(K0),
                                                        _ptr(K1), _ptr(X0), _ptr(X1)))
        return X0, X1

    def epipolar_distance(self, pts0, pts1, K4=None, R10=None, t10=None, F10=None, symmetric=False):
        """calcSampsonDistance / calcSymmetricEpipolarDistance -> dist [n] (float32)."""
        p0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2)
        p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
        n = len(p0)
        if len(p1) != n:
            raise VoError(VO_ERR_SIZE_MISMATCH, "Error in 'fineInliers1PointHistogram()': pts0.size() != pts1.size()")   # motion_estimator.cpp:543
        d = np.zeros(max(n, 1), np.float32)
        if F10 is not None:
            F = np.ascontiguousarray(F10, np.float32)
            check(self.h, self.L.vo_sampson_distance_F(self.h, _ptr(p0), _ptr(p1), n, _ptr(F), _ptr(d)))
        else:
            K, R, t = (np.ascontiguousarray(v, np.float32) for v in (K4, R10, t10))
            fn = self.L.vo_symmetric_epipolar_distance if symmetric else self.L.vo_sampson_distance
            check(self.h, fn(self.h, _ptr(p0), _ptr(p1), n, _ptr(K), _ptr(R), _ptr(t), _ptr(d)))
        return d[:n]

    def inliers_1point_histogram(self, pts0, pts1, K4, thres_1p):
        """findInliers1PointHistogram -> dict(theta_opt, mask, R10, t10, theta)."""
        p0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2)
        p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
        n = len(p0)
        if len(p1) != n:
            raise VoError(VO_ERR_SIZE_MISMATCH, "Error in 'fineInliers1PointHistogram()': pts0.size() != pts1.size()")   # :477
        K = np.ascontiguousarray(K4, np.float32)
        mask, th = np.zeros(max(n, 1), np.uint8), np.zeros(max(n, 1), np.float32)
        th_opt, R, t = np.zeros(1, np.float32), np.zeros((3, 3), np.float32), np.zeros(3, np.float32)
        check(self.h, self.L.vo_inliers_1point_histogram(self.h, _ptr(p0), _ptr(p1), n, _ptr(K), float(thres_1p), _ptr(mask), _ptr(th_opt),
                                                         _ptr(R), _ptr(t), _ptr(th)))
        return dict(theta_opt=float(th_opt[0]), mask=mask[:n].astype(bool), R10=R, t10=t, theta=th[:n])

    def set_detector(self, kind, fast_threshold=20):
        """kind: "harris" (K-det, default) or "orb" (cv::ORB::detect restated, the reference's extractor)."""
        check(self.h, self.L.vo_set_detector(self.h, {"harris": 0, "orb": 1}[kind], int(fast_threshold)))

    def orb_detect(self, slot, fast_threshold, edge=31, max_keypoints=20000):
        """cv::ORB::detect on the slot's image -> (pts [n,2], response [n], octave [n]), unordered."""
        pts, resp = np.zeros((max_keypoints, 2), np.float32), np.zeros(max_keypoints, np.float32)
        octv, n = np.zeros(max_keypoints, np.int32), ctypes.c_int(0)
        check(self.h, self.L.vo_orb_detect(self.h, int(slot), int(fast_threshold), int(edge), int(max_keypoints), _ptr(pts), _ptr(resp),
                                           _ptr(octv), ctypes.byref(n)))
        return pts[:n.value].copy(), resp[:n.value].copy(), octv[:n.value].copy()

    def orb_read_level(self, level, plane=0):
        w, h = ctypes.c_int(0), ctypes.c_int(0)
        check(self.h, self.L.vo_orb_read_level(self.h, int(level), int(plane), None, ctypes.byref(w), ctypes.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        check(self.h, self.L.vo_orb_read_level(self.h, int(level), int(plane), _ptr(out), ctypes.byref(w), ctypes.byref(h)))
        return out

    def pose_5point(self, pts0, pts1, K4, thres_5p, n_hypotheses=0, seed=0):
        """MotionEstimator::calcPose5PointsAlgorithm -> dict(R10, t10, X0, mask, E, n_ransac, n_cheirality)."""
        p0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2)
        p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
        n = len(p0)
        if len(p1) != n:
            raise VoError(VO_ERR_SIZE_MISMATCH, "calcPose5PointsAlgorithm(): pts0.size() != pts1.size()")   # motion_estimator.cpp:26
        K = np.ascontiguousarray(K4, np.float32)
        R, t, E = np.zeros((3, 3), np.float32), np.zeros(3, np.float32), np.zeros((3, 3), np.float32)
        X0, mask, info = np.zeros((max(n, 1), 3), np.float32), np.zeros(max(n, 1), np.uint8), np.zeros(3, np.int32)
        check(self.h, self.L.vo_pose_5point(self.h, _ptr(p0), _ptr(p1), n, _ptr(K), float(thres_5p), int(n_hypotheses), int(seed),
                                            _ptr(R), _ptr(t), _ptr(X0), _ptr(mask), _ptr(E), _ptr(info)))
        return dict(R10=R, t10=t, X0=X0[:n], mask=mask[:n].astype(bool), E=E, n_ransac=int(info[0]), n_cheirality=int(info[1]))

    def five_point_minimal(self, q):
        """Minimal solver alone: q (S, 5, 4) normalised (x0, y0, x1, y1) -> list of (k_s, 3, 3) solution arrays."""
        q = np.ascontiguousarray(q, np.float64).reshape(-1, 5, 4)
        S = len(q)
        E, ns = np.zeros((S, 10, 3, 3)), np.zeros(S, np.int32)
        check(self.h, self.L.vo_five_point_minimal(self.h, _ptr(q), S, _ptr(E), _ptr(ns)))
        return [E[s, :ns[s]] for s in range(S)]

    def depth_filter_normal(self, x_prev, cov_prev, x_curr, cov_curr):
        a = [np.ascontiguousarray(v, np.float64) for v in (x_prev, cov_prev, x_curr, cov_curr)]
        n = len(a[0])
        x, c = np.zeros(n), np.zeros(n)
        check(self.h, self.L.vo_depth_filter_normal(self.h, *[_ptr(v) for v in a], n, _ptr(x), _ptr(c)))
        return x, c

    def depth_filter_student_t(self, x_prev, cov_prev, a, b, x_min, x_max, x_curr, cov_curr):
        xp, cp, xc, cc = [np.ascontiguousarray(v, np.float64) for v in (x_prev, cov_prev, x_curr, cov_curr)]
        a, b, lo, hi = [np.ascontiguousarray(v, np.float64).copy() for v in (a, b, x_min, x_max)]
        n = len(xp)
        x, c = np.zeros(n), np.zeros(n)
        check(self.h, self.L.vo_depth_filter_student_t(self.h, _ptr(xp), _ptr(cp), _ptr(a), _ptr(b), _ptr(lo), _ptr(hi),
                                                       _ptr(xc), _ptr(cc), n, _ptr(x), _ptr(c)))
        return x, c, a, b, lo, hi

    def calc_prior(self, pts0, Xw, Tw1, K4):
        p0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2)
        X = np.ascontiguousarray(Xw, np.float32).reshape(-1, 3)
        n = len(X)
        T = np.ascontiguousarray(Tw1, np.float32)
        K = np.ascontiguousarray(K4, np.float32)
        out = p0.copy()
        check(self.h, self.L.vo_ft_calc_prior(self.h, _ptr(p0), _ptr(X), n, _ptr(T), _ptr(K), _ptr(out)))
        return out

    def compact(self, mask):
        m = np.ascontiguousarray(mask).astype(np.uint8)
        idx = np.zeros(max(len(m), 1), np.int32)
        k = ctypes.c_int(0)
        check(self.h, self.L.vo_compact(self.h, _ptr(m), len(m), _ptr(idx), ctypes.byref(k)))
        return idx[:k.value]

    def ft_track_with_scale(self, slot0, slot1, pts0, scale_est, pts_track, mask=None):
        """FeatureTracker::trackWithScale -> (pts_track, mask)."""
        p0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2)
        n = len(p0)
        pt = np.ascontiguousarray(pts_track, np.float32).reshape(-1, 2).copy()
        if len(pt) != n:
            raise VoError(VO_ERR_SIZE_MISMATCH, "pts_track.size() != pts0.size()")   # feature_tracker.cpp:283
        sc = np.ascontiguousarray(scale_est, np.float32)
        m = np.ones(n, np.uint8) if mask is None else np.ascontiguousarray(mask).astype(np.uint8).copy()
        check(self.h, self.L.vo_ft_track_with_scale(self.h, slot0, slot1, _ptr(p0), _ptr(sc), n, _ptr(pt), _ptr(m)))
        return pt, m.astype(bool)

    def depth_filter_normal_d(self, x_prev_d, cov_prev_d, x_curr_d, cov_curr_d, n, x_upd_d, cov_upd_d):
        """Device-resident DepthFilter::updateNormalDistribution (device pointers as ints; asynchronous)."""
        check(self.h, self.L.vo_depth_filter_normal_d(self.h, vp(x_prev_d), vp(cov_prev_d), vp(x_curr_d), vp(cov_curr_d), int(n),
                                                      vp(x_upd_d), vp(cov_upd_d)))

    def depth_filter_student_t_d(self, x_prev_d, cov_prev_d, a_d, b_d, xmin_d, xmax_d, x_curr_d, cov_curr_d, n, x_upd_d, cov_upd_d):
        check(self.h, self.L.vo_depth_filter_student_t_d(self.h, vp(x_prev_d), vp(cov_prev_d), vp(a_d), vp(b_d), vp(xmin_d), vp(xmax_d),
                                                         vp(x_curr_d), vp(cov_curr_d), int(n), vp(x_upd_d), vp(cov_upd_d)))

    # ---------------------------------------------------------------- local bundle adjustment
    def lba_solve(self, p, dist=False):
        """SparseBundleAdjustmentSolver::solveForFiniteIterations on a flat problem dict
        (layout of synth.lba_problem / vo_lba_problem). Returns (poses, points, avg_err, success)."""
        keep = {}
        for k, dt in (("poses", np.float64), ("opt_index", np.int32), ("points", np.float64), ("obs_ptr", np.int32),
                      ("obs_frame", np.int32), ("obs_right", np.uint8), ("obs_px", np.float64)):
            keep[k] = np.ascontiguousarray(p[k], dt)
        s = LbaProblem()
        s.n_frames, s.n_opt, s.n_points, s.n_obs = int(p["n_frames"]), int(p["n_opt"]), int(p["n_points"]), int(p["n_obs"])
        s.poses = keep["poses"].ctypes.data_as(c_f64_p)
        s.opt_index = keep["opt_index"].ctypes.data_as(c_int_p)
        s.points = keep["points"].ctypes.data_as(c_f64_p)
        s.obs_ptr = keep["obs_ptr"].ctypes.data_as(c_int_p)
        s.obs_frame = keep["obs_frame"].ctypes.data_as(c_int_p)
        s.obs_right = keep["obs_right"].ctypes.data_as(c_u8_p)
        s.obs_px = keep["obs_px"].ctypes.data_as(c_f64_p)
        s.K_l = (ctypes.c_double * 4)(*[float(v) for v in p["K_l"]])
        s.K_r = (ctypes.c_double * 4)(*[float(v) for v in p["K_r"]])
        s.T_lr = (ctypes.c_double * 16)(*[float(v) for v in np.asarray(p["T_lr"], np.float64).ravel()])
        s.is_stereo = int(p["is_stereo"])
        s.huber = float(p["huber"])
        s.lambda_ = float(p["lam"])
        s.max_iter = int(p["max_iter"])
        poses = np.zeros((s.n_frames, 4, 4))
        points = np.zeros((max(s.n_points, 1), 3))
        avg = np.zeros(s.max_iter)
        ok = ctypes.c_int(0)
        fn = self.L.vo_lba_solve_dist if dist else self.L.vo_lba_solve
        check(self.h, fn(self.h, ctypes.byref(s), _ptr(poses), _ptr(points), _ptr(avg), ctypes.byref(ok)))
        return poses, points[:s.n_points], avg, bool(ok.value)

    # ---------------------------------------------------------------- landmark-sharded local BA over NCCL
    def dist_init(self, rank, world, unique_id):
        """Collective: ncclCommInitRank on this context's device. unique_id: the 128 bytes rank 0 got from dist_unique_id()."""
        buf = (ctypes.c_char * 128).from_buffer_copy(bytes(unique_id))
        check(self.h, self.L.vo_dist_init(self.h, int(rank), int(world), buf))

    def dist_finalize(self):
        check(self.h, self.L.vo_dist_finalize(self.h))

    def lba_solve_dist(self, p_local):
        """Collective: p_local = all frames / poses + this rank's landmarks (sharding.split_lba_problem).
        Returns (poses [all, identical on every rank], points [this rank's], avg_err [global], success)."""
        return self.lba_solve(p_local, dist=True)

    # ---------------------------------------------------------------- stereo tracking step (S1)
    def stereo_track_step(self, slot_l0, slot_l1, slot_r1, img_l1, img_r1, pts_l0, pts_r0, Xw, tri, T_wp, dT_pc_prev,
                          K_l, K_r, T_lr, win, max_level, thres_err, thres_poseba, do_scale_refine=True, sampson_y=660.0,
                          want_counts=True):
        """Steady-state part of StereoVO::trackStereoImages (stereo_vo.cpp:475-670), device resident."""
        prm = StereoStepParams()
        prm.window_size, prm.max_level, prm.thres_error, prm.thres_poseba_error = int(win), int(max_level), float(thres_err), float(thres_poseba)
        prm.K_l = (ctypes.c_float * 4)(*[float(v) for v in K_l])
        prm.K_r = (ctypes.c_float * 4)(*[float(v) for v in K_r])
        prm.T_lr = (ctypes.c_float * 16)(*[float(v) for v in np.asarray(T_lr, np.float32).ravel()])
        prm.do_scale_refine = 1 if do_scale_refine else 0
        prm.sampson_y = float(sampson_y)
        l0 = np.ascontiguousarray(pts_l0, np.float32).reshape(-1, 2)
        r0 = np.ascontiguousarray(pts_r0, np.float32).reshape(-1, 2)
        X = np.ascontiguousarray(Xw, np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(tri).astype(np.uint8)
        n = len(l0)
        Twp = np.ascontiguousarray(T_wp, np.float32)
        dTp = np.ascontiguousarray(dT_pc_prev, np.float32)
        T_wc, dT = np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32)
        idx = np.zeros(max(n, 1), np.int32)
        o_l1, o_r1 = np.zeros((max(n, 1), 2), np.float32), np.zeros((max(n, 1), 2), np.float32)
        counts = np.zeros(5, np.int32)
        n_out = ctypes.c_int(0)
        w = h = step = 0
        for im in (img_l1, img_r1):
            if im is not None:
                assert im.dtype == np.uint8 and im.ndim == 2 and im.strides[1] == 1
                h, w, step = im.shape[0], im.shape[1], im.strides[0]
        if w == 0:
            raise ValueError("pass the new images (use upload_image + the raw C ABI to keep them resident)")
        check(self.h, self.L.vo_stereo_track_step(
            self.h, ctypes.byref(prm), slot_l0, slot_l1, slot_r1, _ptr(img_l1), _ptr(img_r1), w, h, step, n, _ptr(l0), _ptr(r0),
            _ptr(X), _ptr(t), _ptr(Twp), _ptr(dTp), _ptr(T_wc), _ptr(dT), ctypes.byref(n_out), _ptr(idx), _ptr(o_l1), _ptr(o_r1),
            _ptr(counts) if want_counts else None))
        k = n_out.value
        return dict(T_wc=T_wc, dT_pc=dT, index=idx[:k].copy(), pts_l1=o_l1[:k].copy(), pts_r1=o_r1[:k].copy(),
                    counts=[int(c) for c in counts])

    @staticmethod
    def _step_params(K_l, K_r, T_lr, win, max_level, thres_err, thres_poseba, do_scale_refine, sampson_y):
        prm = StereoStepParams()
        prm.window_size, prm.max_level, prm.thres_error, prm.thres_poseba_error = int(win), int(max_level), float(thres_err), float(thres_poseba)
        prm.K_l = (ctypes.c_float * 4)(*[float(v) for v in K_l])
        prm.K_r = (ctypes.c_float * 4)(*[float(v) for v in K_r])
        prm.T_lr = (ctypes.c_float * 16)(*[float(v) for v in np.asarray(T_lr, np.float32).ravel()])
        prm.do_scale_refine = 1 if do_scale_refine else 0
        prm.sampson_y = float(sampson_y)
        return prm

    def detect_bucketed(self, slot, pts_occupied, n_bins_u, n_bins_v, edge=31, min_score=0):
        """FeatureExtractor::updateWeightBin + extractORBwithBinning_fast with the K-det response."""
        occ = np.ascontiguousarray(pts_occupied, np.float32).reshape(-1, 2)
        cap = n_bins_u * n_bins_v
        out = np.zeros((cap, 2), np.float32)
        n = ctypes.c_int(0)
        check(self.h, self.L.vo_detect_bucketed(self.h, slot, _ptr(occ) if len(occ) else None, len(occ), n_bins_u, n_bins_v, edge,
                                                int(min_score), _ptr(out), cap, ctypes.byref(n)))
        return out[:n.value].copy()

    def stereo_reconstruct(self, pts_l, pts_r, K_l, K_r, T_lr, T_wc):
        """stereo_vo.cpp:767-797: DLT + reprojection / depth gates + X_w = T_wc X_l. Returns (Xw, ok)."""
        pl = np.ascontiguousarray(pts_l, np.float32).reshape(-1, 2)
        pr = np.ascontiguousarray(pts_r, np.float32).reshape(-1, 2)
        n = len(pl)
        Xw = np.zeros((n, 3), np.float32)
        ok = np.zeros(n, np.uint8)
        Kl, Kr = np.ascontiguousarray(K_l, np.float32), np.ascontiguousarray(K_r, np.float32)
        Tlr, Twc = np.ascontiguousarray(T_lr, np.float32), np.ascontiguousarray(T_wc, np.float32)
        check(self.h, self.L.vo_stereo_reconstruct(self.h, _ptr(pl), _ptr(pr), n, _ptr(Kl), _ptr(Kr), _ptr(Tlr), _ptr(Twc), _ptr(Xw), _ptr(ok)))
        return Xw, ok.astype(bool)

    def stereo_frame_step(self, slot_l0, slot_l1, slot_r1, img_l1, img_r1, pts_l0, pts_r0, Xw, tri, T_wp, dT_pc_prev,
                          K_l, K_r, T_lr, win, max_level, thres_err, thres_poseba, thres_bi, n_bins_u, n_bins_v, det_edge=31,
                          det_min_score=0, new_depth_gate=True, do_scale_refine=True, sampson_y=660.0, want_counts=True):
        """Tracking step + new-feature stage of StereoVO::trackStereoImages (stereo_vo.cpp:475-740), one synchronisation."""
        fp = StereoFrameParams()
        fp.track = self._step_params(K_l, K_r, T_lr, win, max_level, thres_err, thres_poseba, do_scale_refine, sampson_y)
        fp.thres_bidirection, fp.n_bins_u, fp.n_bins_v, fp.det_edge = float(thres_bi), int(n_bins_u), int(n_bins_v), int(det_edge)
        fp.det_min_score, fp.new_depth_gate = int(det_min_score), 1 if new_depth_gate else 0
        l0 = np.ascontiguousarray(pts_l0, np.float32).reshape(-1, 2)
        r0 = np.ascontiguousarray(pts_r0, np.float32).reshape(-1, 2)
        X = np.ascontiguousarray(Xw, np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(tri).astype(np.uint8)
        n = len(l0)
        Twp = np.ascontiguousarray(T_wp if T_wp is not None else np.eye(4), np.float32)
        dTp = np.ascontiguousarray(dT_pc_prev if dT_pc_prev is not None else np.eye(4), np.float32)
        T_wc, dT = np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32)
        idx = np.zeros(max(n, 1), np.int32)
        o_l1, o_r1 = np.zeros((max(n, 1), 2), np.float32), np.zeros((max(n, 1), 2), np.float32)
        nbins = max(1, n_bins_u * n_bins_v)
        n_l1, n_r1 = np.zeros((nbins, 2), np.float32), np.zeros((nbins, 2), np.float32)
        counts = np.zeros(5, np.int32)
        res = StereoFrameResult()
        res.T_wc, res.dT_pc, res.index, res.pts_l1, res.pts_r1 = _ptr(T_wc), _ptr(dT), _ptr(idx), _ptr(o_l1), _ptr(o_r1)
        res.counts = _ptr(counts) if want_counts else None
        res.new_l1, res.new_r1 = _ptr(n_l1), _ptr(n_r1)
        w = h = step = 0
        for im in (img_l1, img_r1):
            if im is not None:
                assert im.dtype == np.uint8 and im.ndim == 2 and im.strides[1] == 1
                h, w, step = im.shape[0], im.shape[1], im.strides[0]
        if w == 0:
            raise ValueError("pass the new images")
        check(self.h, self.L.vo_stereo_frame_step(
            self.h, ctypes.byref(fp), slot_l0, slot_l1, slot_r1, _ptr(img_l1), _ptr(img_r1), w, h, step, n, _ptr(l0), _ptr(r0),
            _ptr(X), _ptr(t), _ptr(Twp), _ptr(dTp), ctypes.byref(res)))
        k, m = res.n_tracked, res.n_new
        return dict(T_wc=T_wc, dT_pc=dT, index=idx[:k].copy(), pts_l1=o_l1[:k].copy(), pts_r1=o_r1[:k].copy(),
                    counts=[int(c) for c in counts], n_detected=res.n_detected, new_l1=n_l1[:m].copy(), new_r1=n_r1[:m].copy())

    def mono_frame_step(self, slot_0, slot_1, img_1, pts0, Xw, triangulated, bundled, T_wc_prev, dT01_prior, K, win, max_level,
                        thres_err, thres_bi, thres_sampson, thres_poseba, use_bundled_only, n_bins_u=0, n_bins_v=0, det_edge=31,
                        det_min_score=0, do_scale_refine=True, want_counts=True, thres_5p=0.0, n_hypotheses=0, seed=0, init_mode=False):
        """Steady-state branch of MonoVO::trackImage (mono_vo.cpp:724-992), one synchronisation; init_mode: the second
        image of a sequence (:562-659; Xw / flags / dT01_prior may be None)."""
        prm = MonoFrameParams()
        prm.thres_5p, prm.n_hypotheses, prm.seed, prm.init_mode = float(thres_5p), int(n_hypotheses), int(seed), int(bool(init_mode))
        prm.window_size, prm.max_level, prm.thres_error = int(win), int(max_level), float(thres_err)
        prm.thres_bidirection, prm.thres_sampson, prm.thres_poseba_error = float(thres_bi), float(thres_sampson), float(thres_poseba)
        prm.K = (ctypes.c_float * 4)(*[float(v) for v in K])
        prm.use_bundled_only, prm.do_scale_refine = int(bool(use_bundled_only)), int(bool(do_scale_refine))
        prm.n_bins_u, prm.n_bins_v, prm.det_edge, prm.det_min_score = int(n_bins_u), int(n_bins_v), int(det_edge), int(det_min_score)
        p0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2)
        n = len(p0)
        if init_mode:
            X, fl, dTp = np.zeros((max(n, 1), 3), np.float32), np.zeros(max(n, 1), np.uint8), np.eye(4, dtype=np.float32)
        else:
            X = np.ascontiguousarray(Xw, np.float32).reshape(-1, 3)
            fl = (np.asarray(triangulated).astype(np.uint8) | (np.asarray(bundled).astype(np.uint8) << 1)).astype(np.uint8)
            dTp = np.ascontiguousarray(dT01_prior, np.float32)
        Twp = np.ascontiguousarray(T_wc_prev, np.float32)
        T_wc, dT01, dT10 = (np.zeros((4, 4), np.float32) for _ in range(3))
        idx = np.zeros(max(n, 1), np.int32)
        o1 = np.zeros((max(n, 1), 2), np.float32)
        nbins = max(1, n_bins_u * n_bins_v)
        n1, n0 = np.zeros((nbins, 2), np.float32), np.zeros((nbins, 2), np.float32)
        counts = np.zeros(5, np.int32)
        res = MonoFrameResult()
        res.T_wc, res.dT01, res.dT10, res.index, res.pts1 = _ptr(T_wc), _ptr(dT01), _ptr(dT10), _ptr(idx), _ptr(o1)
        res.counts = _ptr(counts) if want_counts else None
        res.new_p1, res.new_p0 = _ptr(n1), _ptr(n0)
        assert img_1.dtype == np.uint8 and img_1.ndim == 2 and img_1.strides[1] == 1
        h, w, step = img_1.shape[0], img_1.shape[1], img_1.strides[0]
        check(self.h, self.L.vo_mono_frame_step(self.h, ctypes.byref(prm), slot_0, slot_1, _ptr(img_1), w, h, step, n, _ptr(p0), _ptr(X),
                                                _ptr(fl), _ptr(Twp), _ptr(dTp), ctypes.byref(res)))
        k, m = res.n_tracked, res.n_new
        return dict(T_wc=T_wc, dT01=dT01, dT10=dT10, index=idx[:k].copy(), pts1=o1[:k].copy(), counts=[int(c) for c in counts],
                    n_detected=res.n_detected, new_p1=n1[:m].copy(), new_p0=n0[:m].copy(), used_5point=bool(res.used_5point),
                    n_5p_ransac=int(res.n_5p_ransac))

    # ---------------------------------------------------------------- stereo rectification (camera.cpp:300-546)
    def rectify_init(self, K_l, D_l, K_r, D_r, T_lr, w, h):
        """Build the rectification maps on the device; returns (K_rect [4], T_lr_rect [4,4])."""
        a = [np.ascontiguousarray(v, np.float32) for v in (K_l, D_l, K_r, D_r, T_lr)]
        K_rect, T_rect = np.zeros(4, np.float32), np.zeros((4, 4), np.float32)
        check(self.h, self.L.vo_rectify_init(self.h, *[_ptr(v) for v in a], int(w), int(h), _ptr(K_rect), _ptr(T_rect)))
        self._rect_wh = (int(w), int(h))
        return K_rect, T_rect

    def undistort_init(self, K4, D5, w, h):
        """Single-camera undistortion map (camera.cpp:57-87) into map set 0."""
        K, D = np.ascontiguousarray(K4, np.float32), np.ascontiguousarray(D5, np.float32)
        check(self.h, self.L.vo_undistort_init(self.h, _ptr(K), _ptr(D), int(w), int(h)))
        self._rect_wh = (int(w), int(h))

    def read_rectify_maps(self, right):
        w, h = self._rect_wh
        mu, mv = np.zeros((h, w), np.float32), np.zeros((h, w), np.float32)
        check(self.h, self.L.vo_read_rectify_maps(self.h, 1 if right else 0, _ptr(mu), _ptr(mv)))
        return mu, mv

    def upload_image_rectified(self, slot, right, img):
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        check(self.h, self.L.vo_upload_image_rectified(self.h, slot, 1 if right else 0, _ptr(img), img.shape[1], img.shape[0], img.strides[0]))
