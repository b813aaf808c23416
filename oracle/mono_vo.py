"""MonoVO sequence oracle -- TEST INFRASTRUCTURE ONLY.

CPU restatement of MonoVO::trackImage (core/visual_odometry/mono_vo/mono_vo.cpp:496-1194) with the bookkeeping it relies
on: Landmark incl. the parallax of every observation against the first one (landmark.cpp:28-135), Frame (frame.cpp:3-60),
Keyframes (keyframes.cpp:30-120), MotionEstimator::localBundleAdjustmentSparseSolver (motion_estimator.cpp:1090-1205),
SparseBAParameters::setPosesAndPoints (ba_solver/sparse_ba_parameters.h:292-465) and the solver's write-back
(ba_solver/sparse_bundle_adjustment.cpp:631-718).  The numeric stages are the per-stage oracles (oracle/mono_step.py,
oracle/five_point.py, misc_oracle.c, lba_oracle.c); the keypoint detector is oracle/detect.py.

Restated on purpose (they change results):
  * the first image is a keyframe (the window is empty, keyframes.cpp:52-53); its pose difference is the "initial scale"
    dT10 = [I | (0, 0, -1)] (mono_vo.cpp:547-550);
  * on the second image the observations are added BEFORE the frame gets its pose (:602-603 vs :611), so the parallax of
    the tracked landmarks is computed with an identity rotation; new landmarks see the real poses (:646-647);
  * dT01 of a frame is inverseSE3_f(dT10) (frame.cpp:50-54), not the dT01 the pose-only BA returned;
  * a landmark's parallax reference is its FIRST observation and the CURRENT pose of the frame it was made in
    (landmark.cpp:108-111) -- the local BA moves those poses;
  * initial reconstruction (:660-687) has no reprojection gate, keyframe reconstruction (:1032-1076) has both gates and
    needs more than two keyframe observations;
  * "refine" after the LBA (:1084-1128) rewrites every last observation with itself: nothing.
"""
import numpy as np

from . import detect as odet
from . import klt as oklt
from . import lba as olba
from . import misc as omisc
from . import mono_step as omono
from . import pose as opose
from . import step as ostep

f32 = np.float32
D2R = f32(np.pi / 180.0)


def default_params(**kw):
    """config/mono/kitti_00.yaml (thres_parallax in degrees, converted at mono_vo.cpp:219)."""
    p = dict(window_size=21, max_level=6, thres_error=60.0, thres_bidirection=0.5, thres_sampson=1000.0, thres_parallax_deg=1.0,
             n_bins_u=30, n_bins_v=12, det_edge=31, det_min_score=0, thres_5p=1.0, thres_poseba_error=5.0,
             kf_overlap_ratio=0.6, kf_rot_deg=10.0, kf_trans=4.0, kf_window=9, do_scale_refine=True,
             lba_max_iter=10, lba_huber=0.5, lba_min_kf=3, lba_n_fix=2, detector="harris", fast_threshold=15)
    p.update(kw)
    return p


class _Frame:
    __slots__ = ("id", "Twc", "Tcw", "dT01", "dT10", "pts", "lm_ids")

    def __init__(self, fid):
        self.id = fid
        self.set_pose(np.eye(4, dtype=f32))
        self.dT01 = np.eye(4, dtype=f32)
        self.dT10 = np.eye(4, dtype=f32)
        self.pts = np.zeros((0, 2), f32)
        self.lm_ids = np.zeros(0, np.int64)

    def set_pose(self, Twc):                      # frame.cpp:44-48
        self.Twc = np.asarray(Twc, f32).copy()
        self.Tcw = opose.inverse_se3_f(self.Twc)

    def set_pose_diff10(self, dT10):              # frame.cpp:50-54
        self.dT10 = np.asarray(dT10, f32).copy()
        self.dT01 = opose.inverse_se3_f(self.dT10)


class MonoVOOracle:
    def __init__(self, w, h, K, params=None, lk=oklt.lk_cv2, five_point=None):
        self.w, self.h = w, h
        self.K = np.asarray(K, f32)
        self.p = params or default_params()
        self.lk = lk
        self.five_point = five_point              # callable (frame_id, pts0, pts1) -> (ok, R10, t10, mask); None = cv2
        self.X, self.tri, self.alive, self.bundled = [], [], [], []
        self.last_frame, self.first_frame, self.first_px, self.last_px = [], [], [], []
        self.age, self.last_parallax, self.kf_obs = [], [], []
        self.frames = []                          # all frames (poses are read back by the parallax / reconstruction)
        self.prev = None
        self.prev_img = None
        self.window = []
        self.initialised = False
        self.poses = []
        self.info = []

    # ------------------------------------------------------------------ landmarks
    def _new_landmarks(self, pts, fr):
        base = len(self.X)
        for pt in pts:
            self.X.append(np.zeros(3, f32)); self.tri.append(False); self.alive.append(True); self.bundled.append(False)
            self.last_frame.append(fr.id); self.first_frame.append(fr.id)
            self.first_px.append(np.asarray(pt, f32).copy()); self.last_px.append(np.asarray(pt, f32).copy())
            self.age.append(1); self.last_parallax.append(f32(0)); self.kf_obs.append([])
        return np.arange(base, base + len(pts), dtype=np.int64)

    def _add_observations(self, ids, pts, fr):
        """Landmark::addObservationAndRelatedFrame (landmark.cpp:76-135) for a batch (vectorised per first frame; float32 in
        the reference's operation order)."""
        if len(ids) == 0:
            return
        K = self.K
        fxinv, fyinv = f32(f32(1.0) / K[0]), f32(f32(1.0) / K[1])
        pts = np.asarray(pts, f32).reshape(-1, 2)
        ids = [int(i) for i in ids]
        f0 = np.asarray([self.first_frame[i] for i in ids], np.int64)
        p0 = np.asarray([self.first_px[i] for i in ids], f32).reshape(-1, 2)
        par = np.zeros(len(ids), f32)
        for g in np.unique(f0):
            sel = np.flatnonzero(f0 == g)
            R = ostep.mul4_f32(self.frames[int(g)].Tcw, fr.Twc)[:3, :3]
            x0 = np.stack([(p0[sel, 0] - K[2]) * fxinv, (p0[sel, 1] - K[3]) * fyinv, np.ones(len(sel), f32)], 1).astype(f32)
            b = np.stack([(pts[sel, 0] - K[2]) * fxinv, (pts[sel, 1] - K[3]) * fyinv, np.ones(len(sel), f32)], 1).astype(f32)
            x1 = np.stack([((R[r, 0] * b[:, 0] + R[r, 1] * b[:, 1]) + R[r, 2] * b[:, 2]) for r in range(3)], 1).astype(f32)
            dot = ((x0[:, 0] * x1[:, 0] + x0[:, 1] * x1[:, 1]) + x0[:, 2] * x1[:, 2]).astype(f32)
            n0 = np.sqrt(((x0[:, 0] * x0[:, 0] + x0[:, 1] * x0[:, 1]) + x0[:, 2] * x0[:, 2]).astype(f32)).astype(f32)
            n1 = np.sqrt(((x1[:, 0] * x1[:, 0] + x1[:, 1] * x1[:, 1]) + x1[:, 2] * x1[:, 2]).astype(f32)).astype(f32)
            c = (dot / (n0 * n1).astype(f32)).astype(f32)
            c = np.where(c >= 1.0, f32(0.99999), c)
            c = np.where(c <= -1.0, f32(-0.99999), c).astype(f32)
            par[sel] = np.arccos(c).astype(f32)
        for j, i in enumerate(ids):
            self.age[i] += 1
            self.last_frame[i] = fr.id
            self.last_px[i] = pts[j].copy()
            self.last_parallax[i] = par[j]

    def _extract(self, img, occupied):
        p = self.p
        if p.get("detector", "harris") == "orb":          # the reference's extractor (oracle/orb.py, pinned against cv2.ORB)
            from . import orb as oorb
            return oorb.detect_bucketed(img, occupied, p["n_bins_u"], p["n_bins_v"], p["fast_threshold"], backend="cv2")
        return odet.detect_bucketed(img, occupied, p["n_bins_u"], p["n_bins_v"], p["det_edge"], p["det_min_score"])

    # ------------------------------------------------------------------ keyframes
    def _check_update_rule(self, fr):             # keyframes.cpp:47-120
        if not self.window:
            return True
        kf = self.window[-1]
        last = np.asarray([self.last_frame[i] for i in kf.lm_ids], np.int64)
        with np.errstate(invalid="ignore", divide="ignore"):
            ratio = f32(int((last == fr.id).sum())) / f32(len(kf.lm_ids))
        if ratio <= f32(self.p["kf_overlap_ratio"]):
            return True
        dT = ostep.mul4_f32(kf.Tcw, fr.Twc)
        costheta = f32(f32(f32(f32(dT[0, 0] + dT[1, 1]) + dT[2, 2]) - f32(1.0)) * f32(0.5))
        costheta = min(max(costheta, f32(-0.999999)), f32(0.999999))
        rot = f32(np.arccos(costheta))
        t = dT[:3, 3]
        dtrans = f32(np.sqrt(f32(f32(f32(t[0] * t[0]) + f32(t[1] * t[1])) + f32(t[2] * t[2]))))
        return bool(rot >= f32(self.p["kf_rot_deg"]) * D2R or dtrans >= f32(self.p["kf_trans"]))

    def _add_keyframe(self, fr):                  # keyframes.cpp:30-45
        if len(self.window) == self.p["kf_window"]:
            self.window.pop(0)
        self.window.append(fr)
        for i in fr.lm_ids:
            i = int(i)
            self.kf_obs[i].append((fr.id, f32(self.last_px[i][0]), f32(self.last_px[i][1])))

    def _project(self, X):
        K = self.K
        with np.errstate(divide="ignore", invalid="ignore"):
            invz = f32(1.0) / X[:, 2]
            return np.stack([K[0] * X[:, 0] * invz + K[2], K[1] * X[:, 1] * invz + K[3]], 1).astype(f32)

    def _dlt_groups(self, cand, pt0, pt1, f0_ids, fr1):
        """triangulateDLT of candidates grouped by the frame of their first point; returns X0, X1 and Tw0 per candidate."""
        X0 = np.zeros((len(cand), 3), f32)
        X1 = np.zeros((len(cand), 3), f32)
        for f0 in sorted(set(f0_ids)):
            sel = np.flatnonzero(np.asarray(f0_ids) == f0)
            T10 = ostep.mul4_f32(fr1.Tcw, self.frames[f0].Twc)
            a, b = omisc.triangulate_dlt(pt0[sel], pt1[sel], T10[:3, :3].copy(), T10[:3, 3].copy(), self.K, self.K)
            X0[sel], X1[sel] = a, b
        return X0, X1

    def _reconstruct_initial(self, fr):           # mono_vo.cpp:660-687
        thr = f32(self.p["thres_parallax_deg"]) * D2R
        cand = [int(i) for i in fr.lm_ids if not self.tri[int(i)] and self.last_parallax[int(i)] >= thr]
        if not cand:
            return 0
        pt0 = np.asarray([self.first_px[i] for i in cand], f32)
        pt1 = np.asarray([self.last_px[i] for i in cand], f32)
        f0 = [self.first_frame[i] for i in cand]
        X0, _ = self._dlt_groups(cand, pt0, pt1, f0, self.frames[fr.id])
        n = 0
        for j, i in enumerate(cand):
            if X0[j, 2] > 0:
                self.X[i] = ostep._xform(self.frames[f0[j]].Twc, X0[j:j + 1])[0]
                self.tri[i] = True
                n += 1
        return n

    def _reconstruct_keyframe(self, fr):          # mono_vo.cpp:1032-1076
        thr = f32(self.p["thres_parallax_deg"]) * D2R
        cand = [int(i) for i in fr.lm_ids
                if self.alive[int(i)] and not self.tri[int(i)] and self.last_parallax[int(i)] >= thr and len(self.kf_obs[int(i)]) > 2]
        if not cand:
            return 0
        pt0 = np.asarray([self.kf_obs[i][0][1:] for i in cand], f32)
        pt1 = np.asarray([self.kf_obs[i][-1][1:] for i in cand], f32)
        f0 = [self.kf_obs[i][0][0] for i in cand]
        X0, X1 = self._dlt_groups(cand, pt0, pt1, f0, self.frames[fr.id])   # the last keyframe observation is this frame
        with np.errstate(invalid="ignore", over="ignore"):
            d0 = pt0 - self._project(X0)
            d1 = pt1 - self._project(X1)
            n0 = (d0[:, 0] * d0[:, 0] + d0[:, 1] * d0[:, 1]).astype(f32)
            n1 = (d1[:, 0] * d1[:, 0] + d1[:, 1] * d1[:, 1]).astype(f32)
            ok = ~(n0 > 1.0) & ~(n1 > 1.0) & (X0[:, 2] > 0) & (X1[:, 2] > 0)
        for j in np.flatnonzero(ok):
            i = cand[j]
            self.X[i] = ostep._xform(self.frames[f0[j]].Twc, X0[j:j + 1])[0]
            self.tri[i] = True
        return int(ok.sum())

    # ------------------------------------------------------------------ local BA (mono)
    def _local_ba(self):
        p = self.p
        if len(self.window) < p["lba_min_kf"]:
            return None
        frames = self.window
        nf = len(frames)
        fidx = {fr.id: k for k, fr in enumerate(frames)}
        seen, lmset = set(), []
        for fr in frames:
            for i in fr.lm_ids:
                i = int(i)
                if i not in seen and self.tri[i] and self.alive[i]:
                    seen.add(i); lmset.append(i)
        Twj_ref = frames[0].Twc.astype(np.float64)
        Twj_ref[3] = [0, 0, 0, 1]
        Tjw_ref = np.eye(4)
        Tjw_ref[:3, :3] = Twj_ref[:3, :3].T
        Tjw_ref[:3, 3] = -(Twj_ref[:3, :3].T @ Twj_ref[:3, 3])
        inv_s, s = 1.0 / 10.0, 10.0
        lms, pts, obs_ptr, obs_frame, obs_px = [], [], [0], [], []
        for i in lmset:
            ob = [o for o in self.kf_obs[i] if o[0] in fidx]
            if len(ob) < 2:
                continue
            Xw = self.X[i].astype(np.float64)
            pts.append([(Tjw_ref[r, 0] * Xw[0] + Tjw_ref[r, 1] * Xw[1] + Tjw_ref[r, 2] * Xw[2] + Tjw_ref[r, 3]) * inv_s for r in range(3)])
            lms.append(i)
            for (fid, x, y) in ob:
                obs_frame.append(fidx[fid]); obs_px.append((float(x), float(y)))
            obs_ptr.append(len(obs_frame))
        if not lms:
            return dict(n_points=0, n_obs=0, avg_err=np.zeros(0), ok=False)
        poses = np.zeros((nf, 4, 4))
        for k, fr in enumerate(frames):
            Tjw = fr.Tcw.astype(np.float64)
            Tjw[3] = [0, 0, 0, 1]
            Tj = Tjw @ Twj_ref
            Tj[:3, 3] *= inv_s
            poses[k] = Tj
        opt_index = np.full(nf, -1, np.int32)
        opt_index[p["lba_n_fix"]:] = np.arange(nf - p["lba_n_fix"])
        prob = dict(n_frames=nf, n_opt=nf - p["lba_n_fix"], n_points=len(lms), n_obs=len(obs_frame), poses=poses,
                    opt_index=opt_index, points=np.asarray(pts, np.float64).reshape(-1, 3), obs_ptr=np.asarray(obs_ptr, np.int32),
                    obs_frame=np.asarray(obs_frame, np.int32), obs_right=np.zeros(len(obs_frame), np.uint8),
                    obs_px=np.asarray(obs_px, np.float64).reshape(-1, 2), K_l=self.K.astype(np.float64),
                    K_r=self.K.astype(np.float64), T_lr=np.eye(4), is_stereo=0, huber=p["lba_huber"], lam=1e-5,
                    max_iter=p["lba_max_iter"])
        rc, poses_o, points_o, avg, ok = olba.lba_solve(prob)
        if rc != 0:
            raise RuntimeError(f"Local BA failed rc={rc}")
        for k, fr in enumerate(frames):             # write-back (sparse_bundle_adjustment.cpp:631-718)
            if opt_index[k] < 0:
                continue
            Tjw = poses_o[k].copy()
            Tjw[:3, 3] *= s
            Tjw = Tjw @ Tjw_ref
            Tjw_f = Tjw.astype(f32)
            Tjw_f[3] = [0, 0, 0, 1]
            fr.set_pose(opose.inverse_se3_f(Tjw_f))
        for j, i in enumerate(lms):
            X = points_o[j] * s
            Xf = np.array([Twj_ref[r, 0] * X[0] + Twj_ref[r, 1] * X[1] + Twj_ref[r, 2] * X[2] + Twj_ref[r, 3] for r in range(3)]).astype(f32)
            self.X[i] = Xf
            self.tri[i] = True
            nrm = f32(np.sqrt(f32(f32(f32(Xf[0] * Xf[0]) + f32(Xf[1] * Xf[1])) + f32(Xf[2] * Xf[2]))))
            if nrm <= 3000:
                self.bundled[i] = True
            else:
                self.alive[i] = False
        return dict(n_points=len(lms), n_obs=len(obs_frame), avg_err=avg, ok=ok, problem=prob)

    # ------------------------------------------------------------------ the step
    def track(self, img):
        p = self.p
        fr = _Frame(len(self.frames))
        self.frames.append(fr)
        info = dict(frame=fr.id, keyframe=False, n_new=0, lba=None, used_5point=False, n_recon=0)
        fp = None
        if self.five_point is not None:
            fp = lambda a, b: self.five_point(fr.id, a, b)       # noqa: E731
        if self.prev is None:
            # ---- first image (mono_vo.cpp:528-561)
            pts = self._extract(img, np.zeros((0, 2), f32))
            fr.lm_ids = self._new_landmarks(pts, fr)
            fr.pts = pts.astype(f32)
            T_init = np.eye(4, dtype=f32)
            T_init[2, 3] = -1.0
            fr.set_pose_diff10(T_init)
            info.update(n_extracted=len(pts), n_tracked=0, n_new=len(pts))
        else:
            pv = self.prev
            alive = np.asarray([self.alive[int(i)] for i in pv.lm_ids], bool)        # landmark.cpp:233-270
            ids0, pts0 = pv.lm_ids[alive], pv.pts[alive]
            if not self.initialised:
                # ---- second image (:562-696)
                st = omono.mono_init_step(self.prev_img, img, pts0, pv.Twc, self.K, p["window_size"], p["max_level"], p["thres_error"],
                                          p["thres_bidirection"], p["thres_sampson"], p["thres_5p"], p["n_bins_u"], p["n_bins_v"],
                                          p["det_edge"], p["det_min_score"], lk=self.lk, five_point=fp, detect_fn=self._extract)
                ids = ids0[st["index"]]
                self._add_observations(ids, st["pts1"], fr)                          # :602-603, pose still identity
                fr.set_pose(st["T_wc"])
                fr.set_pose_diff10(st["dT10"])
            else:
                Xw = np.asarray([self.X[int(i)] for i in ids0], f32).reshape(-1, 3)
                tri = np.asarray([self.tri[int(i)] for i in ids0], bool)
                bun = np.asarray([self.bundled[int(i)] for i in ids0], bool)
                st = omono.mono_frame_step(self.prev_img, img, pts0, Xw, tri, bun, pv.Twc, pv.dT01, self.K, p["window_size"], p["max_level"],
                                           p["thres_error"], p["thres_bidirection"], p["thres_sampson"], p["thres_poseba_error"],
                                           len(self.window) > 5, p["n_bins_u"], p["n_bins_v"], p["det_edge"], p["det_min_score"],
                                           do_scale_refine=p["do_scale_refine"], lk=self.lk, five_point=fp, thres_5p=p["thres_5p"], detect_fn=self._extract)
                ids = ids0[st["index"]]
                fr.set_pose(st["T_wc"])                                              # :889 / :947
                fr.set_pose_diff10(st["dT10"])
                self._add_observations(ids, st["pts1"], fr)                          # :966-967
            self.dbg = dict(pts0=pts0, ids0=ids0, step=st, T_wp=pv.Twc.copy(), dT_prev=pv.dT01.copy())
            info["used_5point"] = bool(st.get("used_5point", False))
            n_tracked = len(ids)
            pts = st["pts1"]
            # new features back-tracked into the previous image become landmarks born there (:638-657 / :993-1012)
            if len(st.get("new_p1", ())):
                new_ids = self._new_landmarks(st["new_p0"], pv)
                self._add_observations(new_ids, st["new_p1"], fr)
                ids = np.concatenate([ids, new_ids])
                pts = np.concatenate([pts, st["new_p1"]])
                info["n_new"] = len(new_ids)
            fr.pts, fr.lm_ids = pts.astype(f32), ids
            info.update(n_in=len(ids0), n_tracked=n_tracked, counts=st["counts"], n_extracted=st.get("n_detected", 0))
            if not self.initialised:
                info["n_recon"] = self._reconstruct_initial(fr)
                self.initialised = True
        # ---- keyframe (:1021-1157)
        if self._check_update_rule(fr):
            info["keyframe"] = True
            self._add_keyframe(fr)
            info["n_recon_kf"] = self._reconstruct_keyframe(fr)
            info["lba"] = self._local_ba()
        self.prev = fr
        self.prev_img = img
        self.poses.append(fr.Twc.copy())
        self.info.append(info)
        return fr.Twc.copy(), info

    def all_poses(self):
        """stats_frame[j].Twc refreshed from all_frames_ (:1184-1185)."""
        return np.stack([f.Twc for f in self.frames])
