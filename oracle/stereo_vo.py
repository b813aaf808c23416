"""StereoVO sequence oracle -- TEST INFRASTRUCTURE ONLY.

CPU restatement of StereoVO::trackStereoImages (core/visual_odometry/stereo_vo/stereo_vo.cpp:392-989) with the
bookkeeping it relies on: Landmark (landmark.cpp:59-152), Frame / StereoFrame (frame.cpp:44-60, 196-206),
StereoKeyframes (keyframes.cpp:177-303), MotionEstimator::localBundleAdjustmentSparseSolver_Stereo
(motion_estimator.cpp:1207-1340), SparseBAParameters::setPosesAndPoints (ba_solver/sparse_ba_parameters.h:292-465)
and the solver's write-back (ba_solver/sparse_bundle_adjustment.cpp:631-718).  Every numeric stage is one of the
pinned / restated oracles: cv2.calcOpticalFlowPyrLK, klt_scale_oracle.c, pose_oracle.c, misc_oracle.c, lba_oracle.c.
The keypoint detector is oracle/detect.py (see its header: cv::ORB is third-party and is replaced by an exact
integer Harris response; the bucketing around it is the reference's).

Things restated on purpose (they change results):
  * dT_pc_prev of the next frame is inverseSE3_f(dT.inverse()) (stereo_vo.cpp:643 + frame.cpp:50-54), not dT;
  * new features are appended to lmtrack_final AFTER n_pts was fixed, so a keyframe triangulates only the tracked
    survivors (stereo_vo.cpp:732-734 vs :767);
  * every keyframe re-triangulates and overwrites the 3-D point of every survivor that passes the gates (:794);
  * the first frame never becomes a keyframe (:842-949), the second always does (keyframes.cpp:222);
  * landmarks killed by the LBA (norm > 3000) are dropped at the first compaction of the next frame
    (landmark.cpp:304);
  * Eigen's general 4x4 float inverse (third-party, unpinned) is restated as the adjugate formula.
"""
import numpy as np

from . import detect as odet
from . import klt as oklt
from . import lba as olba
from . import misc as omisc
from . import pose as opose
from . import step as ostep

f32 = np.float32
D2R = f32(np.pi / 180.0)


def default_params(**kw):
    p = dict(window_size=21, max_level=3, thres_error=80.0, thres_bidirection=0.5, sampson_y=660.0,
             thres_poseba_error=3.0, n_bins_u=64, n_bins_v=32, det_edge=31, det_min_score=0,
             kf_overlap_ratio=0.6, kf_rot_deg=15.0, kf_trans=10.0, kf_window=9, do_scale_refine=True,
             lba_max_iter=10, lba_huber=0.5, lba_min_kf=3, lba_n_fix=2, detector="harris", fast_threshold=20)
    p.update(kw)
    return p


def inv4_f32(M):
    """General 4x4 inverse in float32 by the adjugate (restates Eigen's Matrix4f::inverse(), see module header)."""
    m = np.asarray(M, f32).reshape(16)
    inv = np.zeros(16, f32)

    def F(*t):   # sum of signed triple products, float32, left to right
        s = f32(0)
        for sign, a, b, c in t:
            v = f32(f32(m[a] * m[b]) * m[c])
            s = f32(s + v) if sign > 0 else f32(s - v)
        return s
    inv[0] = F((1, 5, 10, 15), (-1, 5, 11, 14), (-1, 9, 6, 15), (1, 9, 7, 14), (1, 13, 6, 11), (-1, 13, 7, 10))
    inv[4] = F((-1, 4, 10, 15), (1, 4, 11, 14), (1, 8, 6, 15), (-1, 8, 7, 14), (-1, 12, 6, 11), (1, 12, 7, 10))
    inv[8] = F((1, 4, 9, 15), (-1, 4, 11, 13), (-1, 8, 5, 15), (1, 8, 7, 13), (1, 12, 5, 11), (-1, 12, 7, 9))
    inv[12] = F((-1, 4, 9, 14), (1, 4, 10, 13), (1, 8, 5, 14), (-1, 8, 6, 13), (-1, 12, 5, 10), (1, 12, 6, 9))
    inv[1] = F((-1, 1, 10, 15), (1, 1, 11, 14), (1, 9, 2, 15), (-1, 9, 3, 14), (-1, 13, 2, 11), (1, 13, 3, 10))
    inv[5] = F((1, 0, 10, 15), (-1, 0, 11, 14), (-1, 8, 2, 15), (1, 8, 3, 14), (1, 12, 2, 11), (-1, 12, 3, 10))
    inv[9] = F((-1, 0, 9, 15), (1, 0, 11, 13), (1, 8, 1, 15), (-1, 8, 3, 13), (-1, 12, 1, 11), (1, 12, 3, 9))
    inv[13] = F((1, 0, 9, 14), (-1, 0, 10, 13), (-1, 8, 1, 14), (1, 8, 2, 13), (1, 12, 1, 10), (-1, 12, 2, 9))
    inv[2] = F((1, 1, 6, 15), (-1, 1, 7, 14), (-1, 5, 2, 15), (1, 5, 3, 14), (1, 13, 2, 7), (-1, 13, 3, 6))
    inv[6] = F((-1, 0, 6, 15), (1, 0, 7, 14), (1, 4, 2, 15), (-1, 4, 3, 14), (-1, 12, 2, 7), (1, 12, 3, 6))
    inv[10] = F((1, 0, 5, 15), (-1, 0, 7, 13), (-1, 4, 1, 15), (1, 4, 3, 13), (1, 12, 1, 7), (-1, 12, 3, 5))
    inv[14] = F((-1, 0, 5, 14), (1, 0, 6, 13), (1, 4, 1, 14), (-1, 4, 2, 13), (-1, 12, 1, 6), (1, 12, 2, 5))
    inv[3] = F((-1, 1, 6, 11), (1, 1, 7, 10), (1, 5, 2, 11), (-1, 5, 3, 10), (-1, 9, 2, 7), (1, 9, 3, 6))
    inv[7] = F((1, 0, 6, 11), (-1, 0, 7, 10), (-1, 4, 2, 11), (1, 4, 3, 10), (1, 8, 2, 7), (-1, 8, 3, 6))
    inv[11] = F((-1, 0, 5, 11), (1, 0, 7, 9), (1, 4, 1, 11), (-1, 4, 3, 9), (-1, 8, 1, 7), (1, 8, 3, 5))
    inv[15] = F((1, 0, 5, 10), (-1, 0, 6, 9), (-1, 4, 1, 10), (1, 4, 2, 9), (1, 8, 1, 6), (-1, 8, 2, 5))
    det = f32(f32(f32(m[0] * inv[0]) + f32(m[1] * inv[4])) + f32(m[2] * inv[8]))
    det = f32(det + f32(m[3] * inv[12]))
    idet = f32(f32(1.0) / det)
    return (inv * idet).astype(f32).reshape(4, 4)


def _project(K, X):
    """Camera::projectToPixel (camera.cpp:208-213), float32."""
    with np.errstate(divide="ignore", invalid="ignore"):
        invz = f32(1.0) / X[:, 2]
        return np.stack([K[0] * X[:, 0] * invz + K[2], K[1] * X[:, 1] * invz + K[3]], 1).astype(f32)


def reconstruct(pts_l, pts_r, K_l, K_r, T_rl, T_wc):
    """stereo_vo.cpp:767-797 / :911-941: DLT, 1-px^2 reprojection gates on both images, both depths > 0,
    X_w = T_wc * X_l.  Returns (Xw [n,3] f32, ok [n] bool)."""
    n = len(pts_l)
    if n == 0:
        return np.zeros((0, 3), f32), np.zeros(0, bool)
    Xl, Xr = omisc.triangulate_dlt(pts_l, pts_r, T_rl[:3, :3], T_rl[:3, 3], K_l, K_r)
    with np.errstate(invalid="ignore", over="ignore"):
        d0 = pts_l - _project(K_l, Xl)
        n0 = (d0[:, 0] * d0[:, 0] + d0[:, 1] * d0[:, 1]).astype(f32)
        d1 = pts_r - _project(K_r, Xr)
        n1 = (d1[:, 0] * d1[:, 0] + d1[:, 1] * d1[:, 1]).astype(f32)
        ok = ~(n0 > 1.0) & ~(n1 > 1.0) & (Xl[:, 2] > 0) & (Xr[:, 2] > 0)
    Xw = ostep._xform(np.asarray(T_wc, f32), Xl)
    return Xw, ok


class _Frame:
    __slots__ = ("id", "Twc", "Tcw", "dT01", "pts_l", "pts_r", "lm_ids")

    def __init__(self, fid):
        self.id = fid
        self.set_pose(np.eye(4, dtype=f32))
        self.dT01 = np.eye(4, dtype=f32)
        self.pts_l = np.zeros((0, 2), f32)
        self.pts_r = np.zeros((0, 2), f32)
        self.lm_ids = np.zeros(0, np.int64)

    def set_pose(self, Twc):          # frame.cpp:44-48
        self.Twc = np.asarray(Twc, f32).copy()
        self.Tcw = opose.inverse_se3_f(self.Twc)


class StereoVOOracle:
    def __init__(self, w, h, K_l, K_r, T_lr, params=None, lk=oklt.lk_cv2):
        self.w, self.h = w, h
        self.K_l, self.K_r = np.asarray(K_l, f32), np.asarray(K_r, f32)
        self.T_lr = np.asarray(T_lr, f32)
        self.T_rl = opose.inverse_se3_f(self.T_lr)
        self.p = params or default_params()
        self.lk = lk
        # landmark table
        self.X = []            # float32[3]
        self.tri = []
        self.alive = []
        self.bundled = []
        self.last_frame = []   # id of the last frame that observed it (== related_frames_.back())
        self.kf_obs = []       # list of (kf_frame_id, is_right, x, y)
        self.prev = None
        self.prev_imgs = None
        self.window = []       # keyframes in the sliding window (_Frame)
        self.n_frames = 0
        self.poses = []        # stats_frame[k].Twc
        self.info = []

    # ------------------------------------------------------------------ landmarks
    def _new_landmarks(self, k, fid):
        base = len(self.X)
        for _ in range(k):
            self.X.append(np.zeros(3, f32)); self.tri.append(False); self.alive.append(True); self.bundled.append(False)
            self.last_frame.append(fid); self.kf_obs.append([])
        return np.arange(base, base + k, dtype=np.int64)

    def _extract(self, img, occupied):
        p = self.p
        if p.get("detector", "harris") == "orb":          # the reference's extractor (oracle/orb.py, pinned against cv2.ORB)
            from . import orb as oorb
            return oorb.detect_bucketed(img, occupied, p["n_bins_u"], p["n_bins_v"], p["fast_threshold"], backend="cv2")
        return odet.detect_bucketed(img, occupied, p["n_bins_u"], p["n_bins_v"], p["det_edge"], p["det_min_score"])

    # ------------------------------------------------------------------ keyframes
    def _check_update_rule(self, fr):             # keyframes.cpp:217-303
        if not self.window:
            return True
        kf = self.window[-1]
        last = np.asarray([self.last_frame[i] for i in kf.lm_ids], np.int64)
        cnt_tracked = int((last == fr.id).sum())
        with np.errstate(invalid="ignore", divide="ignore"):
            ratio = f32(cnt_tracked) / f32(len(kf.lm_ids))
        if ratio <= f32(self.p["kf_overlap_ratio"]):
            return True
        dT = ostep.mul4_f32(kf.Tcw, fr.Twc)
        costheta = f32(f32(f32(f32(dT[0, 0] + dT[1, 1]) + dT[2, 2]) - f32(1.0)) * f32(0.5))
        if costheta >= 0.999999:
            costheta = f32(0.999999)
        if costheta <= -0.999999:
            costheta = f32(-0.999999)
        rot = f32(np.arccos(costheta))
        t = dT[:3, 3]
        dtrans = f32(np.sqrt(f32(f32(f32(t[0] * t[0]) + f32(t[1] * t[1])) + f32(t[2] * t[2]))))
        return bool(rot >= f32(self.p["kf_rot_deg"]) * D2R or dtrans >= f32(self.p["kf_trans"]))

    def _add_keyframe(self, fr):                  # keyframes.cpp:177-215
        if len(self.window) == self.p["kf_window"]:
            self.window.pop(0)
        self.window.append(fr)
        for i, pt in zip(fr.lm_ids, fr.pts_l):
            self.kf_obs[i].append((fr.id, 0, f32(pt[0]), f32(pt[1])))
        for i, pt in zip(fr.lm_ids, fr.pts_r):
            self.kf_obs[i].append((fr.id, 1, f32(pt[0]), f32(pt[1])))

    # ------------------------------------------------------------------ local BA
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
        lms, pts, obs_ptr, obs_frame, obs_right, obs_px = [], [], [0], [], [], []
        for i in lmset:
            ob = [o for o in self.kf_obs[i] if o[0] in fidx]
            if len(ob) < 2:
                continue
            Xw = self.X[i].astype(np.float64)
            Xr = np.array([(Tjw_ref[r, 0] * Xw[0] + Tjw_ref[r, 1] * Xw[1] + Tjw_ref[r, 2] * Xw[2] + Tjw_ref[r, 3]) * inv_s for r in range(3)])
            lms.append(i); pts.append(Xr)
            for (fid, right, x, y) in ob:
                obs_frame.append(fidx[fid]); obs_right.append(right); obs_px.append((float(x), float(y)))
            obs_ptr.append(len(obs_frame))
        poses = np.zeros((nf, 4, 4))
        for k, fr in enumerate(frames):
            Tjw = fr.Tcw.astype(np.float64)
            Tjw[3] = [0, 0, 0, 1]
            Tj = Tjw @ Twj_ref
            Tj[:3, 3] *= inv_s
            poses[k] = Tj
        opt_index = np.full(nf, -1, np.int32)
        opt_index[p["lba_n_fix"]:] = np.arange(nf - p["lba_n_fix"])
        T_lr_s = self.T_lr.astype(np.float64)
        T_lr_s[:3, 3] *= inv_s
        prob = dict(n_frames=nf, n_opt=nf - p["lba_n_fix"], n_points=len(lms), n_obs=len(obs_frame), poses=poses,
                    opt_index=opt_index, points=np.asarray(pts, np.float64).reshape(-1, 3), obs_ptr=np.asarray(obs_ptr, np.int32),
                    obs_frame=np.asarray(obs_frame, np.int32), obs_right=np.asarray(obs_right, np.uint8),
                    obs_px=np.asarray(obs_px, np.float64).reshape(-1, 2), K_l=self.K_l.astype(np.float64),
                    K_r=self.K_r.astype(np.float64), T_lr=T_lr_s, is_stereo=1, huber=p["lba_huber"], lam=1e-5,
                    max_iter=p["lba_max_iter"])
        if prob["n_points"] == 0:
            return dict(n_points=0, n_obs=0, avg_err=np.zeros(0), ok=False)
        rc, poses_o, points_o, avg, ok = olba.lba_solve(prob)
        if rc != 0:
            raise RuntimeError(f"Local BA failed rc={rc}")
        # write-back (sparse_bundle_adjustment.cpp:631-718)
        for k, fr in enumerate(frames):
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
    def track(self, img_l, img_r):
        p = self.p
        fr = _Frame(self.n_frames)
        self.n_frames += 1
        info = dict(frame=fr.id, keyframe=False, n_new=0, lba=None)
        if self.prev is None:
            # ---- first frame (stereo_vo.cpp:842-949)
            pts_l1 = self._extract(img_l, np.zeros((0, 2), f32))
            pts_r1, m = oklt.track_bidirection(self.lk, img_l, img_r, pts_l1, p["window_size"], p["max_level"], p["thres_error"],
                                               p["thres_bidirection"]) if len(pts_l1) else (np.zeros((0, 2), f32), np.zeros(0, bool))
            pl, pr = pts_l1[m], pts_r1[m]
            ids = self._new_landmarks(len(pl), fr.id)
            Xw, ok = reconstruct(pl, pr, self.K_l, self.K_r, self.T_rl, fr.Twc)
            for j, i in enumerate(ids):
                if ok[j]:
                    self.X[i] = Xw[j]; self.tri[i] = True
            fr.pts_l, fr.pts_r, fr.lm_ids = pl, pr, ids
            info.update(n_extracted=len(pts_l1), n_tracked=0, n_new=len(pl), n_recon=int(ok.sum()))
            self.dbg = dict(pts_new=pts_l1, new_l=pl, new_r=pr, Xw_recon=Xw, ok_recon=ok)
        else:
            pv = self.prev
            I0l = self.prev_imgs[0]
            # landmark.cpp:304: dead / untracked landmarks never survive the first compaction
            alive = np.asarray([self.alive[i] for i in pv.lm_ids], bool)
            ids0 = pv.lm_ids[alive]
            pts_l0, pts_r0 = pv.pts_l[alive], pv.pts_r[alive]
            Xw = np.asarray([self.X[i] for i in ids0], f32).reshape(-1, 3)
            tri = np.asarray([self.tri[i] for i in ids0], bool)
            st = ostep.stereo_track_step(I0l, img_l, img_r, pts_l0, pts_r0, Xw, tri, pv.Twc, pv.dT01, self.K_l, self.K_r, self.T_lr,
                                         p["window_size"], p["max_level"], p["thres_error"], p["thres_poseba_error"],
                                         do_scale_refine=p["do_scale_refine"], sampson_y=p["sampson_y"], lk=self.lk)
            self.dbg = dict(pts_l0=pts_l0, pts_r0=pts_r0, Xw=Xw, tri=tri, T_wp=pv.Twc.copy(), dT_prev=pv.dT01.copy(), step=st)
            fr.set_pose(st["T_wc"])
            dT10 = inv4_f32(st["dT_pc"])                       # stereo_vo.cpp:643
            fr.dT01 = opose.inverse_se3_f(dT10)                # frame.cpp:50-54
            ids = ids0[st["index"]]
            pl, pr = st["pts_l1"], st["pts_r1"]
            for i in ids:                                      # [8] addObservationAndRelatedFrame
                self.last_frame[i] = fr.id
            n_tracked = len(ids)
            # ---- [10] new features from empty bins (:690-740)
            pts_new = self._extract(img_l, pl)
            if len(pts_new):
                r_new, m = oklt.track_bidirection(self.lk, img_l, img_r, pts_new, p["window_size"], p["max_level"], p["thres_error"],
                                                  p["thres_bidirection"])
                Xl, Xr = omisc.triangulate_dlt(pts_new, r_new, self.T_rl[:3, :3], self.T_rl[:3, 3], self.K_l, self.K_r)
                with np.errstate(invalid="ignore"):
                    keep = m & (Xl[:, 2] > 0) & (Xr[:, 2] > 0)
                nl, nr = pts_new[keep], r_new[keep]
                new_ids = self._new_landmarks(len(nl), fr.id)
                pl, pr, ids = np.concatenate([pl, nl]), np.concatenate([pr, nr]), np.concatenate([ids, new_ids])
                info["n_new"] = len(nl)
                self.dbg.update(new_l=nl, new_r=nr)
            self.dbg.update(pts_new=pts_new)
            fr.pts_l, fr.pts_r, fr.lm_ids = pl.astype(f32), pr.astype(f32), ids
            info.update(n_in=len(ids0), n_tracked=n_tracked, counts=st["counts"], gn_iters=st["gn_iters"], n_extracted=len(pts_new))
            # ---- [12] keyframe (:755-827)
            if self._check_update_rule(fr):
                info["keyframe"] = True
                self._add_keyframe(fr)
                Xw_new, ok = reconstruct(pl[:n_tracked], pr[:n_tracked], self.K_l, self.K_r, self.T_rl, fr.Twc)
                for j in np.flatnonzero(ok):
                    i = ids[j]
                    self.X[i] = Xw_new[j]; self.tri[i] = True
                info["n_recon"] = int(ok.sum())
                info["lba"] = self._local_ba()
        self.poses.append(fr.Twc.copy())         # stats_frame.back().Twc (:979-980), after the LBA
        self.prev = fr
        self.prev_imgs = (img_l, img_r)
        self.info.append(info)
        return fr.Twc.copy(), info
