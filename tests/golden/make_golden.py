#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.  Run from the repo root: python tests/golden/make_golden.py

* klt_cv2_golden.npz  -- outputs of the reference's OWN library call (cv2 4.13.0:
  calcOpticalFlowPyrLK / buildOpticalFlowPyramid / Sobel) on a small seeded image pair, with the
  reference's argument patterns (feature_tracker.cpp:29,60,69,108,117,186). These pin the oracle
  (and through it the CUDA path) even on a box whose cv2 differs.
* oracle_regression.npz -- outputs of the C restatement for the paths the reference cannot pin
  (pose GN, triangulation, LBA, depth filter, trackWithScale): regression vectors, NOT reference
  outputs ("parity unpinned").
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    import cv2
    from oracle import klt, lba, misc, pose
    from visual_odometry_ros_b200 import synth
    cv2.setNumThreads(1)
    W, H = 200, 136
    rng = np.random.default_rng(77)
    img0 = synth.textured_image(rng, W, H)
    img1 = synth.warp_translate_field(img0, 1.7, -0.9)
    img1 = np.clip(img1.astype(np.float32) + rng.normal(0, 1.0, img1.shape), 0, 255).astype(np.uint8)
    pts0 = synth.grid_features(rng, 120, W, H, border=6, nx=15, ny=10)     # some windows leave the image
    prior = (pts0 + np.array([[1.0, -1.0]], np.float32)).astype(np.float32)
    out = dict(img0=img0, img1=img1, pts0=pts0, prior=prior, cv2_version=np.array(cv2.__version__))
    for win in (21, 15):
        for ml in (0, 2, 6):
            p, s, e = klt.lk_cv2(img0, img1, pts0, win, ml)
            out[f"fwd_w{win}_l{ml}_p"], out[f"fwd_w{win}_l{ml}_s"], out[f"fwd_w{win}_l{ml}_e"] = p, s, np.where(s > 0, e, 0)
            p, s, e = klt.lk_cv2(img0, img1, pts0, win, ml, 4, prior)
            out[f"pri_w{win}_l{ml}_p"], out[f"pri_w{win}_l{ml}_s"], out[f"pri_w{win}_l{ml}_e"] = p, s, np.where(s > 0, e, 0)
    for name, fn in (("track", lambda: klt.track(klt.lk_cv2, img0, img1, pts0, 21, 3, 30.0)),
                     ("track_with_prior", lambda: klt.track_with_prior(klt.lk_cv2, img0, img1, pts0, prior, 21, 3, 30.0)),
                     ("track_bidirection", lambda: klt.track_bidirection(klt.lk_cv2, img0, img1, pts0, 21, 3, 30.0, 0.5)),
                     ("track_bidirection_with_prior",
                      lambda: klt.track_bidirection_with_prior(klt.lk_cv2, img0, img1, pts0, prior, 21, 3, 30.0, 0.5))):
        p, m = fn()
        out[f"ft_{name}_p"], out[f"ft_{name}_m"] = p, m
    nl, pyr = cv2.buildOpticalFlowPyramid(img0, (21, 21), 3, withDerivatives=True)
    out["pyr_nlevels"] = np.array(nl + 1)
    for l in range(nl + 1):
        out[f"pyr_img{l}"], out[f"pyr_der{l}"] = pyr[2 * l], pyr[2 * l + 1]
    out["sobel_du"] = cv2.Sobel(img0, cv2.CV_32F, 1, 0, ksize=3)
    out["sobel_dv"] = cv2.Sobel(img0, cv2.CV_32F, 0, 1, ksize=3)
    np.savez_compressed(os.path.join(HERE, "klt_cv2_golden.npz"), **out)

    reg = {}
    s = synth.pose_scene(seed=1001, n=120)
    K, Tlr = synth.kitti_K(), synth.kitti_T_lr()
    ok, T01, mask, it = pose.pose_gn_stereo(s["X"], s["pts_l1"], s["pts_r1"], K, K, Tlr, 3.0, np.eye(4))
    reg.update(pose_X=s["X"], pose_pl=s["pts_l1"], pose_pr=s["pts_r1"], pose_T01=T01, pose_mask=mask, pose_iters=np.array(it))
    ok, R, t, m, it = pose.pose_gn_mono(s["X"], s["pts_l1"], K, 5, np.eye(3), np.zeros(3), 0)
    reg.update(mono_R=R, mono_t=t, mono_mask=m, mono_iters=np.array(it))
    rng = np.random.default_rng(5)
    X = np.stack([rng.uniform(-10, 10, 64), rng.uniform(-3, 2, 64), rng.uniform(4, 40, 64)], 1)
    t10 = np.array([-synth.BASELINE_M, 0, 0], np.float32)
    p0 = np.stack([K[0] * X[:, 0] / X[:, 2] + K[2], K[1] * X[:, 1] / X[:, 2] + K[3]], 1).astype(np.float32)
    Xr = X + t10
    p1 = np.stack([K[0] * Xr[:, 0] / Xr[:, 2] + K[2], K[1] * Xr[:, 1] / Xr[:, 2] + K[3]], 1).astype(np.float32)
    X0, X1 = misc.triangulate_dlt(p0, p1, np.eye(3), t10, K, K)
    reg.update(tri_p0=p0, tri_p1=p1, tri_X0=X0, tri_X1=X1)
    p = synth.lba_problem(seed=4004, n_kf=5, n_points=40)
    rc, poses, pts, avg, ok = lba.lba_solve(p)
    reg.update(lba_poses=poses, lba_points=pts, lba_avg=avg)
    pt, m = klt.track_with_scale(img0, img1, pts0, np.full(len(pts0), 1.03, np.float32), prior)
    reg.update(kscale_p=pt, kscale_m=m)
    np.savez_compressed(os.path.join(HERE, "oracle_regression.npz"), **reg)
    for f in ("klt_cv2_golden.npz", "oracle_regression.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
