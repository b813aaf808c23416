"""Soak run: StereoVO and MonoVO over a long rendered sequence; per-100-frame means of the frame time (no creep expected)."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
from visual_odometry_ros_b200 import mono_vo as mvo, stereo_vo as svo, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
W, H, K = synth.KITTI_W, synth.KITTI_H, synth.kitti_K()
L, R, T = synth.stereo_sequence(n, W, H, K, seed=3003, device="cuda")
T0inv = np.linalg.inv(T[0])
for name in ("stereo", "mono"):
    vo = (svo.StereoVO(svo.make_parameters(W, H, K, K, synth.kitti_T_lr(), n_bins_u=64, n_bins_v=32)) if name == "stereo"
          else mvo.MonoVO(mvo.make_parameters(W, H, K, max_level=3, n_bins_u=64, n_bins_v=32)))
    ms, kf = [], []
    for k in range(n):
        t0 = time.perf_counter()
        if name == "stereo":
            vo.trackStereoImages(L[k], R[k], 0.1 * k)
        else:
            vo.trackImage(L[k], 0.1 * k)
        ms.append((time.perf_counter() - t0) * 1e3)
        kf.append(vo.frame_info()["keyframe"])
    ms, kf = np.asarray(ms), np.asarray(kf, bool)
    gt = T0inv @ T[n - 1]
    P = vo.pose()
    s = 1.0 if name == "stereo" else np.linalg.norm((T0inv @ T[1])[:3, 3]) / np.linalg.norm(vo.frame_pose(1)[:3, 3])
    drift = np.linalg.norm(s * P[:3, 3] - gt[:3, 3]) / np.linalg.norm(gt[:3, 3])
    print(name, "frames", n, "keyframes", int(kf.sum()), "drift %.4f" % drift, "stats consistent", vo.stats_consistent())
    for a in range(0, n, 100):
        seg, segk = ms[a:a + 100], kf[a:a + 100]
        print("  frames %4d-%4d: mean %.3f ms, non-kf %.3f, kf %.3f, max %.2f" % (a, min(a + 100, n) - 1, seg.mean(), seg[~segk].mean(), seg[segk].mean() if segk.any() else 0, seg.max()))
    vo.close()
