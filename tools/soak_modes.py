"""Soak of both drop-in classes with every reference-arithmetic option on (strict pose sums, trackWithScale stale buffers)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from visual_odometry_ros_b200 import mono_vo as mvo, stereo_vo as svo, synth
n = 200
W, H, K = synth.KITTI_W, synth.KITTI_H, synth.kitti_K()
L, R, T = synth.stereo_sequence(n, W, H, K, seed=3003, device="cuda")
for name in ("stereo", "mono"):
    vo = (svo.StereoVO(svo.make_parameters(W, H, K, K, synth.kitti_T_lr(), n_bins_u=64, n_bins_v=32, pose_strict=True, scale_faithful_borders=True)) if name == "stereo"
          else mvo.MonoVO(mvo.make_parameters(W, H, K, max_level=3, n_bins_u=64, n_bins_v=32, pose_strict=True, scale_faithful_borders=True)))
    ms = []
    for k in range(n):
        t0 = time.perf_counter()
        (vo.trackStereoImages(L[k], R[k], 0.1 * k) if name == "stereo" else vo.trackImage(L[k], 0.1 * k))
        ms.append((time.perf_counter() - t0) * 1e3)
    gt = np.linalg.inv(T[0]) @ T[n - 1]
    P = vo.pose()
    s = 1.0 if name == "stereo" else np.linalg.norm((np.linalg.inv(T[0]) @ T[1])[:3, 3]) / np.linalg.norm(vo.frame_pose(1)[:3, 3])
    print(name, "all quirks on: mean %.3f ms, max %.2f, drift %.4f, stats consistent" % (np.mean(ms[2:]), np.max(ms[2:]), np.linalg.norm(s * P[:3, 3] - gt[:3, 3]) / np.linalg.norm(gt[:3, 3])), vo.stats_consistent())
