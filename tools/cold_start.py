"""Cold start: frame times of the very first frames of a process (no warm-up instance).  The library is the first CUDA user."""
import sys
import time
import numpy as np
sys.path.insert(0, ".")
from visual_odometry_ros_b200 import stereo_vo as svo, synth

W, H, K = synth.KITTI_W, synth.KITTI_H, synth.kitti_K()
t0 = time.perf_counter()
vo = svo.StereoVO(svo.make_parameters(W, H, K, K, synth.kitti_T_lr(), n_bins_u=64, n_bins_v=32))
print("constructor %.1f ms" % ((time.perf_counter() - t0) * 1e3))
L, R, T = synth.stereo_sequence(16, W, H, K, seed=3003, device="cuda")      # torch comes up after the library
ms = []
for k in range(16):
    t0 = time.perf_counter()
    vo.trackStereoImages(L[k], R[k], 0.1 * k)
    ms.append((time.perf_counter() - t0) * 1e3)
print("frame ms:", " ".join("%.2f" % m for m in ms))
