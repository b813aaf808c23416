"""Time of one strict-mode pose-GN call at N = 2000 with exactly 7 iterations (VO_POSE_NO_EARLY_STOP): the cost of the consumer
warp's chain of dependent FP32 additions."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from visual_odometry_ros_b200 import capi, synth
ctx = capi.Context(device=0, max_w=64, max_h=64, n_slots=0, max_feat=8192)
s = synth.pose_scene(seed=1001, n=2000)
K, Tlr = synth.kitti_K(), synth.kitti_T_lr()
fl = capi.VO_POSE_STRICT | capi.VO_POSE_NO_EARLY_STOP
for _ in range(5):
    r = ctx.pose_gn_stereo(s["X"], s["pts_l1"], s["pts_r1"], K, K, Tlr, 3.0, np.eye(4), flags=fl, max_iter=7)
t0 = time.perf_counter()
for _ in range(200):
    r = ctx.pose_gn_stereo(s["X"], s["pts_l1"], s["pts_r1"], K, K, Tlr, 3.0, np.eye(4), flags=fl, max_iter=7)
print("strict pose GN, ms per call (7 iterations, N=2000):", (time.perf_counter() - t0) / 200 * 1e3)
